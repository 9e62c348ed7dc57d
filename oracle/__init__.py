"""CPU oracle: TEST INFRASTRUCTURE ONLY (see oracle/tfhe_ref.c and oracle/cleartext.py headers).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product package tfhe_fbs_map_b200 never does."""
