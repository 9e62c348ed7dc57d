"""The cleartext oracle (oracle/cleartext.py) against golden vectors produced by the reference itself
(oracle/gen_golden.py): reference fbs_exec_env.py:208-229 and bit_exec_env.py:173-194 under the CLI protocol
map_circuit.py:137-140,174-180."""
import io
import os

import numpy as np
import pytest

from conftest import GOLD, load_ref_mapped, out_hash, read_golden_blif, read_golden_lbf, selfcheck_inputs, unpack_outputs, load_lbf_index
from oracle import cleartext
from tfhe_fbs_map_b200.formats import read_lbf

ENTRIES = load_ref_mapped()


@pytest.mark.parametrize("entry", ENTRIES, ids=lambda e: f"{e['circuit']}-p{e['p']}-{e['mapper']}{'-strict' if e.get('strict') else ''}")
def test_lut_eval_matches_reference_outputs(entry):
    env = read_lbf(entry["lbf"])
    inputs = selfcheck_inputs(entry["input_names"])
    got = cleartext.lut_eval(env, inputs)
    want = unpack_outputs(entry)
    assert list(map(str, got.keys())) == list(want.keys())
    for k in got:
        assert np.array_equal(np.asarray(got[k]), want[str(k)]), k
    assert out_hash(got) == entry["out_sha256"]
    assert env.stats() == entry["stats"]


@pytest.mark.parametrize("name", sorted({e["circuit"] for e in ENTRIES}))
def test_bit_eval_matches_reference_outputs(name):
    entry = next(e for e in ENTRIES if e["circuit"] == name)
    env = read_golden_blif(name)
    assert [i.name for i in env.inputs] == entry["input_names"]
    got = cleartext.bit_eval(env, selfcheck_inputs(entry["input_names"]))
    want = unpack_outputs(entry)
    for k in got:
        assert np.array_equal(np.asarray(got[k]), want[str(k)]), k


@pytest.mark.parametrize("item", load_lbf_index(), ids=lambda e: e["file"])
def test_premapped_lbf_fixture_hash(item):
    env = read_golden_lbf(item["file"])
    got = cleartext.lut_eval(env, selfcheck_inputs(item["input_names"]))
    assert out_hash(got) == item["out_sha256"]
    assert env.stats() == item["stats"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/fbs_mapper"), reason="reference tree not mounted")
def test_oracle_pinned_against_live_reference():
    """Where /root/reference is mounted, run the reference's own eval next to the restatement."""
    import sys
    sys.path.insert(0, "/root/reference/fbs_mapper")
    import bit_exec_env as ref_bit
    from tfhe_fbs_map_b200.formats import parse_blif
    for name in ("full_adder", "aes_sbox", "ascon_lut"):
        text = open(os.path.join(GOLD, "blif", f"{name}.blif")).read()
        ref_env = parse_blif(text, env_cls=ref_bit.BitExecEnv)
        inputs = selfcheck_inputs([i.name for i in ref_env.inputs], batch=300, seed=7)
        want = ref_env.eval(inputs)
        got = cleartext.bit_eval(ref_env, inputs)          # restatement on the reference's own classes
        for k in want:
            assert np.array_equal(want[k], got[k])
