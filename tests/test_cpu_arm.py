"""The tuned CPU arm of the benchmark (baseline/cpu_pbs.cpp) computes the specification exactly: its ciphertexts equal the
oracle's bit for bit, and a whole program decrypts to the cleartext interpreter's outputs.  (CPU only.)"""
import numpy as np
import pytest

from conftest import load_ref_mapped, selfcheck_inputs, unpack_outputs
from oracle.tfhe_ref import RefTFHE
from tfhe_fbs_map_b200 import levelize, params
from tfhe_fbs_map_b200.formats import read_lbf


@pytest.mark.parametrize("name", ["toy3v", "toy3u", "toy7u", "toy5v"])
def test_cpu_arm_pbs_bit_exact_against_oracle(name):
    from baseline.cpu_arm import CpuTFHE
    ps = params.get(name)
    ref = RefTFHE(ps, seed=31)
    cpu = CpuTFHE(ps, ref)
    p = 7
    cases = [([0, 1, 1, 0, 1, 0, 0], 1), ([0, 1, 1, 0, 1, 0, 0, 1, 0, 0, 1, 0, 1, 1], 1), ([0, 1, 1, 0, 1, 0, 0, 0], 0), ([1, 1, 0, 0, 1, 0, 1, 1, 1], 2)]
    msgs, rows, lens, modes = [], [], [], []
    for tab, md in cases:
        for m in range(len(tab)):
            row = np.zeros(2 * p, np.uint8); row[:len(tab)] = tab
            msgs.append(m); rows.append(row); lens.append(len(tab)); modes.append(md)
    cts = ref.encrypt(p, np.array(msgs, np.int32), np.arange(len(msgs)), 3)
    got = cpu.pbs_batch(p, cts, np.array(rows), np.array(lens, np.uint8), np.array(modes, np.int32))
    want, _ = ref.pbs_batch(p, cts, np.array(rows), np.array(lens, np.uint8), np.array(modes, np.int32), want_acc=False)
    assert np.array_equal(got, want)
    assert ref.decrypt(p, got).tolist() == [int(rows[i][msgs[i]]) for i in range(len(msgs))]


def test_cpu_arm_program_equals_oracle_and_cleartext():
    from baseline.cpu_arm import CpuTFHE
    ps = params.get("toy3v")
    ref = RefTFHE(ps, seed=32)
    cpu = CpuTFHE(ps, ref)
    e = next(x for x in load_ref_mapped() if x["circuit"] == "aes_sbox" and x["p"] == 11 and x["mapper"] == "search" and not x.get("strict"))
    prog = levelize(read_lbf(e["lbf"]), 11)
    B = 6
    inputs = selfcheck_inputs(e["input_names"])
    bits = np.array([inputs[nm][:B] for nm in prog.input_names], dtype=np.uint8)
    got = cpu.eval_prog(prog, bits, enc_seed=4, threads=2)
    assert np.array_equal(got, ref.eval_prog(prog, bits, enc_seed=4))
    want = unpack_outputs(e, batch=B)
    for nm in prog.output_names:
        assert np.array_equal(got[prog.out_index[nm]], want[str(nm)]), nm
