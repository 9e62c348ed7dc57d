"""levelize(): levels, slot liveness, table modes; the flat program interpreted in numpy must equal the oracle."""
import numpy as np
import pytest

from conftest import load_lbf_index, load_ref_mapped, read_golden_lbf, selfcheck_inputs
from oracle import cleartext
from tfhe_fbs_map_b200 import LutExecEnv, levelize, table_mode, min_fbs_size
from tfhe_fbs_map_b200.formats import read_lbf


def run_program_numpy(prog, in_bits):
    """Independent numpy interpreter of the flat arrays (what every kernel-side consumer must compute)."""
    a = prog.arrays
    B = in_bits.shape[1]
    vals = np.full((prog.n_slots, B), -99, dtype=np.int64)
    for i in range(prog.n_inputs):
        vals[a["in_slot"][i]] = in_bits[i]
    for lv in range(prog.n_levels):
        lcs = {}
        for q in range(a["lc_level_ptr"][lv], a["lc_level_ptr"][lv + 1]):
            acc = np.full(B, a["lc_const"][q], dtype=np.int64)
            for o in range(a["lc_ptr"][q], a["lc_ptr"][q + 1]):
                assert np.all(vals[a["lc_slot"][o]] >= 0), "lincomb reads a slot that holds no live wire"
                acc = acc + a["lc_coef"][o] * vals[a["lc_slot"][o]]
            lcs[q] = acc
        new = {}
        for q in range(a["bs_level_ptr"][lv], a["bs_level_ptr"][lv + 1]):
            tab = a["bs_tab"][a["bs_tab_ptr"][q]:a["bs_tab_ptr"][q + 1]].astype(np.int64)
            idx = lcs[a["bs_lc"][q]]
            assert idx.min() >= 0 and idx.max() < len(tab)
            new[a["bs_slot"][q]] = tab[idx]
        for s, v in new.items():       # all lincombs of a level run before its bootstraps write
            vals[s] = v
    out = []
    for q in range(len(prog.output_names)):
        acc = np.full(B, a["out_const"][q], dtype=np.int64)
        for o in range(a["out_ptr"][q], a["out_ptr"][q + 1]):
            acc = acc + a["out_coef"][o] * vals[a["out_slot"][o]]
        out.append(acc)
    return np.array(out).reshape(len(prog.output_names), B)


def test_table_modes():
    p = 5
    assert table_mode([0, 1, 1], p) == 1
    assert table_mode([0, 1, 1, 0, 1] + [1, 0, 0], p) == 1          # neg
    assert table_mode([0, 0, 1, 0, 1] + [0, 0], p) == 0             # zero
    assert table_mode([1, 1, 1, 0, 1] + [1, 1, 1], p) == 2          # one
    with pytest.raises(ValueError):
        table_mode([0, 1, 1, 0, 1] + [1, 1], p)                     # mixed
    with pytest.raises(ValueError):
        table_mode([0] * 11, p)


@pytest.mark.parametrize("item", load_lbf_index(), ids=lambda e: e["file"])
@pytest.mark.parametrize("mode", ["reuse", "contiguous", "pad4"])
def test_flat_program_equals_oracle(item, mode):
    env = read_golden_lbf(item["file"])
    kw = dict(reuse=dict(reuse_slots=True), contiguous=dict(reuse_slots=False), pad4=dict(shard_pad=4))[mode]
    prog = levelize(env, item["p"], **kw)
    inputs = selfcheck_inputs(item["input_names"], batch=200)
    want = cleartext.lut_eval(env, inputs)
    in_bits = np.array([inputs[nm] for nm in prog.input_names], dtype=np.uint8)
    got = run_program_numpy(prog, in_bits)
    for nm in prog.output_names:
        assert np.array_equal(got[prog.out_index[nm]], np.asarray(want[nm])), nm
    st = item["stats"]
    assert prog.n_boots == st["nb_bootstrap"] and prog.n_inputs == st["nb_inp"]
    assert prog.n_lincombs <= st["nb_linprod"]
    a = prog.arrays
    # bootstraps of a level are sorted by lincomb and only use lincombs of their own level
    for lv in range(prog.n_levels):
        lcs = a["bs_lc"][a["bs_level_ptr"][lv]:a["bs_level_ptr"][lv + 1]]
        assert np.all(np.diff(lcs) >= 0)
        assert lcs.min() >= a["lc_level_ptr"][lv] and lcs.max() < a["lc_level_ptr"][lv + 1]
    if mode == "reuse":
        assert prog.n_slots <= prog.n_inputs + prog.n_boots
    if mode == "pad4":
        for lv in range(prog.n_levels):
            sl = a["bs_slot"][a["bs_level_ptr"][lv]:a["bs_level_ptr"][lv + 1]]
            assert np.array_equal(sl, np.arange(sl[0], sl[0] + len(sl)))     # level-contiguous for in-place allgather


def test_slot_reuse_saves_memory_on_deep_circuits():
    env = read_golden_lbf("adder128_p15.lbf")
    reuse = levelize(env, 15)
    flat = levelize(env, 15, reuse_slots=False)
    assert reuse.n_levels == flat.n_levels == 127
    assert reuse.n_slots < flat.n_slots


def test_min_fbs_size_and_mode_detection():
    for e in load_ref_mapped():
        if e["mapper"] != "search":
            continue
        env = read_lbf(e["lbf"])
        assert min_fbs_size(env) <= e["p"]
        prog = levelize(env, e["p"])
        assert set(prog.arrays["bs_mode"][:prog.n_boots].tolist()) <= {0, 1, 2}


def test_shared_lincomb_is_keyswitched_once():
    e = next(x for x in load_ref_mapped() if x["circuit"] == "_2_input_gates" and x["p"] == 15)
    prog = levelize(read_lbf(e["lbf"]), 15)
    assert prog.n_boots == 10 and prog.n_lincombs == 2       # 10 tables over 2 shared lincombs


def test_outputs_as_lincombs():
    env = LutExecEnv()
    a, b = env.input("a"), env.input("b")
    x = env.bootstrap(env.linear([1, 1], [a, b]), [0, 1, 0])
    env.output("neg", env.linear([-1], [x], 1)); env.output("pass", a); env.output("one", env.const(1)); env.output("x", x)
    prog = levelize(env, 3)
    bits = np.array([[0, 0, 1, 1], [0, 1, 0, 1]], dtype=np.uint8)
    got = run_program_numpy(prog, bits)
    assert got[prog.out_index["neg"]].tolist() == [1, 0, 0, 1]
    assert got[prog.out_index["pass"]].tolist() == [0, 0, 1, 1]
    assert got[prog.out_index["one"]].tolist() == [1, 1, 1, 1]


@pytest.mark.parametrize("circuit,p,want", [("_2_input_gates", 15, (10, 2)), ("_2_input_gates", 11, (10, 2)), ("full_adder", 15, None)])
def test_multi_value_groups_partition_every_level(circuit, p, want):
    """levelize(multi_value=True): groups = maximal runs of a level's bootstraps on one lincomb (reference fbs_exec_env.py:93-100
    de-duplicates LinearProds, so tables on equal lincombs share it); they partition the level and never cross levels."""
    e = next(x for x in load_ref_mapped() if x["circuit"] == circuit and x["p"] == p and x["mapper"] == "search")
    from tfhe_fbs_map_b200.formats import read_lbf
    prog = levelize(read_lbf(e["lbf"]), p, multi_value=True)
    a = prog.arrays
    gf, gl = a["grp_first"], a["grp_level_ptr"]
    assert prog.multi_value and prog.n_rotations == prog.n_groups == len(gf) - 1 and gf[0] == 0 and gf[-1] == prog.n_boots
    if want:
        assert (prog.n_boots, prog.n_groups) == want
    for lv in range(prog.n_levels):
        assert gf[gl[lv]] == a["bs_level_ptr"][lv] or gl[lv] == gl[lv + 1]
        for g in range(gl[lv], gl[lv + 1]):
            q0, q1 = gf[g], gf[g + 1]
            assert q0 < q1 <= a["bs_level_ptr"][lv + 1]
            assert len(set(a["bs_lc"][q0:q1])) == 1
            if g + 1 < gl[lv + 1]:
                assert a["bs_lc"][q1] != a["bs_lc"][q0]
    d = prog.c_desc()
    assert d.n_groups == prog.n_groups
    plain = levelize(read_lbf(e["lbf"]), p)
    assert not plain.multi_value and plain.c_desc().n_groups == 0 and plain.n_rotations == plain.n_boots
