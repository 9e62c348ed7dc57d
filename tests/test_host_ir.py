"""Host-side mirror of the reference IR: builders, text formats, stats (reference fbs_exec_env.py, bit_exec_env.py)."""
import io
import json
import os

import numpy as np
import pytest

from conftest import GOLD, load_ref_mapped
from oracle import cleartext
from tfhe_fbs_map_b200 import BitExecEnv, LutExecEnv, FbsExecEnv
from tfhe_fbs_map_b200.formats import parse_blif, parse_bristol, read_lbf

DEMOS = json.load(open(os.path.join(GOLD, "demos.json")))


def build_demo():
    env = LutExecEnv()
    a, b, c = env.input("a"), env.input("b"), env.const(1)
    d = env.linear([1, 2], [a, b]); e = env.linear([1, 1], [c, d]); f = env.bootstrap(e, [1, 0, 1, 1, 0])
    g = env.linear([2, 1], [a, f]); h = env.bootstrap(g, [1, 1, 0, 2]); env.bootstrap(h, [1, 0, 1])
    env.output("f", f); env.output("g", g); env.output("h", h)
    return env


def test_alias():
    assert FbsExecEnv is LutExecEnv


def test_demo_program_text_and_outputs_match_reference():
    env = build_demo()
    s = io.StringIO(); env.print(os=s)
    assert s.getvalue() == DEMOS["fbs_exec_env_main"]["program"]
    got = cleartext.lut_eval(env, {"a": [1, 0], "b": [1, 0], "c": [1, 0]})
    assert {k: [int(x) for x in v] for k, v in got.items()} == DEMOS["fbs_exec_env_main"]["outputs"]


def test_builder_invariants():
    env = LutExecEnv()
    a, b = env.input("a"), env.input("b")
    l1 = env.linear([1, 2], [a, b])
    l2 = env.linear([1, 2], [a, b])
    assert l1 is l2                                  # structural de-dup (fbs_exec_env.py:93-100)
    assert env.max_val[l1.name] == 3
    with pytest.raises(AssertionError):
        env.bootstrap(l1, [0, 1, 1])                 # len(table) must be max_val+1 (fbs_exec_env.py:150)
    with pytest.raises(AssertionError):
        env.bootstrap(l1, [1, 1, 1, 1])              # min(table) == 0 (fbs_exec_env.py:86)
    bt = env.bootstrap(l1, [0, 1, 1, 0])
    neg = env.linear([-1], [bt], 1)
    assert env.max_val[neg.name] == 1
    nested = env.linear([2, 1], [neg, a], 0)         # flattening of nested lincombs (fbs_exec_env.py:137-140)
    assert [(c, v.name) for c, v in nested.coef_vals] == [(-2, bt.name), (1, "a")] and nested.const_coef == 2


@pytest.mark.parametrize("entry", load_ref_mapped(), ids=lambda e: f"{e['circuit']}-p{e['p']}-{e['mapper']}{'-strict' if e.get('strict') else ''}")
def test_lbf_round_trip_is_byte_identical(entry):
    env = read_lbf(entry["lbf"])
    s = io.StringIO(); env.write_lbf(os=s)
    assert s.getvalue() == entry["lbf"]


def test_fbs_text_format():
    env = LutExecEnv()
    a, b = env.input("a"), env.input("b")
    m1 = env.linear([-1, 1], [a, b], 1)
    m2 = env.bootstrap(m1, [1, 0, 1])
    env.output("out", m2)
    s = io.StringIO(); env.print(os=s, show_outputs=True)
    assert s.getvalue() == "m1 = -1 * a + 1 * b + 1\nm2 = Bootstrap(m1, [1, 0, 1])\nOutput out = m2\n"
    s = io.StringIO(); env.write_lbf(os=s)
    assert s.getvalue() == ".inputs a b\n.outputs out\n.lincomb a b m1\n-1 1 1\n.bootstrap m1 m2\n101\n.lincomb m2 out\n1\n"


def test_remove_dangling_and_stats():
    env = build_demo()
    before = len(env.instructions)
    env.remove_dangling_nodes()
    assert len(env.instructions) == before - 2       # m6 is not an output; m1 was flattened into m2 and is unused
    st = env.stats()
    assert st["nb_bootstrap"] == 2 and st["nb_inp"] == 2 and st["nb_out"] == 3 and st["max_lut_size"] == 5


def test_bit_env_builders_and_blif_round_trip():
    env = BitExecEnv()
    a, b = env.input("a"), env.input("b")
    assert env.op_and(a, env.CONST0) is env.CONST0 and env.op_and(env.CONST1, b) is b
    assert env.op_or(a, env.CONST1) is env.CONST1 and env.op_xor(env.CONST0, b) is b
    assert env.op_not(env.CONST0) is env.CONST1
    x = env.op_xor(a, b); n = env.op_not(x); y = env.op_lut([n, a], [0, 1, 0, 0])
    env.output("x", x); env.output("y", y)
    s = io.StringIO(); env.to_blif(fs=s, model_name="t")
    env2 = parse_blif(s.getvalue())
    iv = {"a": [0, 0, 1, 1], "b": [0, 1, 0, 1]}
    r1, r2 = cleartext.bit_eval(env, iv), cleartext.bit_eval(env2, iv)
    for k in r1:
        assert np.array_equal(r1[k], r2[k])
    assert env.stats()["nb_xor"] == 1 and env.stats()["nb_not"] == 1


def test_blif_polarity_and_constants():
    text = ".model t\n.inputs a b\n.outputs o z one\n.names a b o\n00 0\n11 0\n.names z\n.names one\n1\n.end\n"
    env = parse_blif(text)
    r = cleartext.bit_eval(env, {"a": [0, 0, 1, 1], "b": [0, 1, 0, 1]})
    assert list(r["o"]) == [0, 1, 1, 0] and r["z"] == 0 and r["one"] == 1


def test_bristol_reader():
    text = "3 7\n2 2 2\n1 1\n\n2 1 0 1 4 XOR\n2 1 2 3 5 AND\n2 1 4 5 6 XOR\n"
    env = parse_bristol(text)
    assert [i.name for i in env.inputs] == ["i_0", "i_1", "i_2", "i_3"] and list(env.outputs.keys()) == [6]
    iv = {f"i_{k}": np.array([(v >> k) & 1 for v in range(16)]) for k in range(4)}
    r = cleartext.bit_eval(env, iv)
    want = [((v & 1) ^ ((v >> 1) & 1)) ^ (((v >> 2) & 1) & ((v >> 3) & 1)) for v in range(16)]
    assert list(r[6]) == want


def test_eval_without_gpu_library_fails_loudly(monkeypatch):
    """No CPU fallback: the product path raises when the CUDA library / device is unavailable."""
    from tfhe_fbs_map_b200 import backend
    monkeypatch.setattr(backend, "LIB_PATH", "/nonexistent/libfbs_b200.so")
    monkeypatch.setattr(backend, "_lib", None)
    monkeypatch.setattr(backend, "_default", {})
    env = build_demo()
    with pytest.raises(RuntimeError):
        env.eval({"a": [1], "b": [0]})
