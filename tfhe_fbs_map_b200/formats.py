"""Readers for the circuit formats around the hot path.

* BLIF ``.names`` subset and Bristol Fashion -> ``BitExecEnv`` (reference fbs_mapper/map_circuit.py:12-89, which
  delegates to the third-party ``blifparser==2.0.1`` / ``bfcl==1.0.1`` packages; neither is installed here, so the
  subset the reference actually consumes is parsed directly).
* ``.lbf`` -> ``LutExecEnv``.  The reference only *writes* this format (fbs_exec_env.py:170-206); reading it back
  lets pre-mapped circuits be executed without re-running the mapper.

``env_cls`` lets the same reader build the reference's own classes (used by oracle/gen_golden.py to feed the
reference mapper), since both expose the same builder methods.
"""
from __future__ import annotations

import re

from .bit_env import BitExecEnv
from .lut_env import Bootstrap, LinearProd, LutExecEnv


def _logical_lines(text):
    """BLIF/LBF lines with ``\\`` continuations joined and comments stripped."""
    buf = ""
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].rstrip()
        if not line.strip():
            continue
        if line.endswith("\\"):
            buf += line[:-1] + " "
            continue
        yield (buf + line).strip()
        buf = ""
    if buf.strip():
        yield buf.strip()


def truth_table_from_rows(rows):
    """Rows of a ``.names`` gate -> truth table, MSB-first index (reference map_circuit.py:12-22): all rows share
    the output polarity of the first row; unlisted assignments take the opposite value."""
    first = rows[0]
    n = len(first) - 1
    tte = 0 if first[-1] == "1" else 1
    tt = [tte] * (2 ** n)
    for r in rows:
        assert len(r) == len(first)
        k = sum(int(r[i]) * 2 ** (n - i - 1) for i in range(n))
        tt[k] = 1 - tte
    return tt


def parse_blif(text: str, env_cls=BitExecEnv):
    """BLIF text -> BitExecEnv (1- and 2-input ``.names`` gates and constants, map_circuit.py:25-50)."""
    env = env_cls()
    wires = {}
    outputs = []
    gates = []          # (inputs, output, rows)
    cur = None
    for line in _logical_lines(text):
        tok = line.split()
        if tok[0] == ".inputs":
            for nm in tok[1:]:
                wires[nm] = env.input(nm)
            cur = None
        elif tok[0] == ".outputs":
            outputs.extend(tok[1:])
            cur = None
        elif tok[0] == ".names":
            cur = (tok[1:-1], tok[-1], [])
            gates.append(cur)
        elif tok[0].startswith("."):
            cur = None          # .model / .end / anything else
        elif cur is not None:
            cur[2].append("".join(tok))     # "11 1" -> "111"
    for ins, out, rows in gates:
        if not rows:                        # ".names x" with no rows is constant 0
            wires[out] = env.CONST0
            continue
        tt = truth_table_from_rows(rows)
        if tt == [0]:
            wires[out] = env.CONST0
        elif tt == [1]:
            wires[out] = env.CONST1
        else:
            assert len(tt) in (2, 4), f"only 1- and 2-input gates are supported: {ins} -> {out}"
            wires[out] = env.op_lut([wires[k] for k in ins], tt, name=out)
    for out in outputs:
        env.output(out, wires[out])
    return env


def parse_blif_file(path, env_cls=BitExecEnv):
    with open(path) as f:
        return parse_blif(f.read(), env_cls)


_BRISTOL_OPS = {"AND": [0, 0, 0, 1], "XOR": [0, 1, 1, 0], "OR": [0, 1, 1, 1], "INV": [1, 0], "NOT": [1, 0]}


def parse_bristol(text: str, env_cls=BitExecEnv):
    """Bristol Fashion text -> BitExecEnv with the naming of reference map_circuit.py:53-89 (inputs ``i_<wire>``,
    gate wires ``w_<wire>``, outputs keyed by integer wire index, EQW = wire copy)."""
    lines = [l.split() for l in text.splitlines() if l.strip()]
    n_gates, n_wires = int(lines[0][0]), int(lines[0][1])
    in_counts = [int(x) for x in lines[1][1:]]
    out_counts = [int(x) for x in lines[2][1:]]
    env = env_cls()
    wires = {}
    for idx in range(sum(in_counts)):
        wires[idx] = env.input(f"i_{idx}")
    for g in lines[3:3 + n_gates]:
        n_in, n_out = int(g[0]), int(g[1])
        ins = [int(x) for x in g[2:2 + n_in]]
        out = int(g[2 + n_in])
        op = g[-1].upper()
        if op in ("EQW", "EQ"):
            wires[out] = wires[ins[0]]
        else:
            wires[out] = env.op_lut([wires[i] for i in ins], list(_BRISTOL_OPS[op]), name=f"w_{out}")
    first_out = n_wires - sum(out_counts)
    for idx in range(first_out, n_wires):
        env.output(idx, wires[idx])
    return env


def read_lbf(text: str) -> LutExecEnv:
    """``.lbf`` text (fbs_exec_env.py:170-206) -> LutExecEnv, keeping the node names of the file.

    The writer ends the file with exactly one identity ``.lincomb <node> <output>`` per output, in ``.outputs``
    order (fbs_exec_env.py:204-206); output names may coincide with input names (e.g. ascon_lut), so those
    trailing records are recognised by position, not by name."""
    env = LutExecEnv()
    nodes = {}
    out_names = []
    records = []
    lines = list(_logical_lines(text))
    i = 0
    while i < len(lines):
        tok = lines[i].split()
        if tok[0] == ".inputs":
            for nm in tok[1:]:
                nodes[nm] = env.input(nm)
            i += 1
        elif tok[0] == ".outputs":
            out_names = tok[1:]
            i += 1
        elif tok[0] in (".lincomb", ".bootstrap"):
            records.append((tok, lines[i + 1]))
            i += 2
        else:
            raise ValueError(f"unknown .lbf line: {lines[i]}")
    n_body = len(records) - len(out_names)
    assert n_body >= 0, "missing output records"
    max_id = 0
    for tok, data in records[:n_body]:
        res = tok[-1]
        if tok[0] == ".lincomb":
            ins = tok[1:-1]
            nums = [int(x) for x in data.split()]
            coefs, const = nums[:len(ins)], (nums[len(ins)] if len(nums) > len(ins) else 0)
            nodes[res] = env._add_instr(LinearProd(res, list(zip(coefs, [nodes[k] for k in ins])), const))
        else:
            src = tok[1]
            table = [int(ch) for ch in data.strip()]
            assert len(table) == env.max_val[nodes[src].name] + 1, f"table/lincomb range mismatch at {res}"
            nodes[res] = env._add_instr(Bootstrap(res, nodes[src], table))
        m = re.fullmatch(r"m(\d+)", res)
        if m:
            max_id = max(max_id, int(m.group(1)))
    for (tok, data), name in zip(records[n_body:], out_names):
        assert tok[0] == ".lincomb" and tok[-1] == name and data.split() == ["1"], f"bad output record for {name}"
        src = tok[1]
        if src in nodes:
            env.output(name, nodes[src])
        else:                                   # constant output: the reference writes the Const's name ("0"/"1")
            env.output(name, env.const(int(src)))
    env._unique_id = max_id
    return env


def read_lbf_file(path) -> LutExecEnv:
    with open(path) as f:
        return read_lbf(f.read())
