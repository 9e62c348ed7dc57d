"""Cleartext restatement of the reference interpreters.  TEST INFRASTRUCTURE ONLY.

``lut_eval``  follows reference fbs_mapper/fbs_exec_env.py:208-229 (LutExecEnv.eval)
``bit_eval``  follows reference fbs_mapper/bit_exec_env.py:173-194 (BitExecEnv.eval)

Both are duck-typed on class *names* so they accept the reference's own node classes (when
/root/reference is importable, to pin this file against the reference: tests/test_oracle_pinned.py and
oracle/gen_golden.py) as well as the product's IR classes.  Pinned against the golden vectors in
tests/golden/*.json (generated from the reference by oracle/gen_golden.py).
Plain numpy; the table look-up uses fancy indexing instead of the reference's per-element lambda
(fbs_exec_env.py:219-220), which is value-identical.
"""
import numpy as np


def _kind(instr):
    return type(instr).__name__


def lut_eval(env, input_values):
    wire = {"0": 0, "1": 1}                                           # fbs_exec_env.py:209
    for instr in env.instructions:                                    # :211 build order is topological
        k = _kind(instr)
        if k == "Input":
            val = np.array(input_values[instr.name]).reshape(-1).astype(np.int64)   # :213-214 (int64 like randint's output)
        elif k == "LinearProd":
            val = np.sum([c * wire[v.name] for c, v in instr.coef_vals], axis=0) + instr.const_coef   # :215-217
        elif k == "Bootstrap":
            val = np.asarray(instr.table, dtype=int)[np.asarray(wire[instr.val.name])]                # :218-220
        else:
            raise AssertionError("Unknown instruction")
        wire[instr.name] = val
    return {name: wire[out.name] for name, out in env.outputs.items()}   # :225-229


def lut_eval_literal(env, input_values):
    """The reference's hot loop LITERALLY (fbs_exec_env.py:208-229): per-element Python lambda through ``np.fromiter`` for
    every bootstrap (:218-220), ``np.sum(list(map(lambda ...)))`` for every lincomb (:215-217).  This is what the reference's
    CPU path costs; ``lut_eval`` above is value-identical but vectorised.  Used for bench.py's config-1 baseline line when
    /root/reference is not mounted (the GPU box)."""
    wire_values = {"0": 0, "1": 1}
    for instr in env.instructions:
        k = _kind(instr)
        if k == "Input":
            val = np.array(input_values[instr.name]).reshape(-1)
        elif k == "LinearProd":
            val = np.sum(list(map(lambda cn: cn[0] * wire_values[cn[1].name], instr.coef_vals)), axis=0) + instr.const_coef
        elif k == "Bootstrap":
            table = instr.table
            val = np.fromiter(map(lambda v: table[v], wire_values[instr.val.name]), dtype=int)
        else:
            raise AssertionError("Unknown instruction")
        wire_values[instr.name] = val
    return {name: wire_values[out.name] for name, out in env.outputs.items()}


def bit_eval(env, input_values):
    wire = {"0": 0, "1": 1}                                           # bit_exec_env.py:174
    for instr in env.instructions:
        k = _kind(instr)
        if k in ("Const", "BConst"):
            continue
        if k in ("Input", "BInput"):
            val = np.array(input_values[instr.name]).reshape(-1).astype(np.int64)   # :180-181
        else:                                                         # any LUT subclass, :182-185
            idx = sum(wire[inp.name] * 2 ** e for e, inp in enumerate(instr.inputs[::-1]))   # first input = MSB
            val = np.asarray(instr.truth_table, dtype=int)[np.asarray(idx)]
        wire[instr.name] = val
    return {name: wire[out.name] for name, out in env.outputs.items()}   # :190-194


def selfcheck_inputs(input_names, batch=1000, seed=42):
    """The CLI's random-input protocol, reference fbs_mapper/map_circuit.py:137-139."""
    np.random.seed(seed)
    return {nm: np.random.randint(0, 2, (batch)) for nm in input_names}
