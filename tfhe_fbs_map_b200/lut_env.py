"""FBS-circuit IR with the reference's ``LutExecEnv`` surface, executed on B200.

Host-side mirror of reference ``fbs_mapper/fbs_exec_env.py:11-276``: same builder methods, node classes,
de-duplication, value-bound tracking, ``.fbs`` / ``.lbf`` writers and statistics, so code written against
the reference (the mapper, the CLI) runs unchanged.  What differs is ``eval``: the reference interprets the
circuit in cleartext with a Python loop (fbs_exec_env.py:208-229); here ``eval`` hands the levelised
program to the CUDA library behind ``include/fbs_b200.h``:

* ``eval(inputs)``        -- encrypted evaluation (real TFHE: one batched programmable bootstrap per level);
* ``eval_clear(inputs)``  -- the same program evaluated by the cleartext CUDA kernel (table look-ups), the
  literal GPU counterpart of the reference's hot loop at fbs_exec_env.py:218-220.

Neither has a CPU fallback: without the compiled library and a GPU they raise.
"""
from __future__ import annotations

import logging
import sys
import textwrap

import numpy as np


class Node:
    """Base wire type (reference fbs_exec_env.py:12-20: equality by repr, hash by name)."""

    def __init__(self, name):
        self.name = name

    def __eq__(self, other):
        return repr(self) == repr(other)

    def __hash__(self):
        return hash(self.name)


class Const(Node):
    def __init__(self, value):
        super().__init__(f"{value}")
        self.value = value

    def __str__(self):
        return f"{self.value}"


class Input(Node):
    def __str__(self):
        return f"Input({self.name})"


class LinearProd(Node):
    """sum(coef * wire) + const_coef (reference fbs_exec_env.py:37-49)."""

    def __init__(self, name, coef_vals, const_coef=0):
        super().__init__(name)
        for _, v in coef_vals:
            assert isinstance(v, Node), "Expected 'Node' type"
        self.coef_vals = coef_vals
        self.const_coef = const_coef

    def __str__(self):
        # textual form is also the de-dup key, so it must match the reference byte for byte
        terms = " + ".join(f"{c} * {v.name}" for c, v in self.coef_vals)
        tail = f"+ {self.const_coef}" if self.const_coef != 0 else ""
        return f"{terms} {tail}"


class Bootstrap(Node):
    """table[val]: one functional bootstrap (reference fbs_exec_env.py:51-61)."""

    def __init__(self, name, val, table):
        super().__init__(name)
        assert isinstance(table, list), "Expected list"
        assert isinstance(val, Node), "Expected LutExecEnv.Node"
        self.table = table
        self.val = val

    def __str__(self):
        return f"Bootstrap({self.val.name}, {self.table})"


class LutExecEnv:
    # nested aliases so ``LutExecEnv.Input`` etc. keep working (reference spells them as inner classes)
    Node = Node
    Const = Const
    Input = Input
    LinearProd = LinearProd
    Bootstrap = Bootstrap

    def __init__(self, merge_linear_prods=True):
        self._unique_id = 0
        self.instructions = []
        self.outputs = {}
        self._merge_linear_prods = merge_linear_prods
        self.max_val = {}
        self.instr_cache = {}
        self.logger = logging.getLogger("LutExecEnv")
        self._compiled = None   # cache: (key, CompiledProgram)

    # ------------------------------------------------------------------ builders
    def _new_id(self):
        self._unique_id += 1
        return f"m{self._unique_id}"

    def _bound(self, instr):
        """Largest value the wire can take (reference fbs_exec_env.py:76-91)."""
        assert instr.name not in self.max_val, "Error"
        if isinstance(instr, Input):
            mv = 1
        elif isinstance(instr, LinearProd):
            mv = instr.const_coef + sum(max(0, c * self.max_val[v.name]) for c, v in instr.coef_vals)
        elif isinstance(instr, Bootstrap):
            assert min(instr.table) == 0
            mv = max(instr.table)
        else:
            assert False, "Unknown instruction"
        self.max_val[instr.name] = mv
        self.logger.getChild("_set_value_bounds").info(f"{instr.name} {mv}")

    def _add_instr(self, instr):
        self.logger.getChild("_add_instr").info(f"{instr.name} = {instr}")
        key = str(instr)
        hit = self.instr_cache.get(key)
        if hit is not None:
            return hit
        self.instr_cache[key] = instr
        self.instructions.append(instr)
        self._bound(instr)
        self._compiled = None
        return instr

    def input(self, input_id):
        return self._add_instr(Input(input_id))

    def const(self, value):
        return Const(value)

    def linear(self, coefs, vals, const_coef=0):
        """Flattening / constant folding as reference fbs_exec_env.py:131-145."""
        flat = []
        for coef, val in zip(coefs, vals):
            assert isinstance(val, Node), "Expected LutExecEnv.Node"
            if isinstance(val, LinearProd) and self._merge_linear_prods:
                flat.extend((coef * c1, v1) for c1, v1 in val.coef_vals)
                const_coef += coef * val.const_coef
            elif isinstance(val, Const):
                const_coef += coef * val.value
            else:
                flat.append((coef, val))
        return self._add_instr(LinearProd(self._new_id(), flat, const_coef))

    def bootstrap(self, val, table):
        assert isinstance(val, Node), "Expected LutExecEnv.Node"
        assert isinstance(table, list), "Expected list"
        assert len(table) == self.max_val[val.name] + 1, f"{table} vs {val.name} {self.max_val[val.name]}"
        table = [int(t) for t in table]   # numpy scalars would print as np.int64(..) under numpy>=2
        return self._add_instr(Bootstrap(self._new_id(), val, table))

    def output(self, name, val):
        assert isinstance(val, Node), "Expected LutExecEnv.Node"
        self.outputs[name] = val
        self._compiled = None

    # ------------------------------------------------------------------ text formats
    def print(self, os=sys.stdout, show_inputs=False, show_outputs=False):
        """``.fbs`` text (reference fbs_exec_env.py:158-168)."""
        for instr in self.instructions:
            if isinstance(instr, Input) and not show_inputs:
                continue
            print(f"{instr.name} = {str(instr)}", file=os)
        if show_outputs:
            for name, val in self.outputs.items():
                print(f"Output {name} = {val.name}", file=os)

    def write_lbf(self, os=sys.stdout):
        """``.lbf`` text (reference fbs_exec_env.py:170-206)."""
        def wrapped(line):
            return " \\\n ".join(textwrap.wrap(line))

        in_names = [i.name for i in self.instructions if isinstance(i, Input)]
        print(wrapped(f".inputs {' '.join(in_names)}"), file=os)
        print(wrapped(f".outputs {' '.join(map(str, self.outputs.keys()))}"), file=os)
        for instr in self.instructions:
            if isinstance(instr, Input):
                continue
            if isinstance(instr, LinearProd):
                cv = sorted(instr.coef_vals, key=lambda e: e[1].name)
                tail = f"{instr.const_coef}" if instr.const_coef != 0 else ""
                print(f".lincomb {' '.join(v.name for _, v in cv)} {instr.name}", file=os)
                print(f"{' '.join(str(c) for c, _ in cv)} {tail}", file=os)
            elif isinstance(instr, Bootstrap):
                print(f".bootstrap {instr.val.name} {instr.name}", file=os)
                print("".join(map(str, instr.table)), file=os)
            else:
                assert False, "Unknown instruction"
        for out, val in self.outputs.items():
            print(f".lincomb {val.name} {out}", file=os)
            print("1", file=os)

    # ------------------------------------------------------------------ graph utilities
    def remove_dangling_nodes(self):
        """Dead-code elimination from the outputs (reference fbs_exec_env.py:231-243)."""
        live = {o.name for o in self.outputs.values()}
        for instr in reversed(self.instructions):
            if instr.name not in live:
                continue
            if isinstance(instr, LinearProd):
                live.update(v.name for _, v in instr.coef_vals)
            elif isinstance(instr, Bootstrap):
                live.add(instr.val.name)
        self.instructions = [i for i in self.instructions if i.name in live]
        self._compiled = None

    def stats(self):
        """Same keys as reference fbs_exec_env.py:245-276."""
        nb_inp = nb_linprod = nb_bootstrap = max_lut_size = 0
        norm2 = {}
        for instr in self.instructions:
            if isinstance(instr, Input):
                nb_inp += 1
                norm2[instr.name] = 1
            elif isinstance(instr, LinearProd):
                nb_linprod += 1
                norm2[instr.name] = sum(c * c * norm2[v.name] for c, v in instr.coef_vals)
            elif isinstance(instr, Bootstrap):
                nb_bootstrap += 1
                max_lut_size = max(max_lut_size, len(instr.table))
                norm2[instr.name] = 1
            else:
                assert False, "Unknown instruction"
        return dict(nb_inp=nb_inp, nb_linprod=nb_linprod, nb_bootstrap=nb_bootstrap, max_lut_size=max_lut_size,
                    norm2_linprod=max(norm2.values()), nb_out=len(self.outputs))

    # ------------------------------------------------------------------ execution (GPU only)
    def input_names(self):
        return [i.name for i in self.instructions if isinstance(i, Input)]

    def _gather_inputs(self, input_values):
        names = self.input_names()
        cols = [np.asarray(input_values[nm]).reshape(-1) for nm in names]
        B = max((len(c) for c in cols), default=1)
        mat = np.empty((len(names), B), dtype=np.uint8)
        for r, c in enumerate(cols):
            assert len(c) in (1, B), "inputs must share one batch length"
            assert np.all((c == 0) | (c == 1)), "inputs are bits (reference fbs_exec_env.py:80)"
            mat[r, :] = c
        return names, mat, B

    def _format_outputs(self, prog, out_mat):
        res = {}
        for name, node in self.outputs.items():
            if isinstance(node, Const):          # reference returns the bare Python scalar (fbs_exec_env.py:209,227)
                res[name] = node.value
            else:
                res[name] = out_mat[prog.out_index[name]].astype(np.int64)
        return res

    def eval(self, input_values, fbs_size=None, backend=None, multi_value=False):
        """Encrypted evaluation on the GPU; same contract as reference fbs_exec_env.py:208-229.

        ``fbs_size`` (p) defaults to the smallest p for which every table is realisable; ``backend`` is a
        :class:`tfhe_fbs_map_b200.backend.B200Backend` (keys + device), default = process-wide backend.
        ``multi_value``: evaluate all tables that share a LinearProd (the builder's de-duplication, reference :93-100) with ONE
        blind rotation (multi-value bootstrap); same decrypted results, fewer rotations, slightly more noise
        (``ParamSet.p_fail(p, norm2, mv_norm2=p + 3)``).
        """
        from . import backend as _be
        be = backend if backend is not None else _be.default_backend()
        prog = be.compile(self, fbs_size, multi_value=multi_value)
        _, mat, B = self._gather_inputs(input_values)
        out = be.eval_bits(prog, mat)
        return self._format_outputs(prog, out)

    def eval_clear(self, input_values, backend=None):
        """Cleartext evaluation of the same levelised program by the CUDA table look-up kernel."""
        from . import backend as _be
        be = backend if backend is not None else _be.default_backend(need_keys=False)
        prog = be.compile(self, None, clear=True)
        _, mat, B = self._gather_inputs(input_values)
        out = be.eval_clear(prog, mat)
        return self._format_outputs(prog, out)


# names the north star / BASELINE.json use for the same class
FbsExecEnv = LutExecEnv
B200FbsExecEnv = LutExecEnv
