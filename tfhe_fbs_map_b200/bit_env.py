"""Boolean-circuit IR with the reference's ``BitExecEnv`` surface.

Host-side mirror of reference ``fbs_mapper/bit_exec_env.py:5-279``: gate nodes, constant-propagating
builders, dead-node removal, statistics and the BLIF writer behave identically.  ``eval`` (the source-circuit
truth used by the CLI self-check, reference map_circuit.py:140) runs on the GPU: every LUT gate is a table
look-up on the MSB-first index of its inputs (bit_exec_env.py:183-185), which is the same "linear combination
+ table" program shape as a mapped circuit, so it goes through the same cleartext CUDA kernel.
"""
from __future__ import annotations

import sys

import numpy as np


class BNode:
    def __init__(self, name):
        self.name = name


class BConst(BNode):
    def __init__(self, val):
        super().__init__(str(val))
        self.val = val

    def __str__(self):
        return self.name


class BInput(BNode):
    def __str__(self):
        return f"Input({self.name})"


class LUT(BNode):
    def __init__(self, name, inputs, truth_table):
        super().__init__(name)
        for inp in inputs:
            assert isinstance(inp, BNode), "something is wrong"
        self.inputs = inputs
        self.truth_table = truth_table

    def __str__(self):
        return f"LUT([{', '.join(i.name for i in self.inputs)}], {self.truth_table})"


def _two_input_gate(label, table):
    class _Gate(LUT):
        def __init__(self, name, inp1, inp2):
            super().__init__(name, [inp1, inp2], truth_table=list(table))
            assert inp1.name != inp2.name, "something is wrong"

        def __str__(self):
            return f"{label}({self.inputs[0].name}, {self.inputs[1].name})"
    _Gate.__name__ = _Gate.__qualname__ = label.capitalize()
    return _Gate


And = _two_input_gate("AND", (0, 0, 0, 1))
Xor = _two_input_gate("XOR", (0, 1, 1, 0))
Or = _two_input_gate("OR", (0, 1, 1, 1))


class Not(LUT):
    def __init__(self, name, inp):
        super().__init__(name, [inp], truth_table=[1, 0])

    def __str__(self):
        return f"Not({self.inputs[0].name})"


class BitExecEnv:
    Node = BNode
    Const = BConst
    Input = BInput
    LUT = LUT
    And = And
    Xor = Xor
    Or = Or
    Not = Not
    CONST0 = BConst(0)
    CONST1 = BConst(1)

    def __init__(self):
        self._unique_id = 0
        self.instructions = []
        self.inputs = []
        self.outputs = {}
        self.ids = set()

    # ------------------------------------------------------------------ builders
    def _new_id(self):
        self._unique_id += 1
        return f"n{self._unique_id}"

    def _get_id(self, name):
        if name is None:
            name = self._new_id()
            while name in self.ids:
                name = self._new_id()
        else:
            assert name not in self.ids, "id already exists in circuit"
        self.ids.add(name)
        return name

    def _add_instr(self, instr):
        self.instructions.append(instr)
        return instr

    def input(self, input_id):
        inp = self._add_instr(BInput(input_id))
        self.inputs.append(inp)
        return inp

    def output(self, name, node):
        assert isinstance(node, BNode), "Expected BitExecEnv.Node"
        self.outputs[name] = node

    def op_lut(self, inputs, truth_table, name=None):
        assert 2 ** len(inputs) == len(truth_table), "length miss-match"
        for inp in inputs:
            assert isinstance(inp, BNode), "Error"
        assert min(truth_table) == 0, "truth table wrong values"
        assert max(truth_table) == 1, "truth table wrong values"
        return self._add_instr(LUT(self._get_id(name), inputs, truth_table))

    # constant propagation rules of reference bit_exec_env.py:113-159
    def op_not(self, inp, name=None):
        if inp is self.CONST0:
            return self.CONST1
        if inp is self.CONST1:
            return self.CONST0
        return self._add_instr(Not(self._get_id(name), inp))

    def op_and(self, inp1, inp2, name=None):
        if inp1 is self.CONST0 or inp2 is self.CONST0:
            return self.CONST0
        if inp1 is self.CONST1:
            return inp2
        if inp2 is self.CONST1:
            return inp1
        return self._add_instr(And(self._get_id(name), inp1, inp2))

    def op_xor(self, inp1, inp2, name=None):
        if inp1 is self.CONST0:
            return inp2
        if inp1 is self.CONST1:
            return self.op_not(inp2)
        if inp2 is self.CONST0:
            return inp1
        if inp2 is self.CONST1:
            return self.op_not(inp1)
        return self._add_instr(Xor(self._get_id(name), inp1, inp2))

    def op_or(self, inp1, inp2, name=None):
        if inp1 is self.CONST0:
            return inp2
        if inp1 is self.CONST1 or inp2 is self.CONST1:
            return self.CONST1
        if inp2 is self.CONST0:
            return inp1
        return self._add_instr(Or(self._get_id(name), inp1, inp2))

    # ------------------------------------------------------------------ text
    def print(self, os=sys.stdout, show_inputs=True, show_outputs=True):
        for instr in self.instructions:
            if isinstance(instr, BInput) and not show_inputs:
                continue
            print(f"{instr.name} = {str(instr)}", file=os)
        if show_outputs:
            for name, out in self.outputs.items():
                print(f"Output {name} = {out.name}", file=os)

    def to_blif(self, fs=sys.stdout, model_name="test"):
        """BLIF writer with minority-polarity rows (reference bit_exec_env.py:247-279)."""
        def rows(tt):
            val = 1 if np.mean(tt) <= 0.5 else 0
            width = int(np.log2(len(tt)))
            return "\n".join(f"{idx:0{width}b} {val}" for idx, t in enumerate(tt) if t == val)

        print(f".model {model_name}", file=fs)
        print(f".inputs {' '.join(i.name for i in self.inputs)}", file=fs)
        print(f".outputs {' '.join(self.outputs.keys())}", file=fs)
        for instr in self.instructions:
            if isinstance(instr, BConst):
                print(f".names CONST{instr.name}", file=fs)
                print(f"{instr.name}", file=fs)
            elif isinstance(instr, BInput):
                pass
            elif isinstance(instr, LUT):
                print(f".names {' '.join(i.name for i in instr.inputs)} {instr.name}", file=fs)
                print(rows(instr.truth_table), file=fs)
            else:
                assert False, "Unknown instruction"
        for name, out in self.outputs.items():
            if out.name != name:
                print(f".names {out.name} {name}\n1 1", file=fs)
        print(".end", file=fs)

    # ------------------------------------------------------------------ graph utilities
    def remove_dangling_nodes(self):
        live = {o.name for o in self.outputs.values()}
        for instr in reversed(self.instructions):
            if instr.name in live and isinstance(instr, LUT):
                live.update(i.name for i in instr.inputs)
        self.instructions = [i for i in self.instructions if i.name in live]

    def stats(self):
        d = dict(nb_inp=0, nb_and=0, nb_xor=0, nb_not=0, nb_lut=0, max_lut_inputs=0, max_lut_size=0)
        for instr in self.instructions:
            if isinstance(instr, BInput):
                d["nb_inp"] += 1
            elif isinstance(instr, And):
                d["nb_and"] += 1
            elif isinstance(instr, Xor):
                d["nb_xor"] += 1
            elif isinstance(instr, Not):
                d["nb_not"] += 1
            elif isinstance(instr, LUT):   # reference counts OR gates and generic LUTs here too
                d["nb_lut"] += 1
                d["max_lut_inputs"] = max(d["max_lut_inputs"], len(instr.inputs))
                d["max_lut_size"] = max(d["max_lut_size"], len(instr.truth_table))
            else:
                assert False, "Unknown instruction"
        d["nb_out"] = len(self.outputs)
        return d

    # ------------------------------------------------------------------ execution (GPU only)
    def to_lut_env(self):
        """Gate-per-table program: index = sum in_i * 2^(k-1-i) (first input is the MSB,
        reference bit_exec_env.py:183-184), value = truth_table[index]."""
        from .lut_env import LutExecEnv
        env = LutExecEnv(merge_linear_prods=False)
        wires = {"0": env.const(0), "1": env.const(1)}
        for instr in self.instructions:
            if isinstance(instr, BInput):
                wires[instr.name] = env.input(instr.name)
            elif isinstance(instr, LUT):
                kk = len(instr.inputs)
                lin = env.linear([2 ** (kk - 1 - i) for i in range(kk)], [wires[i.name] for i in instr.inputs])
                tab = [int(t) for t in instr.truth_table]
                tab = tab[: env.max_val[lin.name] + 1]     # constant inputs shrink the reachable index range
                if max(tab) == 0:                         # keep min(table)==0 invariants happy for all-zero rows
                    wires[instr.name] = env.const(0)
                elif min(tab) == 1:
                    wires[instr.name] = env.const(1)
                else:
                    wires[instr.name] = env.bootstrap(lin, tab)
            else:
                assert False, "Unknown instruction"
        for name, out in self.outputs.items():
            env.output(name, wires[out.name])
        return env

    def eval(self, input_values, backend=None):
        """Same contract as reference bit_exec_env.py:173-194, evaluated by the cleartext CUDA kernel."""
        return self.to_lut_env().eval_clear(input_values, backend=backend)
