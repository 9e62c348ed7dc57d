"""Boolean circuit -> FBS circuit mappers with the reference's interface and results.

Own implementation of reference ``fbs_mapper/map_to_fbs.py``: ``MapToFBSBasic`` (:15-51) and ``MapToFBSHeur`` with the
``naive`` / ``search`` cone mergers (:54-547), following SURVEY.md Appendix H.  It is host-side, runs once per circuit and
is not data-parallel, so it stays in Python; what matters is that it produces the *same mapped circuit* as the
reference, because the mapped circuit is the input of the encrypted executor.  tests/test_mapper.py checks byte-identical
``.lbf`` output and statistics against the reference's goldens (tests/golden/ref_mapped.json) and, where the reference
tree is mounted, against the live reference on further circuits.

A *cone* is a not-yet-bootstrapped sub-circuit: ``support`` wires, integer ``coefs``, and per assignment of the support
(support[0] is the most significant bit) the Boolean value ``tt`` and the integer value ``mvt`` of the linear
combination.  A 2-input gate over two cones becomes one cone if integers (a, b) exist such that a*mvt1 + b*mvt2 still
separates the gate's 0s from its 1s and spans a table the message space Z_p can hold -- directly (span <= p) or through
the negacyclic extension to 2p (reference map_to_fbs.py:81-98); otherwise an operand is bootstrapped first.

Deliberate bug-compatibility (documented in SURVEY.md Appendix E / H):
* the gate's truth table is swapped in place when the operand cones are swapped (map_to_fbs.py:403-405), which
  mutates the source BitExecEnv -- evaluate the source circuit *before* mapping, as the CLI does;
* zero-coefficient variables are dropped by indexing the tables with the compacted assignment numbers
  (map_to_fbs.py:286-300).
"""
from __future__ import annotations

import itertools
import logging

import numpy as np

from .lut_env import LutExecEnv


def _kind(node):
    return type(node).__name__


class MapToFBSBasic:
    """One lincomb + bootstrap per 2-input gate, raw gate tables, fbs_size ignored (reference map_to_fbs.py:15-51)."""

    def map(self, env):
        lut = LutExecEnv()
        wires = {"0": lut.const(0), "1": lut.const(1)}
        for instr in env.instructions:
            logging.getLogger("MapToFBSBasic").info(f"{instr}")
            kind = _kind(instr)
            if kind in ("Input", "BInput"):
                wires[instr.name] = lut.input(instr.name)
            elif hasattr(instr, "truth_table"):
                table, ins = list(instr.truth_table), instr.inputs
                assert len(table) == 2 ** len(ins)
                if len(ins) == 1:
                    src = wires[ins[0].name]
                    if table == [1, 0]:
                        wires[instr.name] = lut.linear([-1], [src], const_coef=1)
                    else:
                        assert table == [0, 1]
                        wires[instr.name] = src
                else:
                    weights = [2 ** k for k in range(len(ins))][::-1]
                    lin = lut.linear(weights, [wires[i.name] for i in ins])
                    wires[instr.name] = lut.bootstrap(lin, table)
            else:
                assert False, "Unknown instruction"
        for name, out in env.outputs.items():
            lut.output(name, wires[out.name])
        return lut


class _Cone:
    __slots__ = ("owner", "support", "coefs", "tt", "mvt", "names")

    def __init__(self, owner, support, coefs, tt, mvt):
        self.owner = owner
        self.support = np.array(support)
        self.coefs = np.array(coefs)
        self.tt = np.array(tt)
        self.mvt = np.array(mvt)
        assert owner._is_lut_valid(self.tt, self.mvt), f"{self.tt} {self.mvt}"
        self.names = np.array([s.name for s in self.support])
        if self.size() != len(np.unique(self.mvt)):
            logging.critical(f"Cone with sparse mvt: {self.size()} {len(np.unique(self.mvt))}\t{self}")

    def size(self):
        return self.owner._mvt_size(self.mvt)

    def norm2_squared(self):
        return np.sum(self.coefs * self.coefs)

    def name(self):
        return repr(self)

    def support_names(self):
        return self.names

    def with_tt(self, tt):
        return _Cone(self.owner, self.support, self.coefs, tt, self.mvt)

    def __repr__(self):
        return f"Cone({self.names}, {self.coefs}, {self.mvt}, {self.tt})"


def _var_column(n_vars, pos):
    """Values of variable ``pos`` (0 = most significant) over all 2^n_vars assignments."""
    assert pos < n_vars
    rep = 1 << (n_vars - pos - 1)
    block = np.hstack((np.zeros(rep, dtype=np.uint32), np.ones(rep, dtype=np.uint32)))
    return np.tile(block, 1 << pos)


class MapToFBSHeur:
    """Greedy cone merging with minimal coefficients (reference map_to_fbs.py:54-547)."""

    def __init__(self, cone_merger, fbs_size=8, max_fbs_size=16, max_truth_table_size=16):
        self.fbs_size = fbs_size
        self.max_fbs_size = max_fbs_size
        self.max_truth_table_size = max_truth_table_size
        if cone_merger == "naive":
            self._find_lincomb_coefs = self._coefs_naive
        elif cone_merger == "search":
            self._find_lincomb_coefs = self._coefs_search
        else:
            assert False, f"Unknown cone merger '{cone_merger}'"
        self._coef_cache = {}
        self.logger = logging.getLogger(f"MapToFBS_{cone_merger}")

    # ------------------------------------------------------------------ table legality (map_to_fbs.py:70-121)
    def _mvt_size(self, mvt):
        return np.max(mvt) - np.min(mvt) + 1

    def _table_from(self, tt, mvt, fill):
        """Table over [min(mvt), max(mvt)]: reachable indices take the gate value, the rest ``fill``."""
        lo, hi = int(mvt.min()), int(mvt.max())
        tab = [fill] * (hi - lo + 1)
        for v, t in zip(mvt, tt):
            tab[int(v) - lo] = t
        return tab

    def _is_mvt_valid(self, tt, mvt):
        return len(set(mvt[tt == 0]).intersection(mvt[tt == 1])) == 0

    def _is_test_vector_valid(self, tv):
        if len(tv) <= self.fbs_size:
            return True
        if len(tv) <= self.max_fbs_size:
            tv = np.array(tv)
            head, tail = tv[0:len(tv) - self.fbs_size], tv[self.fbs_size:]
            neg = np.all(head != tail)                                  # f(x) = -f(x + p)
            zero = np.all(head == tail) and np.all(0 == head)
            one = np.all(head == tail) and np.all(1 == head)
            return bool(neg or zero or one)
        return False

    def _is_lut_valid(self, tt, mvt):
        if not self._is_mvt_valid(tt, mvt):
            return False
        if self._mvt_size(mvt) <= self.fbs_size:
            return True
        return self._is_test_vector_valid(self._table_from(tt, mvt, 0)) or self._is_test_vector_valid(self._table_from(tt, mvt, 1))

    def _fbs_table(self, tt, mvt):
        tv = self._table_from(tt, mvt, 0)
        if self._is_test_vector_valid(tv):
            return tv
        tv = self._table_from(tt, mvt, 1)
        assert self._is_test_vector_valid(tv)
        return tv

    # ------------------------------------------------------------------ cones
    def new_cone(self, support, coefs, tt, mvt):
        return _Cone(self, support, coefs, tt, mvt)

    def new_const(self, cst):
        return self.new_cone([], [], [cst], [0])

    def new_input(self, lut_env, name):
        return self.new_cone([lut_env.input(name)], [1], [0, 1], [0, 1])

    def new_bootstrap(self, lut_env, cone):
        if len(cone.support) <= 1:                      # constants and single wires need no bootstrap
            return cone
        shift = -cone.mvt.min()
        cone.mvt += shift                                # in place, as the reference does
        lin = lut_env.linear([int(c) for c in cone.coefs], cone.support, const_coef=int(shift))
        node = lut_env.bootstrap(lin, [int(t) for t in self._fbs_table(cone.tt, cone.mvt)])
        return self.new_cone([node], [1], [0, 1], [0, 1])

    def new_output(self, lut_env, cone):
        if len(cone.support) == 0:
            return lut_env.const(cone.tt[0])
        if len(cone.support) == 1:
            val = cone.support[0]
            if np.all(cone.tt == [1, 0]):
                return lut_env.linear([-1], [val], const_coef=1)
            return val
        return self.new_bootstrap(lut_env, cone).support[0]

    # ------------------------------------------------------------------ joint tables (map_to_fbs.py:415-440)
    def _joint_indices(self, sup1, sup2):
        sup1, sup2 = np.array(sup1), np.array(sup2)
        joint = np.concatenate((sup1, sup2[~np.isin(sup2, sup1)]))
        n = len(joint)
        idx2 = np.zeros(1 << n, dtype=np.uint32)
        for node in sup2:
            pos = np.where(np.equal(node, joint))[0][0]
            idx2 = (idx2 << 1) + _var_column(n, pos)
        idx1 = np.repeat(np.arange(1 << len(sup1)), 1 << (n - len(sup1)))
        return idx1, idx2

    def _joint_tables(self, cone1, cone2, gate_tt):
        idx1, idx2 = self._joint_indices(cone1.support_names(), cone2.support_names())
        xy = np.vstack((cone1.mvt[idx1], cone2.mvt[idx2])).T
        r_tt = np.array(gate_tt)[2 * cone1.tt[idx1] + cone2.tt[idx2]]
        return xy, r_tt

    # ------------------------------------------------------------------ coefficient choice (map_to_fbs.py:336-401)
    def _coefs_naive(self, xy, r_tt):
        a, b = self._mvt_size(xy[:, 1]), 1
        mvt = a * xy[:, 0] + b * xy[:, 1]
        return ((a, b), mvt) if self._is_lut_valid(r_tt, mvt) else (None, None)

    def _candidates_by_span(self, size1, size2):
        if size1 < size2:
            pairs = itertools.product(range(size2 + 1), range(-size1, size1 + 1))
        else:
            pairs = itertools.product(range(-size2, size2 + 1), range(size1 + 1))
        pairs = np.array(list(pairs))
        span = np.abs(pairs[:, 0]) * (size1 - 1) + np.abs(pairs[:, 1]) * (size2 - 1)
        groups = {}
        for s in np.unique(span):
            groups[s] = sorted(map(tuple, pairs[span == s, :]), reverse=True)
        return groups

    def _coefs_search(self, xy, r_tt):
        best_ab, best_mvt, best_span, best_norm = None, None, 1000000000, 1000000000
        c1max, c2max = self._mvt_size(xy[:, 0]) - 1, self._mvt_size(xy[:, 1]) - 1
        for span_m1, cands in self._candidates_by_span(c1max + 1, c2max + 1).items():
            for a, b in cands:
                span = abs(a) * c1max + abs(b) * c2max
                assert span == span_m1
                mvt = a * xy[:, 0] + b * xy[:, 1]
                norm = np.square(mvt).sum()
                if span < best_span or (span == best_span and norm < best_norm):
                    if self._is_lut_valid(r_tt, mvt):
                        best_ab, best_mvt, best_span, best_norm = (a, b), mvt, span, norm
            if best_ab is not None:
                break
        return best_ab, best_mvt

    def _coefs_cached(self, xy, r_tt):
        key = f"{','.join(map(str, xy))}|{','.join(map(str, r_tt))}"
        if key not in self._coef_cache:
            self._coef_cache[key] = self._find_lincomb_coefs(xy, r_tt)
        return self._coef_cache[key]

    # ------------------------------------------------------------------ merging (map_to_fbs.py:286-334)
    def _simplified_cone(self, support, coefs, tt, mvt):
        if np.sum(coefs == 0) > 0:
            r = np.zeros(1 << len(coefs), dtype=np.uint32)
            for pos, c in enumerate(coefs):
                r <<= 1
                if c != 0:
                    r += _var_column(len(coefs), pos)
            keep = np.unique(r)                          # compacted assignment numbers used as indices (bug-compatible)
            support, coefs = support[coefs != 0], coefs[coefs != 0]
            tt, mvt = tt[keep], mvt[keep]
        g = np.gcd.reduce(coefs)
        coefs //= g
        mvt = mvt // g
        return self.new_cone(support, coefs, tt, mvt)

    def _merge(self, cone1, cone2, ab, tt, mvt):
        n1, n2 = cone1.support_names(), cone2.support_names()
        a, b = ab
        c1, c2 = cone1.coefs * a, cone2.coefs * b
        common = list(set(n1).intersection(n2))
        for node in common:                              # shared wires keep one coefficient, in cone1's position
            c1[np.where(node == n1)[0][0]] += c2[np.where(node == n2)[0][0]]
        keep = ~np.isin(n2, common)
        support = np.hstack((cone1.support, cone2.support[keep]))
        coefs = np.hstack((c1, c2[keep].astype(np.int64)))
        return self._simplified_cone(support, coefs, tt, mvt)

    # ------------------------------------------------------------------ gates (map_to_fbs.py:442-547)
    @staticmethod
    def _swap(cone1, cone2, idx1, idx2, truth_table):
        truth_table[1], truth_table[2] = truth_table[2], truth_table[1]     # in place: quirk Q1
        return cone2, cone1, idx2, idx1, truth_table

    def treat_bit_exec_lut_gate(self, lut_env, input_wires, truth_table):
        if len(input_wires) == 1:
            (cone,) = input_wires
            assert len(truth_table) == 2, "error"
            return cone.with_tt(np.array(truth_table)[cone.tt]), {}
        assert len(input_wires) == 2 and len(truth_table) == 4, "error"
        cone1, cone2 = input_wires
        idx1, idx2 = 0, 1
        if cone1.size() < cone2.size() or (cone1.size() == cone2.size() and cone1.norm2_squared() < cone2.norm2_squared()):
            cone1, cone2, idx1, idx2, truth_table = self._swap(cone1, cone2, idx1, idx2, truth_table)
        boot = {}

        def joint_support():
            return len(set(cone1.support_names()).union(cone2.support_names()))

        if joint_support() > self.max_truth_table_size:
            boot[idx1] = cone1 = self.new_bootstrap(lut_env, cone1)
            cone1, cone2, idx1, idx2, truth_table = self._swap(cone1, cone2, idx1, idx2, truth_table)
            if joint_support() > self.max_truth_table_size:
                boot[idx1] = cone1 = self.new_bootstrap(lut_env, cone1)

        for attempt in range(3):
            xy, r_tt = self._joint_tables(cone1, cone2, truth_table)
            if len(np.unique(r_tt)) == 1:
                return self.new_const(r_tt[0]), boot
            ab, r_mvt = self._coefs_cached(xy, r_tt)
            if ab is not None:
                return self._merge(cone1, cone2, ab, r_tt, r_mvt), boot
            if attempt == 0:
                boot[idx1] = cone1 = self.new_bootstrap(lut_env, cone1)
            elif attempt == 1:
                boot[idx2] = cone2 = self.new_bootstrap(lut_env, cone2)
        assert False, "two bootstrapped operands always merge"

    # ------------------------------------------------------------------ driver (map_to_fbs.py:123-175)
    def map_internal(self, env, nodes_to_bootstrap):
        lut_env = LutExecEnv()
        wires = {"0": self.new_const(0), "1": self.new_const(1)}
        for instr in env.instructions:
            kind = _kind(instr)
            if kind in ("Const", "BConst"):
                wire = wires[instr.name]
            elif kind in ("Input", "BInput"):
                wire = self.new_input(lut_env, instr.name)
            elif hasattr(instr, "truth_table"):
                assert len(instr.inputs) <= 2, "only 1 or 2 input gates are supported"
                ins = [wires[i.name] for i in instr.inputs]
                wire, boot = self.treat_bit_exec_lut_gate(lut_env, ins, instr.truth_table)
                for pos, new_wire in boot.items():       # later consumers reuse the bootstrapped operand
                    wires[instr.inputs[pos].name] = new_wire
            else:
                assert False, "Unknown instruction"
            wires[instr.name] = self.new_bootstrap(lut_env, wire) if instr.name in nodes_to_bootstrap else wire
        for name, out in env.outputs.items():
            lut_env.output(name, self.new_output(lut_env, wires[out.name]))
        return lut_env

    def map(self, env):
        return self.map_internal(env, {o.name for o in env.outputs.values()})
