"""Synthetic source circuits for the benchmark configurations whose input files are not available offline.

BASELINE.json names EPFL ``adder.blif`` / ``multiplier.blif`` and Bristol ``aes_128.txt``; the reference fetches
them with git/wget (reference experiments/gen_makefile_epfl.bash:7, gen_makefile_bristol.bash:4) and there is no
network here.  These builders produce circuits of the same function and gate vocabulary (AND / inverter graphs
expressed as 2-input LUTs, as in an EPFL AIG-derived BLIF) through the ``BitExecEnv`` builder API, so a real
file can be substituted by passing its path to the CLI instead.  ``env_cls`` may be the reference's BitExecEnv.
"""
from __future__ import annotations

from .bit_env import BitExecEnv

# 2-input AND with optional input inversions, as truth tables indexed by 2*a+b (first input is the MSB)
_AND = {(0, 0): [0, 0, 0, 1], (1, 0): [0, 1, 0, 0], (0, 1): [0, 0, 1, 0], (1, 1): [1, 0, 0, 0]}


class _Aig:
    """Tiny AIG helper: literals are (node, inverted)."""

    def __init__(self, env):
        self.env = env

    def land(self, x, y):
        (a, ia), (b, ib) = x, y
        return (self.env.op_lut([a, b], list(_AND[(ia, ib)])), 0)

    def lnot(self, x):
        return (x[0], 1 - x[1])

    def lor(self, x, y):
        return self.lnot(self.land(self.lnot(x), self.lnot(y)))

    def lxor(self, x, y):
        # x^y = ~(~(x & ~y) & ~(~x & y))
        return self.lor(self.land(x, self.lnot(y)), self.land(self.lnot(x), y))

    def materialise(self, x):
        node, inv = x
        return self.env.op_lut([node], [1, 0]) if inv else node


def ripple_carry_adder(nbits=128, env_cls=BitExecEnv):
    """nbits + nbits -> nbits+1 adder as an AND-inverter graph (2*nbits inputs, nbits+1 outputs)."""
    env = env_cls()
    g = _Aig(env)
    a = [(env.input(f"a{i}"), 0) for i in range(nbits)]
    b = [(env.input(f"b{i}"), 0) for i in range(nbits)]
    carry = None
    for i in range(nbits):
        axb = g.lxor(a[i], b[i])
        if carry is None:
            s, carry = axb, g.land(a[i], b[i])
        else:
            s = g.lxor(axb, carry)
            carry = g.lor(g.land(a[i], b[i]), g.land(axb, carry))
        env.output(f"f{i}", g.materialise(s))
    env.output(f"f{nbits}", g.materialise(carry))
    return env


def array_multiplier(nbits=16, env_cls=BitExecEnv):
    """nbits x nbits -> 2*nbits array multiplier (AND partial products, ripple rows of AIG full adders)."""
    env = env_cls()
    g = _Aig(env)
    a = [(env.input(f"a{i}"), 0) for i in range(nbits)]
    b = [(env.input(f"b{i}"), 0) for i in range(nbits)]

    def full_add(x, y, c):
        xy = g.lxor(x, y)
        return g.lxor(xy, c), g.lor(g.land(x, y), g.land(xy, c))

    def half_add(x, y):
        return g.lxor(x, y), g.land(x, y)

    row = [g.land(a[j], b[0]) for j in range(nbits)]          # weights 0..n-1
    outs = [row[0]]
    acc = row[1:]                                             # weights 1..n-1 relative to next row's 0
    top = None
    for i in range(1, nbits):
        pp = [g.land(a[j], b[i]) for j in range(nbits)]
        new, carry = [], None
        for j in range(nbits):
            x = pp[j]
            y = acc[j] if j < len(acc) else top
            if y is None:
                s, c = (x, None) if carry is None else half_add(x, carry)
            elif carry is None:
                s, c = half_add(x, y)
            else:
                s, c = full_add(x, y, carry)
            new.append(s)
            carry = c
        outs.append(new[0])
        acc, top = new[1:], carry
    final = acc + ([top] if top is not None else [])
    outs.extend(final)
    for i, o in enumerate(outs):
        env.output(f"f{i}", g.materialise(o))
    return env
