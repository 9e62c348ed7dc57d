"""Levelise a ``LutExecEnv`` into the flat program descriptor of include/fbs_b200.h.

The reference interpreter walks ``LutExecEnv.instructions`` in build order (reference
fbs_exec_env.py:211-223).  For batched bootstrapping the same DAG is cut into levels:
``level(Input) = 0``, ``level(Bootstrap) = 1 + max level of the wires its lincomb reads``.  All bootstraps
of one level are independent and run as ONE batched programmable bootstrap over (nodes x instances).

A *wire* is an Input or a Bootstrap output and owns a ciphertext slot; LinearProd nodes are folded into CSR
rows over wires.  Several bootstraps may share one LinearProd (reference de-duplicates instructions by text,
fbs_exec_env.py:93-100): they then share one key switch.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np


MAX_FANIN = 64


def _kind(node) -> str:
    """Node classes are matched by NAME so that the reference's own LutExecEnv instances (fbs_exec_env.py:22-61) can be
    levelised as well as this package's mirror classes."""
    return type(node).__name__


class CProgDesc(ctypes.Structure):
    """Mirror of ``fbs_prog_desc`` (include/fbs_b200.h)."""
    _P32 = ctypes.POINTER(ctypes.c_int32)
    _P8 = ctypes.POINTER(ctypes.c_uint8)
    _fields_ = [("p", ctypes.c_int32), ("n_inputs", ctypes.c_int32), ("n_lincombs", ctypes.c_int32),
                ("n_boots", ctypes.c_int32), ("n_levels", ctypes.c_int32), ("n_slots", ctypes.c_int32),
                ("n_outputs", ctypes.c_int32), ("contiguous_levels", ctypes.c_int32),
                ("lc_level_ptr", _P32), ("bs_level_ptr", _P32),
                ("lc_ptr", _P32), ("lc_slot", _P32), ("lc_coef", _P32), ("lc_const", _P32),
                ("bs_lc", _P32), ("bs_slot", _P32), ("bs_tab_ptr", _P32), ("bs_tab", _P8), ("bs_mode", _P32),
                ("in_slot", _P32),
                ("out_ptr", _P32), ("out_slot", _P32), ("out_coef", _P32), ("out_const", _P32),
                # multi-value bootstrap (appended: older mirrors of the descriptor are a prefix): groups of bootstraps on one lincomb
                ("grp_level_ptr", _P32), ("grp_first", _P32), ("n_groups", ctypes.c_int32), ("reserved2", ctypes.c_int32)]


def table_mode(table, p):
    """Negacyclic mode s of a table for plaintext modulus p (reference map_to_fbs.py:81-98).

    len <= p: any table works, s = 1.  p < len <= 2p: the upper part must be the "negation" of the lower
    part, tv[x] + tv[x+p] == s for one constant s (s=1: neg, s=0: zero, s=2: one mode).  Otherwise the table
    cannot be evaluated by one bootstrap in Z_p and a ValueError is raised."""
    L = len(table)
    if L <= p:
        return 1
    if L > 2 * p:
        raise ValueError(f"table of length {L} does not fit fbs_size p={p} (max 2p)")
    s_vals = {(int(table[x]) + int(table[x + p])) % (2 * p) for x in range(L - p)}
    if len(s_vals) != 1:
        raise ValueError(f"table {table} is not negacyclic for p={p}: tv[x]+tv[x+p] must be constant")
    return s_vals.pop()


def min_fbs_size(env) -> int:
    """Smallest p for which every bootstrap table of the circuit is realisable."""
    need = 2
    tables = [i.table for i in env.instructions if _kind(i) == "Bootstrap"]
    for tab in tables:
        need = max(need, (len(tab) + 1) // 2, max(tab) + 1)
    p = need
    while True:
        try:
            for tab in tables:
                table_mode(tab, p)
            return p
        except ValueError:
            p += 1


@dataclass
class Program:
    p: int
    n_inputs: int
    n_levels: int
    n_slots: int
    input_names: list
    output_names: list
    out_index: dict
    arrays: dict = field(default_factory=dict)
    level_widths: list = field(default_factory=list)
    contiguous_levels: bool = False
    shard_pad: int = 1
    n_boots: int = 0
    n_lincombs: int = 0
    multi_value: bool = False      # bootstraps that share a lincomb are evaluated by ONE blind rotation
    n_groups: int = 0              # = number of blind rotations per instance when multi_value

    @property
    def n_rotations(self) -> int:
        """Blind rotations per instance: one per bootstrap, or one per (level, lincomb) group with multi_value."""
        return self.n_groups if self.multi_value else self.n_boots

    def c_desc(self) -> CProgDesc:
        a = self.arrays
        d = CProgDesc()
        d.p, d.n_inputs, d.n_lincombs, d.n_boots = self.p, self.n_inputs, self.n_lincombs, self.n_boots
        d.n_levels, d.n_slots, d.n_outputs = self.n_levels, self.n_slots, len(self.output_names)
        d.contiguous_levels = 1 if self.contiguous_levels else 0
        d.n_groups = self.n_groups if self.multi_value else 0
        for name, tp in CProgDesc._fields_[8:]:
            if name in ("n_groups", "reserved2"):
                continue
            if name in ("grp_level_ptr", "grp_first") and not self.multi_value:
                continue                          # NULL: ordinary one-rotation-per-bootstrap program
            arr = a[name]
            ct = ctypes.c_uint8 if arr.dtype == np.uint8 else ctypes.c_int32
            setattr(d, name, arr.ctypes.data_as(ctypes.POINTER(ct)))
        return d


def _flatten(node, scale, acc, const):
    """Expand a node into {wire: coef} over wires (Input/Bootstrap) plus a constant."""
    if _kind(node) == "Const":
        return const + scale * node.value
    if _kind(node) == "LinearProd":
        const += scale * node.const_coef
        for c, v in node.coef_vals:
            const = _flatten(v, scale * c, acc, const)
        return const
    acc[node.name] = acc.get(node.name, 0) + scale
    return const


def levelize(env, p: int | None = None, reuse_slots: bool = True, shard_pad: int = 1, clear: bool = False,
             preserve_inputs: bool = False, multi_value: bool = False) -> Program:
    """Build the program.  ``multi_value``: bootstraps of a level that share a lincomb (reference fbs_exec_env.py:93-100 makes the
    sharing visible by de-duplicating LinearProds) become one group = ONE blind rotation + one cheap finishing step per table.
    ``reuse_slots``: liveness-based slot reuse (instance-sharded / single GPU);
    ``shard_pad`` > 1: level-contiguous slots padded to a multiple of ``shard_pad`` nodes per level so a
    level's outputs can be all-gathered in place across ``shard_pad`` ranks (node-sharded mode);
    ``preserve_inputs``: never recycle the input slots, so the same encrypted inputs can be run repeatedly."""
    instrs = env.instructions
    if p is None:
        if clear:   # cleartext look-ups have no modulus; pick one that passes the loader's table-length check
            p = max([2] + [len(i.table) for i in env.instructions if _kind(i) == "Bootstrap"])
        else:
            p = min_fbs_size(env)
    inputs = [i for i in instrs if _kind(i) == "Input"]
    boots = [i for i in instrs if _kind(i) == "Bootstrap"]
    by_name = {i.name: i for i in instrs}

    level = {i.name: 0 for i in inputs}
    lin_ops = {}                      # LinearProd name -> ({wire: coef}, const)

    def ops_of(node):
        if node.name not in lin_ops:
            acc = {}
            const = _flatten(node, 1, acc, 0)
            lin_ops[node.name] = ({w: c for w, c in acc.items() if c != 0}, const)
        return lin_ops[node.name]

    for b in boots:
        ops, _ = ops_of(b.val)
        for w in ops:
            assert w in level, f"wire {w} used before definition"
        level[b.name] = 1 + max((level[w] for w in ops), default=0)
    n_levels = max((level[b.name] for b in boots), default=0)

    # bootstraps per level, grouped by the lincomb they consume (first-use order), so that one level's
    # bootstraps are sorted by lincomb index and any node sub-range uses a contiguous lincomb range
    lc_index = {}
    lc_rows = []                      # (level, name)
    per_level = [[] for _ in range(n_levels)]
    for b in boots:
        per_level[level[b.name] - 1].append(b)
    boots_sorted = []
    lc_level_ptr, bs_level_ptr = [0], [0]
    for lv, bl in enumerate(per_level):
        order = {}
        for b in bl:
            # the same LinearProd may feed bootstraps of one level only (its level is determined by its operands)
            order.setdefault(b.val.name, len(order))
        for nm in order:
            lc_index[(lv, nm)] = len(lc_rows)
            lc_rows.append((lv, nm))
        bl_sorted = sorted(bl, key=lambda b: order[b.val.name])     # stable: keeps build order inside a group
        boots_sorted.extend(bl_sorted)
        lc_level_ptr.append(len(lc_rows))
        bs_level_ptr.append(len(boots_sorted))

    # outputs as lincombs over wires
    out_names = list(env.outputs.keys())
    out_rows = []
    for nm in out_names:
        node = env.outputs[nm]
        acc = {}
        const = _flatten(node, 1, acc, 0)
        out_rows.append(({w: c for w, c in acc.items() if c != 0}, const))

    # last level at which each wire is read (outputs keep it alive to the end)
    INF = n_levels + 1
    last_use = {w: 0 for w in level}
    for lv, nm in lc_rows:
        for w in lin_ops[nm][0]:
            last_use[w] = max(last_use[w], lv + 1)
    for ops, _ in out_rows:
        for w in ops:
            last_use[w] = INF
    if preserve_inputs:
        for i in inputs:
            last_use[i.name] = INF

    slot = {}
    if shard_pad > 1 or not reuse_slots:
        nxt = 0
        for i in inputs:
            slot[i.name] = nxt
            nxt += 1
        for lv in range(n_levels):
            bl = boots_sorted[bs_level_ptr[lv]:bs_level_ptr[lv + 1]]
            for j, b in enumerate(bl):
                slot[b.name] = nxt + j
            width = len(bl)
            nxt += -(-width // shard_pad) * shard_pad
        n_slots = max(nxt, 1)
    else:
        free, nxt = [], 0
        expiring = {}                 # level -> wires whose last read is at that level
        for w, lu in last_use.items():
            expiring.setdefault(lu, []).append(w)
        for i in inputs:
            slot[i.name] = nxt
            nxt += 1
        for w in expiring.get(0, []):                # inputs nobody reads
            if w in slot:
                free.append(slot[w])
        for lv in range(1, n_levels + 1):
            # lincombs of level lv run before its bootstraps: wires last read here can be overwritten now
            for w in expiring.get(lv, []):
                free.append(slot[w])
            for b in boots_sorted[bs_level_ptr[lv - 1]:bs_level_ptr[lv]]:
                if free:
                    slot[b.name] = free.pop()
                else:
                    slot[b.name] = nxt
                    nxt += 1
        n_slots = max(nxt, 1)

    # ---- arrays
    lc_ptr, lc_slot, lc_coef, lc_const = [0], [], [], []
    for lv, nm in lc_rows:
        ops, const = lin_ops[nm]
        assert len(ops) <= MAX_FANIN, "lincomb fan-in too large"
        for w, c in ops.items():
            lc_slot.append(slot[w])
            lc_coef.append(c)
        lc_ptr.append(len(lc_slot))
        lc_const.append(const)
    bs_lc, bs_slot, bs_tab_ptr, bs_tab, bs_mode = [], [], [0], [], []
    for b in boots_sorted:
        lv = level[b.name] - 1
        bs_lc.append(lc_index[(lv, b.val.name)])
        bs_slot.append(slot[b.name])
        tab = [int(t) for t in b.table]
        if not clear:
            assert len(tab) <= 2 * p, f"table of length {len(tab)} does not fit fbs_size p={p}"
            assert max(tab) < 2 * p and len(tab) <= 64
            bs_mode.append(table_mode(tab, p))
        else:
            bs_mode.append(1)
        bs_tab.extend(tab)
        bs_tab_ptr.append(len(bs_tab))
    out_ptr, out_slot, out_coef, out_const = [0], [], [], []
    for ops, const in out_rows:
        for w, c in ops.items():
            out_slot.append(slot[w])
            out_coef.append(c)
        out_ptr.append(len(out_slot))
        out_const.append(const)

    def i32(x):
        return np.ascontiguousarray(np.asarray(x, dtype=np.int32).reshape(-1))

    arrays = dict(lc_level_ptr=i32(lc_level_ptr), bs_level_ptr=i32(bs_level_ptr), lc_ptr=i32(lc_ptr), lc_slot=i32(lc_slot),
                  lc_coef=i32(lc_coef), lc_const=i32(lc_const), bs_lc=i32(bs_lc), bs_slot=i32(bs_slot),
                  bs_tab_ptr=i32(bs_tab_ptr), bs_tab=np.ascontiguousarray(np.asarray(bs_tab, dtype=np.uint8).reshape(-1)),
                  bs_mode=i32(bs_mode), in_slot=i32([slot[i.name] for i in inputs]),
                  out_ptr=i32(out_ptr), out_slot=i32(out_slot), out_coef=i32(out_coef), out_const=i32(out_const))
    # ctypes needs a valid pointer even for empty arrays
    for k, v in arrays.items():
        if v.size == 0:
            arrays[k] = np.zeros(1, dtype=v.dtype)
    prog = Program(p=p if not clear else max(p, 2), n_inputs=len(inputs), n_levels=n_levels, n_slots=n_slots,
                   input_names=[i.name for i in inputs], output_names=out_names,
                   out_index={nm: k for k, nm in enumerate(out_names)}, arrays=arrays,
                   level_widths=[bs_level_ptr[i + 1] - bs_level_ptr[i] for i in range(n_levels)],
                   contiguous_levels=(shard_pad > 1 or not reuse_slots), shard_pad=shard_pad)
    # true counts (arrays may have been padded to length 1)
    prog.n_boots = len(boots_sorted)
    prog.n_lincombs = len(lc_rows)
    # groups: maximal runs of a level's bootstraps (sorted by lincomb) on the same lincomb
    grp_first, grp_level_ptr = [], [0]
    for lv in range(n_levels):
        for q in range(bs_level_ptr[lv], bs_level_ptr[lv + 1]):
            if q == bs_level_ptr[lv] or bs_lc[q] != bs_lc[q - 1]:
                grp_first.append(q)
        grp_level_ptr.append(len(grp_first))
    prog.n_groups = len(grp_first)
    grp_first.append(len(boots_sorted))
    prog.arrays["grp_first"], prog.arrays["grp_level_ptr"] = i32(grp_first), i32(grp_level_ptr)
    prog.multi_value = bool(multi_value) and not clear
    return prog
