"""N>1 host logic on CPU: world_size-2 gloo.  The partitioning / exchange code of tfhe_fbs_map_b200/dist.py runs
unchanged; the per-level compute is done by the CPU oracle (allowed in tests) instead of the CUDA library."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_ref_mapped, selfcheck_inputs, unpack_outputs
from tfhe_fbs_map_b200 import levelize, params
from tfhe_fbs_map_b200.dist import instance_shard, level_node_range, run_node_sharded
from tfhe_fbs_map_b200.formats import read_lbf

Q = 0x3FFE8001 * 0x3FFF4001


def test_partition_helpers():
    for total in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            parts = [instance_shard(total, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == total
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    for width in (1, 2, 5, 8, 13):
        for world in (1, 2, 4, 8):
            rs = [level_node_range(width, world, r) for r in range(world)]
            covered = sorted(i for b, e, _ in rs for i in range(b, e))
            assert covered == list(range(width))
            assert len({c for _, _, c in rs}) == 1


class OracleEngine:
    """Same interface as dist.B200Engine, computed on the CPU by oracle/tfhe_ref.c."""

    def __init__(self, ref, program, B):
        self.ref, self.prog, self.B = ref, program, B
        self.ct_words = ref.ct_words
        self.wires = torch.zeros(program.n_slots * B * self.ct_words, dtype=torch.int64)
        self.np = self.wires.numpy().view(np.uint64).reshape(program.n_slots, B, self.ct_words)

    def encrypt(self, bits, inst_offset=0, total=None, enc_seed=5):
        total = total or self.B
        a = self.prog.arrays
        for i in range(self.prog.n_inputs):
            ids = np.array([i * total + inst_offset + b for b in range(self.B)], dtype=np.uint64)
            self.np[a["in_slot"][i]] = self.ref.encrypt(self.prog.p, bits[i].astype(np.int32), ids, enc_seed)

    def run_level(self, level, nb, ne):
        a, p = self.prog.arrays, self.prog.p
        delta = (Q + p) // (2 * p)
        b0 = int(a["bs_level_ptr"][level])
        outs = {}
        for q in range(b0 + nb, b0 + ne):
            lc = int(a["bs_lc"][q])
            tab = a["bs_tab"][a["bs_tab_ptr"][q]:a["bs_tab_ptr"][q + 1]]
            for b in range(self.B):
                acc = np.zeros(self.ct_words, dtype=object)
                for o in range(a["lc_ptr"][lc], a["lc_ptr"][lc + 1]):
                    acc = (acc + int(a["lc_coef"][o]) * self.np[a["lc_slot"][o], b].astype(object)) % Q
                acc[-1] = (acc[-1] + int(a["lc_const"][lc]) * delta) % Q
                out, _, _, _ = self.ref.pbs(p, np.array(acc, dtype=np.uint64), tab, int(a["bs_mode"][q]))
                outs[(int(a["bs_slot"][q]), b)] = out
        for (s, b), v in outs.items():
            self.np[s, b] = v

    def decrypt(self):
        a, p = self.prog.arrays, self.prog.p
        delta = (Q + p) // (2 * p)
        res = np.zeros((len(self.prog.output_names), self.B), np.uint8)
        for q in range(len(self.prog.output_names)):
            for b in range(self.B):
                acc = np.zeros(self.ct_words, dtype=object)
                for o in range(a["out_ptr"][q], a["out_ptr"][q + 1]):
                    acc = (acc + int(a["out_coef"][o]) * self.np[a["out_slot"][o], b].astype(object)) % Q
                acc[-1] = (acc[-1] + int(a["out_const"][q]) * delta) % Q
                res[q, b] = self.ref.decrypt(p, np.array(acc, dtype=np.uint64)[None, :])[0]
        return res

    def slot_view(self, slot_begin, n_slots):
        per = self.B * self.ct_words
        return self.wires[slot_begin * per:(slot_begin + n_slots) * per]


def _worker(rank, world, port, lbf, p, names, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.tfhe_ref import RefTFHE
        ref = RefTFHE(params.get("toy3"), seed=31)             # same seed on every rank = replicated keys, no broadcast
        prog = levelize(read_lbf(lbf), p, shard_pad=world)
        eng = OracleEngine(ref, prog, B)
        inputs = selfcheck_inputs(names)
        bits = np.array([inputs[nm][:B] for nm in prog.input_names], dtype=np.uint8)
        eng.encrypt(bits)
        words = run_node_sharded(eng, prog, dist, world, rank, in_place=False)
        out = eng.decrypt()
        # every rank must hold the identical replicated wire buffer for the slots a later level or an output reads
        a = prog.arrays
        live = sorted(set(int(s) for s in a["bs_slot"][:prog.n_boots]))
        digest = torch.tensor([int(eng.np[live].astype(np.uint64).sum(dtype=np.uint64) & np.uint64(0x7FFFFFFFFFFFFFFF))], dtype=torch.int64)
        gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, digest)
        ret[rank] = (out, [int(g.item()) for g in gathered], words)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("circuit,p", [("ascon_lut", 11), ("aes_sbox", 11)])
def test_node_sharded_levels_world2_gloo(circuit, p):
    e = next(x for x in load_ref_mapped() if x["circuit"] == circuit and x["p"] == p and x["mapper"] == "search")
    B, world = 2, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, e["lbf"], p, e["input_names"], B, ret), nprocs=world, join=True)
    want = unpack_outputs(e, batch=B)
    prog = levelize(read_lbf(e["lbf"]), p, shard_pad=world)
    for rank in range(world):
        out, digests, words = ret[rank]
        assert len(set(digests)) == 1, "wire replicas diverged"
        assert words > 0
        for nm in prog.output_names:
            assert np.array_equal(out[prog.out_index[nm]], want[str(nm)]), (rank, nm)


class FakeFusedEngine(OracleEngine):
    """CPU stand-in for dist.FusedB200Engine: `fused = True` makes run_node_sharded call run_level for EVERY level on every rank
    (empty ranges too: on the GPU the call carries the level's device-side wait + signal) and finish() once per run.  The "peer
    stores" are an all-gather inside run_level, so a rank that skipped a level would dead-lock -- exactly what the real
    hand-off would do."""
    fused = True

    def __init__(self, ref, program, B, world, rank):
        super().__init__(ref, program, B)
        self.world, self.rank, self.calls, self.finished = world, rank, [], 0

    def run_level(self, level, nb, ne):
        self.calls.append((level, nb, ne))
        if ne > nb:
            super().run_level(level, nb, ne)
        a = self.prog.arrays
        b0, b1 = int(a["bs_level_ptr"][level]), int(a["bs_level_ptr"][level + 1])
        chunk = -(-(b1 - b0) // self.world)
        first = int(a["bs_slot"][b0])
        parts = [self.slot_view(first + r * chunk, chunk) for r in range(self.world)]
        dist.all_gather(parts, parts[self.rank].clone())

    def finish(self):
        self.finished += 1


def _fused_worker(rank, world, port, lbf, p, names, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.tfhe_ref import RefTFHE
        ref = RefTFHE(params.get("toy3"), seed=31)
        prog = levelize(read_lbf(lbf), p, shard_pad=world)
        eng = FakeFusedEngine(ref, prog, B, world, rank)
        inputs = selfcheck_inputs(names)
        bits = np.array([inputs[nm][:B] for nm in prog.input_names], dtype=np.uint8)
        eng.encrypt(bits)
        run_node_sharded(eng, prog, dist, world, rank)
        ret[rank] = (eng.decrypt(), eng.calls, eng.finished)
    finally:
        dist.destroy_process_group()


def test_fused_exchange_control_flow_world2_gloo():
    """Host logic of the fused node-sharded mode (tfhe_fbs_map_b200/dist.py): every rank issues every level -- aes_sbox has a level
    of width 1, so rank 1's range is empty there -- and closes the run with finish(); outputs equal the reference's."""
    e = next(x for x in load_ref_mapped() if x["circuit"] == "aes_sbox" and x["p"] == 11 and x["mapper"] == "search" and not x.get("strict"))
    B, world = 1, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fused_worker, args=(world, port, e["lbf"], 11, e["input_names"], B, ret), nprocs=world, join=True)
    want = unpack_outputs(e, batch=B)
    prog = levelize(read_lbf(e["lbf"]), 11, shard_pad=world)
    assert min(prog.level_widths) == 1
    for rank in range(world):
        out, calls, finished = ret[rank]
        assert [c[0] for c in calls] == list(range(prog.n_levels)) and finished == 1
        for nm in prog.output_names:
            assert np.array_equal(out[prog.out_index[nm]], want[str(nm)]), (rank, nm)
    assert any(nb == ne for _, nb, ne in ret[1][1]), "rank 1 should have had an empty range at the width-1 level"
