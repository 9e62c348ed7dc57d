"""Multi-GPU partitioning of the encrypted evaluation: one process per GPU, ``torch.distributed`` for plumbing.

Two ways the path shards (BASELINE.json configs[2] and configs[3]):

* **instances** -- a batch of independent encrypted circuit instances is split contiguously across ranks; keys are
  replicated by running the same *seeded* key generation on every device, so there is **no collective** on the data
  path (outputs may be gathered at the end for convenience).
* **nodes** -- ONE circuit (any batch size) whose levels are wide: the bootstraps of each level are split across ranks,
  every rank keeps a full replica of the wire buffer, and after each level the new output ciphertexts are exchanged
  with an in-place **all-gather** (NCCL over NVLink).  ``levelize(shard_pad=world)`` lays the slots of a level out
  contiguously and padded to a multiple of ``world`` so every rank contributes an equal, contiguous chunk.

The level loop is written against a small *engine* interface (``encrypt``, ``run_level``, ``decrypt``, ``wires``) so
that the partitioning / exchange logic is exercised on CPU with the gloo backend (tests/test_dist_gloo.py, where the
engine is backed by the CPU oracle) and on GPUs with NCCL (``B200Engine``).
"""
from __future__ import annotations

import numpy as np


def instance_shard(total: int, world: int, rank: int):
    """Contiguous split of ``total`` instances: returns (offset, count) of this rank (sizes differ by at most 1)."""
    base, rem = divmod(total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def level_node_range(width: int, world: int, rank: int):
    """Bootstraps [begin, end) of a level of ``width`` nodes handled by ``rank``; chunk = ceil(width/world) so that the
    slot region of the level (padded to world*chunk) splits into equal contiguous pieces."""
    chunk = -(-width // world)
    begin = min(rank * chunk, width)
    end = min(begin + chunk, width)
    return begin, end, chunk


class B200Engine:
    """Engine over the CUDA library with the wire buffer held in a torch tensor (device memory + stream plumbing)."""

    def __init__(self, backend, cprog, B, torch_mod):
        self.be, self.cp, self.B, self.torch = backend, cprog, B, torch_mod
        words = backend.wires_bytes(cprog, B) // 8
        self.wires = torch_mod.empty(words, dtype=torch_mod.int64, device=f"cuda:{backend.device}")
        self.ct_words = backend.params.ct_words
        self.stream = torch_mod.cuda.current_stream().cuda_stream

    def encrypt(self, bits: np.ndarray, inst_offset=0, total=None, enc_seed=None):
        d_in = self.torch.from_numpy(np.ascontiguousarray(bits, dtype=np.uint8)).to(self.wires.device)
        self.be.encrypt_inputs(self.cp, d_in.data_ptr(), self.B, self.wires.data_ptr(), stream=self.stream,
                               inst_offset=inst_offset, total=total or self.B, enc_seed=enc_seed)
        self.torch.cuda.current_stream().synchronize()

    def run_level(self, level, node_begin, node_end):
        self.be.run_level(self.cp, level, self.B, self.wires.data_ptr(), node_begin, node_end, stream=self.stream)

    def decrypt(self) -> np.ndarray:
        n_out = len(self.cp.program.output_names)
        d_out = self.torch.empty((n_out, self.B), dtype=self.torch.uint8, device=self.wires.device)
        self.be.decrypt_outputs(self.cp, self.B, self.wires.data_ptr(), d_out.data_ptr(), stream=self.stream)
        self.torch.cuda.current_stream().synchronize()
        return d_out.cpu().numpy()

    def slot_view(self, slot_begin, n_slots):
        per = self.B * self.ct_words
        return self.wires[slot_begin * per:(slot_begin + n_slots) * per]


class FusedB200Engine(B200Engine):
    """Node-sharded engine whose blind-rotation epilogue stores every output ciphertext into all peers' wire replicas
    (peer-mapped NVLink stores): the compute step and the exchange are ONE kernel, and the levels are ordered ON THE
    DEVICE (per-level epoch flags in the peers' flag pages, ``fbs_set_peers``): the host neither synchronises nor barriers
    between levels.  The wire buffer is a cudaMalloc allocation of the library, shared between the per-GPU processes by
    CUDA IPC.  Use as a context manager (or call ``close()``): the peer binding is dropped even if a level raises."""

    fused = True

    def __init__(self, backend, cprog, B, torch_mod, dist, world, rank, handoff="device"):
        """``handoff``: "device" (default) = per-level epoch flags, no host sync between levels; "host" = round-1 behaviour
        kept for A/B measurements: ``torch.cuda.synchronize()`` + ``dist.barrier()`` after every level."""
        assert handoff in ("device", "host")
        self.be, self.cp, self.B, self.torch = backend, cprog, B, torch_mod
        self.ct_words = backend.params.ct_words
        self.stream = torch_mod.cuda.current_stream().cuda_stream
        self.nbytes = backend.wires_bytes(cprog, B)
        self.dist, self.world, self.rank, self.handoff = dist, world, rank, handoff
        self.ptr, self.peer_ptrs = backend.wires_alloc(self.nbytes), []
        try:
            handles = [None] * world
            dist.all_gather_object(handles, backend.ipc_export(self.ptr))
            self.peer_ptrs = [backend.ipc_import(h) for r, h in enumerate(handles) if r != rank]
            backend.set_peers(self.ptr, self.nbytes, self.peer_ptrs, rank if handoff == "device" else -1)
            dist.barrier()                     # every rank's flag page is zeroed before anybody signals
        except Exception:
            self.close(collective=False)
            raise

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close(collective=exc[0] is None)

    class _Ptr:
        def __init__(self, p):
            self.p = p

        def data_ptr(self):
            return self.p

    @property
    def wires(self):
        return FusedB200Engine._Ptr(self.ptr)

    def encrypt(self, bits, inst_offset=0, total=None, enc_seed=None):
        d_in = self.torch.from_numpy(np.ascontiguousarray(bits, dtype=np.uint8)).cuda()
        self.be.encrypt_inputs(self.cp, d_in.data_ptr(), self.B, self.ptr, stream=self.stream, inst_offset=inst_offset, total=total or self.B,
                               enc_seed=enc_seed)
        self.torch.cuda.synchronize()
        self.dist.barrier()                    # inputs are in place on every replica before any level starts

    def run_level(self, level, node_begin, node_end):
        # called for EVERY level on every rank (also with an empty range): the call carries the level's wait + signal
        self.be.run_level(self.cp, level, self.B, self.ptr, node_begin, node_end, stream=self.stream)
        if self.handoff == "host":
            self.torch.cuda.synchronize()      # this rank's peer stores are complete ...
            self.dist.barrier()                # ... and so are everybody else's

    def decrypt(self):
        n_out = len(self.cp.program.output_names)
        d_out = self.torch.empty((n_out, self.B), dtype=self.torch.uint8, device="cuda")
        self.be.decrypt_outputs(self.cp, self.B, self.ptr, d_out.data_ptr(), stream=self.stream)
        self.torch.cuda.synchronize()
        return d_out.cpu().numpy()

    def finish(self):
        """End of a run: this rank's last level has completed AND every peer's stores into this replica have landed."""
        if self.handoff == "host":
            return
        self.be.run_level_sync(self.ptr, stream=self.stream)
        self.torch.cuda.synchronize()
        lost = self.be.sync_status()
        if lost:
            raise RuntimeError(f"device-side level hand-off timed out waiting for rank {lost - 1}")

    def close(self, collective=True):
        if self.ptr is None:
            return
        try:
            self.torch.cuda.synchronize()
        finally:
            self.be.set_peers(None, 0, [], 0)
            for p in self.peer_ptrs:
                self.be.ipc_close(p)
            self.peer_ptrs = []
            if collective:
                self.dist.barrier()            # nobody still stores into a buffer that is about to be freed
            self.be.wires_free(self.ptr)
            self.ptr = None


def run_node_sharded(engine, program, dist, world: int, rank: int, in_place: bool = True):
    """Evaluate all levels with the bootstraps of each level split over ``world`` ranks and an all-gather of the new
    output ciphertexts after every level.  ``program`` must come from ``levelize(..., shard_pad=world)``.
    Returns the number of words exchanged (for reporting)."""
    assert program.contiguous_levels and (world == 1 or program.shard_pad == world), \
        "program must be levelised with shard_pad=world (reuse_slots=False when world == 1)"
    a = program.arrays
    exchanged = 0
    multi = getattr(program, "multi_value", False)
    assert not multi or world == 1 or getattr(engine, "fused", False), \
        "multi-value programs shard by groups: their outputs are not equal contiguous slot chunks, use the fused exchange"
    for lv in range(program.n_levels):
        b0, b1 = int(a["bs_level_ptr"][lv]), int(a["bs_level_ptr"][lv + 1])
        # the sharding unit: bootstraps, or (multi-value) the groups of bootstraps that share one blind rotation
        width = int(a["grp_level_ptr"][lv + 1] - a["grp_level_ptr"][lv]) if multi else b1 - b0
        nb, ne, chunk = level_node_range(width, world, rank)
        fused = getattr(engine, "fused", False) and world > 1
        if ne > nb or fused:                     # fused: an empty range still waits for / signals the level on the device
            engine.run_level(lv, nb, ne)
        if world == 1:
            continue
        if fused:                                # outputs already sit in every replica, levels are ordered by device flags
            exchanged += (b1 - b0 if multi else chunk * world) * engine.B * engine.ct_words
            continue
        first_slot = int(a["bs_slot"][b0])
        region = engine.slot_view(first_slot, chunk * world)
        mine = engine.slot_view(first_slot + rank * chunk, chunk)
        if in_place:
            dist.all_gather_into_tensor(region, mine)
        else:   # backends without an in-place path: gather into views of the region
            parts = [engine.slot_view(first_slot + r * chunk, chunk) for r in range(world)]
            dist.all_gather(parts, mine.clone())
        exchanged += region.numel()
    if getattr(engine, "fused", False) and world > 1:
        engine.finish()
    return exchanged


def eval_instances_sharded(backend, cprog, bits: np.ndarray, world: int, rank: int):
    """Instance-sharded evaluation through the host-buffer call: this rank evaluates its slice, no collective."""
    total = bits.shape[1]
    off, cnt = instance_shard(total, world, rank)
    if cnt == 0:
        return off, np.zeros((len(cprog.program.output_names), 0), np.uint8)
    out = backend.eval_bits(cprog, np.ascontiguousarray(bits[:, off:off + cnt]), inst_offset=off, total=total)
    return off, out
