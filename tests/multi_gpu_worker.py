"""Worker of tests/test_gpu_multi.py: launched by torchrun with one process per GPU.

Every rank evaluates the SAME encrypted instances of one circuit with the bootstraps of each level split across the ranks
(BASELINE.json configs[3]) in three ways -- fused peer-store epilogue + device-side level hand-off, fused + host barrier,
NCCL all-gather -- and checks that its replica of the wire buffer ends up bit-identical to a plain one-GPU run of the same
program on the same encrypted inputs, and that the decrypted outputs equal the cleartext interpreter (oracle)."""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--param-set", default="toy3")
    ap.add_argument("--lbf", default="aes_sbox_p11.lbf")
    ap.add_argument("--p", type=int, default=11)
    ap.add_argument("--batch", type=int, default=4)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from conftest import read_golden_lbf
    from oracle import cleartext
    from tfhe_fbs_map_b200 import levelize
    from tfhe_fbs_map_b200.backend import B200Backend
    from tfhe_fbs_map_b200.dist import B200Engine, FusedB200Engine, run_node_sharded

    env = read_golden_lbf(args.lbf)
    prog = levelize(env, args.p, shard_pad=world)
    be = B200Backend(args.param_set, device=local, seed=33)            # same seed on every rank: identical keys, no broadcast
    cp = be.load(prog)
    B = args.batch
    bits = np.random.default_rng(5).integers(0, 2, (prog.n_inputs, B)).astype(np.uint8)
    want = cleartext.lut_eval(env, {nm: bits[i] for i, nm in enumerate(prog.input_names)})
    want_mat = np.array([np.asarray(want[nm]) for nm in prog.output_names], dtype=np.uint8)
    a = prog.arrays
    boot_slots = [int(s) for s in a["bs_slot"]]
    cudart = ctypes.CDLL("libcudart.so")

    def slots_of(ptr, nbytes):
        t = torch.empty(nbytes // 8, dtype=torch.int64, device="cuda")
        cudart.cudaMemcpy(ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(ptr), ctypes.c_size_t(nbytes), 3)
        return t.cpu().numpy().reshape(prog.n_slots, B, be.params.ct_words)[boot_slots]

    # plain one-GPU run of the same (padded) program on the same encrypted inputs
    solo = B200Engine(be, cp, B, torch)
    solo.encrypt(bits, enc_seed=77)
    be.run(cp, B, solo.wires.data_ptr(), stream=solo.stream)
    torch.cuda.synchronize()
    ref_slots = slots_of(solo.wires.data_ptr(), be.wires_bytes(cp, B))
    assert np.array_equal(solo.decrypt(), want_mat), "one-GPU run differs from the cleartext oracle"

    for mode in ("fused-device", "fused-host", "nccl"):
        if mode == "nccl":
            eng = B200Engine(be, cp, B, torch)
        else:
            eng = FusedB200Engine(be, cp, B, torch, dist, world, rank, handoff=mode.split("-")[1])
        eng.encrypt(bits, enc_seed=77)
        for _ in range(2):                                            # twice: epochs keep counting, results identical
            run_node_sharded(eng, prog, dist, world, rank)
        torch.cuda.synchronize()
        dist.barrier()
        got = slots_of(eng.wires.data_ptr(), be.wires_bytes(cp, B))
        bad = [q for q in range(len(boot_slots)) if not np.array_equal(got[q], ref_slots[q])]
        assert not bad, f"{mode}: rank {rank} replica differs from the one-GPU wire buffer at bootstraps {bad[:8]}"
        assert np.array_equal(eng.decrypt(), want_mat), f"{mode}: decrypted outputs differ from the cleartext oracle"
        if mode != "nccl":
            eng.close()
        dist.barrier()
    print(f"MULTI_OK rank {rank}/{world} {args.param_set} {args.lbf} B={B}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
