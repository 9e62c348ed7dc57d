#!/usr/bin/env python
"""bench.py -- encrypted circuit evaluation throughput (PBS/s) on 1..8 B200, next to the host-CPU baseline.

Metric and workload (BASELINE.json): PBS/sec and encrypted circuit evals/sec.  At N=1 the workload is
configs[1]: a 128-bit adder mapped with --fbs_size 15 --mapper search, evaluated as level-batched programmable
bootstraps.  EPFL adder.blif is not available offline, so the circuit is the synthetic AIG ripple-carry stand-in
(tfhe_fbs_map_b200/circuits.py) pre-mapped by the REFERENCE mapper into tests/golden/lbf/adder128_p15.lbf
(oracle/gen_golden.py).  One "step" = one pass of the whole circuit (255 bootstraps, 127 levels) over a batch of B
independent encrypted instances per GPU; instances shard across GPUs with keys replicated and no collective
(scaling = weak).

  value : PBS/s with the encrypted inputs already resident in HBM (timed: K x fbs_run, CUDA events, max over ranks)
  e2e   : PBS/s through the drop-in call fbs_eval_bits with pinned HOST buffers (H2D bits, encrypt, all levels,
          decrypt, D2H bits inside the timed region)
  roofline     : schema-mandated HBM view of the dominant kernel (k_blind_rotate)
  roofline_int : the bound that actually binds (integer multiply issue): mul32/s vs measured IMAD.WIDE peak
  cpu_baseline : the CPU oracle (oracle/tfhe_ref.c, same parameter set) on a bounded sample, all host cores

`--impl reference` times the reference-side CPU implementation of the path (the oracle port: the reference has no
encrypted path of its own and concrete cannot be built offline) with all host threads.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "adder128_p15": dict(lbf="adder128_p15.lbf", p=15, cpu_sample="adder8_p15.lbf",
                         desc="128-bit AIG ripple-carry adder (synthetic stand-in for EPFL adder.blif), fbs_size 15, search mapper"),
    "mult16_p17": dict(lbf="mult16_p17.lbf", p=17, cpu_sample="mult8_p17.lbf", desc="16x16 array multiplier, fbs_size 17"),
    "aes_sbox_p11": dict(lbf="aes_sbox_p11.lbf", p=11, cpu_sample="aes_sbox_p11.lbf", desc="AES s-box non-linear core, fbs_size 11"),
    # BASELINE configs[2]: the full cipher (tfhe_fbs_map_b200/circuits.py generator, FIPS-197 verified; Bristol aes_128.txt is not
    # available offline), 14 954 bootstraps per instance: time-boxed with a small per-GPU batch, evals/s extrapolates linearly
    "aes128_p11": dict(lbf="aes128_r10_p11.lbf.gz", p=11, cpu_sample="aes_sbox_p11.lbf", batch=16,
                       desc="AES-128 (10 rounds, key schedule included; in-repo generator standing in for Bristol aes_128.txt), fbs_size 11, search mapper"),
}


def load_env(fn):
    from tfhe_fbs_map_b200.formats import read_lbf, read_lbf_file
    path = os.path.join(ROOT, "tests", "golden", "lbf", fn)
    if fn.endswith(".gz"):
        import gzip
        with gzip.open(path, "rt") as f:
            return read_lbf(f.read())
    return read_lbf_file(path)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.stop_flag = gpu_index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 7 and s[3 + i].lower().startswith("active")})
        busy = [v for v in sm if mx and v > 0.5 * mx[0]] or sm
        return dict(sm_mhz=statistics.median(busy) if busy else None, sm_max_mhz=mx[0] if mx else None, reasons=reasons,
                    samples=len(self.samples))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return dict(hbm_gbs=6650.0), "fallback"


def cpu_baseline(ps, sample_fn, p, seed, threads, instances=None):
    """CPU oracle on a bounded sample: `instances` encrypted evaluations of a small circuit of the same kind."""
    from oracle.tfhe_ref import RefTFHE, lib
    from tfhe_fbs_map_b200 import levelize
    env = load_env(sample_fn)
    prog = levelize(env, p)
    if getattr(ps, "bsk_unroll", 1) > 1:
        # the CPU runs the classic one-bit-per-step blind rotation of the same shape: the oracle's literal key-unrolled
        # restatement does 2^m - 1 external products per key group and would make the CPU look slower than it is
        from dataclasses import asdict
        from tfhe_fbs_map_b200.params import ParamSet
        d = asdict(ps); d.update(bsk_unroll=1, name=ps.name + " shape, classic blind rotation")
        ps = ParamSet(**d)
    L = lib()
    if not threads:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = L.ref_max_threads()
    cores = threads
    B = instances or cores
    ref = RefTFHE(ps, seed=seed)
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2, (prog.n_inputs, B)).astype(np.uint8)
    t0 = time.time()
    out = ref.eval_prog(prog, bits, enc_seed=7, threads=cores)
    dt = time.time() - t0
    from oracle import cleartext
    want = cleartext.lut_eval(env, {nm: bits[i] for i, nm in enumerate(prog.input_names)})
    ok = all(np.array_equal(out[prog.out_index[nm]], np.asarray(want[nm])) for nm in prog.output_names)
    # the reference's literal path: cleartext LutExecEnv.eval restatement, 1 core
    big = rng.integers(0, 2, (prog.n_inputs, 20000)).astype(np.uint8)
    t1 = time.time()
    cleartext.lut_eval(env, {nm: big[i] for i, nm in enumerate(prog.input_names)})
    dtc = time.time() - t1
    return dict(value=prog.n_boots * B / dt, unit="PBS/s", cores=cores, kind="port", wall_s=dt,
                sample=f"{sample_fn}: {prog.n_boots} PBS/instance x {B} instances, parameter set {ps.name}, {dt:.1f} s wall; decrypted == cleartext: {ok}",
                cleartext_lookups_per_s=prog.n_boots * 20000 / dtc, evals_per_s=B / dt)


def run_reference(args, wl, ps, rank, world):
    if rank != 0:
        return
    from oracle.tfhe_ref import lib
    # every host thread this process may use: torchrun exports OMP_NUM_THREADS=1, which ref_max_threads() would obey
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or lib().ref_max_threads()
    vals = []
    for step in range(args.warmup + args.steps):
        cb = cpu_baseline(ps, wl["cpu_sample"], wl["p"], args.seed, cores, instances=max(1, cores // 2) if step < args.warmup else cores)
        if step >= args.warmup:
            vals.append(cb)
    v = statistics.mean(x["value"] for x in vals)
    line = dict(metric="PBS/sec", value=v, unit="PBS/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * statistics.mean(x["wall_s"] for x in vals),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64 (q = p1*p2, two 30-bit NTT primes; RNS u32x2 in the NTT)", data="synthetic",
                impl="reference", config=dict(workload=args.workload, param_set=ps.name, note="CPU port of the path (oracle/tfhe_ref.c); the reference has no encrypted executor and concrete cannot be built offline"),
                cpu_baseline=dict(vals[-1], value=v),
                e2e=dict(value=v, unit="PBS/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def run_nodes(args, wl, ps, be, rank, world, local, torch, dist):
    """BASELINE configs[3]: ONE circuit, node-sharded levels, NCCL all-gather of output LWEs per level."""
    from tfhe_fbs_map_b200 import levelize
    from tfhe_fbs_map_b200.dist import B200Engine, FusedB200Engine, run_node_sharded
    from oracle import cleartext
    env = load_env(wl["lbf"])
    prog = levelize(env, wl["p"], shard_pad=world, reuse_slots=False)
    cp = be.load(prog)
    B = args.batch
    rng = np.random.default_rng(77)                      # every rank evaluates the SAME instances
    bits = rng.integers(0, 2, (prog.n_inputs, B)).astype(np.uint8)
    want = cleartext.lut_eval(env, {nm: bits[i] for i, nm in enumerate(prog.input_names)})
    want_mat = np.array([np.asarray(want[nm]) for nm in prog.output_names], dtype=np.uint8)
    fused = args.exchange == "fused" and world > 1
    eng = FusedB200Engine(be, cp, B, torch, dist, world, rank) if fused else B200Engine(be, cp, B, torch)
    eng.encrypt(bits)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        run_node_sharded(eng, prog, dist, world, rank)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    words = 0
    for _ in range(args.steps):
        words += run_node_sharded(eng, prog, dist, world, rank)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sampler.stop_flag.set(); sampler.join(timeout=2)
    out = eng.decrypt()
    mism = int((out != want_mat).sum())
    if rank == 0:
        n_pbs = prog.n_boots * B * args.steps
        print(json.dumps(dict(
            metric="PBS/sec", value=n_pbs / (ms * 1e-3), unit="PBS/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="u64 (q = p1*p2, two 30-bit NTT primes; RNS u32x2 in the NTT)",
            data="synthetic", config=dict(workload=args.workload, desc=wl["desc"], param_set=ps.name, fbs_size=wl["p"], instances=B,
                                          pbs_per_instance=prog.n_boots, levels=prog.n_levels, level_width_median=int(np.median(prog.level_widths)),
                                          sharding="nodes of each level split across GPUs; " + ("sample-extract epilogue stores output LWE ciphertexts into every peer replica over NVLink (fused compute+exchange), host barrier per level" if fused else "NCCL in-place all-gather of output LWE ciphertexts per level")),
            evals_per_s=B * args.steps / (ms * 1e-3), mismatches=mism,
            allgather_bytes_per_step=words * 8 // max(1, args.steps), gpu_launches=3 * prog.n_levels * args.steps, clocks=sampler.summary())))
    if fused:
        eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="adder128_p15", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="encrypted instances per GPU per step (default 296; 16 for aes128_p11)")
    ap.add_argument("--param-set", default="A3", help="A3 / A2 = set A with three (default) / two key bits per blind-rotation step; A = classic")
    ap.add_argument("--seed", type=int, default=20241018)
    ap.add_argument("--shard", default="instances", choices=["instances", "nodes"],
                    help="instances: batch split across GPUs, no collective (weak scaling); nodes: one circuit, each level's "
                         "bootstraps split across GPUs + NCCL all-gather of the output LWE ciphertexts per level (strong scaling)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="--shard nodes: 'fused' = sample-extract epilogue stores into all peers' replicas (one kernel does "
                         "compute + exchange), 'nccl' = separate in-place all-gather per level")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from tfhe_fbs_map_b200 import params, levelize
    ps = params.get(args.param_set)
    wl = WORKLOADS[args.workload]
    if args.batch is None:
        args.batch = wl.get("batch", 296)

    if args.impl == "reference":
        run_reference(args, wl, ps, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the encrypted executor has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tfhe_fbs_map_b200.backend import B200Backend, RunStats

    be = B200Backend(ps, device=local, seed=args.seed)          # identical seeded keys on every rank: no key broadcast
    if args.shard == "nodes":
        run_nodes(args, wl, ps, be, rank, world, local, torch, dist)
        return
    env = load_env(wl["lbf"])
    prog = levelize(env, wl["p"], preserve_inputs=True)      # steps re-run the same resident encrypted inputs
    cp = be.load(prog)
    B = args.batch
    n_in, n_out, n_pbs_step = prog.n_inputs, len(prog.output_names), prog.n_boots * B
    rng = np.random.default_rng(1000 + rank)
    bits = rng.integers(0, 2, (n_in, B)).astype(np.uint8)
    from oracle import cleartext                                    # checker only: decrypted outputs vs cleartext semantics
    want = cleartext.lut_eval(env, {nm: bits[i] for i, nm in enumerate(prog.input_names)})
    want_mat = np.array([np.asarray(want[nm]) for nm in prog.output_names], dtype=np.uint8)

    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream
    wires = torch.empty(be.wires_bytes(cp, B) // 8, dtype=torch.int64, device="cuda")
    d_in = torch.from_numpy(bits).cuda()
    d_out = torch.empty((n_out, B), dtype=torch.uint8, device="cuda")
    be.encrypt_inputs(cp, d_in.data_ptr(), B, wires.data_ptr(), stream=sp, inst_offset=rank * B, total=world * B)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident: K x fbs_run ----------------------------------------------------------------------------
    for _ in range(args.warmup):
        be.run(cp, B, wires.data_ptr(), stream=sp)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    st = RunStats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        be.run(cp, B, wires.data_ptr(), stream=sp, stats=st)
    e1.record(stream)
    barrier()
    ms_res = max_over_ranks(e0.elapsed_time(e1))
    be.decrypt_outputs(cp, B, wires.data_ptr(), d_out.data_ptr(), stream=sp)
    torch.cuda.synchronize()
    mism_res = int((d_out.cpu().numpy() != want_mat).sum())

    # ---- end to end through the drop-in call, pinned host buffers ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_in = torch.from_numpy(bits).pin_memory()
        h_out = torch.empty((n_out, B), dtype=torch.uint8).pin_memory()
        for _ in range(min(args.warmup, 1)):
            be.eval_bits(cp, None, in_ptr=h_in.data_ptr(), out_ptr=h_out.data_ptr(), B=B, inst_offset=rank * B, total=world * B)
        barrier()
        t0 = time.perf_counter()
        st2_launch, st2_ms = 0, 0.0
        for _ in range(args.steps):
            be.eval_bits(cp, None, in_ptr=h_in.data_ptr(), out_ptr=h_out.data_ptr(), B=B, inst_offset=rank * B, total=world * B)
            st2_launch += be.last_stats["n_launches"]
            st2_ms += be.last_stats["ms_total"]
        barrier()
        ms_e2e = max_over_ranks(st2_ms)                       # device time (CUDA events inside fbs_eval_bits), incl. copies
        wall_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3)
        mism_e2e = int((h_out.numpy() != want_mat).sum())
        e2e = dict(value=world * n_pbs_step * args.steps / (max(ms_e2e, wall_e2e) * 1e-3), unit="PBS/s", h2d_bytes_per_step=n_in * B,
                   d2h_bytes_per_step=n_out * B, ms_per_step=max(ms_e2e, wall_e2e) / args.steps, mismatches=mism_e2e,
                   evals_per_s=world * B * args.steps / (max(ms_e2e, wall_e2e) * 1e-3))
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        int_peak = be.measure_int_peak()
        n_br = prog.n_levels * args.steps
        br_ms_avg = st.ms_blind_rotate / n_br
        pbs_per_launch = n_pbs_step / prog.n_levels
        info = be.info()
        # algorithmic HBM bytes of one blind-rotate launch: BSK once (shared by every PBS of the launch via L2),
        # per PBS the mod-switched LWE in and the extracted LWE out
        alg_bytes = info["bsk_bytes"] + pbs_per_launch * ((ps.n + 1) * 2 + ps.ct_words * 8)
        achieved = alg_bytes / (br_ms_avg * 1e-3) / 1e9
        mul32 = ps.mul32_per_pbs() * n_pbs_step * args.steps / (ms_res * 1e-3)
        value = world * n_pbs_step * args.steps / (ms_res * 1e-3)
        line = dict(
            metric="PBS/sec", value=value, unit="PBS/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_res / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="u64 (q = p1*p2, two 30-bit NTT primes; RNS u32x2 in the NTT)", data="synthetic",
            config=dict(workload=args.workload, desc=wl["desc"], param_set=ps.name, n=ps.n, k=ps.k, N=ps.N, bsk_l=ps.bsk_l, ks_l=ps.ks_l, bsk_unroll=ps.bsk_unroll,
                        fbs_size=wl["p"], instances_per_gpu=B, pbs_per_instance=prog.n_boots, levels=prog.n_levels,
                        p_fail_per_pbs=ps.p_fail(wl["p"], env.stats()["norm2_linprod"]), sharding="instances, keys replicated, no collective",
                        l2="wire buffer %.2f GB per GPU > 126 MB L2; BSK+KSK (%.0f MB) are re-streamed every level" % (wires.numel() * 8 / 1e9, (info["bsk_bytes"] + info["ksk_bytes"]) / 1e6)),
            evals_per_s=world * B * args.steps / (ms_res * 1e-3), mismatches=mism_res,
            phase_ms_per_step=dict(lincomb=st.ms_lincomb / args.steps, keyswitch=st.ms_keyswitch / args.steps, blind_rotate=st.ms_blind_rotate / args.steps),
            roofline=dict(bound="hbm", achieved=achieved, peak=peaks.get("hbm_gbs"), unit="GB/s", frac=achieved / peaks.get("hbm_gbs"),
                          traffic={3: 238.35e6, 2: 154.63e6}.get(ps.bsk_unroll, 61.95e6) / 1e9,
                          traffic_note="GB per launch: dram__bytes_read+write of a 592-PBS launch (two waves), profiles/" + ("r1_v13_hot_kernels_summary.txt; the key-unrolled BSK (114 MB at 3 bits per step) does not stay L2-resident between waves and is re-read from HBM once per wave" if ps.bsk_unroll > 1 else "r1_v5_hot_kernels_summary.txt"),
                          kernel="k_blind_rotate2" if ps.bsk_unroll > 1 else "k_blind_rotate", launch_ms=br_ms_avg, peak_source=peak_kind,
                          note="kernel is integer-issue bound by design (accumulator on chip, keys L2-resident); see roofline_int"),
            roofline_int=dict(bound="int32-multiply", achieved=mul32 / 1e12, peak=int_peak / 1e12, unit="T mul32/s", frac=mul32 / int_peak,
                              mul32_per_pbs=ps.mul32_per_pbs(), modmul_per_pbs=ps.modmul_per_pbs(), peak_source="measured (fbs_measure_int_peak: mad.wide.u32 chains)"),
            e2e=e2e, gpu_launches=int(st.n_launches), clocks=sampler.summary())
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(ps, wl["cpu_sample"], wl["p"], args.seed, 0)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
