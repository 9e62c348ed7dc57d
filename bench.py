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
  roofline     : schema-mandated HBM view of the dominant kernel (k_blind_rotate2)
  roofline_int : the bound that actually binds (integer multiply issue): blind-rotation mul32/s over blind-rotation time vs
                 the measured IMAD.WIDE peak, as executed-work fraction and as fraction of SURVEY 8(d)'s canonical count
  cpu_baseline : the TUNED CPU arm (baseline/cpu_pbs.cpp: RNS Harvey/Shoup NTTs, -O3 -march=native, OpenMP) on a bounded
                 sample of the SAME workload, all host cores; plus the parity oracle's rate and the reference's literal
                 cleartext LutExecEnv.eval (BASELINE configs[0]) on 1 core and on all cores
  nodes        : BASELINE configs[3]: ONE 64x64 multiplier (fbs_size 17) with the bootstraps of every level split across the
                 GPUs, fused peer-store exchange + device-side level hand-off, at batch 1 and 64 (strong scaling)

`--impl reference` times the CPU implementation of the path (the tuned arm: the reference has no encrypted path of its
own and concrete cannot be built offline) with all host threads on the same workload, parameter set and metric.
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
    # BASELINE configs[3]: 64x64 -> 128 multiplier (EPFL multiplier.blif is not available offline: array multiplier from
    # tfhe_fbs_map_b200/circuits.py, mapped by the REFERENCE mapper, oracle/gen_mult64.py), 8 126 bootstraps, 250 levels, median width 42
    "mult64_p17": dict(lbf="mult64_p17.lbf.gz", p=17, cpu_sample="mult8_p17.lbf", batch=64,
                       desc="64x64 array multiplier (synthetic stand-in for EPFL multiplier.blif), fbs_size 17, search mapper"),
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


# DRAM bytes (read + write) of one blind-rotate launch, from ncu --set full captures under profiles/ (NOT measured in the bench run itself)
TRAFFIC_GB = {3: 0.23771, 2: 0.15463, 1: 0.06195, 0: 0.06195}
TRAFFIC_SOURCE = ("GB per 592-PBS launch from profiles/r2_v7_hot_kernels_summary.txt (ncu dram__bytes_read.sum + dram__bytes_write.sum), not measured in this run; "
                  "the key-unrolled BSK (114 MB at 3 bits per step) does not stay L2-resident between waves and is re-read from HBM once per wave")


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return dict(hbm_gbs=6650.0), "fallback"


def host_cores():
    """Every host thread this process may use (torchrun exports OMP_NUM_THREADS=1: do not trust omp_get_max_threads)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_CPU_ARM = {}


def cpu_arm(ps, seed):
    """(oracle context, tuned CPU context) with the same seeded keys; built on first use (-march=native: per box)."""
    if ps.bsk_unroll not in (2, 3):      # the CPU arm is written for the key-unrolled one-level shape: use the set's three-bit twin
        from dataclasses import asdict
        from tfhe_fbs_map_b200.params import ParamSet
        d = asdict(ps); d.update(bsk_unroll=3, name=ps.name + " shape, three key bits per step")
        ps = ParamSet(**d)
    key = (ps.name, seed)
    if key not in _CPU_ARM:
        from oracle.tfhe_ref import RefTFHE
        from baseline.cpu_arm import CpuTFHE
        ref = RefTFHE(ps, seed=seed)
        _CPU_ARM[key] = (ref, CpuTFHE(ps, ref))
    return _CPU_ARM[key]


CPU_ARM_DESC = ("tuned CPU arm baseline/cpu_pbs.cpp: same specification and parameter set as the GPU path (bit-identical ciphertexts, "
                "tests/test_cpu_arm.py), RNS over the two 30-bit NTT primes, Harvey/Shoup butterflies auto-vectorised by g++ -O3 -march=native, "
                "Montgomery point-wise products, three key bits per blind-rotation step, OpenMP over instances; concrete's Rust/FFT CPU PBS "
                "cannot be built offline and is not substituted")


def cpu_sample_step(ps, wl, seed, cores, levels=32, enc_seed=7):
    """One bounded CPU step of the benchmark workload: the first `levels` levels of the SAME program (the per-bootstrap cost does
    not depend on the level) on `cores` encrypted instances, one instance per host thread."""
    from tfhe_fbs_map_b200 import levelize
    _, cpu = cpu_arm(ps, seed)
    prog = levelize(load_env(wl["lbf"]), wl["p"])
    bits = np.random.default_rng(1).integers(0, 2, (prog.n_inputs, cores)).astype(np.uint8)
    t0 = time.time()
    cpu.eval_prog(prog, bits, enc_seed=enc_seed, threads=cores, max_levels=levels)
    dt = time.time() - t0
    lv = min(levels, prog.n_levels)
    return dict(value=cpu.last_pbs * cores / dt, wall_s=dt, pbs=cpu.last_pbs * cores,
                sample=f"levels 0..{lv - 1} of the {prog.n_levels} levels of {wl['lbf']} ({cpu.last_pbs} of {prog.n_boots} PBS per instance) x {cores} instances "
                       f"(one per host thread), parameter set {ps.name}, {dt:.1f} s wall")


def cpu_verify(ps, wl, seed, cores):
    """Correctness of the tuned CPU arm on this box: a whole small circuit of the same kind decrypts to the cleartext outputs."""
    from tfhe_fbs_map_b200 import levelize
    from oracle import cleartext
    _, cpu = cpu_arm(ps, seed)
    env = load_env(wl["cpu_sample"])
    prog = levelize(env, wl["p"])
    B = max(2, min(cores, 8))
    bits = np.random.default_rng(2).integers(0, 2, (prog.n_inputs, B)).astype(np.uint8)
    out = cpu.eval_prog(prog, bits, enc_seed=9, threads=cores)
    want = cleartext.lut_eval(env, {nm: bits[i] for i, nm in enumerate(prog.input_names)})
    return all(np.array_equal(out[prog.out_index[nm]], np.asarray(want[nm])) for nm in prog.output_names)


def _literal_worker(args):
    lbf_text, p, B, seed, use_ref = args
    env, evaluate = literal_evaluator(lbf_text, use_ref)
    rng = np.random.default_rng(seed)
    inputs = {nm: rng.integers(0, 2, B) for nm in [i.name for i in env.instructions if type(i).__name__ == "Input"]}
    t0 = time.time()
    evaluate(inputs)
    return time.time() - t0


def literal_evaluator(lbf_text, use_ref):
    """The reference's own LutExecEnv.eval (imported from /root/reference/fbs_mapper when mounted) on the program of a golden
    .lbf, else the line-by-line restatement oracle/cleartext.py:lut_eval_literal (the GPU box has no /root/reference)."""
    from tfhe_fbs_map_b200.formats import read_lbf
    env = read_lbf(lbf_text)
    if use_ref:
        sys.path.insert(0, "/root/reference/fbs_mapper")
        import fbs_exec_env as ref_lut
        r, m = ref_lut.LutExecEnv(), {}
        for ins in env.instructions:
            k = type(ins).__name__
            if k == "Input":
                m[ins.name] = r.input(ins.name)
            elif k == "LinearProd":
                m[ins.name] = r.linear([c for c, _ in ins.coef_vals], [m[v.name] for _, v in ins.coef_vals], ins.const_coef)
            elif k == "Bootstrap":
                m[ins.name] = r.bootstrap(m[ins.val.name], list(ins.table))
        for name, out in env.outputs.items():
            r.output(name, m[out.name])
        return env, r.eval
    from oracle import cleartext
    return env, (lambda iv: cleartext.lut_eval_literal(env, iv))


def reference_literal(cores, B=100000):
    """BASELINE configs[0]: smallest generated circuit (half_adder) mapped with --fbs_size 15 --mapper search, evaluated by the
    reference's cleartext LutExecEnv.eval on random inputs: 1 core (the reference is single-threaded) and all cores
    (independent processes, inputs sharded)."""
    import multiprocessing as mp
    entries = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_mapped.json")))
    e = next(x for x in entries if x["circuit"] == "half_adder" and x["p"] == 15 and x["mapper"] == "search")
    use_ref = os.path.isdir("/root/reference/fbs_mapper")
    n_fbs = e["stats"]["nb_bootstrap"]
    t1 = _literal_worker((e["lbf"], 15, B, 1, use_ref))
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.time()
        pool.map(_literal_worker, [(e["lbf"], 15, B, 10 + i, use_ref) for i in range(cores)])
        tall = time.time() - t0
    return dict(circuit="half_adder, fbs_size 15, search mapper (reference-mapped golden)", fbs_per_eval=n_fbs, batch=B,
                implementation="reference fbs_mapper/fbs_exec_env.py LutExecEnv.eval (imported from /root/reference)" if use_ref else
                "oracle/cleartext.py lut_eval_literal (line-by-line restatement of fbs_exec_env.py:208-229; /root/reference is not mounted on this box)",
                lookups_per_s_1core=n_fbs * B / t1, evals_per_s_1core=B / t1, cores=cores,
                lookups_per_s_all_cores=n_fbs * B * cores / tall, evals_per_s_all_cores=B * cores / tall)


def run_config1(args, ps):
    """BASELINE configs[0]: the reference's own CPU path (cleartext LutExecEnv.eval, fbs_exec_env.py:208-229) on the smallest
    generated circuit, 1 core and all cores; when a GPU is present also the same program through this library."""
    cores = host_cores()
    lit = reference_literal(cores)
    line = dict(metric="cleartext FBS look-ups/sec", value=lit["lookups_per_s_all_cores"], unit="look-ups/s", n_gpus=0, steps=1, warmup=0,
                higher_is_better=True, data="synthetic (uniform random input bits)", impl="reference",
                config=dict(workload="half_adder_p15_search_cleartext", desc="BASELINE configs[0]: smallest circuit of experiments/generate_benchmarks.py "
                            "(half_adder, 2 FBS, 1 level) mapped with --fbs_size 15 --mapper search, cleartext LutExecEnv.eval", batch=lit["batch"]),
                cpu_baseline=dict(lit, value=lit["lookups_per_s_all_cores"], unit="look-ups/s", kind="reference" if "imported" in lit["implementation"] else "port",
                                  sample=f"{lit['batch']} random input vectors per process, {cores} processes"))
    try:
        import torch
        if torch.cuda.is_available():
            from tfhe_fbs_map_b200.backend import B200Backend
            from tfhe_fbs_map_b200.formats import read_lbf
            entries = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_mapped.json")))
            e = next(x for x in entries if x["circuit"] == "half_adder" and x["p"] == 15 and x["mapper"] == "search")
            env = read_lbf(e["lbf"])
            be = B200Backend(ps, device=0, seed=args.seed)
            rng = np.random.default_rng(1)
            B = 1 << 20
            iv = {nm: rng.integers(0, 2, B) for nm in e["input_names"]}
            env.eval_clear(iv, backend=be)
            t0 = time.time(); got = env.eval_clear(iv, backend=be); dt = time.time() - t0
            from oracle import cleartext
            want = cleartext.lut_eval(env, iv)
            ok = all(np.array_equal(np.asarray(got[k]), np.asarray(want[k])) for k in want)
            Be = 4096
            ive = {nm: v[:Be] for nm, v in iv.items()}
            env.eval(ive, fbs_size=15, backend=be)
            t0 = time.time(); gote = env.eval(ive, fbs_size=15, backend=be); dte = time.time() - t0
            oke = all(np.array_equal(np.asarray(gote[k]), np.asarray(want[k])[:Be]) for k in want)
            line["gpu"] = dict(cleartext_kernel_lookups_per_s=e["stats"]["nb_bootstrap"] * B / dt, cleartext_kernel_batch=B, cleartext_kernel_matches_oracle=ok,
                               encrypted_pbs_per_s=e["stats"]["nb_bootstrap"] * Be / dte, encrypted_batch=Be, encrypted_matches_cleartext=oke,
                               note="wall clock through the Python API incl. host<->device copies (and encrypt/decrypt for the encrypted run)", param_set=ps.name)
    except Exception as ex:
        line["gpu"] = f"unavailable: {ex}"
    print(json.dumps(line))


def cpu_baseline(ps, wl, seed, threads=0):
    """cpu_baseline object of the bench line: the tuned CPU arm on a bounded sample of the same workload, all host cores."""
    cores = threads or host_cores()
    ok = cpu_verify(ps, wl, seed, cores)
    st = cpu_sample_step(ps, wl, seed, cores)
    out = dict(value=st["value"], unit="PBS/s", cores=cores, kind="port", wall_s=st["wall_s"], sample=st["sample"],
               implementation=CPU_ARM_DESC, verified=f"{wl['cpu_sample']} whole program on the CPU arm: decrypted == cleartext: {ok}",
               evals_per_s_extrapolated=st["value"] / max(1, load_prog_boots(wl)))
    # the parity oracle (oracle/tfhe_ref.c, every modmul a 128-bit %): NOT a performance baseline, shown for continuity with round 1
    try:
        ref, _ = cpu_arm(ps, seed)
        from tfhe_fbs_map_b200 import levelize
        env = load_env(wl["cpu_sample"])
        prog = levelize(env, wl["p"])
        nb = min(cores, 8)
        bits = np.random.default_rng(3).integers(0, 2, (prog.n_inputs, nb)).astype(np.uint8)
        t0 = time.time()
        # two levels are enough for a rate
        sub = min(prog.n_boots, 4)
        cts_p = wl["p"]
        low = np.random.default_rng(4).integers(0, 2, (nb * sub, cts_p)).astype(np.uint8)
        tabs = np.concatenate([low, 1 - low], axis=1)
        cts = ref.encrypt(cts_p, np.zeros(nb * sub, np.int32), np.arange(nb * sub), 5)
        ref.pbs_batch(cts_p, cts, tabs, np.full(nb * sub, 2 * cts_p, np.uint8), None, want_acc=False, threads=cores)
        out["oracle_checker_pbs_per_s"] = nb * sub / (time.time() - t0)
    except Exception as ex:                                  # the oracle rate is informational only
        out["oracle_checker_pbs_per_s"] = f"unavailable: {ex}"
    try:
        out["reference_literal"] = reference_literal(cores)
    except Exception as ex:
        out["reference_literal"] = f"unavailable: {ex}"
    return out


def load_prog_boots(wl):
    return load_env(wl["lbf"]).stats()["nb_bootstrap"]


def run_reference(args, wl, ps, rank, world):
    """`--impl reference`: the CPU implementation of the path on the box's host cores, same workload / parameter set / metric;
    each step a bounded sample (cpu_sample_step).  Rank 0 alone runs and prints."""
    if rank != 0:
        return
    cores = host_cores()
    ok = cpu_verify(ps, wl, args.seed, cores)
    vals = []
    for step in range(args.warmup + args.steps):
        st = cpu_sample_step(ps, wl, args.seed, cores, enc_seed=100 + step)
        if step >= args.warmup:
            vals.append(st)
    v = sum(x["pbs"] for x in vals) / sum(x["wall_s"] for x in vals)
    env = load_env(wl["lbf"])
    line = dict(metric="PBS/sec", value=v, unit="PBS/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * statistics.mean(x["wall_s"] for x in vals),
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u64 (q = p1*p2, two 30-bit NTT primes; RNS u32x2 in the NTT)", data="synthetic",
                impl="reference",
                config=bench_config(args, wl, ps, env),
                evals_per_s=v / env.stats()["nb_bootstrap"],
                cpu_baseline=dict(value=v, unit="PBS/s", cores=cores, kind="port", sample=vals[-1]["sample"], implementation=CPU_ARM_DESC,
                                  verified=f"{wl['cpu_sample']} whole program on the CPU arm: decrypted == cleartext: {ok}"),
                e2e=dict(value=v, unit="PBS/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def bench_config(args, wl, ps, env):
    """`config` of the bench line.  Computed from the arguments only (nothing measured), so the GPU arm and the reference arm
    print the SAME object (the driver compares them); keys that describe one arm say so in their name."""
    from tfhe_fbs_map_b200 import levelize
    st = env.stats()
    prog = levelize(env, wl["p"], preserve_inputs=True)
    B = args.batch
    wires_gb = prog.n_slots * B * ps.ct_words * 8 / 1e9
    return dict(workload=args.workload, desc=wl["desc"], param_set=ps.name, n=ps.n, k=ps.k, N=ps.N, bsk_l=ps.bsk_l, ks_l=ps.ks_l, bsk_unroll=ps.bsk_unroll,
                fbs_size=wl["p"], pbs_per_instance=st["nb_bootstrap"], levels=prog.n_levels, p_fail_per_pbs=ps.p_fail(wl["p"], st["norm2_linprod"]),
                gpu_instances_per_gpu=B, gpu_sharding="instances, keys replicated, no collective",
                cpu_sample="first 32 levels of the same program x one instance per host thread per step",
                l2="GPU arm: wire buffer %.2f GB per GPU > 126 MB L2; BSK+KSK (%.0f MB) are re-streamed every level" % (wires_gb, (ps.bsk_bytes + ps.ksk_bytes) / 1e6))


def measure_nodes(be, wl_name, B, exchange, warmup, steps, rank, world, local, torch, dist, one_gpu_too=True):
    """BASELINE configs[3]: ONE circuit, the bootstraps of every level split across the ranks (strong scaling).  Every rank
    holds a full replica of the wire buffer; `exchange`: "fused" = the sample-extract epilogue stores each output ciphertext
    into every peer replica over NVLink and the levels are ordered by device-side epoch flags (no host sync between levels),
    "fused-host" = same stores, host barrier per level (round-1 behaviour), "nccl" = in-place all-gather per level.
    Returns the measurement dict (on every rank; times are max over ranks, CUDA events)."""
    from tfhe_fbs_map_b200 import levelize
    from tfhe_fbs_map_b200.dist import B200Engine, FusedB200Engine, run_node_sharded
    from oracle import cleartext
    wl = WORKLOADS[wl_name]
    env = load_env(wl["lbf"])
    prog = levelize(env, wl["p"], shard_pad=world, reuse_slots=False)
    cp = be.load(prog)
    rng = np.random.default_rng(77)                      # every rank evaluates the SAME instances
    bits = rng.integers(0, 2, (prog.n_inputs, B)).astype(np.uint8)
    want = cleartext.lut_eval(env, {nm: bits[i] for i, nm in enumerate(prog.input_names)})
    want_mat = np.array([np.asarray(want[nm]) for nm in prog.output_names], dtype=np.uint8)

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

    def timed(eng, w, r, n_warm, n_steps):
        for _ in range(n_warm):
            run_node_sharded(eng, prog, dist, w, r)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        words = 0
        for _ in range(n_steps):
            words += run_node_sharded(eng, prog, dist, w, r)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / n_steps, words // max(1, n_steps)

    res = dict(workload=wl_name, desc=wl["desc"], fbs_size=wl["p"], instances=B, pbs_per_instance=prog.n_boots, levels=prog.n_levels,
               level_width_median=int(np.median(prog.level_widths)), level_width_max=int(max(prog.level_widths)), n_gpus=world, scaling="strong",
               p_fail_per_pbs=be.params.p_fail(wl["p"], env.stats()["norm2_linprod"]))
    ms1 = None
    if world > 1 and one_gpu_too:
        # the same program on ONE GPU (every rank runs it on its own GPU at the same time; max over ranks): the strong-scaling reference point
        solo = B200Engine(be, cp, B, torch)
        solo.encrypt(bits, enc_seed=5)
        ms1, _ = timed(solo, 1, 0, 1 if B > 8 else warmup, 1 if B > 8 else steps)
        res["one_gpu_mismatches"] = int((solo.decrypt() != want_mat).sum())
        del solo
        torch.cuda.empty_cache()
    fused = exchange.startswith("fused") and world > 1
    eng = FusedB200Engine(be, cp, B, torch, dist, world, rank, handoff="host" if exchange == "fused-host" else "device") if fused else B200Engine(be, cp, B, torch)
    try:
        eng.encrypt(bits, enc_seed=5)
        barrier()
        ms, words = timed(eng, world, rank, warmup, steps)
        mism = int((eng.decrypt() != want_mat).sum())
    finally:
        if fused:
            eng.close()
    n_pbs = prog.n_boots * B
    res.update(value=n_pbs / (ms * 1e-3), unit="PBS/s", ms_per_step=ms, evals_per_s=B / (ms * 1e-3), ms_per_level=ms / prog.n_levels, mismatches=mism,
               steps=steps, warmup=warmup, exchange_bytes_per_step=words * 8,
               exchange=("none (one GPU)" if world == 1 else
                         {"fused": "fused: blind-rotation epilogue stores output LWE ciphertexts into every peer replica over NVLink (peer-mapped memory) + device-side per-level epoch flags, no host sync between levels",
                          "fused-host": "fused peer stores + host barrier per level", "nccl": "NCCL in-place all-gather of output LWE ciphertexts per level"}[exchange]))
    if ms1 is not None:
        res.update(one_gpu_ms_per_step=ms1, one_gpu_value=n_pbs / (ms1 * 1e-3), speedup_vs_1gpu=ms1 / ms, efficiency_vs_1gpu=ms1 / (ms * world))
    return res


def run_nodes(args, wl, ps, be, rank, world, local, torch, dist):
    """`--shard nodes`: the node-sharded measurement alone, as its own bench line (strong scaling)."""
    sampler = ClockSampler(local)
    sampler.start()
    r = measure_nodes(be, args.workload, args.batch, args.exchange, args.warmup, args.steps, rank, world, local, torch, dist)
    sampler.stop_flag.set(); sampler.join(timeout=2)
    if rank == 0:
        print(json.dumps(dict(
            metric="PBS/sec", value=r["value"], unit="PBS/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="strong", vs_baseline=None, dtype="u64 (q = p1*p2, two 30-bit NTT primes; RNS u32x2 in the NTT)",
            data="synthetic", config=dict(workload=args.workload, desc=wl["desc"], param_set=ps.name, fbs_size=wl["p"], instances=args.batch,
                                          pbs_per_instance=r["pbs_per_instance"], levels=r["levels"], level_width_median=r["level_width_median"], sharding="nodes: " + r["exchange"]),
            evals_per_s=r["evals_per_s"], mismatches=r["mismatches"], nodes=r, gpu_launches=3 * r["levels"] * args.steps, clocks=sampler.summary())))
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
    ap.add_argument("--exchange", default="fused", choices=["fused", "fused-host", "nccl"],
                    help="node sharding: 'fused' = sample-extract epilogue stores into all peers' replicas (one kernel does "
                         "compute + exchange) + device-side level hand-off, 'fused-host' = same stores, host barrier per level, "
                         "'nccl' = separate in-place all-gather per level")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-nodes", action="store_true", help="skip the node-sharded 64x64 multiplier object ('nodes') of the line")
    ap.add_argument("--nodes-batches", default="1,64", help="instance batches of the node-sharded object")
    ap.add_argument("--config1", action="store_true",
                    help="BASELINE configs[0] as its own line: smallest generated circuit (half_adder), --fbs_size 15 --mapper search, "
                         "cleartext LutExecEnv.eval on random inputs on the host cores; next to it the same program on the GPU (cleartext "
                         "look-up kernel and encrypted)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from tfhe_fbs_map_b200 import params, levelize
    ps = params.get(args.param_set)
    wl = WORKLOADS[args.workload]
    if args.batch is None:
        args.batch = wl.get("batch", 296)

    if args.config1:
        if rank == 0:
            run_config1(args, ps)
        return
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

    # ---- BASELINE configs[3] on the same launch: one 64x64 multiplier, node-sharded levels (strong scaling) -----------------------
    nodes = None
    if not args.no_nodes:
        del wires
        torch.cuda.empty_cache()
        nodes = {}
        for nb in [int(x) for x in args.nodes_batches.split(",") if x]:
            nodes[f"batch{nb}"] = measure_nodes(be, "mult64_p17", nb, args.exchange, 1, 2 if nb > 8 else 3, rank, world, local, torch, dist)
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
        # integer roofline of the blind rotation alone: its multiplies over ITS time (the key switch runs on the tensor cores)
        from dataclasses import asdict
        from tfhe_fbs_map_b200.params import ParamSet
        dcl = asdict(ps); dcl.update(bsk_unroll=1, name="classic")
        canon_modmul = ParamSet(**dcl).modmul_per_pbs()                 # one key bit per step
        br_pbs_per_s = n_pbs_step * args.steps / (st.ms_blind_rotate * 1e-3)
        mul32_exec = 4 * ps.modmul_per_pbs() * br_pbs_per_s
        mul32_canon = 4 * canon_modmul * br_pbs_per_s
        value = world * n_pbs_step * args.steps / (ms_res * 1e-3)
        cfg = bench_config(args, wl, ps, env)
        line = dict(
            metric="PBS/sec", value=value, unit="PBS/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_res / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="u64 (q = p1*p2, two 30-bit NTT primes; RNS u32x2 in the NTT)", data="synthetic",
            config=cfg,
            evals_per_s=world * B * args.steps / (ms_res * 1e-3), mismatches=mism_res,
            phase_ms_per_step=dict(lincomb=st.ms_lincomb / args.steps, keyswitch=st.ms_keyswitch / args.steps, blind_rotate=st.ms_blind_rotate / args.steps),
            roofline=dict(bound="hbm", achieved=achieved, peak=peaks.get("hbm_gbs"), unit="GB/s", frac=achieved / peaks.get("hbm_gbs"),
                          traffic=TRAFFIC_GB.get(ps.bsk_unroll), traffic_source=TRAFFIC_SOURCE,
                          kernel="k_blind_rotate2" if ps.bsk_unroll > 1 else "k_blind_rotate", launch_ms=br_ms_avg, pbs_per_launch=pbs_per_launch, peak_source=peak_kind,
                          note="kernel is integer-issue bound by design (accumulator on chip, keys streamed from L2 by TMA); see roofline_int"),
            roofline_int=dict(bound="int32-multiply", kernel="blind rotation only (key switch = int8 tensor-core GEMM, not counted)",
                              achieved=mul32_exec / 1e12, peak=int_peak / 1e12, unit="T mul32/s", frac=mul32_exec / int_peak,
                              frac_canonical=mul32_canon / int_peak, achieved_canonical=mul32_canon / 1e12,
                              modmul_per_pbs_executed=ps.modmul_per_pbs(), modmul_per_pbs_canonical=canon_modmul, blind_rotate_pbs_per_s=br_pbs_per_s,
                              note="frac: multiplies the kernel executes (4 mul32 per modular multiply of the key-unrolled count) / measured peak; "
                                   "frac_canonical: SURVEY 8(d)'s classic one-bit-per-step count at the same PBS rate (the difference is the algorithmic saving of key unrolling)",
                              peak_source="measured in this run (fbs_measure_int_peak: mad.wide.u32 chains)"),
            e2e=e2e, gpu_launches=int(st.n_launches), clocks=sampler.summary())
        if nodes is not None:
            line["nodes"] = nodes
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(ps, wl, args.seed)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
