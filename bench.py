#!/usr/bin/env python
"""bench.py -- population env-steps/s of the signal-gated market-making rollout on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full population rollout (BASELINE.json configs[1]: P = 4096 individuals per GPU x
T = 14 400 bars = 60 synthetic 510300-shaped days, H = 32, no fee, phi = 1e-4): every individual's
policy MLP + quantisation + FPT env step for every bar, fitness and trade count out.  At N > 1 the
population is sharded by contiguous global index (weak scaling: 4096 individuals per GPU) and each
step ends with the all-gather of the fitness / trade slices that a sharded GA generation performs.

  value   whole-job env-steps/s, genomes and bars resident in HBM, CUDA-event timed per step
  e2e     the same through the host-buffer C-ABI entry (sgmm_rollout_population_host): pinned host
          genomes H2D + kernel + fitness/trades D2H inside the timed region
  roofline  FP32 CUDA-core roofline (compute-bound; 2368 algorithmic FLOP / env-step)
            against the FFMA peak measured live on this device; hbm sub-object for the bar/genome stream
  tensor_core_h32   the same workload through the tensor-core rollout (sgmm_tc32.cu, SGMM_PRECISION_BF16):
            device-timed value, e2e, tensor roofline (algorithmic and executed TFLOP/s), GA generations/s
  cpu_baseline / --impl reference   the CPU oracle port (oracle/sgmm_oracle.c, pthreads over all host
            cores) on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

P_PER_GPU = 4096
N_DAYS = 60
HIDDEN = 32
PHI, TICK, FEE = 1e-4, 0.001, 0.0
FLOP_PER_STEP = 2368.0               # 2*(3H + H^2 + 2H) at H = 32
METRIC = "population_env_steps_per_sec"
UNIT = "env-steps/s"


def workload_config(n_gpus):
    return {"workload": "configs[1]: DRL agent, population 4096 per GPU, 60 synthetic 510300-shaped days "
                        "(T=14400 bars), H=32 MLP 3-32-32-2, no fee, phi=1e-4, tick=0.001",
            "population_per_gpu": P_PER_GPU, "population_total": P_PER_GPU * n_gpus, "bars": N_DAYS * 240,
            "hidden": HIDDEN, "fee_rate": FEE, "phi": PHI, "sharding": f"population x{n_gpus} (weak)",
            "l2": "256 MiB device memset between timed steps (flushes the 126 MB L2); per-step CUDA events "
                  "exclude the flush"}


def make_inputs(p_total):
    from sgmm_b200 import synthetic
    bundle = synthetic.synthetic_bundle(N_DAYS)
    stats = synthetic.train_stats_of(bundle)
    master, genomes = synthetic.policy_like_genomes(p_total, HIDDEN, seed=0, sigma=0.05, out_scale=1.0)
    return bundle, stats, master, genomes


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows[-3:]]
        sm, smax, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_rate(bundle, stats, genomes, target_seconds, threads=0):
    """env-steps/s of the CPU oracle port on a bounded sample (all host threads)."""
    from oracle import oracle
    z1, z2 = oracle.normalise(bundle, stats)
    bz = (z1, z2) + tuple(bundle[2:])
    cores = threads or oracle.max_threads()
    T = len(z1)
    probe_n = max(cores, 8)
    t0 = time.perf_counter()
    oracle.rollout_population(bz, PHI, TICK, FEE, genomes=genomes[:probe_n], nthreads=cores)
    probe = time.perf_counter() - t0
    n = int(max(probe_n, min(genomes.shape[0], probe_n * target_seconds / max(probe, 1e-3))))
    n = max(cores, (n // cores) * cores)
    n = min(n, genomes.shape[0])
    t0 = time.perf_counter()
    fit, trd = oracle.rollout_population(bz, PHI, TICK, FEE, genomes=genomes[:n], nthreads=cores)
    dt = time.perf_counter() - t0
    return n * T / dt, cores, n, T, dt, fit


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Python
    and does not travel to the GPU box, so this is the C oracle port (kind 'port') on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bundle, stats, master, genomes = make_inputs(P_PER_GPU)
    rates = []
    sample = None
    for i in range(args.warmup + args.steps):
        rate, cores, n, T, dt, _ = cpu_port_rate(bundle, stats, genomes, target_seconds=max(2.0, 20.0 / max(1, args.steps)))
        sample = (cores, n, T)
        if i >= args.warmup:
            rates.append((n * T, dt))
    steps_done = sum(r[0] for r in rates)
    total = sum(r[1] for r in rates)
    value = steps_done / total
    cores, n, T = sample
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * total / max(1, len(rates)), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32 policy / f64 env", "data": "synthetic",
           "config": workload_config(args.gpus),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"{n} individuals x {T} bars per step (C oracle port of the reference's "
                                      f"evaluate_individual, pthreads over {cores} host threads)"},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import sgmm_b200
    from sgmm_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    p_total = P_PER_GPU * n_gpus
    bundle, stats, master, genomes_all = make_inputs(p_total)
    first = rank * P_PER_GPU
    genomes = genomes_all[first:first + P_PER_GPU]
    T = len(bundle[0])
    bun = sgmm_b200.Bundle.from_arrays(bundle, stats, TICK, device=local)
    g_dev = torch.from_numpy(genomes).to(dev)
    g_pin = torch.from_numpy(genomes).pin_memory()
    fit_pin = torch.empty(P_PER_GPU, dtype=torch.float64).pin_memory()
    trd_pin = torch.empty(P_PER_GPU, dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fit_all = torch.empty(p_total, dtype=torch.float64, device=dev) if world > 1 else None
    trd_all = torch.empty(p_total, dtype=torch.int32, device=dev) if world > 1 else None

    def step_device():
        f, t = sgmm_b200.rollout_population(bun, g_dev, phi=PHI, fee_rate=FEE)
        if world > 1:
            dist.all_gather_into_tensor(fit_all, f)
            dist.all_gather_into_tensor(trd_all, t)
        return f, t

    import ctypes as C
    L = _lib.lib()
    mm = _lib.Population(HIDDEN, 0, P_PER_GPU, g_pin.data_ptr(), None, 0.0, 0.0, 0, 0, 0)
    prm = _lib.RolloutParams(PHI, FEE, 0, 0, 0, 0)

    def step_e2e():
        st = C.c_void_p(torch.cuda.current_stream(local).cuda_stream)
        _lib.check(L.sgmm_rollout_population_host(bun.handle, C.byref(mm), None, C.byref(prm), fit_pin.data_ptr(),
                                                  trd_pin.data_ptr(), st))
        return float(fit_pin[0])

    # SGMM_PRECISION_TF32 / _F16 / _BF16: the tensor-core rollout (sgmm_tc32.cu)
    TC_MODES = (("tf32", 2), ("f16", 3), ("bf16", 1))
    prm_tc = {name: _lib.RolloutParams(PHI, FEE, code, 0, 0, 0) for name, code in TC_MODES}

    def step_device_tc(mode):
        f, t = sgmm_b200.rollout_population(bun, g_dev, phi=PHI, fee_rate=FEE, precision=mode)
        if world > 1:
            dist.all_gather_into_tensor(fit_all, f)
            dist.all_gather_into_tensor(trd_all, t)
        return f, t

    def step_e2e_tc(mode):
        st = C.c_void_p(torch.cuda.current_stream(local).cuda_stream)
        _lib.check(L.sgmm_rollout_population_host(bun.handle, C.byref(mm), None, C.byref(prm_tc[mode]), fit_pin.data_ptr(),
                                                  trd_pin.data_ptr(), st))
        return float(fit_pin[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -------------------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_device()
        step_e2e()
        for mode, _ in TC_MODES:
            step_device_tc(mode)
            step_e2e_tc(mode)
    barrier()
    fp32_peak = sgmm_b200.measure_fp32_peak(local)

    # ---- device-timed steps -------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    f_last = None
    for (e0, e1) in evs:
        flush.fill_(1)                                  # L2 flush, outside the step's event pair
        e0.record()
        f_last, t_last = step_device()
        e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    step_ms = [e0.elapsed_time(e1) for (e0, e1) in evs]
    dev_ms = sum(step_ms)
    # ---- end-to-end steps (host buffers, copies inside the timed region) ----------------------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    # ---- the same two measurements for the tensor-core rollout (stated tolerance), per precision ----
    tc_raw = {}
    tc_extra = 0.0
    for mode, _ in TC_MODES:
        evs_tc = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        f_tc = None
        for (e0, e1) in evs_tc:
            flush.fill_(1)
            e0.record()
            f_tc, _ = step_device_tc(mode)
            e1.record()
        barrier()
        ms = [e0.elapsed_time(e1) for (e0, e1) in evs_tc]
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e_tc(mode)
        barrier()
        e2e_t = time.perf_counter() - t0
        tc_raw[mode] = [ms, e2e_t, float(f_tc.sum().item())]
        tc_extra += e2e_t + 1e-3 * sum(ms)
    clocks = sampler.stop(t_wall0, t_wall1 + e2e_s + tc_extra)

    vals = [dev_ms, e2e_s] + [x for mode, _ in TC_MODES for x in (sum(tc_raw[mode][0]), tc_raw[mode][1])]
    tt = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, e2e_s = tt[0].item(), tt[1].item()
    for k, (mode, _) in enumerate(TC_MODES):
        tc_raw[mode] += [tt[2 + 2 * k].item(), tt[3 + 2 * k].item()]          # max over ranks: device ms, e2e s

    # ---- GA generations/s (secondary metric; CUDA-graph replay of ask+rollout+tell+validate+select) ----
    ga_rate, ga_small, ga_rate_tc = None, None, None
    if rank == 0 and world == 1:
        from sgmm_b200 import synthetic
        from sgmm_b200.engine import DeviceGA

        def ga_rate_of(train_b, val_b, pop, ngen, precision=None):
            ga = DeviceGA(master, None, pop_size=pop, sigma=0.05, phi=PHI, fee_rate=FEE, use_arl=False, seed=0,
                          max_generations=2 * ngen + 4, device=local, precision=precision)
            ga.generation(train_b, val_b)
            torch.cuda.synchronize()
            graph = ga.capture(train_b, val_b)
            graph.replay()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(ngen):
                graph.replay()
            torch.cuda.synchronize()
            r = ngen / (time.perf_counter() - t0)
            ga.close()
            return r

        val = sgmm_b200.Bundle.from_arrays(synthetic.synthetic_bundle(12, first_day=N_DAYS), stats, TICK, device=local)
        ga_rate = ga_rate_of(bun, val, P_PER_GPU, 5)
        ga_rate_tc = {mode: ga_rate_of(bun, val, P_PER_GPU, 5, precision=mode) for mode, _ in TC_MODES}
        # the reference's own scale (BASELINE configs[0]): population 50, one training day, one validation day
        d1 = synthetic.synthetic_bundle(1, first_day=200)
        st1 = synthetic.train_stats_of(d1)
        t1 = sgmm_b200.Bundle.from_arrays(d1, st1, TICK, device=local)
        v1 = sgmm_b200.Bundle.from_arrays(synthetic.synthetic_bundle(1, first_day=201), st1, TICK, device=local)
        ga_small = ga_rate_of(t1, v1, 50, 200)

    # ---- secondary: the H=256 tensor-core path (BASELINE config 4 shape, shortened), rank 0 only ----
    h256 = None
    if rank == 0 and world == 1:
        try:
            from sgmm_b200 import synthetic
            m256, _ = synthetic.policy_like_genomes(1, hidden=256, seed=0)
            m256 = torch.from_numpy(m256).to(dev)
            p256 = 4 * torch.cuda.get_device_properties(local).multi_processor_count
            def run256():
                return sgmm_b200.rollout_seeded(bun, m256, count=p256, sigma=0.05, seed=1, generation=0, phi=PHI,
                                                fee_rate=3e-4, hidden=256)
            run256(); torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); run256(); a1.record(); torch.cuda.synchronize()
            ms256 = a0.elapsed_time(a1)
            st256 = p256 * T
            h256 = {"kernel": "spec256_kernel (tcgen05, bf16 x bf16 -> fp32 TMEM, 5-inventory speculation)",
                    "population": p256, "bars": T, "fee_rate": 3e-4, "ms": ms256, "env_steps_per_sec": st256 / ms256 * 1e3,
                    "algorithmic_tflops": st256 * 133632 / ms256 / 1e9,
                    "executed_hidden_tflops": st256 * (128 / 25) * 131072 / ms256 / 1e9}
        except Exception as e:          # secondary metric must never take the headline down
            h256 = {"error": str(e)}

    if rank == 0:
        K = args.steps
        steps_per_step = p_total * T
        value = steps_per_step * K / (dev_ms * 1e-3)
        e2e = steps_per_step * K / e2e_s
        kernel_s = (dev_ms / K) * 1e-3                       # the rollout kernel is the whole device step
        achieved_tflops = FLOP_PER_STEP * P_PER_GPU * T / kernel_s / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        G = HIDDEN * HIDDEN + 7 * HIDDEN + 2
        alg_bytes = P_PER_GPU * G * 4 + T * 48 + P_PER_GPU * 12       # genomes once + bars once + outputs
        cpu_rate, cores, n, Tc, dt, _ = cpu_port_rate(bundle, stats, genomes, target_seconds=12.0)
        # tensor-core rollout: algorithmic FLOPs as above; EXECUTED tensor FLOPs per unit of two 128-row tiles
        # (= 50 env-steps): L1 128x64x16, L2 2 x 128x32x48, L3 2 x 128x16x48, x2 FLOP per MAC
        tc_exec_flop_per_step = {"bf16": 2.0 * (128 * 64 * 16 + 2 * 128 * 32 * 48 + 2 * 128 * 16 * 48) / 50.0,
                                 "tf32": 2.0 * (128 * 64 * 16 + 2 * 128 * 32 * 40 + 2 * 128 * 16 * 40) / 50.0}
        tc_exec_flop_per_step["f16"] = tc_exec_flop_per_step["bf16"]
        tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        tol = {"bf16": "policy outputs within 0.12 tick of the fp32 oracle; 3 of 1920 offsets of the golden ARL audit set flip, all at "
                       "near-ties of the fp32 result",
               "f16": "f16 operands, f16 accumulators for layers 1-2 (fp32 for the output layer): policy outputs within 0.03 tick of "
                      "the fp32 oracle (measured 0.0105); 0 of 1920 offsets of the golden ARL audit set flip",
               "tf32": "policy outputs within 0.03 tick of the fp32 oracle (measured 0.011); 0 of 1920 offsets of the golden ARL "
                       "audit set flip: the reference's shipped backtest is reproduced exactly (actions, trades, fitness)"}
        tc = {"kernel": "tc32_kernel (sgmm_tc32.cu): all three policy layers on tcgen05 (fp32 accumulate in TMEM, A operands chained "
                        "through tensor memory), 5-inventory speculation, grouped GEMM over 14 individuals per CTA; "
                        "tests/test_gpu_tc32.py states the tolerances; the env step given the offsets is bit-exact"}
        for mode, _ in TC_MODES:
            ms_list, _, csum, dev_ms_tc, e2e_tc_s = tc_raw[mode]
            tc_value = steps_per_step * K / (dev_ms_tc * 1e-3)
            tc_kernel_s = (dev_ms_tc / K) * 1e-3
            tc_alg_tflops = FLOP_PER_STEP * P_PER_GPU * T / tc_kernel_s / 1e12
            tc[mode] = {"value": tc_value, "unit": UNIT, "ms_per_step": dev_ms_tc / K, "per_step_ms": ms_list,
                        "speedup_vs_exact_kernel": tc_value / value, "tolerance": tol[mode],
                        "e2e": {"value": steps_per_step * K / e2e_tc_s, "unit": UNIT, "ms_per_step": 1e3 * e2e_tc_s / K,
                                "h2d_bytes_per_step": P_PER_GPU * G * 4 * n_gpus, "d2h_bytes_per_step": P_PER_GPU * 12 * n_gpus},
                        "roofline": {"bound": "tensor", "achieved": tc_alg_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                                     "frac": tc_alg_tflops / tensor_peak, "traffic": None,
                                     "executed_tflops": tc_exec_flop_per_step[mode] * P_PER_GPU * T / tc_kernel_s / 1e12,
                                     "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (tf32 nominal peak is half)"
                                                     if "bf16_tflops_sustained" in peaks else "fallback 1400"),
                                     "note": "algorithmic = 2368 FLOP/env-step; executed counts the 5-inventory speculation and the "
                                             "K / N padding; the kernel is bound by the CUDA-core side (TMEM <-> register "
                                             "conversion, speculative fp64 env step), see profiles/r1_tc32_ncu_full.txt"},
                        "ga_generations_per_sec": (ga_rate_tc or {}).get(mode), "checksum": csum}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 policy (SGMM-F32 order, FFMA2) / int32 fills / f64 env", "data": "synthetic",
            "config": workload_config(n_gpus),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": P_PER_GPU * G * 4 * n_gpus,
                    "d2h_bytes_per_step": P_PER_GPU * 12 * n_gpus, "ms_per_step": 1e3 * e2e_s / K,
                    "entry": "sgmm_rollout_population_host (pinned host genomes in, fitness/trades out; bundle resident)"},
            "gpu_launches": (1 + len(TC_MODES)) * K,    # K exact-kernel + K tensor-core rollouts per mode in the device-timed regions
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved_tflops / fp32_peak if fp32_peak else None,
                         "traffic": 20922880,
                         "kernel": "rollout_kernel_h32<4,false,false>",
                         "algorithmic_flop_per_env_step": FLOP_PER_STEP,
                         "peak_source": "FFMA/FFMA2 peak measured live on this device by sgmm_measure_fp32_peak "
                                        "(MEASURED_PEAKS.json has no fp32 figure; theoretical 74.4 TFLOP/s at 1965 MHz)",
                         "note": "compute-bound on the FP32 CUDA-core pipe (SURVEY.md 8d); bound is neither hbm nor tensor. "
                                 "traffic = dram__bytes_read+write of one launch from profiles/r1_rollout_accwarp_ncu_full.txt "
                                 "(P=4096: 20.9 MB read, i.e. the genomes once; bars stay in L2; nothing written but the results)",
                         "hbm": {"achieved": alg_bytes / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / kernel_s / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes,
                                 "peak_source": hbm_src}},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / kernel_s / 1e9 / hbm_peak, "traffic": 20922880,
                             "note": "reported for completeness: the bar/genome stream is <0.1% of HBM peak by construction"},
            "roofline_tensor_h256": (None if not h256 or "error" in h256 else
                                     {"bound": "tensor", "achieved": h256["algorithmic_tflops"],
                                      "peak": peaks.get("bf16_tflops_sustained", 1400.0), "unit": "TFLOP/s",
                                      "frac": h256["algorithmic_tflops"] / peaks.get("bf16_tflops_sustained", 1400.0),
                                      "executed_tflops": h256["executed_hidden_tflops"],
                                      "note": "secondary kernel spec256_kernel (H=256, BASELINE config 4 shape); algorithmic = "
                                              "133632 FLOP/env-step, executed = 5.12x hidden-layer FLOPs (5-inventory speculation)"}),
            "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} individuals x {Tc} bars in {dt:.1f} s (C oracle port, pthreads)"},
            "per_step_ms": step_ms,
            "ga_generations_per_sec": ga_rate,
            "ga_generations_per_sec_config0_pop50_1day": ga_small,
            "h256_tensor_core": h256,
            "tensor_core_h32": tc,
            "checksum": float(f_last.sum().item()) if f_last is not None else None,
        }
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line of the contract goes to the real stdout; everything libraries print while the
    benchmark runs (NCCL's version banner, for instance) was redirected to stderr."""
    line = json.dumps(obj) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(line); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line.encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the duration of the run
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
