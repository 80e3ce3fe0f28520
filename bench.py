#!/usr/bin/env python
"""bench.py — BASELINE.json metric: env-steps/s of PPO LunarLander-v3 at 4096 vectorised envs per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one full PPO iteration of the reference's hot path over one batch of synthetic (self-generated)
experience: a T=128-step lockstep rollout of 4096 LunarLander-v3 copies (policy forward -> categorical
sample -> env step with auto-reset), GAE, advantage normalisation and the whole update (10 epochs x 32
minibatches of 16,384: forward, fused clipped-surrogate loss, backward, global-norm clip, Adam) — nothing
skipped.  value = env-steps / second over K timed steps, CUDA events on the launching stream, max over ranks
(weak scaling: 4096 envs per GPU).  `e2e` is the same quantity through the public trainer API
(PPOTrainer.train_iteration): host wall clock including the LR write (H2D) and the metrics / episode-
statistics read-back (D2H).  Parts of the working set (rollout 26 MB, weights < 1 MB) are smaller than the
126 MB L2, so every timed step is preceded by an L2 flush (a 256 MB buffer write) outside the timed events;
within a step the kernels see the cache state the real workload produces.

Extra objects: `roofline` for the dominant kernel (the dense-layer GEMM), `cpu_baseline` (the oracle port of
the reference loop on the host cores), `clocks`, `gpu_launches`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_ENVS, T_STEPS, N_MB, N_EPOCHS = 4096, 128, 32, 10
WORKLOAD = "PPO LunarLander-v3, 4096 vectorised envs/GPU, T=128, 10 epochs x 32 minibatches of 16384 (BASELINE configs[1])"


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.samples, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- reference arm
def _cpu_arm_rows(n_steps: int, warmup: int, with_single_rows: bool):
    """The reference's CPU implementation of the path on this box's host cores (oracle/cpu_arm.py): the unmodified reference
    file on oracle/gymnasium_shim where /root/reference exists (kind "reference"), its hand port elsewhere (kind "port").
    P pinned single-thread workers created once; every step = one PPO iteration (2048 env steps + full update) per worker."""
    from oracle import cpu_arm
    kind = cpu_arm.pick_kind()
    host = cpu_arm.host_info()
    cores = host["worker_cores"]
    pool = cpu_arm.WorkerPool(kind, cores)
    try:
        rounds = [pool.run_step() for _ in range(warmup + n_steps)]
    finally:
        pool.close()
    timed = rounds[warmup:]
    value = sum(r["env_steps"] for r in timed) / sum(r["seconds"] for r in timed)
    rows = {f"{len(cores)}proc_pinned_1thread_each": round(value, 1)}
    if with_single_rows:
        rows.update(cpu_arm.single_process_rows(kind))
    return {"value": value, "unit": "env-steps/s", "cores": len(cores), "kind": kind, "sample": cpu_arm.sample_text(kind, len(cores)),
            "rows": rows, "per_core": round(value / len(cores), 1),
            "host": {k: host[k] for k in ("nproc", "os_cpu_count", "physical_cores_in_mask", "cgroup_quota_cores", "model")},
            "per_step_values": [round(r["value"], 1) for r in timed]}, timed


def run_reference(args, rank, world):
    if rank != 0:
        return
    t0 = time.perf_counter()
    cpu, timed = _cpu_arm_rows(args.steps, args.warmup, with_single_rows=True)
    v = cpu["value"]
    ms = 1000.0 * sum(r["seconds"] for r in timed) / len(timed)
    out = {"metric": "env_steps_per_sec", "value": v, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": {"workload": WORKLOAD, "note": "reference path is single-env: each step = every pinned worker runs one 2048-step "
                      "rollout + full update (10 epochs x 32 minibatches of 64), the reference's own sizes"},
           "cpu_baseline": cpu,
           "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def gemm_roofline(torch, ops, peaks, peaks_src):
    """Dominant kernel: gemm_kernel<128,128,8,8> on the widest layer of the update (M=16384, N=512, K=256).
    Timed with CUDA events on the launching stream over operand sets that together exceed L2."""
    M, N, K = N_ENVS * T_STEPS // N_MB, 512, 256
    sets = 6  # 6 x (16 + 32 + 0.5) MB = 291 MB > 126 MB L2
    xs = [torch.randn(M, K, device="cuda") for _ in range(sets)]
    ys = [torch.empty(M, N, device="cuda") for _ in range(sets)]
    # the weights live in a flat parameter buffer with pre-split tf32 images registered, exactly as in the trainer: the launch
    # takes the shipping path (TMA-fed warp-specialised 2-CTA kernel).  The pass over the operand sets is recorded in a CUDA
    # graph like the trainer's epoch graph (the ws path encodes a TMA descriptor per launch on the host).
    flat = torch.zeros(N * K + 8, device="cuda")
    flat[4:4 + N * K] = (torch.randn(N, K, device="cuda") / 16).reshape(-1)
    w, b = flat[4:4 + N * K].view(N, K), torch.zeros(N, device="cuda")
    images = ops.weight_images_register(flat, [(4, N, K)])

    def one_pass():
        for i in range(sets):
            ops.linear_forward(xs[i], w, b, 1, out=ys[i])

    one_pass()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        one_pass()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        one_pass()
    g.replay()
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    dur_ms = e0.elapsed_time(e1) / (reps * sets)
    ops.weight_images_unregister(flat)
    del images
    flops = 2.0 * M * N * K
    achieved = flops / (dur_ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    traffic = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("gemm_fwd_dram_bytes_per_launch")
        except Exception:
            traffic = None
    from gymrl_b200 import _ffi
    tc = _ffi.load().gymrl_get_gemm_mode() == 1
    kernel = ("gemm3x_ws_kernel<2, 256> (tcgen05.mma.cta_group::2 kind::tf32, 3 MMAs per fp32 product; B = weights by TMA from pre-split images, "
              "A register-split; persistent, 2 TMEM accumulators, overlapped epilogue; M=16384 N=512 K=256, +bias+tanh)"
              if tc else "gemm_kernel<128,128,8,8,kmajor,kmajor> (fp32 FFMA, M=16384 N=512 K=256, +bias+tanh)")
    return {"bound": "tensor", "kernel": kernel,
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "launch_ms": dur_ms, "flops_per_launch": flops, "peak_source": peaks_src + ": dense bf16 cuBLAS, sustained",
            "tensor_flops_executed_per_launch": 3 * flops if tc else 0.0,
            "frac_of_3xtf32_ceiling": (achieved / (peak / 6.0)) if tc else None,
            "note": "algorithmic flops = 2MNK of the fp32 GEMM. The kernel keeps the reference's fp32 accuracy by running three "
                    "tf32 MMAs per product (tf32 dense peak = bf16/2), so its ceiling is peak/6; `frac` is still quoted against "
                    "the measured bf16 peak as the rules ask."}


def run_ours(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: gymrl_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from gymrl_b200 import _ffi, ops
    from gymrl_b200.algorithms import ppo_lunarlander as P

    cfg = P.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.num_epochs = N_ENVS, T_STEPS, N_MB, N_EPOCHS
    cfg.seed, cfg.max_train_steps = 0, 10 ** 12
    torch.manual_seed(0)
    tr = P.PPOTrainer(cfg)
    flush = torch.empty(64 * 1024 * 1024, device="cuda", dtype=torch.float32)  # 256 MB > L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed region: K steps, inputs (env state, parameters) resident in HBM ----
    for _ in range(args.warmup):
        tr.collect_rollout(); tr.update(None, read_metrics=False)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = tr.total_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the timed events)
        ev[k][0].record()
        tr.collect_rollout()
        tr.update(None, read_metrics=False)
        ev[k][1].record()
    barrier()
    launches = tr.total_launches() - l0   # our kernels only (torch's flush memset is not counted)
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    steps_total = N_ENVS * T_STEPS * args.steps * world
    value = steps_total / (dev_ms * 1e-3)

    # ---- e2e: the public API, host wall clock, LR H2D + metrics/episode-stats D2H every step ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tr.train_iteration()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = steps_total / float(t.item())
    h2d = 8                                  # float64 learning rate
    d2h = 8 * 4 + (4 + 4) * 1024 + 8         # metrics[8] + episode ring (returns, lengths) + its counter

    if rank != 0:
        return
    peaks, peaks_src = measured_peaks()
    roof = gemm_roofline(torch, ops, peaks, peaks_src)
    # HBM view of the whole step, from SURVEY §8(d): 550 algorithmic bytes per env-step for C2
    hbm_achieved = 550.0 * value / 1e9
    roof["path_hbm"] = {"algorithmic_bytes_per_env_step": 550, "achieved_GBps": hbm_achieved, "peak_GBps": peaks["hbm_gbs"],
                        "frac": hbm_achieved / peaks["hbm_gbs"],
                        "note": "the path is compute/latency bound (12.4 MFLOP per env-step), not HBM bound"}
    cpu = None
    if world == 1:
        cpu, _ = _cpu_arm_rows(2, 1, with_single_rows=False)   # bench's cpu_baseline leg: the oracle is the thing timed here
    out = {"metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "envs_per_gpu": N_ENVS, "rollout_steps": T_STEPS, "epochs": N_EPOCHS,
                      "minibatches_per_epoch": N_MB, "minibatch": N_ENVS * T_STEPS // N_MB, "params": 200965,
                      "parallelism": f"dp{world} (env shards, 1 gradient sum / optimizer step: " + ("one-shot NVLink peer-memory reduction fused with the clip norm, csrc/comm.cu)" if getattr(tr, "comm", None) is not None else ("ncclAllReduce)" if world > 1 else "none on 1 GPU)")),
                      "l2": "flushed (256 MB write) before every timed step", "cuda_graphs": True},
           "e2e": {"value": e2e_val, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "note": "PPOTrainer.train_iteration(): observations never exist on the host in this design; the per-step "
                           "host traffic is the LR scalar in and the metrics + episode statistics out"},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------- other BASELINE configs
CONFIGS = {
    "c2": WORKLOAD,
    "c3": "Rainbow DQN CartPole-v1, 8192 vectorised envs/GPU, PER sum-tree (capacity 2^21) + 5-step returns on device, one update of "
          "B = 8192 per lockstep, U*B/N = 1 (BASELINE configs[2])",
    "c4": "SAC Pendulum-v1, 4096 vectorised envs/GPU, replay capacity 2^20, twin-Q + auto-alpha update of B = 4096 per lockstep, "
          "U*B/N = 1 (BASELINE configs[3])",
    "c5": "PPO-full (mHC backbone) LunarLander-v3, 4096 vectorised envs/GPU (32768 over 8 GPUs), T=128, 4 epochs x 4 minibatches of "
          "131072 (BASELINE configs[4])",
}


def _make_config_runner(which, torch):
    """(trainer, step_fn, env_steps_per_step, public_step_fn, roofline_fn) for one of c3 / c4 / c5; a "step" = K_LOCK locksteps
    (c3, c4: act -> env step -> store -> update, one CUDA graph each) or one PPO-full iteration (c5)."""
    if which == "c3":
        from gymrl_b200.algorithms import rainbow_dqn_cartpole as R
        cfg = R.Config()
        cfg.num_envs, cfg.batch_size, cfg.memory_capacity, cfg.seed = 8192, 8192, 1 << 21, 0
        cfg.max_episodes = 10 ** 6      # keeps the LR / beta schedules away from their end points during the run
        tr = R.RainbowDQNTrainer(cfg)
        tr.env.reset(out=tr.cur)
        K = 200

        def step():
            for _ in range(K):
                tr.lockstep()

        def roof(value, peaks, src):
            # 377 + 548 * (U*B/N) algorithmic bytes per env-step, tree depth 21 -> 925 B at U*B/N = 1
            ach = 925.0 * value / 1e9
            return {"bound": "hbm", "kernel": "whole lockstep (sumtree_sample / sumtree_update / nstep_push / replay gather + the Q-net GEMMs)",
                    "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None,
                    "algorithmic_bytes_per_env_step": 925, "peak_source": src,
                    "note": "pointer-chasing tree walks over a 32 MB float64 tree and 42 kernels of <= 8192 threads per lockstep (parallel graph branches): "
                            "L2-latency / launch bound, nowhere near the HBM roofline (stated, not hidden)"}
        return tr, step, cfg.num_envs * K, step, roof
    if which == "c4":
        from gymrl_b200.algorithms import sac_pendulum as S
        cfg = S.Config()
        cfg.num_envs, cfg.batch_size, cfg.memory_capacity, cfg.seed = 4096, 4096, 1 << 20, 0
        tr = S.SACTrainer(cfg)
        tr.env.reset(out=tr.cur)
        K = 200

        def step():
            for _ in range(K):
                tr.lockstep()

        def roof(value, peaks, src):
            # ~2.5 MFLOP of dense layers per sampled transition at U*B/N = 1 (5 forwards, 3 backwards)
            peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
            ach = 2.5e6 * value / 1e12
            return {"bound": "tensor", "kernel": "whole lockstep (actor / twin-critic dense layers, B = 4096, 3xTF32 tcgen05 + skinny heads)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                    "algorithmic_flops_per_env_step": 2.5e6, "peak_source": src + ": dense bf16 cuBLAS, sustained",
                    "path_hbm": {"algorithmic_bytes_per_env_step": 72, "achieved_GBps": 72.0 * value / 1e9, "peak_GBps": peaks["hbm_gbs"]},
                    "note": "~84 kernels of M = 4096 per lockstep recorded as a DAG of parallel graph branches: launch-latency bound (0.31 ms per lockstep), far from either roofline"}
        return tr, step, cfg.num_envs * K, step, roof
    if which == "c5":
        from gymrl_b200.algorithms import ppo_full_lunarlander as F
        cfg = F.Config()
        cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.seed, cfg.max_train_steps = N_ENVS, T_STEPS, 4, 0, 10 ** 12
        tr = F.PPOTrainer(cfg)

        def step():
            tr.collect_experience()
            tr.update(None, read_metrics=False)

        def roof(value, peaks, src):
            peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
            ach = 3.7e6 * value / 1e12       # 3.7 MFLOP per env-step for PPO-full
            return {"bound": "tensor", "kernel": "whole iteration (mHC stage kernels are instruction bound; dense layers 3xTF32 tcgen05)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                    "algorithmic_flops_per_env_step": 3.7e6, "peak_source": src + ": dense bf16 cuBLAS, sustained",
                    "path_hbm": {"algorithmic_bytes_per_env_step": 278, "achieved_GBps": 278.0 * value / 1e9, "peak_GBps": peaks["hbm_gbs"]}}
        return tr, step, N_ENVS * T_STEPS, tr.train_iteration, roof
    raise SystemExit(f"unknown config {which}")


def run_config(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: gymrl_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world > 1 and args.config != "c5":
        raise SystemExit("c3 / c4 shard by independent replicas with per-rank replay (no exchange step): run them with --gpus 1")
    from gymrl_b200 import _ffi
    torch.manual_seed(0)
    tr, step, units, public_step, roof_fn = _make_config_runner(args.config, torch)
    flush = torch.empty(64 * 1024 * 1024, device="cuda", dtype=torch.float32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches():
        return _ffi.launch_count() + getattr(tr, "graph_launches", 0)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        step()
        ev[k][1].record()
    barrier()
    n_launch = launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    total = units * args.steps * world
    value = total / (dev_ms * 1e-3)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        public_step()
        if args.config != "c5":
            _ = tr.env.episode_stats(100)        # the per-log-line D2H of the public train() loop
    barrier()
    t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = total / float(t.item())
    if rank != 0:
        return
    peaks, src = measured_peaks()
    out = {"metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": CONFIGS[args.config], "env_steps_per_step": units, "l2": "flushed (256 MB write) before every timed step",
                      "cuda_graphs": True,
                      "parallelism": f"dp{world}" + (" (env shards; gradient sum per optimizer step: " + ("one-shot NVLink peer-memory reduction)" if getattr(tr, "comm", None) is not None else "ncclAllReduce)") if world > 1 else "")},
           "e2e": {"value": e2e_val, "unit": "env-steps/s", "h2d_bytes_per_step": 8 if args.config == "c5" else 16,
                   "d2h_bytes_per_step": 8232 if args.config == "c5" else 24,
                   "note": "public trainer API (train_iteration / lockstep + episode statistics), host wall clock; observations never "
                           "exist on the host in this design, the host traffic is schedule scalars in and metrics out"},
           "gpu_launches": int(n_launch), "clocks": clocks, "roofline": roof_fn(value, peaks, src), "cpu_baseline": None}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json config (default c2 = the metric's own)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    if args.config == "c2":
        run_ours(args, rank, world, local_rank)
    else:
        run_config(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
