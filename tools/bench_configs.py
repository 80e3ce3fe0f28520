"""Throughput of the other BASELINE configs at SURVEY §8(d)'s sizes (bench.py measures C2 only; these are the parity-test
configs, measured here so the design notes can quote them), device-timed with CUDA events:

  C3  Rainbow DQN CartPole-v1, 8192 envs, PER capacity 2^21, one update of B = 8192 per lockstep (U*B/N = 1)
  C4  SAC Pendulum-v1, 4096 envs, replay capacity 2^20, one update of B = 4096 per lockstep (U*B/N = 1)
  C5  PPO-full LunarLander-v3, one 4096-env shard, T = 128, 4 epochs x 4 minibatches

    python tools/bench_configs.py [c3|c4|c5 ...] [--locksteps 300]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def timed_locksteps(step_fn, warm, n):
    for _ in range(warm):
        step_fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        step_fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


def c3(a):
    from gymrl_b200 import _ffi
    from gymrl_b200.algorithms import rainbow_dqn_cartpole as R
    cfg = R.Config()
    cfg.num_envs, cfg.batch_size, cfg.memory_capacity, cfg.seed = 8192, 8192, 1 << 21, 0
    cfg.max_episodes = 10 ** 6   # keeps the LR / beta schedules away from their end points during the probe
    cfg.use_cuda_graph = not a.eager
    tr = R.RainbowDQNTrainer(cfg)
    tr.env.reset(out=tr.cur)
    step = tr.lockstep

    c0 = _ffi.launch_count()
    ms, wall = timed_locksteps(step, 20, a.locksteps)
    launches = (_ffi.launch_count() - c0 + tr.graph_launches) / (a.locksteps + 20)
    return {"config": "C3 Rainbow CartPole-v1, 8192 envs, B=8192, capacity 2^21, U=1", "ms_per_lockstep": round(ms, 4),
            "host_ms_per_lockstep": round(wall, 4), "env_steps_per_s": round(cfg.num_envs / (ms * 1e-3)), "launches_per_lockstep": round(launches, 1),
            "cuda_graph": not a.eager}


def c4(a):
    from gymrl_b200 import _ffi
    from gymrl_b200.algorithms import sac_pendulum as S
    cfg = S.Config()
    cfg.num_envs, cfg.batch_size, cfg.memory_capacity, cfg.seed = 4096, 4096, 1 << 20, 0
    cfg.use_cuda_graph = not a.eager
    tr = S.SACTrainer(cfg)
    tr.env.reset(out=tr.cur)
    step = tr.lockstep

    c0 = _ffi.launch_count()
    ms, wall = timed_locksteps(step, 20, a.locksteps)
    launches = (_ffi.launch_count() - c0 + tr.graph_launches) / (a.locksteps + 20)
    return {"config": "C4 SAC Pendulum-v1, 4096 envs, B=4096, capacity 2^20, U=1", "ms_per_lockstep": round(ms, 4),
            "host_ms_per_lockstep": round(wall, 4), "env_steps_per_s": round(cfg.num_envs / (ms * 1e-3)), "launches_per_lockstep": round(launches, 1),
            "cuda_graph": not a.eager}


def c5(a):
    from gymrl_b200.algorithms import ppo_full_lunarlander as F
    cfg = F.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.seed = 4096, 128, 4, 0
    tr = F.PPOTrainer(cfg)
    for _ in range(3):
        tr.collect_experience(); tr.update(None, read_metrics=False)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tot_r = tot_u = 0.0
    iters = 4
    for _ in range(iters):
        ev[0].record(); tr.collect_experience(); ev[1].record(); tr.update(None, read_metrics=False); ev[2].record()
        torch.cuda.synchronize()
        tot_r += ev[0].elapsed_time(ev[1]); tot_u += ev[1].elapsed_time(ev[2])
    n = 4096 * 128 * iters
    return {"config": "C5 shard: PPO-full LunarLander-v3, 4096 envs, T=128, 4 epochs x 4 minibatches of 131072", "rollout_ms": round(tot_r / iters, 2),
            "update_ms": round(tot_u / iters, 2), "env_steps_per_s": round(n / ((tot_r + tot_u) * 1e-3))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c3", "c4", "c5"])
    ap.add_argument("--locksteps", type=int, default=300)
    ap.add_argument("--eager", action="store_true", help="off-policy locksteps as eager launches instead of one CUDA graph")
    a = ap.parse_args()
    for w in a.which:
        print(json.dumps({"c3": c3, "c4": c4, "c5": c5}[w](a)), flush=True)


if __name__ == "__main__":
    main()
