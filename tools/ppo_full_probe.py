"""Throughput probe of the C5 per-GPU shard (PPO-full LunarLander-v3, 4096 envs, T = 128, 4 epochs x 4 minibatches of
131072): device-timed env-steps/s of rollout + GAE + update.   python tools/ppo_full_probe.py [iters]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from gymrl_b200.algorithms import ppo_full_lunarlander as F
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    cfg = F.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.seed = 4096, 128, 4, 0
    tr = F.PPOTrainer(cfg)
    for _ in range(2):
        tr.collect_experience(); tr.update(None, read_metrics=False)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tot_r = tot_u = 0.0
    for _ in range(iters):
        ev[0].record(); tr.collect_experience(); ev[1].record(); tr.update(None, read_metrics=False); ev[2].record()
        torch.cuda.synchronize()
        tot_r += ev[0].elapsed_time(ev[1]); tot_u += ev[1].elapsed_time(ev[2])
    n = 4096 * 128 * iters
    print(f"ppo_full C5 shard: rollout {tot_r / iters:.1f} ms, update {tot_u / iters:.1f} ms, {n / ((tot_r + tot_u) * 1e-3):,.0f} env-steps/s")
    m, avg, total = tr.train_iteration()
    print("metrics", m, "avg reward", avg, "episodes", total)


if __name__ == "__main__":
    main()
