"""Per-iteration rollout / update time of the C2 trainer while it LEARNS, for each LunarLander solver arrangement.

The arrangements are bit-identical, so three trainers with the same seed see the same policies and the same env states at every
iteration: any difference in the rollout time of iteration k is the step kernel's.  (bench.py's device-timed region is iterations
4-8 after start, its end-to-end region iterations 9-13: the env population shifts towards ground contact as PPO learns.)

    python tools/solver_drift.py [iterations=14] [solvers=0,2,3]
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    solvers = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,2,3").split(",")]
    from gymrl_b200.algorithms import ppo_lunarlander as P
    out = {}
    for v in solvers:
        cfg = P.Config()
        cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.num_epochs = 4096, 128, 32, 10
        cfg.seed, cfg.max_train_steps = 0, 10 ** 12
        torch.manual_seed(0)
        tr = P.PPOTrainer(cfg)
        tr.env.set_solver(v)
        roll, upd, ret = [], [], []
        for k in range(iters):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            tr._anneal()
            a.record()
            tr.collect_rollout()
            b.record()
            tr.update(None, read_metrics=False)
            c.record()
            torch.cuda.synchronize()
            roll.append(round(a.elapsed_time(b), 2)); upd.append(round(b.elapsed_time(c), 2))
            ret.append(round(tr._refresh_episode_rewards()[0], 1))
        out[f"solver{v}"] = {"rollout_ms": roll, "update_ms": upd, "avg_return": ret}
        del tr
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
