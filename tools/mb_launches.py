"""One eager PPO minibatch (C2 shape: 16384 rows) between cudaProfilerStart/Stop, for a per-kernel launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/mb_launches.csv python tools/mb_launches.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from gymrl_b200.algorithms import ppo_lunarlander as P
    cfg = P.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.num_epochs, cfg.use_cuda_graph = 4096, 128, 32, 1, False
    tr = P.PPOTrainer(cfg)
    tr.collect_rollout()
    tr.update(None, read_metrics=False)
    tr.ctr_mb.zero_()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    tr._minibatch_body()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
