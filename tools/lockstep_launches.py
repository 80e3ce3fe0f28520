"""One eager off-policy lockstep (C3: Rainbow CartPole 8192 envs, or C4: SAC Pendulum 4096 envs) between
cudaProfilerStart/Stop, for a per-kernel launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/c3_launches.csv python tools/lockstep_launches.py c3
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "c3"
    if which == "c3":
        from gymrl_b200.algorithms import rainbow_dqn_cartpole as R
        cfg = R.Config()
        cfg.num_envs, cfg.batch_size, cfg.memory_capacity, cfg.seed, cfg.use_cuda_graph = 8192, 8192, 1 << 21, 0, False
        cfg.max_episodes = 10 ** 6
        tr = R.RainbowDQNTrainer(cfg)
    else:
        from gymrl_b200.algorithms import sac_pendulum as S
        cfg = S.Config()
        cfg.num_envs, cfg.batch_size, cfg.memory_capacity, cfg.seed, cfg.use_cuda_graph = 4096, 4096, 1 << 20, 0, False
        tr = S.SACTrainer(cfg)
    tr.env.reset(out=tr.cur)
    for _ in range(12):
        tr.lockstep()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    tr._lockstep_body()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
