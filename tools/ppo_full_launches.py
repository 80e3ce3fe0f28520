"""Launch list of the C5 shard (PPO-full LunarLander-v3, 4096 envs): one rollout step and one minibatch (131072 rows) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv`."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from gymrl_b200.algorithms import ppo_full_lunarlander as F
    cfg = F.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.seed, cfg.use_cuda_graph = 4096, 128, 4, 0, False
    tr = F.PPOTrainer(cfg)
    tr.collect_experience()
    tr.update(None, read_metrics=False)
    buf, N = tr.buffer, tr.N
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    # one lockstep of the rollout (policy forward, sample, env step)
    tr.net.forward(buf.obs[0], tr.acts_roll, N)
    tr._sample_step(tr.acts_roll, 0)
    tr.env.step(buf.action[0], obs=buf.obs[1], reward=buf.reward[0], terminated=tr.term, truncated=tr.trunc, want_next_obs=False, done=buf.done[0])
    torch.cuda.synchronize()
    # one minibatch
    tr.ctr_mb.zero_()
    tr._minibatch_body()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
