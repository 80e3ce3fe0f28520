"""One LunarLander step of 4096 envs (mid-rollout state, random actions) between cudaProfilerStart/Stop:

    ncu --profile-from-start off --set full --import-source on -k regex:lunar_step -o gpurun_out/env_step python tools/env_profile.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from gymrl_b200 import ops
    N = 4096
    env = ops.VecEnv("LunarLander-v3", N, seed=3)
    obs = torch.empty(N, 8, device="cuda")
    rew = torch.empty(N, device="cuda")
    te = torch.empty(N, dtype=torch.uint8, device="cuda")
    tu, dn = torch.empty_like(te), torch.empty_like(te)
    env.reset(out=obs)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
        a = torch.randint(0, 4, (N,), device="cuda", dtype=torch.int32, generator=g)
        env.step(a, obs=obs, reward=rew, terminated=te, truncated=tu, want_next_obs=False, done=dn)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    env.step(a, obs=obs, reward=rew, terminated=te, truncated=tu, want_next_obs=False, done=dn)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
