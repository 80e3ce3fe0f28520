"""LunarLander step latency vs envs-per-warp (GYMRL_LL_LANES) at a given env count, after a warm-up of random steps.
    GYMRL_LL_LANES=1 python tools/env_lanes_probe.py 512
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from gymrl_b200 import ops
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    env = ops.VecEnv("LunarLander-v3", N, seed=3)
    obs = torch.empty(N, 8, device="cuda")
    rew = torch.empty(N, device="cuda")
    te = torch.empty(N, dtype=torch.uint8, device="cuda")
    tu, dn = torch.empty_like(te), torch.empty_like(te)
    env.reset(out=obs)
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.randint(0, 4, (400, N), device="cuda", dtype=torch.int32, generator=g)
    for t in range(200):
        env.step(acts[t], obs=obs, reward=rew, terminated=te, truncated=tu, want_next_obs=False, done=dn)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for t in range(200, 400):
        env.step(acts[t], obs=obs, reward=rew, terminated=te, truncated=tu, want_next_obs=False, done=dn)
    b.record()
    torch.cuda.synchronize()
    print(f"N={N} us/step={a.elapsed_time(b) / 200 * 1000:.1f}")


if __name__ == "__main__":
    main()
