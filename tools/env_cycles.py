"""Per-env cycle counters of the LunarLander step kernel (gymrl_env_set_profile): where the slowest envs of a step
spend their time.  Random policy, 4096 envs, counters sampled over a window of steps after a warm-up.
    python tools/env_cycles.py [N] [warm] [steps]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from gymrl_b200 import ops
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    env = ops.VecEnv("LunarLander-v3", N, seed=3)
    obs = torch.empty(N, 8, device="cuda")
    rew = torch.empty(N, device="cuda")
    te = torch.empty(N, dtype=torch.uint8, device="cuda")
    tu, dn = torch.empty_like(te), torch.empty_like(te)
    env.reset(out=obs)
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.randint(0, 4, (warm + steps, N), device="cuda", dtype=torch.int32, generator=g)
    for t in range(warm):
        env.step(acts[t], obs=obs, reward=rew, terminated=te, truncated=tu, want_next_obs=False, done=dn)
    prof = torch.zeros(N, 8, dtype=torch.int64, device="cuda")
    env.set_profile(prof)
    rows = []
    for t in range(warm, warm + steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.step(acts[t], obs=obs, reward=rew, terminated=te, truncated=tu, want_next_obs=False, done=dn)
        b.record()
        torch.cuda.synchronize()
        p = prof.cpu()
        rows.append((a.elapsed_time(b) * 1e3, p.clone()))
    env.set_profile(None)
    names = ["step", "collide", "setup", "vel", "pos", "nc", "pos_it", "slot"]
    import numpy as np
    us = np.array([r[0] for r in rows])
    print(f"kernel+tick us/step: median {np.median(us):.1f} min {us.min():.1f} max {us.max():.1f}")
    allp = torch.stack([r[1] for r in rows]).numpy()   # [steps, N, 8]
    mx = allp[:, :, 0].max(axis=1)
    print(f"slowest env per step (cycles): median {np.median(mx):.0f}  -> {np.median(mx) / 1.965e3:.1f} us at 1965 MHz")
    # the slowest env of each step: phase split
    idx = allp[:, :, 0].argmax(axis=1)
    top = allp[np.arange(len(rows)), idx]
    print("slowest env of each step, mean over steps:", {n: float(top[:, k].mean()) for k, n in enumerate(names)})
    flat = allp.reshape(-1, 8)
    for nc in range(0, 7):
        m = flat[:, 5] == nc
        if m.sum() == 0:
            continue
        f = flat[m]
        print(f"nc={nc}: count {m.sum():7d}  step {f[:, 0].mean():9.0f} (p99 {np.percentile(f[:, 0], 99):9.0f})  collide {f[:, 1].mean():7.0f}  "
              f"setup {f[:, 2].mean():7.0f}  vel {f[:, 3].mean():9.0f}  pos {f[:, 4].mean():8.0f}  pos_it {f[:, 6].mean():5.1f}")
    # contact layout of the heavy envs: (leg-1 contacts, leg-2 contacts, 2-point blocks, lander contacts) -> count, mean cycles
    from collections import Counter
    hv = flat[flat[:, 5] > 0]
    lay = Counter()
    cyc = Counter()
    for row in hv:
        k = int(row[7])
        key = (k & 15, (k >> 4) & 15, (k >> 8) & 15, (k >> 12) & 15)
        lay[key] += 1
        cyc[key] += int(row[0])
    for key, n in sorted(lay.items(), key=lambda kv: -kv[1])[:16]:
        print(f"  legs {key[0]}+{key[1]} (2-pt blocks {key[2]}, lander {key[3]}): {n:6d} envs-steps, mean {cyc[key] / n:9.0f} cycles")
    top_k = top[:, 7]
    print("  slowest env of each step, layouts:", Counter((int(k) & 15, (int(k) >> 4) & 15, (int(k) >> 8) & 15) for k in top_k).most_common(8))
    light = flat[flat[:, 7] >= 0]
    print("all: mean step cycles", light[:, 0].mean())


if __name__ == "__main__":
    main()
