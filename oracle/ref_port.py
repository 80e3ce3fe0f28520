"""CPU port of the reference's PPO LunarLander path, for timing only.  TEST/BENCH INFRASTRUCTURE.

The unmodified reference (algorithms/ppo_lunarlander.py) cannot travel to the GPU box (/root/reference
exists only in the build container, gymnasium/Box2D are not installed anywhere), so bench.py's
`cpu_baseline` / `--impl reference` legs time this restatement of the same single-process loop:
one env copy stepped one transition at a time (oracle/lunar_lander.c standing in for gymnasium+Box2D),
batch-1 policy forward + Categorical sample per step (ref :198-231), Python-loop fp64 GAE (ref :179-196),
10 epochs x 32 minibatches of 64 with torch CPU autograd, clip_grad_norm_(0.5), Adam(3e-4, eps 1e-5)
(ref :233-330).  kind = "port": same algorithm, same torch eager CPU kernels as the reference would run,
our env restatement underneath.  P independent worker processes (one per host core) give the
"reference on all host cores" figure.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


class PortTrainer:
    """One single-env PPO learner of the reference's shape; step() = collect_rollout + update (one iteration)."""

    def __init__(self, seed: int = 0, update_freq: int = 2048):
        import torch
        import torch.nn as nn

        from .lunar import LunarLanderVec
        torch.set_num_threads(1)
        torch.manual_seed(seed)
        np.random.seed(seed)

        def lin(i, o, std=np.sqrt(2)):
            l = nn.Linear(i, o)
            nn.init.orthogonal_(l.weight, gain=std)
            nn.init.constant_(l.bias, 0)
            return l

        self.shared = nn.Sequential(lin(8, 256), nn.Tanh(), lin(256, 256), nn.Tanh())
        self.actor = nn.Sequential(lin(256, 256), nn.Tanh(), lin(256, 4, 0.01))
        self.critic = nn.Sequential(lin(256, 256), nn.Tanh(), lin(256, 1, 1.0))
        self.params = list(self.shared.parameters()) + list(self.actor.parameters()) + list(self.critic.parameters())
        self.opt = torch.optim.Adam(self.params, lr=3e-4, eps=1e-5)
        self.env = LunarLanderVec(1, seed=seed, first_env_id=seed)
        self.update_freq = update_freq

    def step(self) -> int:
        import torch
        import torch.nn as nn
        from torch.distributions import Categorical
        shared, actor, critic, opt, env, params = self.shared, self.actor, self.critic, self.opt, self.env, self.params
        gamma, lam, clip, dual, ent_c, vf_c = 0.99, 0.95, 0.2, 3.0, 0.01, 0.5
        S, Aa, LP, V, R, D = [], [], [], [], [], []
        state = env.reset()[0]
        for _ in range(self.update_freq):
            st = torch.tensor(state, dtype=torch.float32).unsqueeze(0)
            with torch.no_grad():
                f = shared(st)
                dist = Categorical(logits=actor(f))
                a = dist.sample()
                lp, v = dist.log_prob(a).item(), critic(f).squeeze().item()
            obs, _, r, te, tr = env.step(np.array([a.item()], dtype=np.int32))
            S.append(state); Aa.append(a.item()); LP.append(lp); V.append(v); R.append(float(r[0])); D.append(bool(te[0] or tr[0]))
            state = obs[0]
        with torch.no_grad():
            nv = critic(shared(torch.tensor(state, dtype=torch.float32).unsqueeze(0))).squeeze().item()
        rew, done, vals = np.array(R), np.array(D, dtype=np.float32), np.array(V + [nv])
        adv = np.zeros_like(rew)
        last = 0.0
        for t in reversed(range(len(rew))):
            delta = rew[t] + gamma * vals[t + 1] * (1 - done[t]) - vals[t]
            adv[t] = last = delta + gamma * lam * (1 - done[t]) * last
        ret = adv + vals[:-1]
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        St = torch.tensor(np.array(S), dtype=torch.float32)
        At = torch.tensor(Aa, dtype=torch.long)
        LPt, ADt, RTt = (torch.tensor(x, dtype=torch.float32) for x in (LP, adv, ret))
        idx = np.arange(len(rew))
        for _ in range(10):
            np.random.shuffle(idx)
            for s in range(0, len(idx), 64):
                mb = idx[s:s + 64]
                f = shared(St[mb])
                dist = Categorical(logits=actor(f))
                nlp, vv, ent = dist.log_prob(At[mb]), critic(f).squeeze(-1), dist.entropy()
                ratio = torch.exp(nlp - LPt[mb])
                s1, s2 = ratio * ADt[mb], torch.clamp(ratio, 1 - clip, 1 + clip) * ADt[mb]
                ms = torch.min(s1, s2)
                pl = -torch.mean(torch.where(ADt[mb] < 0, torch.max(ms, dual * ADt[mb]), ms))
                loss = pl + vf_c * torch.mean((vv - RTt[mb]).pow(2)) - ent_c * ent.mean()
                opt.zero_grad()
                loss.backward()
                nn.utils.clip_grad_norm_(params, 0.5)
                opt.step()
                _ = (pl.item(), loss.item())  # the reference syncs 5 scalars per minibatch (:309-322)
        return self.update_freq


def _worker(args):
    seed, n_rollouts, update_freq = args
    tr = PortTrainer(seed, update_freq)
    steps = 0
    t0 = time.perf_counter()
    for _ in range(n_rollouts):
        steps += tr.step()
    return steps, time.perf_counter() - t0


def run_ppo_port(n_rollouts: int = 1, update_freq: int = 2048, processes: int | None = None, seed: int = 0):
    """Returns dict(value=env-steps/s summed over worker processes, cores=processes, sample=...)."""
    P = processes or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(P) as pool:
        res = pool.map(_worker, [(seed + i, n_rollouts, update_freq) for i in range(P)])
    wall = time.perf_counter() - t0
    total_steps = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return {"value": total_steps / slowest, "cores": P, "wall_s": wall,
            "sample": f"{P} independent single-env PPO workers x {n_rollouts} rollout(s) of {update_freq} steps + full update "
                      f"(10 epochs x 32 minibatches of 64), torch CPU 1 thread each, oracle C env"}


if __name__ == "__main__":
    print(run_ppo_port(1, 2048, processes=2))
