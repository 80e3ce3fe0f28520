"""Learning evidence: run the trainers to the reference scripts' own stop criteria and record the curves.

    python tools/converge.py [ppo dqn rainbow sac ...] [--out gpurun_out/converge.json] [--budget-s 60]

Stop criteria are the reference's own: PPO LunarLander avg(100 episodes) >= 200 (ref ppo_lunarlander.py:361), DQN / Rainbow
CartPole avg(100) >= 495 (ref dqn_cartpole.py:207, rainbow_dqn_cartpole.py:400), SAC Pendulum avg(100) >= -200
(ref sac_pendulum.py:303).  After training each policy is scored by a deterministic evaluation on fresh env copies.
The curves land in the JSON (one row per log point); tests/test_gpu_converge.py runs the same functions with asserts.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def _quiet_eval(trainer, n):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        r = trainer.eval(n)
    return [float(x) for x in r]


def ppo(budget_s=90.0, seed=0, num_envs=4096, num_steps=128, num_minibatches=32, horizon=150_000_000, lr=None, target=200.0,
        verbose=False):
    """C2: PPO LunarLander-v3, 4096 envs x 128 steps, 10 epochs x 32 minibatches of 16384 (the bench configuration)."""
    from gymrl_b200.algorithms import ppo_lunarlander as P
    cfg = P.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.seed = num_envs, num_steps, num_minibatches, seed
    cfg.max_train_steps = horizon
    if lr is not None:
        cfg.lr = lr
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        tr = P.PPOTrainer(cfg)
    curve, solved_at = [], None
    t0 = time.time()
    while tr.step_count < cfg.max_train_steps and time.time() - t0 < budget_s:
        m, avg, total = tr.train_iteration()
        curve.append([tr.step_count, round(avg, 2), total, round(m["approx_kl"], 5), round(m["entropy"], 4), round(m["value_loss"], 3)])
        if verbose:
            print(curve[-1], flush=True)
        if total >= 100 and avg >= target:
            solved_at = tr.step_count
            break
    wall = time.time() - t0
    ev = _quiet_eval(tr, 256)
    return {"algo": "ppo_lunarlander", "config": f"{num_envs} envs x {num_steps} steps, {cfg.num_epochs} epochs x {num_minibatches} minibatches, lr {cfg.lr}",
            "criterion": f"avg100 >= {target} (ref ppo_lunarlander.py:361)", "solved_at_step": solved_at, "train_wall_s": round(wall, 2),
            "final_avg100": curve[-1][1] if curve else None, "eval_mean_256_deterministic": round(float(np.mean(ev)), 2),
            "dropped_manifold_events": tr.env.overflow_count(),
            "eval_frac_ge_200": round(float(np.mean(np.asarray(ev) >= 200.0)), 3),
            "curve_columns": ["env_steps", "avg100", "episodes", "approx_kl", "entropy", "value_loss"], "curve": curve}


def _offpolicy(tr, name, config_s, criterion, target, budget_s, log_every, lockstep_fn, verbose=False, min_episodes=100,
               extra=None):
    env = tr.env
    curve, solved_at = [], None
    t0 = time.time()
    step = 0
    last_total = -1
    while time.time() - t0 < budget_s:
        if getattr(tr, "max_train_steps", None) and getattr(tr, "total_steps", 0) + log_every > tr.max_train_steps:
            break   # the reference's LR / beta schedules end here (ref rainbow_dqn_cartpole.py:354-357 goes negative beyond it)
        for _ in range(log_every):
            lockstep_fn()
        step += log_every
        avg, _, total = env.episode_stats(100)
        if total != last_total:
            last_total = total
            curve.append([step * tr.N, round(avg, 2), total])
            if verbose:
                print(name, curve[-1], flush=True)
            if total >= min_episodes and avg >= target:
                solved_at = step * tr.N
                break
    wall = time.time() - t0
    ev = _quiet_eval(tr, 64)
    out = {"algo": name, "config": config_s, "criterion": criterion, "solved_at_step": solved_at, "train_wall_s": round(wall, 2),
           "final_avg100": curve[-1][1] if curve else None, "eval_mean_64_deterministic": round(float(np.mean(ev)), 2),
           "curve_columns": ["env_steps", "avg100", "episodes"], "curve": curve}
    if extra:
        out.update(extra)
    return out


def dqn(budget_s=60.0, seed=0, num_envs=64, batch_size=256, target_sync_updates=200, epsilon_decay=None, verbose=False):
    """DQN CartPole-v1 (reference hyper-parameters; N envs in lockstep, one update per lockstep)."""
    import contextlib
    import io
    from gymrl_b200.algorithms import dqn_cartpole as D
    cfg = D.Config()
    cfg.num_envs, cfg.seed, cfg.batch_size, cfg.target_sync_updates = num_envs, seed, batch_size, target_sync_updates
    if epsilon_decay is not None:
        cfg.epsilon_decay = epsilon_decay
    with contextlib.redirect_stdout(io.StringIO()):
        tr = D.DQNTrainer(cfg)
    tr.env.reset(out=tr.cur)
    lockstep = tr.lockstep       # act -> env step -> store -> update (+ the hard target sync schedule), one CUDA graph per lockstep

    return _offpolicy(tr, "dqn_cartpole", f"{num_envs} envs, B={batch_size}, target sync every {target_sync_updates} updates" if num_envs > 1
                      else "1 env, reference schedule", "avg100 >= 495 (ref dqn_cartpole.py:207)", 495.0, budget_s,
                      50 if num_envs > 1 else 200, lockstep, verbose)


def rainbow(budget_s=60.0, seed=0, num_envs=64, batch_size=256, capacity=None, max_episodes=None, verbose=False):
    """Rainbow DQN CartPole-v1 (reference hyper-parameters; N envs in lockstep, one PER update per lockstep)."""
    import contextlib
    import io
    from gymrl_b200.algorithms import rainbow_dqn_cartpole as R
    cfg = R.Config()
    cfg.num_envs, cfg.seed, cfg.batch_size = num_envs, seed, batch_size
    if capacity:
        cfg.memory_capacity = capacity
    if max_episodes:
        cfg.max_episodes = max_episodes
    with contextlib.redirect_stdout(io.StringIO()):
        tr = R.RainbowDQNTrainer(cfg)
    tr.env.reset(out=tr.cur)
    return _offpolicy(tr, "rainbow_dqn_cartpole", f"{num_envs} envs, B={batch_size}, capacity {cfg.memory_capacity}, max_episodes {cfg.max_episodes}",
                      "avg100 >= 495 (ref rainbow_dqn_cartpole.py:400)", 495.0, budget_s, 100, tr.lockstep, verbose)


def sac(budget_s=60.0, seed=0, num_envs=16, batch_size=256, verbose=False):
    """SAC Pendulum-v1 (reference hyper-parameters)."""
    import contextlib
    import io
    from gymrl_b200.algorithms import sac_pendulum as S
    cfg = S.Config()
    cfg.num_envs, cfg.seed, cfg.batch_size = num_envs, seed, batch_size
    with contextlib.redirect_stdout(io.StringIO()):
        tr = S.SACTrainer(cfg)
    tr.env.reset(out=tr.cur)
    return _offpolicy(tr, "sac_pendulum", f"{num_envs} envs, B={batch_size}", "avg100 >= -200 (ref sac_pendulum.py:303)", -200.0, budget_s, 200,
                      tr.lockstep, verbose)


def td3(budget_s=60.0, seed=0, num_envs=16, batch_size=256, verbose=False):
    import contextlib
    import io
    from gymrl_b200.algorithms import td3_pendulum as T
    cfg = T.Config()
    cfg.num_envs, cfg.seed, cfg.batch_size = num_envs, seed, batch_size
    with contextlib.redirect_stdout(io.StringIO()):
        tr = T.TD3Trainer(cfg)
    tr.env.reset(out=tr.cur)
    return _offpolicy(tr, "td3_pendulum", f"{num_envs} envs, B={batch_size}", "avg100 >= -200", -200.0, budget_s, 200, tr.lockstep, verbose)


RUNS = {"ppo": ppo, "dqn": dqn, "rainbow": rainbow, "sac": sac, "td3": td3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["ppo", "dqn", "rainbow", "sac"])
    ap.add_argument("--out", default="gpurun_out/converge.json")
    ap.add_argument("--budget-s", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--kw", default="{}", help="JSON of keyword overrides applied to every selected run")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    kw = json.loads(a.kw)
    results = []
    for w in a.which:
        try:
            r = RUNS[w](budget_s=a.budget_s, seed=a.seed, verbose=a.verbose, **kw)
        except Exception as e:   # keep the other runs' evidence
            import traceback
            traceback.print_exc()
            r = {"algo": w, "error": repr(e)}
        results.append(r)
        print(json.dumps({k: v for k, v in r.items() if k != "curve"}), flush=True)
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        prev = []
        if Path(a.out).exists():
            try:
                prev = json.loads(Path(a.out).read_text())
            except Exception:
                prev = []
        Path(a.out).write_text(json.dumps(prev + [r]))


if __name__ == "__main__":
    main()
