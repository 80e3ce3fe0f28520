"""Run the UNMODIFIED reference trainer files on oracle/gymnasium_shim.  TEST / BENCH INFRASTRUCTURE ONLY.

    python -m oracle.run_reference dqn_cartpole --steps 5000
    python -m oracle.run_reference ppo_lunarlander --steps 40960 --threads 1
    python -m oracle.run_reference all-cores ppo_lunarlander --steps 4096      # P pinned worker processes, summed

The files are executed where they lie under /root/reference (importlib; nothing is copied), with `cfg.device = "cpu"`,
torch / numpy / random seeded by this harness (the reference seeds none of them), and a step budget enforced
from OUTSIDE the script: `env.step` is wrapped by a counter that raises once the budget is spent, so schedules that depend
on `cfg.max_train_steps` / `cfg.max_episodes` (LR anneal, beta anneal) are the reference's own.  Returns the env-steps/s of
the whole loop (policy forward + env + buffer + update()), the finished episodes' returns, and what the script printed last.

/root/reference exists only in the build container: on the GPU box `available()` is False and bench.py falls back to
oracle/ref_port.py (the hand port of the same loop) — see bench.py `--impl reference`.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import random
import sys
import time
from pathlib import Path

import numpy as np

from . import ref_loader

SCRIPTS = {
    # name: (file, trainer class, attr holding the env, off-policy?)
    "dqn_cartpole": ("algorithms/dqn_cartpole.py", "DQNTrainer"),
    "ppo_lunarlander": ("algorithms/ppo_lunarlander.py", "PPOTrainer"),
    "rainbow_dqn_cartpole": ("algorithms/rainbow_dqn_cartpole.py", "RainbowDQNTrainer"),
    "sac_pendulum": ("algorithms/sac_pendulum.py", "SACTrainer"),
    "td3_pendulum": ("algorithms/td3_pendulum.py", "TD3Trainer"),
    "ppo_full_lunarlander": ("algorithms/ppo_full_lunarlander.py", "PPOTrainer"),
}


class BudgetReached(Exception):
    pass


def available() -> bool:
    return ref_loader.available()


def _wrap_env(env, budget, log):
    """Count env.step calls, record finished episodes' returns, raise BudgetReached when the budget is spent."""
    step0, reset0 = env.step, env.reset
    st = {"steps": 0, "ret": 0.0, "t_first": None}

    def step(a):
        if st["steps"] >= budget:
            raise BudgetReached()
        if st["t_first"] is None:
            st["t_first"] = time.perf_counter()
        out = step0(a)
        st["steps"] += 1
        st["ret"] += float(out[1])
        if out[2] or out[3]:
            log.append(st["ret"])
            st["ret"] = 0.0
        return out

    def reset(*a, **k):
        st["ret"] = 0.0
        return reset0(*a, **k)

    env.step, env.reset = step, reset
    return st


def run(script: str, steps: int, seed: int = 0, threads: int | None = 1, quiet: bool = True, cfg_overrides: dict | None = None):
    """Execute `script`'s Trainer.train() for `steps` env steps.  Returns a dict (env_steps_per_s, returns, ...)."""
    if not available():
        raise FileNotFoundError("/root/reference not present")
    import torch
    from . import gymnasium_shim
    gymnasium_shim.install()
    if threads:
        torch.set_num_threads(int(threads))
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    path, cls = SCRIPTS[script]
    mod = ref_loader.load(path, name=f"ref_run_{script}")
    cfg = mod.Config()
    cfg.device = "cpu"
    for k, v in (cfg_overrides or {}).items():
        setattr(cfg, k, v)
    returns: list = []
    out = io.StringIO()
    with contextlib.redirect_stdout(out if quiet else sys.stdout):
        trainer = getattr(mod, cls)(cfg)
        st = _wrap_env(trainer.env, steps, returns)
        t0 = time.perf_counter()
        try:
            trainer.train()
            finished = True
        except BudgetReached:
            finished = False
        wall = time.perf_counter() - t0
    lines = [l for l in out.getvalue().splitlines() if l.strip()]
    return {"script": path, "kind": "reference", "steps": st["steps"], "wall_s": wall, "env_steps_per_s": st["steps"] / max(wall, 1e-9),
            "episodes": len(returns), "returns": returns, "train_returned": finished, "threads": threads or torch.get_num_threads(),
            "last_line": lines[-1] if lines else "", "solved": any("solved" in l.lower() for l in lines)}


def _worker(args):
    script, steps, seed, core = args
    if core is not None:
        try:
            os.sched_setaffinity(0, {core})
        except Exception:
            pass
    r = run(script, steps, seed=seed, threads=1)
    return r["steps"], r["wall_s"]


def run_all_cores(script: str, steps_per_worker: int, processes: int | None = None, seed: int = 0, pool=None):
    """P independent pinned single-thread processes of the unmodified script; value = summed steps
    over the slowest worker's wall time."""
    import multiprocessing as mp
    cores = sorted(os.sched_getaffinity(0))
    P = processes or len(cores)
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(P)
    try:
        res = pool.map(_worker, [(script, steps_per_worker, seed + i, cores[i % len(cores)]) for i in range(P)])
    finally:
        if own:
            pool.close(); pool.join()
    total, slowest = sum(r[0] for r in res), max(r[1] for r in res)
    return {"value": total / slowest, "cores": P, "kind": "reference", "script": SCRIPTS[script][0],
            "sample": f"{P} pinned single-thread processes x {steps_per_worker} env steps of the unmodified {SCRIPTS[script][0]} on oracle/gymnasium_shim"}


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("script")
    ap.add_argument("rest", nargs="*")
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--processes", type=int, default=None)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.script == "all-cores":
        r = run_all_cores(a.rest[0], a.steps, a.processes, a.seed)
    else:
        r = run(a.script, a.steps, a.seed, a.threads, quiet=not a.verbose)
    if a.out:
        Path(a.out).write_text(json.dumps(r))
    r = dict(r)
    if "returns" in r and len(r["returns"]) > 12:
        r["returns"] = r["returns"][:4] + ["..."] + r["returns"][-8:]
    print(json.dumps(r))


if __name__ == "__main__":
    main()
