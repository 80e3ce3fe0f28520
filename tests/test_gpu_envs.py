"""GPU parity: device envs (through the C ABI) vs the CPU oracle restatements, same Philox seeds.

CartPole / Pendulum: float64 state, observations must agree to float32 rounding (libm vs CUDA sin/cos can
differ in the last float64 bit) and every flag exactly.  LunarLander: float32 solver built with -fmad=false
on both sides -> bit-exact observations, rewards, flags and full physics state.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from gymrl_b200 import ops
    return ops


def _run_pair(dev_env, ora_env, steps, action_fn, discrete, check):
    ops = _ops()
    obs_d = dev_env.reset().cpu().numpy()
    obs_o = ora_env.reset()
    check("reset", obs_d, obs_o)
    for t in range(steps):
        a = action_fn(t, obs_o)
        if discrete:
            a_t = torch.as_tensor(a.astype(np.int32), device="cuda")
        else:
            a_t = torch.as_tensor(a.astype(np.float32), device="cuda").reshape(dev_env.num_envs, -1)
        o, r, te, tr, no = dev_env.step(a_t)
        oo, ono, orr, ote, otr = ora_env.step(a)
        check(f"obs@{t}", o.cpu().numpy(), oo)
        check(f"next_obs@{t}", no.cpu().numpy(), ono)
        check(f"reward@{t}", r.cpu().numpy(), orr)
        assert np.array_equal(te.cpu().numpy(), ote), f"terminated@{t}"
        assert np.array_equal(tr.cpu().numpy(), otr), f"truncated@{t}"
        obs_o = oo


def test_cartpole_parity():
    from oracle.envs_np import CartPoleVec
    ops = _ops()
    N = 300
    dev, ora = ops.VecEnv("CartPole-v1", N, seed=7, first_env_id=11), CartPoleVec(N, seed=7, first_env_id=11)
    rng = np.random.default_rng(0)

    def check(tag, a, b):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-6, err_msg=tag)
    # a stabilising-ish policy so that some episodes run into the 500-step TimeLimit
    def act(t, obs):
        good = (obs[:, 2] * 10 + obs[:, 3] * 2 + obs[:, 0] * 0.5 + obs[:, 1] * 1.0 > 0).astype(np.int64)
        rnd = rng.integers(0, 2, N)
        return np.where(np.arange(N) % 3 == 0, good, rnd)
    _run_pair(dev, ora, 650, act, True, check)
    np.testing.assert_allclose(dev.get_state().cpu().numpy(), ora.get_state(), rtol=0, atol=1e-9)
    _, _, total = dev.episode_stats(64)
    assert total == len(ora.finished_returns) and total > 100
    assert max(ora.finished_lengths) == 500  # truncation path exercised


def test_episode_ring_single_warp():
    """One warp => ring order == env index order == the oracle's bookkeeping order: last-k means must agree."""
    from oracle.envs_np import CartPoleVec
    ops = _ops()
    N = 32
    dev, ora = ops.VecEnv("CartPole-v1", N, seed=21), CartPoleVec(N, seed=21)
    dev.reset(); ora.reset()
    rng = np.random.default_rng(4)
    for t in range(300):
        a = rng.integers(0, 2, N)
        dev.step(torch.as_tensor(a.astype(np.int32), device="cuda"))
        ora.step(a)
    for k in (1, 10, 100):
        mean_ret, mean_len, total = dev.episode_stats(k)
        assert total == len(ora.finished_returns)
        kk = min(k, total)
        assert abs(mean_ret - np.mean(ora.finished_returns[-kk:])) < 1e-4
        assert abs(mean_len - np.mean(ora.finished_lengths[-kk:])) < 1e-4


def test_pendulum_parity():
    from oracle.envs_np import PendulumVec
    ops = _ops()
    N = 257
    dev, ora = ops.VecEnv("Pendulum-v1", N, seed=3), PendulumVec(N, seed=3)
    rng = np.random.default_rng(1)

    def check(tag, a, b):
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-6, err_msg=tag)
    _run_pair(dev, ora, 450, lambda t, obs: rng.uniform(-3, 3, N), False, check)
    np.testing.assert_allclose(dev.get_state().cpu().numpy(), ora.get_state(), rtol=0, atol=1e-9)


def _heuristic(s):
    angle_targ = np.clip(s[:, 0] * 0.5 + s[:, 2] * 1.0, -0.4, 0.4)
    hover_targ = 0.55 * np.abs(s[:, 0])
    angle_todo = (angle_targ - s[:, 4]) * 0.5 - s[:, 5] * 1.0
    hover_todo = (hover_targ - s[:, 1]) * 0.5 - s[:, 3] * 0.5
    legs = (s[:, 6] > 0) | (s[:, 7] > 0)
    angle_todo = np.where(legs, 0.0, angle_todo)
    hover_todo = np.where(legs, -s[:, 3] * 0.5, hover_todo)
    a = np.zeros(len(s), np.int64)
    a = np.where(angle_todo > 0.05, 1, a)
    a = np.where(angle_todo < -0.05, 3, a)
    a = np.where((hover_todo > np.abs(angle_todo)) & (hover_todo > 0.05), 2, a)
    return a


@pytest.mark.parametrize("solver", [0, 2, 3])
def test_lunarlander_bit_exact(solver):
    from oracle.lunar import LunarLanderVec
    ops = _ops()
    N = 96
    dev, ora = ops.VecEnv("LunarLander-v3", N, seed=5, first_env_id=1000), LunarLanderVec(N, seed=5, first_env_id=1000)
    dev.set_solver(solver)
    assert dev.get_solver() == solver
    rng = np.random.default_rng(2)

    def check(tag, a, b):
        assert np.array_equal(a, b), f"{tag}: max |diff| = {np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))}"

    def act(t, obs):  # one third heuristic (lands, sleeps -> +100), one third random, one third no-op (crashes)
        return np.where(np.arange(N) % 3 == 0, _heuristic(obs), np.where(np.arange(N) % 3 == 1, rng.integers(0, 4, N), 0))
    _run_pair(dev, ora, 700, act, True, check)
    sd, so = dev.get_state().cpu().numpy(), ora.get_state()
    assert np.array_equal(sd, so), f"state planes differ at {np.argwhere(sd != so)[:5]}"
    _, _, total = dev.episode_stats(100)
    assert total >= N  # every env finished at least once (crash, landing or timeout paths all exercised)


@pytest.mark.parametrize("solver", [0, 2, 3])
def test_lunarlander_teacher_forced_single_steps(solver):
    """set_state from the oracle, one step, compare: isolates single-step arithmetic from trajectory divergence."""
    from oracle.lunar import LunarLanderVec
    ops = _ops()
    N = 64
    ora = LunarLanderVec(N, seed=9)
    dev = ops.VecEnv("LunarLander-v3", N, seed=9)
    dev.set_solver(solver)
    ora.reset(); dev.reset()
    rng = np.random.default_rng(3)
    for t in range(120):
        a = np.where(np.arange(N) % 2 == 0, _heuristic(ora.obs), rng.integers(0, 4, N))
        dev.set_state(torch.as_tensor(ora.get_state()))
        o, r, te, tr, no = dev.step(torch.as_tensor(a.astype(np.int32), device="cuda"))
        oo, ono, orr, ote, otr = ora.step(a)
        assert np.array_equal(no.cpu().numpy(), ono) and np.array_equal(r.cpu().numpy(), orr)
        assert np.array_equal(te.cpu().numpy(), ote)


def test_lunarlander_solver_variants_agree_at_size():
    """The arrangements of the solver loops (gymrl_env_set_solver) give the same bits: 2048 copies x 400 steps of a heuristic /
    random / no-op mix (landings, crashes, sleeping copies, time-outs), every output of every step and the final state snapshot.
    Variants 2 and 3 replace div.rn by its fast sequence inside the position iterations and repeat the phase with the plain
    operator when an operand leaves the sequence's exponent window - this is the test of that claim at size."""
    ops = _ops()
    N = 2048
    variants = (0, 2, 3)
    envs = [ops.VecEnv("LunarLander-v3", N, seed=21) for _ in variants]
    for v, env in zip(variants, envs):
        env.set_solver(v)
    o = [env.reset().clone() for env in envs]
    assert all(torch.equal(o[0], x) for x in o[1:])
    rng = np.random.default_rng(4)
    idx = np.arange(N)
    obs = o[0]
    for t in range(400):
        ob = obs.cpu().numpy()
        act = np.where(idx % 3 == 0, _heuristic(ob), np.where(idx % 3 == 1, rng.integers(0, 4, N), 0)).astype(np.int32)
        a = torch.as_tensor(act, device="cuda")
        r = [env.step(a) for env in envs]
        for k in range(1, len(envs)):
            for x, y in zip(r[0], r[k]):
                assert torch.equal(x, y), f"step {t}, solver {variants[k]}"
        obs = r[0][0].clone()
    s0 = envs[0].get_state()
    assert all(torch.equal(s0, env.get_state()) for env in envs[1:])
    totals = [env.episode_stats(100)[2] for env in envs]
    assert totals[0] >= N // 2 and all(t == totals[0] for t in totals)
    with pytest.raises(RuntimeError, match="unknown solver variant"):
        envs[0].set_solver(1)


def test_shard_independence():
    """Global env ids key the RNG: a shard reproduces the same envs of a bigger batch."""
    ops = _ops()
    big = ops.VecEnv("LunarLander-v3", 200, seed=1, first_env_id=0)
    part = ops.VecEnv("LunarLander-v3", 32, seed=1, first_env_id=100)
    ob, op = big.reset().clone(), part.reset().clone()
    assert torch.equal(ob[100:132], op)
    a = torch.randint(0, 4, (200,), device="cuda", dtype=torch.int32)
    for _ in range(50):
        ob = big.step(a)[0]
        op = part.step(a[100:132].contiguous())[0]
    assert torch.equal(ob[100:132], op)


def test_lunarlander_full_size_properties():
    """N = 4096 (BASELINE config): invariants that do not need the oracle."""
    ops = _ops()
    N = 4096
    env = ops.VecEnv("LunarLander-v3", N, seed=0)
    obs = env.reset().clone()
    assert torch.isfinite(obs).all()
    assert (obs[:, 1] > 1.3).all() and (obs[:, 1] < 1.5).all()       # spawn height
    assert (obs[:, 6:] == 0).all()
    done_total = torch.zeros(N, device="cuda")
    ret = torch.zeros(N, device="cuda", dtype=torch.float64)
    for t in range(400):
        a = torch.randint(0, 4, (N,), device="cuda", dtype=torch.int32)
        o, r, te, tr, no = env.step(a)
        assert torch.isfinite(o).all() and torch.isfinite(r).all()
        d = (te | tr).bool()
        # terminal rewards are exactly +-100; legs flags are binary
        assert ((r[te.bool()] == -100) | (r[te.bool()] == 100)).all()
        assert ((no[:, 6:] == 0) | (no[:, 6:] == 1)).all()
        # auto-reset: finished envs show a fresh spawn observation
        if d.any():
            assert (o[d][:, 1] > 1.3).all()
        done_total += d
    assert (done_total > 0).float().mean() > 0.95   # random policy crashes within ~100-150 steps
    mean_ret, mean_len, total = env.episode_stats(1000)
    assert -600 < mean_ret < -50 and 50 < mean_len < 200


def test_cartpole_pendulum_full_size_properties():
    ops = _ops()
    env = ops.VecEnv("CartPole-v1", 8192, seed=0)
    obs = env.reset()
    assert (obs.abs() <= 0.05).all()
    for _ in range(60):
        o, r, te, tr, no = env.step(torch.randint(0, 2, (8192,), device="cuda", dtype=torch.int32))
        assert (r == 1).all()
        assert ((no[:, 0].abs() > 2.4) | (no[:, 2].abs() > 12 * 2 * np.pi / 360))[te.bool()].all()
    pen = ops.VecEnv("Pendulum-v1", 4096, seed=0)
    obs = pen.reset()
    assert torch.allclose(obs[:, 0] ** 2 + obs[:, 1] ** 2, torch.ones(4096, device="cuda"), atol=1e-6)
    n_trunc = 0
    for t in range(200):
        o, r, te, tr, no = pen.step(torch.zeros(4096, 1, device="cuda"))
        assert (r <= 0).all() and (te == 0).all()
        n_trunc += int(tr.sum())
    assert n_trunc == 4096  # every env hits the 200-step TimeLimit exactly once


def test_lunarlander_reports_dropped_manifolds():
    """The device solver keeps 8 contact slots per env copy; a ninth touching manifold is dropped (as in the CPU oracle) and
    COUNTED (gymrl_env_overflow_count) instead of vanishing silently.  2048 copies x 300 random-action steps (many crash
    landings): the counter reads back and, with 3 bodies over 10 terrain edges, stays at 0."""
    import torch
    from gymrl_b200 import ops
    env = ops.VecEnv("LunarLander-v3", 2048, seed=11)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(300):
        env.step(torch.randint(0, 4, (2048,), device="cuda", dtype=torch.int32, generator=g), want_next_obs=False)
    n = env.overflow_count()
    _, _, total = env.episode_stats(100)
    assert isinstance(n, int) and total > 1000
    assert n == 0, f"{n} touching manifolds were dropped in {total} episodes"
    assert ops.VecEnv("CartPole-v1", 4).overflow_count() == 0
