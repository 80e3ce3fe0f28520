"""X1 (BASELINE configs[0]): the UNMODIFIED reference trainer files run end to end on
oracle/gymnasium_shim, and the shim's scalar envs agree with the vector oracles the CUDA envs are checked against.

CPU only.  The reference-script tests are skipped where /root/reference is absent (the GPU box)."""
import math

import numpy as np
import pytest

from oracle import envs_np, gymnasium_shim as shim, ref_loader, run_reference

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


# --------------------------------------------------------------------------------------------- shim vs vector oracle
def test_shim_api_surface():
    g = shim.install()
    import gymnasium
    assert gymnasium is g and gymnasium.spaces.Box is shim.Box
    from gymnasium.wrappers import AtariPreprocessing  # noqa: F401  (import-only, ref utils/runner.py:6)
    for name, d, steps in (("CartPole-v1", 4, 500), ("Pendulum-v1", 3, 200), ("LunarLander-v3", 8, 1000)):
        env = gymnasium.make(name, render_mode=None)
        assert env.observation_space.shape == (d,) and env.spec.max_episode_steps == steps
        obs, info = env.reset(seed=3)
        assert obs.dtype == np.float32 and obs.shape == (d,) and info == {}
        a = env.action_space.sample()
        obs2, r, te, tr, info = env.step(a)
        assert obs2.dtype == np.float32 and isinstance(te, bool) and isinstance(tr, bool) and isinstance(float(r), float)
        assert env.unwrapped is env
        env.close()
    assert gymnasium.make("CartPole-v1").action_space.n == 2
    pend = gymnasium.make("Pendulum-v1")
    assert isinstance(pend.action_space, gymnasium.spaces.Box) and pend.action_space.high[0] == 2.0 and pend.action_space.shape == (1,)
    with pytest.raises(ValueError):
        gymnasium.make("FlappyBird-v0")


def test_shim_seed_semantics():
    """reset(seed=s) restarts the stream (same first state every time, ref dqn_cartpole.py:174 with cfg.seed set);
    reset() continues it."""
    for name in ("CartPole-v1", "Pendulum-v1", "LunarLander-v3"):
        env = shim.make(name)
        a, _ = env.reset(seed=11)
        b, _ = env.reset(seed=11)
        c, _ = env.reset()
        assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_shim_cartpole_matches_vector_oracle():
    env = shim.make("CartPole-v1")
    vec = envs_np.CartPoleVec(1, seed=0)
    rng = np.random.default_rng(0)
    env.reset(seed=5)
    for ep in range(20):
        obs, _ = env.reset()
        vec.state[0] = env.state        # same float64 start state (the two draw from different generators)
        vec.elapsed[0] = 0
        for t in range(600):
            a = int(rng.integers(2))
            o, r, te, tr, _ = env.step(a)
            _, nobs, vr, vte, vtr = vec.step(np.array([a]))
            assert np.array_equal(o, nobs[0]) and r == float(vr[0]) and te == bool(vte[0]) and tr == bool(vtr[0])
            if te or tr:
                break
        assert te or tr


def test_shim_cartpole_time_limit():
    env = shim.make("CartPole-v1")
    env.reset(seed=0)
    # a bang-bang controller on the pole angle + angular velocity keeps the pole up long enough to hit TimeLimit(500)
    obs, _ = env.reset()
    for t in range(500):
        a = 1 if obs[2] + 0.5 * obs[3] + 0.02 * obs[0] + 0.05 * obs[1] > 0 else 0
        obs, r, te, tr, _ = env.step(a)
        if te or tr:
            break
    assert t == 499 and tr and not te


def test_shim_pendulum_matches_vector_oracle():
    env = shim.make("Pendulum-v1")
    vec = envs_np.PendulumVec(1, seed=0)
    rng = np.random.default_rng(1)
    env.reset(seed=2)
    vec.state[0] = env.state
    vec.elapsed[0] = 0
    for t in range(200):
        a = rng.uniform(-3, 3, size=(1,)).astype(np.float32)
        o, r, te, tr, _ = env.step(a)
        _, nobs, vr, vte, vtr = vec.step(a)
        assert np.array_equal(o, nobs[0]) and np.float32(r) == vr[0] and not te and tr == bool(vtr[0])
    assert tr


def test_shim_lunarlander_matches_vector_oracle():
    from oracle.lunar import LunarLanderVec
    env = shim.make("LunarLander-v3")
    obs, _ = env.reset(seed=9)
    vec = LunarLanderVec(1, seed=9, first_env_id=0)
    assert np.array_equal(obs, vec.reset()[0])
    rng = np.random.default_rng(3)
    done_seen = 0
    for t in range(1500):
        a = int(rng.integers(4))
        o, r, te, tr, _ = env.step(a)
        vo, vn, vr, vte, vtr = vec.step(np.array([a], np.int32))
        assert np.array_equal(o, vn[0]) and np.float32(r) == vr[0] and te == bool(vte[0]) and tr == bool(vtr[0])
        if te or tr:
            done_seen += 1
            o, _ = env.reset()                      # the vector oracle auto-resets into the same next episode
            assert np.array_equal(o, vo[0])
    assert done_seen >= 3


def test_shim_lunarlander_heuristic_lands():
    """gymnasium's own heuristic controller (lunar_lander.py::heuristic) lands on the shim env: the external anchor the
    unpinnable env arithmetic has."""
    def heuristic(s):
        angle_targ = s[0] * 0.5 + s[2] * 1.0
        angle_targ = max(-0.4, min(0.4, angle_targ))
        hover_targ = 0.55 * abs(s[0])
        angle_todo = (angle_targ - s[4]) * 0.5 - s[5] * 1.0
        hover_todo = (hover_targ - s[1]) * 0.5 - s[3] * 0.5
        if s[6] or s[7]:
            angle_todo = 0
            hover_todo = -(s[3]) * 0.5
        a = 0
        if hover_todo > abs(angle_todo) and hover_todo > 0.05:
            a = 2
        elif angle_todo < -0.05:
            a = 3
        elif angle_todo > +0.05:
            a = 1
        return a
    env = shim.make("LunarLander-v3")
    rets = []
    env.reset(seed=123)
    for ep in range(10):
        s, _ = env.reset()
        tot = 0.0
        while True:
            s, r, te, tr, _ = env.step(heuristic(s))
            tot += r
            if te or tr:
                break
        rets.append(tot)
    assert np.mean(rets) > 150.0, rets


# --------------------------------------------------------------------------------------------- unmodified reference scripts
@needs_ref
def test_reference_dqn_cartpole_runs_and_learns():
    """algorithms/dqn_cartpole.py, untouched, 1 env, CPU (BASELINE configs[0]): runs through select_action / env.step /
    ReplayBuffer.push / update() / the hard target sync, and the returns rise within the first 2500 steps."""
    r = run_reference.run("dqn_cartpole", steps=2500, seed=0)
    assert r["steps"] == 2500 and r["episodes"] >= 10
    early, late = np.mean(r["returns"][:10]), np.mean(r["returns"][-5:])
    print(f"reference dqn_cartpole on the shim: {r['env_steps_per_s']:.0f} env-steps/s, first-10 mean {early:.1f}, last-5 mean {late:.1f}")
    assert late > 2.0 * early, r["returns"]


@needs_ref
def test_reference_ppo_lunarlander_runs():
    """algorithms/ppo_lunarlander.py, untouched: one full rollout of 2048 steps + compute_gae + 320 minibatch updates."""
    r = run_reference.run("ppo_lunarlander", steps=2049, seed=0)
    assert r["steps"] == 2049 and r["episodes"] >= 5
    assert "Updates: 1" in r["last_line"] and "KL:" in r["last_line"], r["last_line"]
    assert all(math.isfinite(x) and -1500.0 < x < 400.0 for x in r["returns"])


@needs_ref
@pytest.mark.parametrize("script,steps", [("rainbow_dqn_cartpole", 320), ("sac_pendulum", 260), ("td3_pendulum", 260),
                                          ("ppo_full_lunarlander", 200)])
def test_reference_other_scripts_run(script, steps):
    r = run_reference.run(script, steps=steps, seed=0)
    assert r["steps"] == steps and r["episodes"] >= 1 and all(math.isfinite(x) for x in r["returns"])


@needs_ref
@pytest.mark.parametrize("script,agent_cls", [("legacy/LunarLander(PPO).py", "PPO"), ("legacy/CartPole(NDQN).py", "DQN")])
def test_legacy_scripts_import_against_our_utils_surface(script, agent_cls):
    """The two in-scope `utils` consumers execute their module bodies UNCHANGED against gymrl_b200.utils.install():
    every name they take from `utils.model / utils.buffer / utils.runner` resolves, their Config(BasicConfig) builds, and their agent
    classes carry the protocol utils.runner.train drives (choose_action / evaluate / update, ModelLoader save/load).  Running
    BenchMark.train needs the device env, i.e. a GPU, and /root/reference is not on the GPU box — tests/test_gpu_utils.py drives the
    same loop there with agents of this shape."""
    import runpy
    import sys
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.buffer", "utils.model", "utils.normalization", "utils.runner")}
    try:
        import gymrl_b200.utils as U
        U.install()
        ns = runpy.run_path(str(ref_loader.REFERENCE_ROOT / script), run_name="legacy_under_test")
        cfg = ns["Config"]()
        assert cfg.env_name in ("LunarLander-v3", "CartPole-v1") and hasattr(cfg, "batch_size") and hasattr(cfg, "gamma")
        agent = ns[agent_cls]
        for m in ("choose_action", "evaluate", "update", "save_model", "load_model"):
            assert callable(getattr(agent, m)), m
        assert ns["BenchMark"].train.__module__ == "gymrl_b200.utils.runner"
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
