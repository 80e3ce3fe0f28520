"""The shipped CUDA LunarLander solver SOURCE, compiled for the host, against the C oracle — no GPU needed.

tests/test_gpu_envs.py proves kernel == oracle on a B200; this file proves the same statements (gymrl_b200/csrc/env_lunar.cu,
built with -DGYMRL_HOSTSIM by tests/hostsim.py) == oracle/lunar_lander.c on the CPU box, so a change to the solver loops is
checked before any GPU time is spent.  Bit-exact: observations, rewards, flags and the complete 128-double state snapshot.
"""
import numpy as np
import pytest

import hostsim  # tests/hostsim.py (pytest puts tests/ on sys.path: rootdir conftest, no package)

pytestmark = pytest.mark.skipif(hostsim.nvcc() is None, reason="nvcc not available")


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


def _policy(N, rng):
    def act(obs):  # one third heuristic (lands, sleeps -> +100), one third random, one third no-op (crashes)
        return np.where(np.arange(N) % 3 == 0, _heuristic(obs), np.where(np.arange(N) % 3 == 1, rng.integers(0, 4, N), 0))
    return act


@pytest.mark.parametrize("variant", sorted(hostsim.BUILDS))
def test_hostsim_free_running_bit_exact(variant):
    """Both sides run their own trajectory from reset for 400 steps (auto-reset included): everything stays identical."""
    from oracle.lunar import LunarLanderVec
    N = 48
    ora = LunarLanderVec(N, seed=5, first_env_id=1000)
    sim = hostsim.HostSimLunarVec(N, seed=5, first_env_id=1000, variant=variant)
    o0, s0 = ora.reset(), sim.reset()
    assert np.array_equal(o0, s0)
    assert np.array_equal(ora.get_state(), sim.get_state())
    act = _policy(N, np.random.default_rng(2))
    finished, contact_steps, two_leg = 0, 0, 0
    for t in range(400):
        a = act(ora.obs)
        ro, rs = ora.step(a), sim.step(a)
        for name, x, y in zip(("obs", "next_obs", "reward", "terminated", "truncated"), ro, rs):
            assert np.array_equal(x, y), f"step {t}: {name} differs at envs {np.argwhere(x != y)[:4].ravel()}"
        finished += int((ro[3] | ro[4]).sum())
        contact_steps += int((sim.prof[:, 4] > 0).sum())
        lay = sim.prof[:, 8]
        two_leg += int((((lay & 15) > 0) & (((lay >> 4) & 15) > 0)).sum())
    so, ss = ora.get_state(), sim.get_state()
    assert np.array_equal(so, ss), f"state planes differ at {np.argwhere(so != ss)[:5]}"
    assert finished >= N // 2 and contact_steps > 500 and two_leg > 100   # crash / landing / both-leg contact paths all exercised
    assert int(sim.prof[:, 7].sum()) == 0                                 # no dropped manifolds


@pytest.mark.parametrize("variant", sorted(hostsim.BUILDS))
def test_hostsim_teacher_forced_single_steps(variant):
    """Oracle state in, one step, compare: isolates the single-step arithmetic (contact-heavy states included)."""
    from oracle.lunar import LunarLanderVec
    N = 32
    ora = LunarLanderVec(N, seed=9)
    sim = hostsim.HostSimLunarVec(N, seed=9, variant=variant)
    ora.reset()
    rng = np.random.default_rng(3)
    for t in range(300):
        a = np.where(np.arange(N) % 2 == 0, _heuristic(ora.obs), rng.integers(0, 4, N))
        sim.set_state(ora.get_state())
        ro, rs = ora.step(a), sim.step(a)
        for name, x, y in zip(("obs", "next_obs", "reward", "terminated", "truncated"), ro, rs):
            assert np.array_equal(x, y), f"step {t}: {name} differs"
        assert np.array_equal(ora.get_state(), sim.get_state()), f"step {t}: state differs"


def test_hostsim_is_not_in_the_product():
    """The host build lives under tests/_build: the package never loads it, the header does not declare its entry points and
    the product library does not export them."""
    import pathlib
    import subprocess
    root = pathlib.Path(__file__).resolve().parent.parent
    for p in (root / "gymrl_b200").rglob("*.py"):
        t = p.read_text()
        assert "gymrl_hostsim" not in t and "import hostsim" not in t and "libhostsim" not in t, p
    assert "gymrl_hostsim" not in (root / "include" / "gymrl.h").read_text()
    lib = root / "gymrl_b200" / "lib" / "libgymrl_b200.so"
    if lib.exists():
        syms = subprocess.run(["nm", "-D", "--defined-only", str(lib)], capture_output=True, text=True).stdout
        assert "hostsim" not in syms
