"""oracle/gymnasium_shim — the subset of the `gymnasium` API the in-scope reference scripts touch, over this repo's
CPU env restatements.  TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).

Why it exists: gymnasium and box2d-py are third-party, un-pinned
(reference requirements.txt:3-6), absent from /root/reference, from this image and from the GPU box, so the
UNMODIFIED reference files (algorithms/dqn_cartpole.py, ppo_lunarlander.py, ... and utils/runner.py) cannot import.
`install()` registers this package as `sys.modules["gymnasium"]`; the scripts then run exactly as written:

    gym.make(name, render_mode=None, **kw)                (ref dqn_cartpole.py:94, ppo_lunarlander.py:160, utils/runner.py:53)
    env.observation_space.shape / env.action_space.{n, shape, high, low, sample()}
    env.spec.max_episode_steps                            (ref rainbow_dqn_cartpole.py:273, utils/runner.py:77)
    env.reset(seed=None) -> (obs, info);  env.step(a) -> (obs, reward, terminated, truncated, info)
    env.close(), env.render(), env.unwrapped
    gym.spaces.{Box, Discrete}, gym.ObservationWrapper, gym.Wrapper, gymnasium.wrappers.AtariPreprocessing (import only)

PARITY UNPINNED, like the vector oracles: the arithmetic restates gymnasium >= 1.0's cartpole.py / pendulum.py and (through
oracle/lunar_lander.c) lunar_lander.py + Box2D 2.3; tests/test_reference_scripts.py checks the scalar envs here against
oracle/envs_np.py step for step.  Single-env semantics are gymnasium's: NO auto-reset, TimeLimit sets `truncated`,
`np_random` is a PCG64 Generator re-seeded by reset(seed=...), `action_space.sample()` draws from the space's own Generator.
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np

__version__ = "1.0.0+gymrl_b200.shim"
_gymrl_stub = True     # oracle/ref_loader.install_gymnasium_stub() leaves an installed shim alone


# ----------------------------------------------------------------------------------------------- spaces
class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self.shape, self.dtype = shape, dtype
        self._np_random = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._np_random = np.random.default_rng(seed)
        return [seed]


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        low_a, high_a = np.asarray(low), np.asarray(high)
        if shape is None:
            shape = low_a.shape if low_a.shape else high_a.shape
        shape = tuple(shape)
        super().__init__(shape, np.dtype(dtype), seed)
        self.low = np.broadcast_to(low_a, shape).astype(dtype).copy()
        self.high = np.broadcast_to(high_a, shape).astype(dtype).copy()

    def sample(self):
        finite = np.isfinite(self.low) & np.isfinite(self.high)
        out = self._np_random.normal(size=self.shape)
        u = self._np_random.uniform(np.where(finite, self.low, 0.0), np.where(finite, self.high, 1.0), size=self.shape)
        return np.where(finite, u, out).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        super().__init__((), np.dtype(np.int64), seed)
        self.n, self.start = int(n), int(start)

    def sample(self):
        return int(self.start + self._np_random.integers(self.n))

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class EnvSpec:
    def __init__(self, id, max_episode_steps):
        self.id, self.max_episode_steps = id, max_episode_steps


# ----------------------------------------------------------------------------------------------- env base
class Env:
    metadata = {"render_modes": []}
    observation_space: Space
    action_space: Space
    spec: EnvSpec = None
    render_mode = None

    def __init__(self):
        self.np_random = np.random.default_rng()
        self._elapsed = 0

    @property
    def unwrapped(self):
        return self

    def _seed(self, seed):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)

    def render(self):
        return None

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(action)


class ObservationWrapper(Wrapper):
    def reset(self, **kw):
        obs, info = self.env.reset(**kw)
        return self.observation(obs), info

    def step(self, action):
        obs, r, te, tr, info = self.env.step(action)
        return self.observation(obs), r, te, tr, info

    def observation(self, obs):
        raise NotImplementedError


# ----------------------------------------------------------------------------------------------- CartPole-v1
class CartPoleEnv(Env):
    """gymnasium classic_control/cartpole.py CartPoleEnv under TimeLimit(500): float64 state, euler integrator."""
    gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    theta_threshold_radians = 12 * 2 * math.pi / 360
    x_threshold = 2.4

    def __init__(self, render_mode=None, **kw):
        super().__init__()
        self.render_mode = render_mode
        high = np.array([self.x_threshold * 2, np.inf, self.theta_threshold_radians * 2, np.inf], dtype=np.float32)
        self.observation_space = Box(-high, high, dtype=np.float32)
        self.action_space = Discrete(2)
        self.spec = EnvSpec("CartPole-v1", 500)
        self.state = None

    def reset(self, seed=None, options=None):
        self._seed(seed)
        self.state = self.np_random.uniform(low=-0.05, high=0.05, size=(4,))
        self._elapsed = 0
        return np.array(self.state, dtype=np.float32), {}

    def step(self, action):
        x, x_dot, theta, theta_dot = self.state
        total_mass = self.masspole + self.masscart
        polemass_length = self.masspole * self.length
        force = self.force_mag if action == 1 else -self.force_mag
        costheta, sintheta = np.cos(theta), np.sin(theta)
        temp = (force + polemass_length * np.square(theta_dot) * sintheta) / total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * np.square(costheta) / total_mass))
        xacc = temp - polemass_length * thetaacc * costheta / total_mass
        x = x + self.tau * x_dot
        x_dot = x_dot + self.tau * xacc
        theta = theta + self.tau * theta_dot
        theta_dot = theta_dot + self.tau * thetaacc
        self.state = np.array((x, x_dot, theta, theta_dot), dtype=np.float64)
        terminated = bool(x < -self.x_threshold or x > self.x_threshold
                          or theta < -self.theta_threshold_radians or theta > self.theta_threshold_radians)
        self._elapsed += 1
        truncated = self._elapsed >= self.spec.max_episode_steps
        return np.array(self.state, dtype=np.float32), 1.0, terminated, truncated, {}


# ----------------------------------------------------------------------------------------------- Pendulum-v1
class PendulumEnv(Env):
    """gymnasium classic_control/pendulum.py PendulumEnv under TimeLimit(200)."""
    max_speed, max_torque, dt, g, m, l = 8.0, 2.0, 0.05, 10.0, 1.0, 1.0

    def __init__(self, render_mode=None, g=10.0, **kw):
        super().__init__()
        self.render_mode, self.g = render_mode, g
        high = np.array([1.0, 1.0, self.max_speed], dtype=np.float32)
        self.action_space = Box(-self.max_torque, self.max_torque, shape=(1,), dtype=np.float32)
        self.observation_space = Box(-high, high, dtype=np.float32)
        self.spec = EnvSpec("Pendulum-v1", 200)
        self.state = None

    def _get_obs(self):
        theta, thetadot = self.state
        return np.array([np.cos(theta), np.sin(theta), thetadot], dtype=np.float32)

    def reset(self, seed=None, options=None):
        self._seed(seed)
        high = np.array([np.pi, 1.0])
        self.state = self.np_random.uniform(low=-high, high=high)
        self._elapsed = 0
        return self._get_obs(), {}

    def step(self, u):
        th, thdot = self.state
        u = np.clip(np.asarray(u, dtype=np.float32).reshape(-1), -self.max_torque, self.max_torque)[0]
        u = float(u)
        an = ((th + np.pi) % (2 * np.pi)) - np.pi
        costs = an ** 2 + 0.1 * thdot ** 2 + 0.001 * (u ** 2)
        newthdot = thdot + (3 * self.g / (2 * self.l) * np.sin(th) + 3.0 / (self.m * self.l ** 2) * u) * self.dt
        newthdot = np.clip(newthdot, -self.max_speed, self.max_speed)
        newth = th + newthdot * self.dt
        self.state = np.array([newth, newthdot])
        self._elapsed += 1
        truncated = self._elapsed >= self.spec.max_episode_steps
        return self._get_obs(), -costs, False, truncated, {}


# ----------------------------------------------------------------------------------------------- LunarLander-v3
class LunarLanderEnv(Env):
    """LunarLander-v3 under TimeLimit(1000): one LLEnv of oracle/lunar_lander.c (gymnasium lunar_lander.py + Box2D 2.3
    restated).  np_random is replaced by the C file's Philox stream keyed by (seed, episode counter); reset(seed=s)
    restarts that stream, so — as with gymnasium — a fixed seed gives the same terrain / initial push every episode."""

    def __init__(self, render_mode=None, continuous=False, **kw):
        super().__init__()
        if continuous:
            raise NotImplementedError("shim LunarLander-v3: discrete actions only (all the in-scope scripts use)")
        from .. import lunar
        self._L = lunar.lib()
        import ctypes as C
        self._C = C
        self._L.ll_step_single.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        self.render_mode = render_mode
        low = np.array([-2.5, -2.5, -10.0, -10.0, -2 * math.pi, -10.0, -0.0, -0.0], dtype=np.float32)
        high = np.array([2.5, 2.5, 10.0, 10.0, 2 * math.pi, 10.0, 1.0, 1.0], dtype=np.float32)
        self.observation_space = Box(low, high, dtype=np.float32)
        self.action_space = Discrete(4)
        self.spec = EnvSpec("LunarLander-v3", 1000)
        self._seed_val = int(np.random.SeedSequence().entropy & 0x7FFFFFFF)
        self._h = self._L.ll_create(self._seed_val, 0)
        self._obs = np.zeros(8, np.float32)
        self._r = np.zeros(1, np.float32)
        self._te = np.zeros(1, np.uint8)
        self._tr = np.zeros(1, np.uint8)

    def reset(self, seed=None, options=None):
        if seed is not None:
            self._L.ll_destroy(self._h)
            self._seed_val = int(seed)
            self._h = self._L.ll_create(self._seed_val, 0)
        self._L.ll_reset(self._h, self._obs.ctypes.data)
        return self._obs.copy(), {}

    def step(self, action):
        self._L.ll_step_single(self._h, int(action), self._obs.ctypes.data, self._r.ctypes.data, self._te.ctypes.data,
                               self._tr.ctypes.data)
        return self._obs.copy(), float(self._r[0]), bool(self._te[0]), bool(self._tr[0]), {}

    def close(self):
        if getattr(self, "_h", None):
            self._L.ll_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_REGISTRY = {"CartPole-v1": CartPoleEnv, "Pendulum-v1": PendulumEnv, "LunarLander-v3": LunarLanderEnv}


def make(id, render_mode=None, **kwargs):
    if id not in _REGISTRY:
        raise ValueError(f"gymnasium shim: unknown env id {id!r} (have {sorted(_REGISTRY)})")
    if render_mode == "human":
        raise RuntimeError("gymnasium shim: no renderer (render_mode='human' is the scripts' visual test, out of scope)")
    return _REGISTRY[id](render_mode=render_mode, **kwargs)


# ----------------------------------------------------------------------------------------------- sub-modules + install
spaces = types.ModuleType("gymnasium.spaces")
spaces.Space, spaces.Box, spaces.Discrete = Space, Box, Discrete

wrappers = types.ModuleType("gymnasium.wrappers")


class _AtariPreprocessing(Wrapper):      # imported by the reference's utils/runner.py:6, never constructed in scope
    def __init__(self, *a, **k):
        raise NotImplementedError("gymnasium shim: Atari is out of scope")


wrappers.AtariPreprocessing = _AtariPreprocessing


def install():
    """Register this package as `gymnasium` (idempotent).  Returns the module object."""
    me = sys.modules[__name__]
    sys.modules["gymnasium"] = me
    sys.modules["gymnasium.spaces"] = spaces
    sys.modules["gymnasium.wrappers"] = wrappers
    return me


def uninstall():
    for k in ("gymnasium", "gymnasium.spaces", "gymnasium.wrappers"):
        if getattr(sys.modules.get(k), "_gymrl_stub", False) or k != "gymnasium":
            sys.modules.pop(k, None)
