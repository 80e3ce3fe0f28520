"""A single-env, Gymnasium-shaped view of the device env (N = 1), so code written against ``gym.make(...)`` — the
reference's utils.runner loop and its unmodified algorithms/*.py scripts — can step the CUDA env.

API subset (complete for the in-scope files): observation_space.shape, action_space.n | .shape | .high |
.low | .sample(), spec.max_episode_steps, reset(seed=None) -> (obs, info), step(a) -> (obs, reward, terminated,
truncated, info), close(), render(), unwrapped.  Each call is one kernel launch plus one small D2H copy — this is the
compatibility path, not the fast one (the vectorised trainers never leave the device).
"""
from __future__ import annotations

import types

import numpy as np
import torch

from .. import ops


class _Discrete:
    def __init__(self, n, rng):
        self.n, self._rng, self.shape = int(n), rng, ()

    def sample(self):
        return int(self._rng.integers(self.n))


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, rng=None):
        self.low, self.high = np.asarray(low, dtype), np.asarray(high, dtype)
        self.shape = tuple(shape) if shape is not None else self.low.shape
        self.dtype, self._rng = dtype, rng or np.random.default_rng()

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)


class DeviceGymEnv:
    def __init__(self, env_name: str, render_mode=None, seed=None, **kwargs):
        self.env_name, self.render_mode = env_name, render_mode
        self._seed = int(seed) if seed is not None else int(np.random.randint(1, 2 ** 31 - 1))
        self._make()
        v = self._vec
        rng = np.random.default_rng(self._seed)
        hi = np.full(v.obs_dim, np.inf, np.float32)
        self.observation_space = Box(-hi, hi, (v.obs_dim,))
        if v.discrete:
            self.action_space = _Discrete(v.n_actions, rng)
        else:
            b = np.full(v.act_dim, v.action_bound, np.float32)
            self.action_space = Box(-b, b, (v.act_dim,), rng=rng)
        self.spec = types.SimpleNamespace(max_episode_steps=v.max_episode_steps, id=env_name)
        self.unwrapped = self
        self._next_first_obs = None

    def _make(self):
        self._vec = ops.VecEnv(self.env_name, 1, seed=self._seed)
        dev = self._vec.device
        self._obs = torch.empty(1, self._vec.obs_dim, device=dev)
        self._act = torch.zeros(1, device=dev, dtype=torch.int32) if self._vec.discrete else torch.zeros(1, self._vec.act_dim, device=dev)

    def reset(self, seed=None, options=None):
        if seed is not None:                       # gymnasium: a seeded reset restarts the env's generator
            self._seed = int(seed)
            self._vec.close()
            self._make()
            self._next_first_obs = None
        if self._next_first_obs is not None:       # the device env auto-reset at the previous episode's end
            obs, self._next_first_obs = self._next_first_obs, None
            return obs, {}
        self._vec.reset(out=self._obs)
        return self._obs[0].cpu().numpy(), {}

    def step(self, action):
        if self._vec.discrete:
            self._act.fill_(int(action))
        else:
            self._act.copy_(torch.as_tensor(np.asarray(action, np.float32).reshape(1, -1)))
        _, r, te, tr, nxt = self._vec.step(self._act, obs=self._obs, want_next_obs=True)
        host = torch.cat([nxt.reshape(-1), r.reshape(-1), te.reshape(-1).float(), tr.reshape(-1).float()]).cpu().numpy()
        D = self._vec.obs_dim
        obs, reward, terminated, truncated = host[:D].astype(np.float32), float(host[D]), bool(host[D + 1]), bool(host[D + 2])
        if terminated or truncated:
            self._next_first_obs = self._obs[0].cpu().numpy()
        return obs, reward, terminated, truncated, {}

    def render(self):
        return None

    def close(self):
        self._vec.close()


def make(env_name: str, render_mode=None, **kwargs) -> DeviceGymEnv:
    return DeviceGymEnv(env_name, render_mode=render_mode, **kwargs)
