"""NumPy restatement of gymnasium's CartPole-v1 and Pendulum-v1, vectorised over N lockstep copies.

TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED: the reference (Starlight0798/gymRL) only calls gym.make("CartPole-v1") /
gym.make("Pendulum-v1") (algorithms/dqn_cartpole.py:94,174,180; rainbow_dqn_cartpole.py:270,367,373;
sac_pendulum.py:154,273,280; td3_pendulum.py:124,234,241); the equations live in third-party
`gymnasium` (un-pinned in requirements.txt:5, absent from /root/reference and from this image) and the
reference holds no tests or golden vectors for them.  Restated from gymnasium >= 1.0's
classic_control/cartpole.py (CartPoleEnv.step/reset, euler integrator, TimeLimit 500) and pendulum.py
(PendulumEnv.step/reset, TimeLimit 200): float64 state, float32 observations.  np_random (PCG64) is
replaced by the Philox streams of oracle/philox.py so that the device env can be compared draw for draw.
Semantics of the vector API mirror include/gymrl.h: auto-reset on done, `next_obs` = true post-step obs.
"""
import numpy as np

from . import philox as px


class _VecBase:
    max_steps = 0

    def __init__(self, num_envs, seed=0, first_env_id=0):
        self.n = int(num_envs)
        self.seed = int(seed)
        self.ids = np.arange(self.n, dtype=np.uint64) + np.uint64(first_env_id)
        self.episode = np.zeros(self.n, dtype=np.uint32)
        self.stepctr = np.zeros(self.n, dtype=np.uint32)
        self.elapsed = np.zeros(self.n, dtype=np.int32)
        self.ep_return = np.zeros(self.n, dtype=np.float64)
        self.finished_returns = []
        self.finished_lengths = []

    def _finish(self, done, ret, el):
        idx = np.nonzero(done)[0]
        self.finished_returns.extend(np.asarray(ret, dtype=np.float64)[idx].astype(np.float32).tolist())
        self.finished_lengths.extend(np.asarray(el)[idx].tolist())


class CartPoleVec(_VecBase):
    obs_dim, n_actions, max_steps = 4, 2, 500

    def __init__(self, num_envs, seed=0, first_env_id=0):
        super().__init__(num_envs, seed, first_env_id)
        self.state = np.zeros((self.n, 4), dtype=np.float64)

    def _draw_reset(self, idx):
        r0 = px.philox(self.seed, self.ids[idx], self.episode[idx].astype(np.uint64) * 8 + 0, px.STREAM_ENV_RESET)
        r1 = px.philox(self.seed, self.ids[idx], self.episode[idx].astype(np.uint64) * 8 + 1, px.STREAM_ENV_RESET)
        u = np.stack([px.u01_f64(r0[:, 0], r0[:, 1]), px.u01_f64(r0[:, 2], r0[:, 3]),
                      px.u01_f64(r1[:, 0], r1[:, 1]), px.u01_f64(r1[:, 2], r1[:, 3])], axis=1)
        return -0.05 + (0.05 - -0.05) * u

    def reset(self, mask=None):
        idx = np.arange(self.n) if mask is None else np.nonzero(mask)[0]
        self.state[idx] = self._draw_reset(idx)
        self.episode[idx] += 1
        self.elapsed[idx] = 0
        self.ep_return[idx] = 0.0
        return self.state.astype(np.float32)

    def step(self, action):
        gravity, masscart, masspole, length, force_mag, tau = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
        total_mass = masspole + masscart
        polemass_length = masspole * length
        theta_thr = 12 * 2 * np.pi / 360
        x_thr = 2.4
        x, x_dot, theta, theta_dot = (self.state[:, k].copy() for k in range(4))
        force = np.where(np.asarray(action) == 1, force_mag, -force_mag)
        costheta, sintheta = np.cos(theta), np.sin(theta)
        temp = (force + polemass_length * np.square(theta_dot) * sintheta) / total_mass
        thetaacc = (gravity * sintheta - costheta * temp) / (length * (4.0 / 3.0 - masspole * np.square(costheta) / total_mass))
        xacc = temp - polemass_length * thetaacc * costheta / total_mass
        x = x + tau * x_dot
        x_dot = x_dot + tau * xacc
        theta = theta + tau * theta_dot
        theta_dot = theta_dot + tau * thetaacc
        self.state = np.stack([x, x_dot, theta, theta_dot], axis=1)
        terminated = (x < -x_thr) | (x > x_thr) | (theta < -theta_thr) | (theta > theta_thr)
        self.elapsed += 1
        truncated = self.elapsed >= self.max_steps
        self.ep_return += 1.0
        self.stepctr += 1
        next_obs = self.state.astype(np.float32)
        reward = np.ones(self.n, dtype=np.float32)
        done = terminated | truncated
        self._finish(done, self.ep_return, self.elapsed)
        if done.any():
            self.reset(done)
        return self.state.astype(np.float32), next_obs, reward, terminated.astype(np.uint8), truncated.astype(np.uint8)

    def get_state(self):
        return np.concatenate([self.state, self.elapsed[:, None], self.episode[:, None], self.stepctr[:, None],
                               self.ep_return[:, None]], axis=1).astype(np.float64)


class PendulumVec(_VecBase):
    obs_dim, act_dim, max_steps, action_bound = 3, 1, 200, 2.0

    def __init__(self, num_envs, seed=0, first_env_id=0):
        super().__init__(num_envs, seed, first_env_id)
        self.state = np.zeros((self.n, 2), dtype=np.float64)

    def _obs(self):
        th, thdot = self.state[:, 0], self.state[:, 1]
        return np.stack([np.cos(th), np.sin(th), thdot], axis=1).astype(np.float32)

    def reset(self, mask=None):
        idx = np.arange(self.n) if mask is None else np.nonzero(mask)[0]
        r0 = px.philox(self.seed, self.ids[idx], self.episode[idx].astype(np.uint64) * 8 + 0, px.STREAM_ENV_RESET)
        pi = 3.141592653589793
        self.state[idx, 0] = -pi + (pi - -pi) * px.u01_f64(r0[:, 0], r0[:, 1])
        self.state[idx, 1] = -1.0 + (1.0 - -1.0) * px.u01_f64(r0[:, 2], r0[:, 3])
        self.episode[idx] += 1
        self.elapsed[idx] = 0
        self.ep_return[idx] = 0.0
        return self._obs()

    def step(self, action):
        max_speed, dt, g, m, l, pi = 8.0, 0.05, 10.0, 1.0, 1.0, 3.141592653589793
        th, thdot = self.state[:, 0].copy(), self.state[:, 1].copy()
        u = np.clip(np.asarray(action, dtype=np.float32).reshape(self.n), np.float32(-2.0), np.float32(2.0)).astype(np.float64)
        an = np.mod(th + pi, 2 * pi) - pi
        costs = an * an + 0.1 * (thdot * thdot) + 0.001 * (u * u)
        newthdot = thdot + (3 * g / (2 * l) * np.sin(th) + 3.0 / (m * (l * l)) * u) * dt
        newthdot = np.clip(newthdot, -max_speed, max_speed)
        newth = th + newthdot * dt
        self.state = np.stack([newth, newthdot], axis=1)
        next_obs = self._obs()
        self.elapsed += 1
        truncated = self.elapsed >= self.max_steps
        terminated = np.zeros(self.n, dtype=bool)
        self.ep_return += -costs
        self.stepctr += 1
        reward = (-costs).astype(np.float32)
        self._finish(truncated, self.ep_return, self.elapsed)
        if truncated.any():
            self.reset(truncated)
        return self._obs(), next_obs, reward, terminated.astype(np.uint8), truncated.astype(np.uint8)

    def get_state(self):
        return np.concatenate([self.state, self.elapsed[:, None], self.episode[:, None], self.stepctr[:, None],
                               self.ep_return[:, None]], axis=1).astype(np.float64)
