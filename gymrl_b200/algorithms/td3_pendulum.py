"""TD3 for Pendulum-v1 on the B200 engine — same surface as the reference ``algorithms/td3_pendulum.py``
(Config, Actor, Critic, ReplayBuffer, TD3Trainer.train/eval/test/update/select_action/soft_update).

    select_action (ref :157-170) -> actor GEMMs + tanh*bound + Gaussian exploration noise / clip kernel
    update        (ref :172-228) -> target smoothing noise kernel, twin target critic, y kernel, critic fwd/loss/bwd/Adam;
                                    every policy_freq-th update: -mean Q1(s, pi(s)) through the critic's input gradient,
                                    tanh*bound backward kernel, actor bwd/Adam, Polyak on both flat buffers.
DDPG (algorithms/ddpg_pendulum.py:154-195) is this trainer with policy_freq = 1, policy_noise = 0 and the Q1 head only.
"""
from __future__ import annotations

import time
from collections import deque

import numpy as np
import torch
import torch.nn as nn

from .. import _ffi, ops, ops_offpolicy as off
from ..graphs import LockstepGraphs
from ..mlp import Chain
from ..nn import FlatParams, FusedAdam
from .sac_pendulum import Critic, ReplayBuffer  # identical twin-critic module and replay ring

f32, i32, u8, f64 = torch.float32, torch.int32, torch.uint8, torch.float64
RELU, NONE = _ffi.ACT_RELU, _ffi.ACT_NONE


class Config:
    def __init__(self):
        self.env_name = "Pendulum-v1"
        self.seed = None
        self.max_episodes = 500
        self.max_steps = 200
        self.batch_size = 128
        self.gamma = 0.99
        self.lr_actor = 1e-3
        self.lr_critic = 1e-3
        self.tau = 0.005
        self.policy_noise = 0.2
        self.noise_clip = 0.5
        self.exploration_noise = 0.1
        self.policy_freq = 2
        self.memory_capacity = 100000
        self.hidden_dim = 256
        self.device = "cuda"
        # ---- engine extras ----
        self.num_envs = 1
        self.max_locksteps = None
        self.use_cuda_graph = True   # train(): one captured graph per lockstep phase (critic-only / critic + actor)


class Actor(nn.Module):
    def __init__(self, state_dim, action_dim, hidden_dim, action_bound):
        super().__init__()
        self.action_bound = action_bound
        self.fc1 = nn.Linear(state_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.fc3 = nn.Linear(hidden_dim, action_dim)

    SPECS = [("fc1.weight", "fc1.bias", RELU), ("fc2.weight", "fc2.bias", RELU), ("fc3.weight", "fc3.bias", NONE)]


class TD3Trainer(LockstepGraphs):
    def __init__(self, config: Config):
        _ffi.require_cuda()
        self.cfg = cfg = config
        self.device = dev = torch.device("cuda", torch.cuda.current_device())
        self.N = N = int(cfg.num_envs)
        self.seed = int(cfg.seed) if cfg.seed is not None else int(time.time_ns() & 0x7FFFFFFF)
        self.env = ops.VecEnv(cfg.env_name, N, seed=self.seed)
        D, A, H, B = self.env.obs_dim, self.env.act_dim, cfg.hidden_dim, int(cfg.batch_size)
        self.D, self.A, self.B = D, A, B
        self.action_bound = float(self.env.action_bound)
        self.actor = Actor(D, A, H, self.action_bound).to(dev)
        self.actor_target = Actor(D, A, H, self.action_bound).to(dev)
        self.actor_target.load_state_dict(self.actor.state_dict())
        self.critic = Critic(D, A, H).to(dev)
        self.critic_target = Critic(D, A, H).to(dev)
        self.critic_target.load_state_dict(self.critic.state_dict())
        self.fp_a, self.fp_at = FlatParams(self.actor, device=dev), FlatParams(self.actor_target, device=dev)
        self.fp_c, self.fp_ct = FlatParams(self.critic, device=dev), FlatParams(self.critic_target, device=dev)
        self.actor_optimizer = FusedAdam(self.fp_a, lr=cfg.lr_actor)
        self.critic_optimizer = FusedAdam(self.fp_c, lr=cfg.lr_critic)
        self.pi_act = Chain.from_names(self.fp_a, Actor.SPECS, N, False)
        self.pi_upd = Chain.from_names(self.fp_a, Actor.SPECS, B, True)
        self.pi_tgt = Chain.from_names(self.fp_at, Actor.SPECS, B, False)
        self.q1 = Chain.from_names(self.fp_c, Critic.Q1, B, True)
        self.q2 = Chain.from_names(self.fp_c, Critic.Q2, B, True)
        self.q1t = Chain.from_names(self.fp_ct, Critic.Q1, B, False)
        self.q2t = Chain.from_names(self.fp_ct, Critic.Q2, B, False)
        from ..graphs import Branches
        self.branches = Branches(1)
        self.memory = ReplayBuffer(cfg.memory_capacity, D, A, dev)
        z = lambda *s, dt=f32: torch.zeros(*s, device=dev, dtype=dt)
        self.idx = z(B, dt=i32)
        self.sa, self.sa2 = z(B, D + A), z(B, D + A)
        self.mu_b, self.act_b, self.noise_b = z(B, A), z(B, A), z(B, A)
        self.y = z(B)
        self.closs, self.aloss = z(2), z(1)
        self.mu_n, self.action = z(N, A), z(N, A)
        self.done = z(N, dt=u8)
        self.total_updates = 0
        self.act_count = 0
        # device mirrors of act_count / total_updates: the RNG kernels add them to their `draw`, so captured graphs draw afresh
        self.ctr_act, self.ctr_upd = z(1, dt=i32), z(1, dt=i32)
        self.cur = z(N, D)
        self.graph_launches = 0
        self.episode_rewards = deque(maxlen=100)
        print(f"Device: {dev}")
        print(f"State dim: {D}, Action dim: {A}")
        print(f"Action bound: [-{self.action_bound}, {self.action_bound}]")

    def soft_update(self, target=None, source=None):
        ops.polyak(self.fp_at.flat, self.fp_a.flat, self.cfg.tau)
        ops.polyak(self.fp_ct.flat, self.fp_c.flat, self.cfg.tau)

    def act(self, obs: torch.Tensor, deterministic: bool = False, noise: torch.Tensor = None) -> torch.Tensor:
        z = self.pi_act.forward(obs, self.N)
        off.tanh_bound(z, self.action_bound, out=self.mu_n)
        if deterministic:
            return self.mu_n
        self.act_count += 1
        a = ops.add_gaussian_noise_clip(self.mu_n, self.cfg.exploration_noise * self.action_bound, self.action_bound, 0.0, noise,
                                        seed=self.seed, first_id=0, draw=1, draw_base=self.ctr_act, action=self.action)
        ops.counter_add(self.ctr_act, 1)
        return a

    @torch.no_grad()
    def select_action(self, state: np.ndarray, deterministic: bool = False) -> np.ndarray:
        obs = torch.as_tensor(np.asarray(state, np.float32), device=self.device).reshape(1, -1)
        chain = getattr(self, "_pi_one", None) or Chain.from_names(self.fp_a, Actor.SPECS, 1, False)
        self._pi_one = chain
        mu = off.tanh_bound(chain.forward(obs, 1), self.action_bound)
        if not deterministic:
            self.act_count += 1
            mu = ops.add_gaussian_noise_clip(mu, self.cfg.exploration_noise * self.action_bound, self.action_bound, 0.0,
                                             seed=self.seed, first_id=1 << 40, draw=self.act_count)
        return mu[0].cpu().numpy()

    def update(self, idx: torch.Tensor = None, noise: torch.Tensor = None):
        cfg, B, A, D, mem = self.cfg, self.B, self.A, self.D, self.memory
        if len(mem) < B:
            return 0.0, 0.0
        self.total_updates += 1
        u = self.total_updates
        if idx is None:
            idx = mem.sample_indices(B, seed=self.seed, draw=1, draw_base=self.ctr_upd, out=self.idx)
        bound = self.action_bound
        # ---- target with smoothing noise (ref :193-204) ----
        off.tanh_bound(self.pi_tgt.forward(mem.next_obs, B, row_index=idx), bound, out=self.mu_b)
        nz = noise if noise is not None else off.fill_normal(self.noise_b, seed=self.seed, entity0=0, draw=1, draw_base=self.ctr_upd)
        ops.add_gaussian_noise_clip(self.mu_b, cfg.policy_noise, bound, cfg.noise_clip, nz, action=self.act_b)
        off.gather_concat(mem.next_obs, idx, self.act_b, None, out=self.sa2, n=B)
        q1t, q2t = self.branches.run(lambda: self.q1t.forward(self.sa2, B), lambda: self.q2t.forward(self.sa2, B))   # twins in parallel
        off.twin_q_target(mem.reward, mem.done, q1t, q2t, cfg.gamma, row_index=idx, out=self.y)
        # ---- critic (ref :206-213) ----
        off.gather_concat(mem.obs, idx, mem.action, idx, out=self.sa, n=B)
        q1, q2 = self.branches.run(lambda: self.q1.forward(self.sa, B), lambda: self.q2.forward(self.sa, B))
        self.closs.zero_()
        off.twin_q_loss(q1, q2, self.y, self.q1.dout, self.q2.dout, self.closs)
        self.branches.run(lambda: self.q1.backward(self.sa, B), lambda: self.q2.backward(self.sa, B))
        self.critic_optimizer.step()
        # ---- delayed actor + target sync (ref :215-226) ----
        if u % cfg.policy_freq == 0:
            zpre = self.pi_upd.forward(mem.obs, B, row_index=idx)
            off.tanh_bound(zpre, bound, out=self.act_b)
            off.gather_concat(mem.obs, idx, self.act_b, None, out=self.sa2, n=B)
            q1 = self.q1.forward(self.sa2, B)
            self.aloss.zero_()
            off.min_q_grad(q1, None, self.q1.dout, None, q1_only=True, acc=self.aloss)
            dx = self.q1.backward(self.sa2, B, param_grads=False, input_grad=True)
            off.tanh_bound_grad(self.act_b, dx[:, D:], self.pi_upd.dout, bound)
            self.pi_upd.backward(mem.obs, B, row_index=idx)
            self.actor_optimizer.step()
            self.soft_update()
        ops.counter_add(self.ctr_upd, 1)
        return self.aloss, self.closs

    # ---------------------------------------------------------------- one lockstep (ref train() loop body :239-254)
    def _lockstep_phases(self) -> int:
        return int(self.cfg.policy_freq)

    def _lockstep_body(self):
        env, mem, cur = self.env, self.memory, self.cur
        a = self.act(cur)
        obs, r, te, tr, nobs = env.step(a, done=self.done)
        mem.store(cur, a, r, nobs, self.done)
        self.update()
        cur.copy_(obs)

    def train(self):
        print("Starting training...")
        cfg, env, mem = self.cfg, self.env, self.memory
        env.reset(out=self.cur)
        max_lock = cfg.max_locksteps or int(cfg.max_episodes * cfg.max_steps / self.N)
        t0, last_total = time.time(), 0
        for step in range(max_lock):
            self.lockstep()
            if step % cfg.max_steps == cfg.max_steps - 1:
                avg, _, total = env.episode_stats(100)
                if total != last_total:
                    last_total = total
                    self.episode_rewards.extend([avg] * min(self.N, 100))
                    sps = (step + 1) * self.N / max(time.time() - t0, 1e-9)
                    print(f"Episodes {total} | Avg(100): {avg:.1f} | Critic: {self.closs[0].item():.3f} | {sps:,.0f} steps/s")
                    if avg >= -200.0 and total >= 100:
                        print(f"\nEnvironment solved in {total} episodes!")
                        break
        print("Training completed!")

    def eval(self, num_episodes: int = 10):
        print(f"\nEvaluating for {num_episodes} episodes...")
        env = ops.VecEnv(self.cfg.env_name, num_episodes, seed=self.seed + 999, first_env_id=1 << 32)
        chain = Chain.from_names(self.fp_a, Actor.SPECS, num_episodes, False)
        obs = env.reset()
        ret = torch.zeros(num_episodes, device=self.device, dtype=f64)
        for _ in range(env.max_episode_steps):
            a = off.tanh_bound(chain.forward(obs, num_episodes), self.action_bound)
            obs, r, te, tr, _ = env.step(a, want_next_obs=False)
            ret += r.double()
        rewards = ret.tolist()
        for i, r in enumerate(rewards):
            print(f"  Episode {i + 1}: Reward = {r:.1f}")
        print(f"Evaluation Results: Mean = {np.mean(rewards):.1f} +/- {np.std(rewards):.1f}")
        env.close()
        return rewards

    def test(self):
        self.eval(num_episodes=5)
        print("\n(visual test skipped: the device env has no renderer)")


if __name__ == "__main__":
    config = Config()
    config.num_envs = 4096
    config.batch_size = 4096
    config.memory_capacity = 1 << 20
    trainer = TD3Trainer(config)
    trainer.train()
    trainer.test()
