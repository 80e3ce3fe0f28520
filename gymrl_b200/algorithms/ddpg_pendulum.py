"""DDPG for Pendulum-v1 on the B200 engine — same surface as the reference ``algorithms/ddpg_pendulum.py``
(Config, Actor, Critic, ReplayBuffer, DDPGTrainer.train/eval/test/update/select_action/soft_update).

    select_action (ref :140-152) -> actor GEMMs + tanh*bound + N(0, noise_std*bound) exploration noise, clipped
    update        (ref :154-195) -> y = r + gamma (1-d) Q_t(s', pi_t(s')); critic MSE fwd/bwd/Adam; actor step through the
                                    critic's input gradient (-mean Q(s, pi(s))); Polyak on both flat buffers — every update.
It is the TD3 kernel set with a single Q head, no target smoothing and no actor delay.
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
from .sac_pendulum import ReplayBuffer
from .td3_pendulum import Actor

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
        self.noise_std = 0.1
        self.memory_capacity = 100000
        self.hidden_dim = 256
        self.device = "cuda"
        # ---- engine extras ----
        self.num_envs = 1
        self.max_locksteps = None
        self.use_cuda_graph = True   # train(): act -> env step -> store -> update as ONE captured graph per lockstep


class Critic(nn.Module):
    def __init__(self, state_dim, action_dim, hidden_dim):
        super().__init__()
        self.fc1 = nn.Linear(state_dim + action_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.fc3 = nn.Linear(hidden_dim, 1)

    Q = [("fc1.weight", "fc1.bias", RELU), ("fc2.weight", "fc2.bias", RELU), ("fc3.weight", "fc3.bias", NONE)]


class DDPGTrainer(LockstepGraphs):
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
        self.actor, self.actor_target = Actor(D, A, H, self.action_bound).to(dev), Actor(D, A, H, self.action_bound).to(dev)
        self.actor_target.load_state_dict(self.actor.state_dict())
        self.critic, self.critic_target = Critic(D, A, H).to(dev), Critic(D, A, H).to(dev)
        self.critic_target.load_state_dict(self.critic.state_dict())
        self.fp_a, self.fp_at = FlatParams(self.actor, device=dev), FlatParams(self.actor_target, device=dev)
        self.fp_c, self.fp_ct = FlatParams(self.critic, device=dev), FlatParams(self.critic_target, device=dev)
        self.actor_optimizer = FusedAdam(self.fp_a, lr=cfg.lr_actor)
        self.critic_optimizer = FusedAdam(self.fp_c, lr=cfg.lr_critic)
        self.pi_act = Chain.from_names(self.fp_a, Actor.SPECS, N, False)
        self.pi_upd = Chain.from_names(self.fp_a, Actor.SPECS, B, True)
        self.pi_tgt = Chain.from_names(self.fp_at, Actor.SPECS, B, False)
        self.q = Chain.from_names(self.fp_c, Critic.Q, B, True)
        self.qt = Chain.from_names(self.fp_ct, Critic.Q, B, False)
        self.memory = ReplayBuffer(cfg.memory_capacity, D, A, dev)
        z = lambda *s, dt=f32: torch.zeros(*s, device=dev, dtype=dt)
        self.idx = z(B, dt=i32)
        self.sa, self.sa2 = z(B, D + A), z(B, D + A)
        self.act_b, self.y, self.dq_unused = z(B, A), z(B), z(B, 1)
        self.closs, self.aloss = z(2), z(1)
        self.mu_n, self.action = z(N, A), z(N, A)
        self.done = z(N, dt=u8)
        self.total_updates = self.act_count = 0
        self.ctr_act, self.ctr_upd = z(1, dt=i32), z(1, dt=i32)   # device mirrors of act_count / total_updates (RNG draws)
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
        off.tanh_bound(self.pi_act.forward(obs, self.N), self.action_bound, out=self.mu_n)
        if deterministic:
            return self.mu_n
        self.act_count += 1
        a = ops.add_gaussian_noise_clip(self.mu_n, self.cfg.noise_std * self.action_bound, self.action_bound, 0.0, noise,
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
            mu = ops.add_gaussian_noise_clip(mu, self.cfg.noise_std * self.action_bound, self.action_bound, 0.0,
                                             seed=self.seed, first_id=1 << 40, draw=self.act_count)
        return mu[0].cpu().numpy()

    def update(self, idx: torch.Tensor = None):
        cfg, B, D, mem, bound = self.cfg, self.B, self.D, self.memory, self.action_bound
        if len(mem) < B:
            return 0.0, 0.0
        self.total_updates += 1
        if idx is None:
            idx = mem.sample_indices(B, seed=self.seed, draw=1, draw_base=self.ctr_upd, out=self.idx)
        # target (ref :176-179)
        off.tanh_bound(self.pi_tgt.forward(mem.next_obs, B, row_index=idx), bound, out=self.act_b)
        off.gather_concat(mem.next_obs, idx, self.act_b, None, out=self.sa2, n=B)
        qt = self.qt.forward(self.sa2, B)
        off.twin_q_target(mem.reward, mem.done, qt, qt, cfg.gamma, row_index=idx, out=self.y)
        # critic (ref :181-186): the twin-loss kernel with both heads on the same Q leaves d(mse)/dq in q.dout
        off.gather_concat(mem.obs, idx, mem.action, idx, out=self.sa, n=B)
        q = self.q.forward(self.sa, B)
        self.closs.zero_()
        off.twin_q_loss(q, q, self.y, self.q.dout, self.dq_unused, self.closs)
        self.q.backward(self.sa, B)
        self.critic_optimizer.step()
        # actor (ref :188-192) + target sync every update (ref :194-195)
        off.tanh_bound(self.pi_upd.forward(mem.obs, B, row_index=idx), bound, out=self.act_b)
        off.gather_concat(mem.obs, idx, self.act_b, None, out=self.sa2, n=B)
        q = self.q.forward(self.sa2, B)
        self.aloss.zero_()
        off.min_q_grad(q, None, self.q.dout, None, q1_only=True, acc=self.aloss)
        dx = self.q.backward(self.sa2, B, param_grads=False, input_grad=True)
        off.tanh_bound_grad(self.act_b, dx[:, D:], self.pi_upd.dout, bound)
        self.pi_upd.backward(mem.obs, B, row_index=idx)
        self.actor_optimizer.step()
        self.soft_update()
        ops.counter_add(self.ctr_upd, 1)
        return self.aloss, self.closs     # closs[0] = 2 x mse (both "heads" are the same Q)

    # ---------------------------------------------------------------- one lockstep (ref train() loop body)
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
                    print(f"Episodes {total} | Avg(100): {avg:.1f} | Critic: {0.5 * self.closs[0].item():.3f} | {sps:,.0f} steps/s")
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
            obs, r, te, tr, _ = env.step(off.tanh_bound(chain.forward(obs, num_episodes), self.action_bound), want_next_obs=False)
            ret += r.double()
        rewards = ret.tolist()
        print(f"Evaluation Results: Mean = {np.mean(rewards):.1f} +/- {np.std(rewards):.1f}")
        env.close()
        return rewards

    def test(self):
        self.eval(num_episodes=5)
        print("\n(visual test skipped: the device env has no renderer)")


if __name__ == "__main__":
    config = Config()
    config.num_envs, config.batch_size, config.memory_capacity = 4096, 4096, 1 << 20
    trainer = DDPGTrainer(config)
    trainer.train()
    trainer.test()
