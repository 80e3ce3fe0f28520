"""Discrete Soft Actor-Critic for CartPole-v1 on the B200 engine — same surface as the reference
``algorithms/sac_cartpole.py`` (Config, ReplayBuffer, Actor, Critic, SACTrainer.train/eval/test/update/select_action/soft_update).
SURVEY §8f rank 2.

    select_action (ref :136-147) -> actor logits -> Categorical sample (gymrl_sample_categorical; softmax(logits) is the policy)
    update        (ref :155-221) -> target  y = r + gamma (1 - d) (sum_a p' min(Q1t, Q2t) + alpha H(p'))      gymrl_sac_discrete_target
                                    critics mse(Q1(s)[a], y), mse(Q2(s)[a], y), one Adam each                   gymrl_sac_discrete_critic_loss
                                    actor   mean(-alpha H - sum_a p min(Q1, Q2)) with the UPDATED critics        gymrl_sac_discrete_actor_grad
                                    alpha   mean(exp(log_alpha) (H - target_entropy)), float32 Adam              gymrl_sac_discrete_alpha_step
                                    Polyak of both critic targets                                                gymrl_polyak
Four optimisers like the reference (actor, critic1, critic2, log_alpha); log_alpha is a float32 scalar here (ref :123-126).
Vectorisation: N envs in lockstep, one update of ``batch_size`` per lockstep, captured as one CUDA graph (`lockstep()`).
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

f32, i32, u8, f64 = torch.float32, torch.int32, torch.uint8, torch.float64
RELU, NONE = _ffi.ACT_RELU, _ffi.ACT_NONE


class Config:
    def __init__(self):
        self.env_name = "CartPole-v1"
        self.seed = None
        self.max_episodes = 500
        self.max_steps = 2000
        self.batch_size = 128
        self.gamma = 0.9
        self.tau = 0.005
        self.lr_actor = 2e-4
        self.lr_critic = 1e-3
        self.lr_alpha = 1e-3
        self.memory_capacity = 10000
        self.hidden_dim = 256
        self.target_entropy = -1.0
        self.device = "cuda"
        # ---- engine extras ----
        self.num_envs = 1
        self.max_locksteps = None
        self.use_cuda_graph = True


SPECS = [("fc1.weight", "fc1.bias", RELU), ("fc2.weight", "fc2.bias", RELU), ("fc3.weight", "fc3.bias", NONE)]


class Actor(nn.Module):
    """Parameter container with the reference's keys (ref :78-89); its softmax lives in the loss / sampling kernels."""

    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int = 256):
        super().__init__()
        self.fc1 = nn.Linear(state_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.fc3 = nn.Linear(hidden_dim, action_dim)


class Critic(nn.Module):
    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int = 256):
        super().__init__()
        self.fc1 = nn.Linear(state_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.fc3 = nn.Linear(hidden_dim, action_dim)


class ReplayBuffer(off.ReplayRing):
    def __init__(self, capacity: int, obs_dim: int = 4, device=None):
        super().__init__(capacity, obs_dim, 1, True, device or torch.device("cuda", torch.cuda.current_device()))

    def push(self, state, action, reward, next_state, done):
        dev = self.state.device
        t = lambda x, dt: torch.as_tensor(np.asarray(x), device=dev).to(dt)
        self.store(t(state, f32).reshape(1, -1), t([action], i32).reshape(1, 1), t([reward], f32), t(next_state, f32).reshape(1, -1),
                   t([bool(done)], u8))


class SACTrainer(LockstepGraphs):
    def __init__(self, config: Config):
        _ffi.require_cuda()
        self.cfg = cfg = config
        self.device = dev = torch.device("cuda", torch.cuda.current_device())
        self.N = N = int(cfg.num_envs)
        self.seed = int(cfg.seed) if cfg.seed is not None else int(time.time_ns() & 0x7FFFFFFF)
        self.env = ops.VecEnv(cfg.env_name, N, seed=self.seed)
        self.D, self.A = D, A = self.env.obs_dim, self.env.n_actions
        H, self.B = cfg.hidden_dim, int(cfg.batch_size)
        B = self.B
        self.actor = Actor(D, A, H).to(dev)
        self.critic1, self.critic2 = Critic(D, A, H).to(dev), Critic(D, A, H).to(dev)
        self.critic1_target, self.critic2_target = Critic(D, A, H).to(dev), Critic(D, A, H).to(dev)
        self.critic1_target.load_state_dict(self.critic1.state_dict())
        self.critic2_target.load_state_dict(self.critic2.state_dict())
        self.fp_a = FlatParams(self.actor, device=dev)
        self.fp_c1, self.fp_c2 = FlatParams(self.critic1, device=dev), FlatParams(self.critic2, device=dev)
        self.fp_c1t, self.fp_c2t = FlatParams(self.critic1_target, device=dev), FlatParams(self.critic2_target, device=dev)
        self.actor_optim = FusedAdam(self.fp_a, lr=cfg.lr_actor)
        self.critic1_optim = FusedAdam(self.fp_c1, lr=cfg.lr_critic)
        self.critic2_optim = FusedAdam(self.fp_c2, lr=cfg.lr_critic)
        self.log_alpha = torch.tensor([np.log(0.01)], device=dev, dtype=f32)
        self.alpha_state = torch.zeros(3, device=dev, dtype=f32)
        self.pi_act = Chain.from_names(self.fp_a, SPECS, N, False)
        self.pi_upd = Chain.from_names(self.fp_a, SPECS, B, True)
        self.q1, self.q2 = Chain.from_names(self.fp_c1, SPECS, B, True), Chain.from_names(self.fp_c2, SPECS, B, True)
        self.q1t, self.q2t = Chain.from_names(self.fp_c1t, SPECS, B, False), Chain.from_names(self.fp_c2t, SPECS, B, False)
        self.memory = ReplayBuffer(cfg.memory_capacity, D, dev)
        z = lambda *s, dt=f32: torch.zeros(*s, device=dev, dtype=dt)
        self.idx, self.y = z(B, dt=i32), z(B)
        self.closs, self.acc, self.aloss_alpha = z(2), z(2), z(1)   # critic losses | actor loss, sum H | alpha loss
        self.action, self.logp = z(N, dt=i32), z(N)
        self.done = z(N, dt=u8)
        self.cur = z(N, D)
        self.ctr_act, self.ctr_upd = z(1, dt=i32), z(1, dt=i32)
        self.total_updates = 0
        self.act_count = 0
        self.graph_launches = 0
        self.episode_rewards = deque(maxlen=100)
        print(f"Device: {dev}")
        print(f"State dim: {D}, Action dim: {A}")

    @property
    def alpha(self) -> torch.Tensor:
        return self.log_alpha.exp()

    def soft_update(self, target=None, source=None):
        ops.polyak(self.fp_c1t.flat, self.fp_c1.flat, self.cfg.tau)
        ops.polyak(self.fp_c2t.flat, self.fp_c2.flat, self.cfg.tau)

    def act(self, obs: torch.Tensor, deterministic: bool = False) -> torch.Tensor:
        logits = self.pi_act.forward(obs, self.N)
        self.act_count += 1
        ops.sample_categorical(logits, seed=self.seed, draw=1, draw_base=self.ctr_act, deterministic=deterministic, action=self.action,
                               logp=self.logp)
        ops.counter_add(self.ctr_act, 1)
        return self.action

    @torch.no_grad()
    def select_action(self, state: np.ndarray, deterministic: bool = False) -> int:
        obs = torch.as_tensor(np.asarray(state, np.float32), device=self.device).reshape(1, -1)
        chain = getattr(self, "_pi_one", None) or Chain.from_names(self.fp_a, SPECS, 1, False)
        self._pi_one = chain
        self.act_count += 1
        a, _, _ = ops.sample_categorical(chain.forward(obs, 1), seed=self.seed, first_id=1 << 40, draw=self.act_count, deterministic=deterministic)
        return int(a.item())

    def update(self, idx: torch.Tensor = None):
        """One update (ref :155-221).  `idx` lets the parity test feed the reference's own sample."""
        cfg, B, mem = self.cfg, self.B, self.memory
        if len(mem) < B:
            return 0.0, 0.0, 0.0, 0.0
        self.total_updates += 1
        if idx is None:
            idx = mem.sample_indices(B, seed=self.seed, draw=1, draw_base=self.ctr_upd, out=self.idx)
        # ---- target (ref :163-176): actor and target critics on s' ----
        zn = self.pi_upd.forward(mem.next_obs, B, row_index=idx)
        q1t, q2t = self.q1t.forward(mem.next_obs, B, row_index=idx), self.q2t.forward(mem.next_obs, B, row_index=idx)
        off.sac_discrete_target(zn, q1t, q2t, mem.reward, mem.done, self.log_alpha, cfg.gamma, row_index=idx, out=self.y)
        # ---- critics (ref :178-189) ----
        q1, q2 = self.q1.forward(mem.obs, B, row_index=idx), self.q2.forward(mem.obs, B, row_index=idx)
        self.closs.zero_()
        off.sac_discrete_critic_loss(q1, q2, mem.action, self.y, self.q1.dout, self.q2.dout, row_index=idx, loss_acc=self.closs)
        self.q1.backward(mem.obs, B, row_index=idx)
        self.critic1_optim.step()
        self.q2.backward(mem.obs, B, row_index=idx)
        self.critic2_optim.step()
        # ---- actor (ref :191-200): the critics just updated, evaluated on s ----
        z = self.pi_upd.forward(mem.obs, B, row_index=idx)
        q1n, q2n = self.q1.forward(mem.obs, B, row_index=idx), self.q2.forward(mem.obs, B, row_index=idx)
        self.acc.zero_()
        off.sac_discrete_actor_grad(z, q1n, q2n, self.log_alpha, self.pi_upd.dout, self.acc)
        self.pi_upd.backward(mem.obs, B, row_index=idx)
        self.actor_optim.step()
        # ---- alpha (ref :202-207) and target sync (ref :209-210) ----
        off.sac_discrete_alpha_step(self.log_alpha, self.alpha_state, self.acc, B, cfg.target_entropy, cfg.lr_alpha, loss_out=self.aloss_alpha)
        self.soft_update()
        ops.counter_add(self.ctr_upd, 1)
        return self.acc, self.closs, self.aloss_alpha

    def losses(self):
        """(actor_loss, critic1_loss, critic2_loss, alpha_loss) as floats — one small D2H, call at log time."""
        a, c, al = self.acc.tolist(), self.closs.tolist(), self.aloss_alpha.tolist()
        return a[0], c[0], c[1], al[0]

    # ---------------------------------------------------------------- one lockstep (ref train() loop body :228-246)
    def _lockstep_body(self):
        env, mem, cur = self.env, self.memory, self.cur
        a = self.act(cur)
        obs, r, te, tr, nobs = env.step(a, done=self.done)
        mem.store(cur, a.view(-1, 1), r, nobs, self.done)       # done = terminated or truncated (ref :232-234)
        self.update()
        cur.copy_(obs)

    def train(self):
        print("Starting training...")
        cfg, env = self.cfg, self.env
        env.reset(out=self.cur)
        max_lock = cfg.max_locksteps or int(cfg.max_episodes * 500 / self.N)
        t0, last_total = time.time(), 0
        for step in range(max_lock):
            self.lockstep()
            if step % 100 == 99:
                avg, _, total = env.episode_stats(100)
                if total != last_total:
                    last_total = total
                    self.episode_rewards.extend([avg] * min(self.N, 100))
                    al, c1, c2, el = self.losses()
                    sps = (step + 1) * self.N / max(time.time() - t0, 1e-9)
                    print(f"Episodes {total} | Avg(100): {avg:.1f} | Alpha: {float(self.alpha):.4f} | Actor: {al:.3f} | "
                          f"Critic: {c1:.3f}/{c2:.3f} | {sps:,.0f} steps/s")
                    if avg >= 495.0 and total >= 100:
                        print(f"\nEnvironment solved in {total} episodes!")
                        break
        print("Training completed!")

    def eval(self, num_episodes: int = 10):
        print(f"\nEvaluating for {num_episodes} episodes...")
        env = ops.VecEnv(self.cfg.env_name, num_episodes, seed=self.seed + 777, first_env_id=1 << 32)
        chain = Chain.from_names(self.fp_a, SPECS, num_episodes, False)
        action = torch.zeros(num_episodes, device=self.device, dtype=i32)
        logp = torch.zeros(num_episodes, device=self.device, dtype=f32)
        obs = env.reset()
        ret = torch.zeros(num_episodes, device=self.device, dtype=f64)
        alive = torch.ones(num_episodes, device=self.device, dtype=torch.bool)
        for _ in range(env.max_episode_steps):
            ops.sample_categorical(chain.forward(obs, num_episodes), deterministic=True, action=action, logp=logp)   # argmax (ref :143-144)
            obs, r, te, tr, _ = env.step(action, want_next_obs=False)
            ret += torch.where(alive, r.double(), torch.zeros_like(ret))
            alive &= ~((te | tr).bool())
            if not bool(alive.any()):
                break
        rewards = ret.tolist()
        for i, r in enumerate(rewards):
            print(f"  Episode {i + 1}: Reward = {r:.0f}")
        print(f"Evaluation Results: Mean = {np.mean(rewards):.1f} +/- {np.std(rewards):.1f}")
        env.close()
        return rewards

    def test(self):
        self.eval(num_episodes=5)
        print("\n(visual test skipped: the device env has no renderer)")


def main():
    config = Config()
    config.num_envs, config.batch_size, config.memory_capacity = 1024, 1024, 1 << 18
    trainer = SACTrainer(config)
    try:
        trainer.train()
    except KeyboardInterrupt:
        print("\nTraining interrupted.")
    trainer.test()


if __name__ == "__main__":
    main()
