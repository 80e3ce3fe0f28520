"""DQN for CartPole-v1 on the B200 engine — same surface as the reference ``algorithms/dqn_cartpole.py``
(Config, QNetwork, ReplayBuffer, DQNTrainer.train/eval/test/update/select_action/get_epsilon).

    select_action (ref :124-133)  -> Q forward + epsilon-greedy kernel for all N envs (exp. decay per call, ref :117-122)
    ReplayBuffer  (ref :68-88)    -> SoA ring in HBM, sample = keyed bijection (without replacement)
    update        (ref :135-168)  -> gather fused into the first GEMM, target max-Q + MSE grad kernel, backward,
                                     per-element clamp(+-1) + Adam fused; hard target copy = polyak(tau=1)

Vectorisation (extra Config attrs; num_envs = 1 reproduces the reference schedule): N envs in lockstep, one
update() of ``batch_size`` per lockstep, target sync every ``target_update_freq`` finished episodes when N == 1
(ref :193-194) and every ``target_sync_updates`` updates when N > 1 (episodes end every step at N = 8192).
"""
from __future__ import annotations

import time
from collections import deque

import numpy as np
import torch
import torch.nn as nn

from .. import _ffi, ops, ops_offpolicy as off
from ..graphs import HostScheduledLockstep
from ..mlp import Chain
from ..nn import FlatParams, FusedAdam, layer_init

f32, i32, u8 = torch.float32, torch.int32, torch.uint8


class Config:
    def __init__(self):
        self.env_name = "CartPole-v1"
        self.seed = None
        self.max_episodes = 500
        self.max_steps = 500
        self.batch_size = 64
        self.gamma = 0.99
        self.lr = 1e-3
        self.epsilon_start = 0.95
        self.epsilon_end = 0.01
        self.epsilon_decay = 800
        self.target_update_freq = 4
        self.memory_capacity = 100000
        self.hidden_dim = 256
        self.device = "cuda"
        # ---- engine extras ----
        self.num_envs = 1
        self.target_sync_updates = 128
        self.max_locksteps = None      # stop criterion for vectorised runs (default: max_episodes * max_steps / num_envs)
        self.use_cuda_graph = True     # lockstep(): act -> env step -> store -> update as ONE captured graph per lockstep


class QNetwork(nn.Module):
    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int = 256):
        super().__init__()
        self.net = nn.Sequential(
            layer_init(nn.Linear(state_dim, hidden_dim)), nn.ReLU(),
            layer_init(nn.Linear(hidden_dim, hidden_dim)), nn.ReLU(),
            layer_init(nn.Linear(hidden_dim, action_dim), std=0.01))

    SPECS = [("net.0.weight", "net.0.bias", _ffi.ACT_RELU), ("net.2.weight", "net.2.bias", _ffi.ACT_RELU),
             ("net.4.weight", "net.4.bias", _ffi.ACT_NONE)]


class ReplayBuffer(off.ReplayRing):
    """Reference-compatible facade: push(s, a, r, s2, done) for one transition, sample(B) -> 5 numpy arrays."""

    def __init__(self, capacity: int, obs_dim: int = 4, device=None):
        super().__init__(capacity, obs_dim, 1, True, device or torch.device("cuda", torch.cuda.current_device()))

    def push(self, state, action, reward, next_state, done):
        dev = self.state.device
        self.store(torch.as_tensor(np.asarray(state, np.float32), device=dev).reshape(1, -1),
                   torch.tensor([[int(action)]], device=dev, dtype=i32), torch.tensor([float(reward)], device=dev, dtype=f32),
                   torch.as_tensor(np.asarray(next_state, np.float32), device=dev).reshape(1, -1),
                   torch.tensor([int(bool(done))], device=dev, dtype=u8))

    def sample(self, batch_size: int, seed: int = 0, draw: int = 0):
        batch_size = min(batch_size, len(self))
        idx = self.sample_indices(batch_size, seed=seed, draw=draw).long()
        return (self.obs[idx].cpu().numpy(), self.action[idx, 0].cpu().numpy(), self.reward[idx].cpu().numpy(),
                self.next_obs[idx].cpu().numpy(), self.done[idx].cpu().numpy().astype(bool))


class DQNTrainer(HostScheduledLockstep):
    def __init__(self, config: Config):
        _ffi.require_cuda()
        self.cfg = cfg = config
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.N = N = int(cfg.num_envs)
        self.seed = int(cfg.seed) if cfg.seed is not None else int(time.time_ns() & 0x7FFFFFFF)
        self.env = ops.VecEnv(cfg.env_name, N, seed=self.seed)
        D, A, B = self.env.obs_dim, self.env.n_actions, int(cfg.batch_size)
        self.policy_net = QNetwork(D, A, cfg.hidden_dim).to(self.device)
        self.target_net = QNetwork(D, A, cfg.hidden_dim).to(self.device)
        self.target_net.load_state_dict(self.policy_net.state_dict())
        self.target_net.eval()
        self.fp = FlatParams(self.policy_net, device=self.device)
        self.fp_t = FlatParams(self.target_net, device=self.device)
        self.optimizer = FusedAdam(self.fp, lr=cfg.lr)
        self.q_act = Chain.from_names(self.fp, QNetwork.SPECS, N, backward=False)      # acting
        self.q_upd = Chain.from_names(self.fp, QNetwork.SPECS, B, backward=True)       # update
        self.q_tgt = Chain.from_names(self.fp_t, QNetwork.SPECS, B, backward=False)
        self.memory = ReplayBuffer(cfg.memory_capacity, D, self.device)
        self.idx = torch.zeros(B, device=self.device, dtype=i32)
        self.loss_acc = torch.zeros(2, device=self.device, dtype=f32)
        self.action = torch.zeros(N, device=self.device, dtype=i32)
        self.done = torch.zeros(N, device=self.device, dtype=u8)
        self.cur = torch.zeros(N, D, device=self.device, dtype=f32)           # current observations of the lockstep
        self.eps_t = torch.zeros(1, device=self.device, dtype=f32)            # host-scheduled epsilon in a device slot
        self.ctr_act = torch.zeros(1, device=self.device, dtype=i32)          # device mirrors of sample_count / update_count
        self.ctr_upd = torch.zeros(1, device=self.device, dtype=i32)
        self.graph_launches = 0
        self._episodes_synced = 0
        self.epsilon = cfg.epsilon_start
        self.sample_count = 0
        self.update_count = 0
        self.episode_rewards = deque(maxlen=100)
        self._episodes_seen = 0
        print(f"Device: {self.device}")
        print(f"State dim: {D}, Action dim: {A}")

    def get_epsilon(self) -> float:
        self.sample_count += 1
        self.epsilon = self.cfg.epsilon_end + (self.cfg.epsilon_start - self.cfg.epsilon_end) * np.exp(
            -1.0 * self.sample_count / self.cfg.epsilon_decay)
        return self.epsilon

    def act(self, obs: torch.Tensor, deterministic: bool = False) -> torch.Tensor:
        """Vector select_action: one epsilon (decayed once per call, ref :117-122) for the whole lockstep."""
        eps = 0.0 if deterministic else self.get_epsilon()
        q = self.q_act.forward(obs, self.N)
        a = ops.select_eps_greedy(q, eps, seed=self.seed, draw=1, draw_base=self.ctr_act, action=self.action)
        if not deterministic:
            ops.counter_add(self.ctr_act, 1)
        return a

    @torch.no_grad()
    def select_action(self, state: np.ndarray, deterministic: bool = False) -> int:
        obs = torch.as_tensor(np.asarray(state, np.float32), device=self.device).reshape(1, -1)
        if self.N != 1:
            chain = getattr(self, "_q_one", None) or Chain.from_names(self.fp, QNetwork.SPECS, 1, backward=False)
            self._q_one = chain
            eps = 0.0 if deterministic else self.get_epsilon()
            return int(ops.select_eps_greedy(chain.forward(obs, 1), eps, seed=self.seed, draw=self.sample_count).item())
        return int(self.act(obs, deterministic).item())

    def update(self, idx: torch.Tensor = None) -> float:
        if len(self.memory) < int(self.cfg.batch_size):
            return 0.0
        self.update_count += 1
        self.optimizer.sync_lr()
        self._update_device(idx)
        return self.loss_acc[0]                  # device scalar; .item() only when the caller wants it

    def _update_device(self, idx: torch.Tensor = None):
        """The device side of update() (ref :135-168), capture-safe: the sampling draw comes from a device counter."""
        cfg, B, mem = self.cfg, int(self.cfg.batch_size), self.memory
        if idx is None:
            idx = mem.sample_indices(B, seed=self.seed, draw=1, draw_base=self.ctr_upd, out=self.idx)
        q = self.q_upd.forward(mem.obs, B, row_index=idx)
        qn = self.q_tgt.forward(mem.next_obs, B, row_index=idx)
        self.loss_acc.zero_()
        off.dqn_loss(q, qn, mem.action, mem.reward, mem.done, cfg.gamma, row_index=idx, dq=self.q_upd.dout, loss_acc=self.loss_acc)
        self.q_upd.backward(mem.obs, B, row_index=idx)
        self.optimizer.launch(clamp=1.0)         # param.grad.clamp_(-1, 1) then Adam (ref :161-166)
        ops.counter_add(self.ctr_upd, 1)

    # ---------------------------------------------------------------- one lockstep (ref train() loop body :178-191)
    def _lockstep_ready(self) -> bool:
        return len(self.memory) + self.N >= int(self.cfg.batch_size)

    def _before_lockstep(self):
        self.eps_t.fill_(float(self.get_epsilon()))      # one decay per select_action call (ref :117-122)
        self.optimizer.sync_lr()

    def _lockstep_body(self):
        mem, cur = self.memory, self.cur
        q = self.q_act.forward(cur, self.N)
        ops.select_eps_greedy(q, self.eps_t, seed=self.seed, draw=1, draw_base=self.ctr_act, action=self.action)
        ops.counter_add(self.ctr_act, 1)
        obs, r, te, tr, nobs = self.env.step(self.action, done=self.done)
        mem.store(cur, self.action.view(-1, 1), r, nobs, self.done)
        if len(mem) >= int(self.cfg.batch_size):
            self._update_device()
        cur.copy_(obs)

    def _host_mirrors(self):
        return self.memory._size_host

    def _set_host_mirrors(self, m):
        self.memory._size_host = m

    def _advance_host_mirrors(self):
        self.memory._size_host = min(self.memory.capacity, self.memory._size_host + self.N)

    def _after_lockstep(self):
        cfg = self.cfg
        if len(self.memory) >= int(cfg.batch_size):
            self.update_count += 1
        if self.N > 1:
            if self.update_count and self.update_count % cfg.target_sync_updates == 0:
                self.sync_target()
        elif bool(self.done.item()):      # reference schedule: hard sync every `target_update_freq` finished episodes (ref :193-194)
            self._episodes_synced += 1
            if self._episodes_synced % cfg.target_update_freq == 0:
                self.sync_target()

    def sync_target(self):
        ops.polyak(self.fp_t.flat, self.fp.flat, 1.0)

    def _episode_stats(self):
        mean_ret, _, total = self.env.episode_stats(100)
        return mean_ret, total

    def train(self):
        print("Starting training...")
        cfg, env = self.cfg, self.env
        env.reset(out=self.cur)
        max_lock = cfg.max_locksteps or int(cfg.max_episodes * cfg.max_steps / self.N)
        last_total, t0 = 0, time.time()
        for step in range(max_lock):
            self.lockstep()
            if self.N == 1 or step % 50 == 49:
                avg, total = self._episode_stats()
                if total != last_total:
                    self.episode_rewards.extend([avg] * min(total - last_total, 100))
                    last_total = total
                    if self.N > 1 or total % 10 == 0:
                        sps = (step + 1) * self.N / max(time.time() - t0, 1e-9)
                        print(f"Episodes {total} | Avg(100): {avg:.1f} | Epsilon: {self.epsilon:.3f} | {sps:,.0f} steps/s")
                    if avg >= 495.0 and total >= 100:
                        print(f"\nEnvironment solved in {total} episodes!")
                        break
        print("Training completed!")

    def eval(self, num_episodes: int = 10):
        print(f"\nEvaluating for {num_episodes} episodes...")
        env = ops.VecEnv(self.cfg.env_name, num_episodes, seed=self.seed + 999, first_env_id=1 << 32)
        chain = Chain.from_names(self.fp, QNetwork.SPECS, num_episodes, backward=False)
        obs = env.reset()
        ret = torch.zeros(num_episodes, device=self.device, dtype=torch.float64)
        alive = torch.ones(num_episodes, device=self.device, dtype=torch.bool)
        for _ in range(env.max_episode_steps):
            a = ops.select_eps_greedy(chain.forward(obs, num_episodes), 0.0)
            obs, r, te, tr, _ = env.step(a, want_next_obs=False)
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


if __name__ == "__main__":
    config = Config()
    config.num_envs = 1024
    trainer = DQNTrainer(config)
    trainer.train()
    trainer.test()
