"""Double DQN + prioritized replay for CartPole-v1 on the B200 engine — same surface as the reference
``algorithms/ddqn_per_cartpole.py`` (Config, QNetwork, SumTree, PrioritizedReplayBuffer, DDQNPERTrainer.train/eval/test/
update/select_action/get_epsilon).  SURVEY §8f rank 2 ("dialect-B PER"): it reuses the kernels of rows a10-a17.

    SumTree / PrioritizedReplayBuffer (ref :67-150) -> SoA ring + float64 sum-tree in the reference's heap layout
        push    : priority = max leaf (1.0 while the tree is empty)                      gymrl_sumtree_store_new
        sample  : stratified v ~ U(seg i, seg (i+1)), `v <= left` descent, beta += 0.001 per call (host scalar -> device),
                  w = (size p / total)^-beta / max w                                      gymrl_sumtree_sample
        update  : p = min(|td| + 1e-4, error_max) ^ alpha                                 gymrl_sumtree_update(clip_max)
    update (ref :206-247) -> online Q(s), double-Q target from the online argmax on s' evaluated by the target net,
                             loss = mean(td^2 w), priorities from |td|, grad clamp(+-1), Adam
    target sync (ref :266-267) -> every `target_update_freq` episodes at N = 1, every `target_sync_updates` updates at N > 1
Vectorisation: N envs in lockstep, one update of ``batch_size`` per lockstep; epsilon decays per select_action call.
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
from ..nn import FlatParams, FusedAdam

f32, i32, u8, f64 = torch.float32, torch.int32, torch.uint8, torch.float64
RELU, NONE = _ffi.ACT_RELU, _ffi.ACT_NONE


class Config:
    def __init__(self):
        self.env_name = "CartPole-v1"
        self.seed = None
        self.max_episodes = 500
        self.max_steps = 10000
        self.batch_size = 64
        self.gamma = 0.9
        self.lr = 0.001
        self.epsilon_start = 0.95
        self.epsilon_end = 0.01
        self.epsilon_decay = 800
        self.target_update_freq = 4
        self.memory_capacity = 65536
        self.hidden_dim = 256
        self.alpha = 0.6
        self.beta = 0.4
        self.beta_increment = 0.001
        self.error_max = 1.0
        self.eps = 1e-4
        self.device = "cuda"
        # ---- engine extras ----
        self.num_envs = 1
        self.max_locksteps = None
        self.use_cuda_graph = True   # lockstep(): act -> env step -> store -> update as ONE captured graph
        self.target_sync_updates = 200


class QNetwork(nn.Module):
    """Same module tree / state_dict keys as the reference QNetwork (ref :54-64)."""

    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int = 256):
        super().__init__()
        self.fc1 = nn.Linear(state_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.fc3 = nn.Linear(hidden_dim, action_dim)

    DUELING = False
    PARAM_ORDER = None

    @staticmethod
    def chain(fp: FlatParams, M: int, backward: bool) -> Chain:
        return Chain.from_names(fp, [("fc1.weight", "fc1.bias", RELU), ("fc2.weight", "fc2.bias", RELU), ("fc3.weight", "fc3.bias", NONE)],
                                M, backward)


class SumTree(off.DeviceSumTree):
    """Reference-compatible facade (update(tree_index, p) / get_leaf(v) / total_priority) over the device tree; this dialect
    addresses leaves by TREE index (ref :75-107), the kernels by data index."""

    def __init__(self, capacity: int, device=None):
        super().__init__(capacity, device or torch.device("cuda", torch.cuda.current_device()))
        self._ring1 = torch.tensor([0, 1], device=self.tree.device, dtype=i32)
        self._beta0 = torch.zeros(1, device=self.tree.device, dtype=f64)

    def update(self, index, priority=None, **kw):
        if isinstance(index, torch.Tensor):
            return super().update(index, priority, **kw)
        dev = self.tree.device
        super().update(torch.tensor([int(index) - self.capacity + 1], device=dev, dtype=i32),
                       torch.tensor([float(priority)], device=dev, dtype=f64))

    def get_leaf(self, v: float):
        u = torch.tensor([float(v)], device=self.tree.device, dtype=f64)
        prio = torch.zeros(1, device=self.tree.device, dtype=f64)
        idx, _ = self.sample(1, self._ring1, self._beta0, uniforms=u, out_prio=prio, raw_values=True, tree_index=True)
        return int(idx.item()), float(prio.item())

    def total_priority(self) -> float:
        return float(self.tree[0].item())


class PrioritizedReplayBuffer:
    def __init__(self, config: Config, state_dim: int = 4, device=None):
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.cfg, self.device, self.capacity = config, dev, int(config.memory_capacity)
        self.tree = SumTree(self.capacity, dev)
        self.ring = off.ReplayRing(self.capacity, state_dim, 1, True, dev)
        B = int(config.batch_size)
        self.beta_t = torch.zeros(1, device=dev, dtype=f64)
        self.batch_index = torch.zeros(B, device=dev, dtype=i32)      # data indices
        self.is_weight = torch.zeros(B, device=dev, dtype=f32)

    def store(self, obs, action, reward, next_obs, done_u8):
        """n transitions at once (one per env of a lockstep): leaves get the current max priority (ref :116-119)."""
        n = obs.shape[0]
        L = self.ring
        from .._ffi import check, load, ptr, stream_ptr
        s, st = stream_ptr(), L.state
        check(load().gymrl_replay_store(ptr(L.obs), ptr(obs, f32), n, L.obs_dim, 0, L.capacity, ptr(st, i32), s))
        check(load().gymrl_replay_store(ptr(L.next_obs), ptr(next_obs, f32), n, L.obs_dim, 0, L.capacity, ptr(st, i32), s))
        check(load().gymrl_replay_store(ptr(L.action), ptr(action), n, 1, 0, L.capacity, ptr(st, i32), s))
        check(load().gymrl_replay_store(ptr(L.reward), ptr(reward, f32), n, 1, 0, L.capacity, ptr(st, i32), s))
        check(load().gymrl_replay_store(ptr(L.done), ptr(done_u8, u8), n, 1, 1, L.capacity, ptr(st, i32), s))
        self.tree.store_new(n, L.state)
        L.advance(n)

    def push(self, transition):
        """Single-transition facade with the reference signature: push((state, action, reward, next_state, done))."""
        s, a, r, s2, d = transition
        dev = self.device
        t = lambda x, dt: torch.as_tensor(np.asarray(x), device=dev).to(dt)
        self.store(t(s, f32).reshape(1, -1), t([a], i32).reshape(1, 1), t([r], f32), t(s2, f32).reshape(1, -1), t([bool(d)], u8))

    def advance_beta(self):
        """beta += beta_increment per sample() call (ref :125-130): host value -> device slot (outside any captured graph)."""
        self.cfg.beta = min(1.0, self.cfg.beta + self.cfg.beta_increment)
        self.beta_t.fill_(self.cfg.beta)

    def sample(self, batch_size: int = None, uniforms=None, seed=0, draw=0, draw_base=None, advance_beta=True):
        """Returns (data indices [B] int32, IS weights [B] float32) — device tensors; beta advances per call (ref :130)."""
        if advance_beta:
            self.advance_beta()
        self.tree.sample(int(batch_size or self.cfg.batch_size), self.ring.state, self.beta_t, uniforms=uniforms, out_idx=self.batch_index,
                         out_w=self.is_weight, seed=seed, draw=draw, draw_base=draw_base)
        return self.batch_index, self.is_weight

    def update_priorities(self, indices, errors):
        """p = min(|td| + eps, error_max) ^ alpha (ref :142-150); `errors` may be signed TD errors (the kernel takes |.|)."""
        self.tree.update(indices, td_error=errors, eps=self.cfg.eps, alpha=self.cfg.alpha, clip_max=self.cfg.error_max)

    def __len__(self) -> int:
        return len(self.ring)


class DDQNPERTrainer(HostScheduledLockstep):
    NET = QNetwork

    def __init__(self, config: Config):
        _ffi.require_cuda()
        self.cfg = cfg = config
        self.device = dev = torch.device("cuda", torch.cuda.current_device())
        self.N = N = int(cfg.num_envs)
        self.seed = int(cfg.seed) if cfg.seed is not None else int(time.time_ns() & 0x7FFFFFFF)
        self.env = ops.VecEnv(cfg.env_name, N, seed=self.seed)
        self.state_dim, self.action_dim = D, A = self.env.obs_dim, self.env.n_actions
        B, Net = int(cfg.batch_size), self.NET
        self.policy_net = Net(D, A, cfg.hidden_dim).to(dev)
        self.target_net = Net(D, A, cfg.hidden_dim).to(dev)
        self.target_net.load_state_dict(self.policy_net.state_dict())
        self.target_net.eval()
        self.fp = FlatParams(self.policy_net, Net.PARAM_ORDER, dev)
        self.fp_t = FlatParams(self.target_net, Net.PARAM_ORDER, dev)
        self.optimizer = FusedAdam(self.fp, lr=cfg.lr)
        self.q_act = Net.chain(self.fp, N, False)
        self.q_upd = Net.chain(self.fp, B, True)
        self.q_nxt = Net.chain(self.fp, B, False)       # online net on s' (double-Q argmax, no gradient)
        self.q_tgt = Net.chain(self.fp_t, B, False)
        self.memory = PrioritizedReplayBuffer(cfg, D, dev)
        self.td = torch.zeros(B, device=dev, dtype=f32)
        self.loss_acc = torch.zeros(2, device=dev, dtype=f32)
        self.action = torch.zeros(N, device=dev, dtype=i32)
        self.done = torch.zeros(N, device=dev, dtype=u8)
        self.cur = torch.zeros(N, D, device=dev, dtype=f32)
        self.eps_t = torch.zeros(1, device=dev, dtype=f32)        # host-scheduled epsilon in a device slot
        self.ctr_act = torch.zeros(1, device=dev, dtype=i32)      # device mirrors of sample_count / update_count
        self.ctr_upd = torch.zeros(1, device=dev, dtype=i32)
        self.graph_launches = 0
        self._episodes_synced = 0
        self.epsilon = cfg.epsilon_start
        self.sample_count = 0
        self.update_count = 0
        self.episode_rewards = deque(maxlen=100)
        print(f"Device: {dev}")
        print(f"State dim: {D}, Action dim: {A}")

    def get_epsilon(self) -> float:
        self.sample_count += 1
        self.epsilon = self.cfg.epsilon_end + (self.cfg.epsilon_start - self.cfg.epsilon_end) * np.exp(
            -1.0 * self.sample_count / self.cfg.epsilon_decay)
        return self.epsilon

    def _greedy_columns(self, out: torch.Tensor) -> torch.Tensor:
        # dueling: argmax(V + A - mean A) == argmax A
        return out[:, :self.action_dim]

    def act(self, obs: torch.Tensor, deterministic: bool = False) -> torch.Tensor:
        """Vector select_action: one epsilon (decayed once per call, ref :180-186) for the whole lockstep."""
        eps = 0.0 if deterministic else self.get_epsilon()
        q = self.q_act.forward(obs, self.N)
        a = ops.select_eps_greedy(self._greedy_columns(q), eps, seed=self.seed, draw=1, draw_base=self.ctr_act, action=self.action)
        if not deterministic:
            ops.counter_add(self.ctr_act, 1)
        return a

    @torch.no_grad()
    def select_action(self, state: np.ndarray, deterministic: bool = False) -> int:
        obs = torch.as_tensor(np.asarray(state, np.float32), device=self.device).reshape(1, -1)
        chain = getattr(self, "_q_one", None) or self.NET.chain(self.fp, 1, False)
        self._q_one = chain
        eps = 0.0 if deterministic else self.get_epsilon()
        return int(ops.select_eps_greedy(self._greedy_columns(chain.forward(obs, 1)), eps, seed=self.seed, draw=self.sample_count).item())

    def update(self, uniforms=None) -> float:
        """One update (ref :206-247).  `uniforms` [B] float64 lets the parity test feed the reference's random.uniform draws."""
        if len(self.memory) < int(self.cfg.batch_size):
            return 0.0
        self.update_count += 1
        self.memory.advance_beta()
        self.optimizer.sync_lr()
        self._update_device(uniforms)
        return self.loss_acc[0]

    def _update_device(self, uniforms=None):
        """Device side of update(), capture-safe: PER beta and the learning rate sit in device slots, the stratified draws
        come from a device counter."""
        cfg, B, A, mem = self.cfg, int(self.cfg.batch_size), self.action_dim, self.memory
        ring = mem.ring
        idx, w = mem.sample(B, uniforms=uniforms, seed=self.seed, draw=1, draw_base=self.ctr_upd, advance_beta=False)
        q = self.q_upd.forward(ring.obs, B, row_index=idx)
        qo = self.q_nxt.forward(ring.next_obs, B, row_index=idx)
        qt = self.q_tgt.forward(ring.next_obs, B, row_index=idx)
        self.loss_acc.zero_()
        duel = self.NET.DUELING
        off.dqn_loss(q[:, :A], qt[:, :A], ring.action, ring.reward, ring.done, cfg.gamma, v=q[:, A:] if duel else None,
                     vnext_target=qt[:, A:] if duel else None, qnext_online=qo[:, :A], vnext_online=qo[:, A:] if duel else None,
                     row_index=idx, is_weight=w, dq=self.q_upd.dout[:, :A], dv=self.q_upd.dout[:, A:] if duel else None,
                     td_error=self.td, loss_acc=self.loss_acc)
        mem.update_priorities(idx, self.td)                      # ref :236-237 (before the optimizer step)
        self.q_upd.backward(ring.obs, B, row_index=idx)
        self.optimizer.launch(clamp=1.0)                         # param.grad.clamp_(-1, 1) then Adam (ref :241-245)
        ops.counter_add(self.ctr_upd, 1)

    # ---------------------------------------------------------------- one lockstep (ref train() loop body)
    def _lockstep_ready(self) -> bool:
        return len(self.memory) + self.N >= int(self.cfg.batch_size)

    def _before_lockstep(self):
        self.eps_t.fill_(float(self.get_epsilon()))
        if self._lockstep_ready():
            self.memory.advance_beta()                           # beta += beta_increment per sample() call (ref :125-130)
        self.optimizer.sync_lr()

    def _lockstep_body(self):
        mem, cur = self.memory, self.cur
        q = self.q_act.forward(cur, self.N)
        ops.select_eps_greedy(self._greedy_columns(q), self.eps_t, seed=self.seed, draw=1, draw_base=self.ctr_act, action=self.action)
        ops.counter_add(self.ctr_act, 1)
        obs, r, te, tr, nobs = self.env.step(self.action, done=self.done)
        mem.store(cur, self.action.view(-1, 1), r, nobs, self.done)
        if len(mem) >= int(self.cfg.batch_size):
            self._update_device()
        cur.copy_(obs)

    def _host_mirrors(self):
        return self.memory.ring._size_host

    def _set_host_mirrors(self, m):
        self.memory.ring._size_host = m

    def _advance_host_mirrors(self):
        ring = self.memory.ring
        ring._size_host = min(ring.capacity, ring._size_host + self.N)

    def _after_lockstep(self):
        cfg = self.cfg
        if len(self.memory) >= int(cfg.batch_size):
            self.update_count += 1
        if self.N > 1:
            if self.update_count and self.update_count % cfg.target_sync_updates == 0:
                self.sync_target()
        elif bool(self.done.item()):
            self._episodes_synced += 1
            if self._episodes_synced % cfg.target_update_freq == 0:
                self.sync_target()

    def sync_target(self):
        ops.polyak(self.fp_t.flat, self.fp.flat, 1.0)

    def train(self):
        print("Starting training...")
        cfg, env = self.cfg, self.env
        env.reset(out=self.cur)
        max_lock = cfg.max_locksteps or int(cfg.max_episodes * 500 / self.N)
        last_total, t0 = 0, time.time()
        for step in range(max_lock):
            self.lockstep()
            if self.N == 1 or step % 50 == 49:
                avg, _, total = env.episode_stats(100)
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
        chain = self.NET.chain(self.fp, num_episodes, False)
        obs = env.reset()
        ret = torch.zeros(num_episodes, device=self.device, dtype=f64)
        alive = torch.ones(num_episodes, device=self.device, dtype=torch.bool)
        for _ in range(env.max_episode_steps):
            a = ops.select_eps_greedy(self._greedy_columns(chain.forward(obs, num_episodes)), 0.0)
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


def main():
    config = Config()
    config.num_envs, config.batch_size, config.memory_capacity = 1024, 1024, 1 << 18
    trainer = DDQNPERTrainer(config)
    try:
        trainer.train()
    except KeyboardInterrupt:
        print("\nTraining interrupted.")
    trainer.test()


if __name__ == "__main__":
    main()
