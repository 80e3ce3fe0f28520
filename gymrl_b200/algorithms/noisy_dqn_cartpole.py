"""NoisyNet DQN (dueling, double-Q) for CartPole-v1 on the B200 engine — same surface as the reference
``algorithms/noisy_dqn_cartpole.py`` (Config, NoisyLinear, NoisyDuelingQNetwork, ReplayBuffer,
NoisyDQNTrainer.train/eval/test/update/select_action).  SURVEY §8f rank 2: it reuses the kernels of rows a10-a16.

    NoisyLinear x 4          (ref :51-103)  -> factorised-noise kernel + W = mu + sigma * outer(eps_out, eps_in) composed on
                                               the device per forward (fresh noise per forward in training mode, ref :96-98);
                                               value | advantage streams as ONE [A+1, H] GEMM, dueling mean inside the loss
    ReplayBuffer             (ref :139-161) -> SoA ring + device sampling without replacement (gymrl_replay_*)
    update                   (ref :206-253) -> online forward on s (noise draw 1), online forward on s' (noise draw 2) for
                                               the double-Q argmax, target net in eval mode (mu only), MSE, Adam (no clipping),
                                               hard target sync every `target_update_freq` updates
Vectorisation: N envs in lockstep, one update of ``batch_size`` per lockstep (ref: one per env step).
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
from .rainbow_dqn_cartpole import NoisyLinear

f32, i32, u8, f64 = torch.float32, torch.int32, torch.uint8, torch.float64
RELU, NONE = _ffi.ACT_RELU, _ffi.ACT_NONE


class Config:
    def __init__(self):
        self.env_name = "CartPole-v1"
        self.seed = None
        self.max_episodes = 500
        self.max_steps = 10000
        self.batch_size = 64
        self.gamma = 0.99
        self.lr = 0.001
        self.target_update_freq = 500
        self.memory_capacity = 10000
        self.hidden_dim = 64
        self.sigma_init = 0.5
        self.device = "cuda"
        # ---- engine extras ----
        self.num_envs = 1
        self.max_locksteps = None
        self.use_cuda_graph = True   # lockstep(): act -> env step -> store -> update as ONE captured graph


class NoisyDuelingQNetwork(nn.Module):
    """Parameter container with the reference's module tree / state_dict keys (ref :106-136)."""

    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int = 64, sigma_init: float = 0.5):
        super().__init__()
        self.fc1 = NoisyLinear(state_dim, hidden_dim, sigma_init)
        self.fc2 = NoisyLinear(hidden_dim, hidden_dim, sigma_init)
        self.value_stream = NoisyLinear(hidden_dim, 1, sigma_init)
        self.advantage_stream = NoisyLinear(hidden_dim, action_dim, sigma_init)

    LAYERS = ("fc1", "fc2", "value_stream", "advantage_stream")   # the order the reference draws noise in (ref :126-133)


class ReplayBuffer(off.ReplayRing):
    def __init__(self, capacity: int, obs_dim: int = 4, device=None):
        super().__init__(capacity, obs_dim, 1, True, device or torch.device("cuda", torch.cuda.current_device()))

    def push(self, state, action, reward, next_state, done):
        dev = self.state.device
        t = lambda x, dt: torch.as_tensor(np.asarray(x), device=dev).to(dt)
        self.store(t(state, f32).reshape(1, -1), t([action], i32).reshape(1, 1), t([reward], f32), t(next_state, f32).reshape(1, -1),
                   t([bool(done)], u8))


class NoisyDuelingEngine:
    """Forward / backward of NoisyDuelingQNetwork for a fixed maximum batch: every layer's effective weight is composed on
    the device from (mu, sigma, eps_in, eps_out); the two streams share one [A+1, H] GEMM (rows :A advantage, row A value)."""

    def __init__(self, fp: FlatParams, D: int, A: int, H: int, M: int, backward: bool, seed: int, entity: int):
        dev = fp.flat.device
        self.fp, self.D, self.A, self.H, self.M, self.seed, self.entity = fp, D, A, H, M, seed, entity
        z = lambda *s: torch.zeros(*s, device=dev, dtype=f32)
        dims = {"fc1": (H, D), "fc2": (H, H), "value_stream": (1, H), "advantage_stream": (A, H)}
        self.eps = {n: (z(dims[n][1]), z(dims[n][0])) for n in NoisyDuelingQNetwork.LAYERS}       # (eps_in, eps_out)
        self.W1, self.b1, self.W2, self.b2 = z(H, D), z(H), z(H, H), z(H)
        self.Wh, self.bh = z(A + 1, H), z(A + 1)
        self.dW1, self.db1, self.dW2, self.db2 = z(H, D), z(H), z(H, H), z(H)
        self.trunk = Chain(fp, [(self.W1, self.b1, self.dW1, self.db1, RELU), (self.W2, self.b2, self.dW2, self.db2, RELU)], M, backward)
        self.out = z(M, A + 1)
        self.ctr = torch.zeros(1, device=dev, dtype=i32)     # noise draw counter (device: captured graphs draw afresh)
        if backward:
            self.dout = z(M, A + 1)
            self.dWh, self.dbh = z(A + 1, H), z(A + 1)
            self.ws = torch.empty(ops.backward_weight_workspace(M, A + 1, H), device=dev, dtype=torch.uint8)

    def _targets(self):
        A = self.A
        return {"fc1": (self.W1, self.b1), "fc2": (self.W2, self.b2), "value_stream": (self.Wh[A:], self.bh[A:]),
                "advantage_stream": (self.Wh[:A], self.bh[:A])}

    def compose(self, noisy: bool, xi=None):
        """xi: optional {layer: (xi_in, xi_out)} of pre-drawn N(0,1) values (parity tests feed the reference's own draws)."""
        P = self.fp.p
        if noisy:
            for k, n in enumerate(NoisyDuelingQNetwork.LAYERS):
                e_in, e_out = self.eps[n]
                off.noisy_sample(e_in, xi[n][0] if xi is not None else None, seed=self.seed, entity=self.entity * 64 + 2 * k,
                                 draw=1, draw_base=self.ctr)
                off.noisy_sample(e_out, xi[n][1] if xi is not None else None, seed=self.seed, entity=self.entity * 64 + 2 * k + 1,
                                 draw=1, draw_base=self.ctr)
            ops.counter_add(self.ctr, 1)
        else:
            for e_in, e_out in self.eps.values():
                e_in.zero_(); e_out.zero_()
        for n, (W, b) in self._targets().items():
            e_in, e_out = self.eps[n]
            off.noisy_compose(P(n + ".weight_mu"), P(n + ".weight_sigma"), e_in, e_out, P(n + ".bias_mu"), P(n + ".bias_sigma"), W, b)

    def forward(self, x, M, row_index=None, noisy=True, xi=None):
        self.compose(noisy, xi)
        h = self.trunk.forward(x, M, row_index=row_index)
        return ops.linear_forward(h, self.Wh, self.bh, NONE, out=self.out, M=M)

    def backward(self, x, M, row_index=None):
        """Given self.dout = dL/d[advantage | value]: gradients of every (mu, sigma) through the composed weights."""
        G, A = self.fp.g, self.A
        ops.linear_backward_weight(self.dout, self.trunk.out, self.dWh, self.dbh, workspace=self.ws, M=M)
        ops.linear_backward_input(self.dout[:M], self.Wh, self.trunk.out, RELU, out=self.trunk.dout)
        self.trunk.backward(x, M, row_index=row_index)
        grads = {"fc1": (self.dW1, self.db1), "fc2": (self.dW2, self.db2), "value_stream": (self.dWh[A:], self.dbh[A:]),
                 "advantage_stream": (self.dWh[:A], self.dbh[:A])}
        for n, (dW, db) in grads.items():
            e_in, e_out = self.eps[n]
            off.noisy_backward(dW, db, e_in, e_out, G(n + ".weight_mu"), G(n + ".weight_sigma"), G(n + ".bias_mu"), G(n + ".bias_sigma"))


class NoisyDQNTrainer(HostScheduledLockstep):
    def __init__(self, config: Config):
        _ffi.require_cuda()
        self.cfg = cfg = config
        self.device = dev = torch.device("cuda", torch.cuda.current_device())
        self.N = N = int(cfg.num_envs)
        self.seed = int(cfg.seed) if cfg.seed is not None else int(time.time_ns() & 0x7FFFFFFF)
        self.env = ops.VecEnv(cfg.env_name, N, seed=self.seed)
        self.state_dim, self.action_dim = D, A = self.env.obs_dim, self.env.n_actions
        H, B = cfg.hidden_dim, int(cfg.batch_size)
        self.policy_net = NoisyDuelingQNetwork(D, A, H, cfg.sigma_init).to(dev)
        self.target_net = NoisyDuelingQNetwork(D, A, H, cfg.sigma_init).to(dev)
        self.target_net.load_state_dict(self.policy_net.state_dict())
        self.target_net.eval()
        self.fp, self.fp_t = FlatParams(self.policy_net, device=dev), FlatParams(self.target_net, device=dev)
        self.optimizer = FusedAdam(self.fp, lr=cfg.lr)
        self.net_act = NoisyDuelingEngine(self.fp, D, A, H, N, False, self.seed, 1)
        self.net_upd = NoisyDuelingEngine(self.fp, D, A, H, B, True, self.seed, 2)
        self.net_nxt = NoisyDuelingEngine(self.fp, D, A, H, B, False, self.seed, 3)
        self.net_tgt = NoisyDuelingEngine(self.fp_t, D, A, H, B, False, self.seed, 4)
        self.memory = ReplayBuffer(cfg.memory_capacity, D, dev)
        self.idx = torch.zeros(B, device=dev, dtype=i32)
        self.td = torch.zeros(B, device=dev, dtype=f32)
        self.loss_acc = torch.zeros(2, device=dev, dtype=f32)
        self.q_sum = torch.zeros(1, device=dev, dtype=f32)
        self.action = torch.zeros(N, device=dev, dtype=i32)
        self.done = torch.zeros(N, device=dev, dtype=u8)
        self.ctr_upd = torch.zeros(1, device=dev, dtype=i32)
        self.cur = torch.zeros(N, D, device=dev, dtype=f32)
        self.graph_launches = 0
        self.learn_step = 0
        self.episode_rewards = deque(maxlen=100)
        print(f"Device: {dev}")
        print(f"State dim: {D}, Action dim: {A}")

    def act(self, obs: torch.Tensor, deterministic: bool = False) -> torch.Tensor:
        out = self.net_act.forward(obs, self.N, noisy=not deterministic)
        # argmax(V + A - mean A) == argmax A : the greedy kernel runs on the advantage columns
        return ops.select_eps_greedy(out[:, :self.action_dim], 0.0, action=self.action)

    @torch.no_grad()
    def select_action(self, state: np.ndarray, deterministic: bool = False) -> int:
        obs = torch.as_tensor(np.asarray(state, np.float32), device=self.device).reshape(1, -1)
        eng = getattr(self, "_net_one", None) or NoisyDuelingEngine(self.fp, self.state_dim, self.action_dim, self.cfg.hidden_dim, 1,
                                                                    False, self.seed, 5)
        self._net_one = eng
        out = eng.forward(obs, 1, noisy=not deterministic)
        return int(ops.select_eps_greedy(out[:, :self.action_dim], 0.0).item())

    def update(self, idx: torch.Tensor = None, xi_cur=None, xi_next=None) -> dict:
        """One update (ref :206-253).  idx / xi_* let the parity test feed the reference's own sample and noise draws."""
        if len(self.memory) < int(self.cfg.batch_size):
            return {}
        self.optimizer.sync_lr()
        self._update_device(idx, xi_cur, xi_next)
        self._after_update()
        return {"loss": self.loss_acc[0]}

    def _update_device(self, idx=None, xi_cur=None, xi_next=None):
        """Device side of update(), capture-safe (sampling and noise draws come from device counters)."""
        cfg, B, A, mem = self.cfg, int(self.cfg.batch_size), self.action_dim, self.memory
        if idx is None:
            idx = mem.sample_indices(B, seed=self.seed, draw=1, draw_base=self.ctr_upd, out=self.idx)
        q = self.net_upd.forward(mem.obs, B, row_index=idx, noisy=True, xi=xi_cur)            # ref :229 (first noise draw)
        qo = self.net_nxt.forward(mem.next_obs, B, row_index=idx, noisy=True, xi=xi_next)     # ref :232 (second noise draw)
        qt = self.net_tgt.forward(mem.next_obs, B, row_index=idx, noisy=False)                # target net is .eval(): mu only
        self.loss_acc.zero_()
        off.dqn_loss(q[:, :A], qt[:, :A], mem.action, mem.reward, mem.done, cfg.gamma, v=q[:, A:], vnext_target=qt[:, A:],
                     qnext_online=qo[:, :A], vnext_online=qo[:, A:], row_index=idx, dq=self.net_upd.dout[:, :A],
                     dv=self.net_upd.dout[:, A:], td_error=self.td, loss_acc=self.loss_acc)
        self.net_upd.backward(mem.obs, B, row_index=idx)
        self.optimizer.launch()
        ops.counter_add(self.ctr_upd, 1)

    def _after_update(self):
        self.learn_step += 1
        if self.learn_step % self.cfg.target_update_freq == 0:
            ops.polyak(self.fp_t.flat, self.fp.flat, 1.0)     # hard sync (ref :248-249); buffers (epsilon) are not parameters

    # ---------------------------------------------------------------- one lockstep (ref train() loop body :255-270)
    def _lockstep_ready(self) -> bool:
        return len(self.memory) + self.N >= int(self.cfg.batch_size)

    def _before_lockstep(self):
        self.optimizer.sync_lr()

    def _lockstep_body(self):
        mem, cur = self.memory, self.cur
        out = self.net_act.forward(cur, self.N, noisy=True)
        ops.select_eps_greedy(out[:, :self.action_dim], 0.0, action=self.action)
        obs, r, te, tr, nobs = self.env.step(self.action, done=self.done)
        mem.store(cur, self.action.view(-1, 1), r, nobs, self.done)   # done = terminated or truncated (ref :263-265)
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
        if len(self.memory) >= int(self.cfg.batch_size):
            self._after_update()

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
                    sps = (step + 1) * self.N / max(time.time() - t0, 1e-9)
                    print(f"Episodes {total} | Avg(100): {avg:.1f} | Loss: {self.loss_acc[0].item():.4f} | {sps:,.0f} steps/s")
                    if avg >= 495.0 and total >= 100:
                        print(f"\nEnvironment solved in {total} episodes!")
                        break
        print("Training completed!")

    def eval(self, num_episodes: int = 10):
        """Deterministic (mu-only) episodes, one env copy per episode, stepped in lockstep (ref eval)."""
        print(f"\nEvaluating for {num_episodes} episodes...")
        env = ops.VecEnv(self.cfg.env_name, num_episodes, seed=self.seed + 4242, first_env_id=1 << 32)
        eng = NoisyDuelingEngine(self.fp, self.state_dim, self.action_dim, self.cfg.hidden_dim, num_episodes, False, self.seed, 6)
        action = torch.zeros(num_episodes, device=self.device, dtype=i32)
        obs = env.reset()
        ret = torch.zeros(num_episodes, device=self.device, dtype=f64)
        alive = torch.ones(num_episodes, device=self.device, dtype=torch.bool)
        for _ in range(env.max_episode_steps):
            out = eng.forward(obs, num_episodes, noisy=False)
            ops.select_eps_greedy(out[:, :self.action_dim], 0.0, action=action)
            obs, r, te, tr, _ = env.step(action, want_next_obs=False)
            ret += torch.where(alive, r.double(), torch.zeros_like(ret))
            alive &= ~((te | tr).bool())
            if not bool(alive.any()):
                break
        rewards = ret.tolist()
        for i, r in enumerate(rewards):
            print(f"  Episode {i + 1}: Reward = {r:.1f}")
        print(f"Evaluation Results: Mean = {np.mean(rewards):.1f} +/- {np.std(rewards):.1f}")
        env.close()
        return rewards

    def test(self):
        self.eval(num_episodes=5)
        print("\n(visual test skipped: the device env has no renderer)")


def main():
    config = Config()
    config.num_envs = 1024
    config.batch_size = 1024
    config.memory_capacity = 1 << 18
    trainer = NoisyDQNTrainer(config)
    try:
        trainer.train()
    except KeyboardInterrupt:
        print("\nTraining interrupted.")
    trainer.test()


if __name__ == "__main__":
    main()
