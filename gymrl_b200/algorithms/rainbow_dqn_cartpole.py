"""Rainbow DQN for CartPole-v1 on the B200 engine — same surface as the reference
``algorithms/rainbow_dqn_cartpole.py`` (Config, NoisyLinear, DuelingNoisyNetwork, SumTree,
PrioritizedNStepBuffer, RainbowDQNTrainer.train/eval/test/update/select_action).

    NoisyLinear / dueling head (ref :51-113) -> factorised-noise kernel, W = mu + sigma*outer composed on device, both heads
                                                as ONE [A+1, H] GEMM; dueling aggregation fused into the loss kernel
    PrioritizedNStepBuffer   (ref :155-264) -> per-env n-step windows + SoA ring + float64 sum-tree with the reference's
                                                exact heap layout; stratified sampling and IS weights on device
    update                   (ref :311-361) -> double-Q target, IS-weighted TD loss, priorities written back from td BEFORE
                                                the backward pass (q6), global-norm clip 10 + Adam, Polyak tau, LR decay
Vectorisation: N envs in lockstep, one update of ``batch_size`` per lockstep; ``total_steps`` counts lockstep
select_action calls exactly like the reference counts single-env calls.
"""
from __future__ import annotations

import math
import time
from collections import deque

import numpy as np
import torch
import torch.nn as nn

from .. import _ffi, ops, ops_offpolicy as off
from ..mlp import Chain
from ..nn import FlatParams, FusedAdam

f32, i32, u8, f64 = torch.float32, torch.int32, torch.uint8, torch.float64
RELU, NONE = _ffi.ACT_RELU, _ffi.ACT_NONE


class Config:
    def __init__(self):
        self.env_name = "CartPole-v1"
        self.seed = None
        self.max_episodes = 500
        self.max_steps = 500
        self.batch_size = 256
        self.gamma = 0.9
        self.tau = 0.005
        self.lr = 1e-3
        self.memory_capacity = 20000
        self.hidden_dim = 256
        self.n_steps = 5
        self.alpha = 0.6
        self.beta_init = 0.4
        self.grad_clip = 10.0
        self.device = "cuda"
        # ---- engine extras ----
        self.num_envs = 1
        self.max_locksteps = None
        self.use_cuda_graph = True   # train(): act -> env step -> n-step/PER store -> update as ONE captured graph per lockstep


class NoisyLinear(nn.Module):
    """Parameter container with the reference's names/initialisation (ref :51-74); the math runs in the kernels."""

    def __init__(self, in_features: int, out_features: int, sigma_init: float = 0.5):
        super().__init__()
        self.in_features, self.out_features, self.sigma_init = in_features, out_features, sigma_init
        self.weight_mu = nn.Parameter(torch.empty(out_features, in_features))
        self.weight_sigma = nn.Parameter(torch.empty(out_features, in_features))
        self.register_buffer("weight_epsilon", torch.zeros(out_features, in_features))
        self.bias_mu = nn.Parameter(torch.empty(out_features))
        self.bias_sigma = nn.Parameter(torch.empty(out_features))
        self.register_buffer("bias_epsilon", torch.zeros(out_features))
        mu_range = 1 / math.sqrt(in_features)
        self.weight_mu.data.uniform_(-mu_range, mu_range)
        self.bias_mu.data.uniform_(-mu_range, mu_range)
        self.weight_sigma.data.fill_(sigma_init / math.sqrt(in_features))
        self.bias_sigma.data.fill_(sigma_init / math.sqrt(out_features))


class DuelingNoisyNetwork(nn.Module):
    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int = 256):
        super().__init__()
        self.fc1 = nn.Linear(state_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.advantage = NoisyLinear(hidden_dim, action_dim)
        self.value = NoisyLinear(hidden_dim, 1)

    TRUNK = [("fc1.weight", "fc1.bias", RELU), ("fc2.weight", "fc2.bias", RELU)]


class DuelingEngine:
    """Forward/backward of DuelingNoisyNetwork: trunk chain + one composed [A+1, H] head GEMM."""

    def __init__(self, net: DuelingNoisyNetwork, fp: FlatParams, A: int, H: int, M: int, backward: bool, seed: int, entity: int):
        dev = fp.flat.device
        self.fp, self.A, self.H, self.M, self.seed, self.entity = fp, A, H, M, seed, entity
        self.trunk = Chain.from_names(fp, DuelingNoisyNetwork.TRUNK, M, backward)
        z = lambda *s: torch.zeros(*s, device=dev, dtype=f32)
        self.W, self.b = z(A + 1, H), z(A + 1)            # composed head (rows :A advantage, row A value)
        self.eps_in_a, self.eps_out_a, self.eps_in_v, self.eps_out_v = z(H), z(A), z(H), z(1)
        self.out = z(M, A + 1)
        self.draws = 0
        self.ctr = torch.zeros(1, device=dev, dtype=i32)   # device mirror of `draws`: captured graphs draw fresh noise per replay
        if backward:
            self.dout = z(M, A + 1)
            self.dW, self.db = z(A + 1, H), z(A + 1)
            self.ws = torch.empty(ops.backward_weight_workspace(M, A + 1, H), device=dev, dtype=torch.uint8)
            from ..graphs import Branches
            self._br = Branches(1)

    def _layers(self):
        fp, A = self.fp, self.A
        return [dict(w_mu=fp.p("advantage.weight_mu"), w_sigma=fp.p("advantage.weight_sigma"), b_mu=fp.p("advantage.bias_mu"),
                     b_sigma=fp.p("advantage.bias_sigma"), eps_in=self.eps_in_a, eps_out=self.eps_out_a, w=self.W[:A], b=self.b[:A]),
                dict(w_mu=fp.p("value.weight_mu"), w_sigma=fp.p("value.weight_sigma"), b_mu=fp.p("value.bias_mu"),
                     b_sigma=fp.p("value.bias_sigma"), eps_in=self.eps_in_v, eps_out=self.eps_out_v, w=self.W[A:], b=self.b[A:])]

    def compose(self, noisy: bool, xi=None):
        """reset_noise() + the weight composition of both NoisyLinear heads: one fused launch (gymrl_noisy_refresh), or the
        separate sample / compose entry points when pre-drawn normals are given.
        xi: optional dict of pre-drawn normals {in_a, out_a, in_v, out_v} (parity tests)."""
        fp, A = self.fp, self.A
        if xi is None:
            if noisy:
                self.draws += 1
            off.noisy_refresh(self._layers(), self.H, noisy=noisy, seed=self.seed, entity_base=self.entity * 16, entity_stride=4096,
                              draw=1, draw_base=self.ctr if noisy else None, counter_inc=1 if noisy else 0)
            return
        self.draws += 1
        for k, (eps, ent) in enumerate(((self.eps_in_a, 0), (self.eps_out_a, 1), (self.eps_in_v, 2), (self.eps_out_v, 3))):
            key = ("in_a", "out_a", "in_v", "out_v")[k]
            off.noisy_sample(eps, xi[key], seed=self.seed, entity=self.entity * 16 + ent * 4096, draw=1, draw_base=self.ctr)
        ops.counter_add(self.ctr, 1)
        off.noisy_compose(fp.p("advantage.weight_mu"), fp.p("advantage.weight_sigma"), self.eps_in_a, self.eps_out_a,
                          fp.p("advantage.bias_mu"), fp.p("advantage.bias_sigma"), self.W[:A], self.b[:A])
        off.noisy_compose(fp.p("value.weight_mu"), fp.p("value.weight_sigma"), self.eps_in_v, self.eps_out_v,
                          fp.p("value.bias_mu"), fp.p("value.bias_sigma"), self.W[A:], self.b[A:])

    def forward(self, x, M, row_index=None, noisy=True, xi=None):
        self.compose(noisy, xi)
        h = self.trunk.forward(x, M, row_index=row_index)
        return ops.linear_forward(h, self.W, self.b, NONE, out=self.out, M=M)

    def backward(self, x, M, row_index=None):
        """Given self.dout = dL/d[adv | value].  The head's parameter gradients do not feed the trunk's backward pass: two branches."""
        fp, A = self.fp, self.A

        def head_grads():
            ops.linear_backward_weight(self.dout, self.trunk.out, self.dW, self.db, workspace=self.ws, M=M)
            off.noisy_backward2(dict(dw=self.dW[:A], db=self.db[:A], eps_in=self.eps_in_a, eps_out=self.eps_out_a,
                                     dw_mu=fp.g("advantage.weight_mu"), dw_sigma=fp.g("advantage.weight_sigma"),
                                     db_mu=fp.g("advantage.bias_mu"), db_sigma=fp.g("advantage.bias_sigma")),
                                dict(dw=self.dW[A:], db=self.db[A:], eps_in=self.eps_in_v, eps_out=self.eps_out_v,
                                     dw_mu=fp.g("value.weight_mu"), dw_sigma=fp.g("value.weight_sigma"),
                                     db_mu=fp.g("value.bias_mu"), db_sigma=fp.g("value.bias_sigma")), self.H)

        def trunk():
            ops.linear_backward_input(self.dout[:M], self.W, self.trunk.out, RELU, out=self.trunk.dout)
            self.trunk.backward(x, M, row_index=row_index)

        self._br.run(trunk, head_grads)


class SumTree(off.DeviceSumTree):
    """Reference-compatible facade (update(i, p) / get_index(v) / priority_sum / priority_max) over the device tree."""

    def __init__(self, capacity: int, device=None):
        super().__init__(capacity, device or torch.device("cuda", torch.cuda.current_device()))
        self._ring1 = torch.tensor([0, 1], device=self.tree.device, dtype=i32)
        self._beta0 = torch.zeros(1, device=self.tree.device, dtype=f64)

    def update(self, data_index, priority=None, **kw):
        if isinstance(data_index, torch.Tensor):
            return super().update(data_index, priority, **kw)
        dev = self.tree.device
        super().update(torch.tensor([int(data_index)], device=dev, dtype=i32), torch.tensor([float(priority)], device=dev, dtype=f64))

    def get_index(self, v: float):
        u = torch.tensor([float(v)], device=self.tree.device, dtype=f64)
        prio = torch.zeros(1, device=self.tree.device, dtype=f64)
        idx, _ = self.sample(1, self._ring1, self._beta0, uniforms=u, out_prio=prio, raw_values=True)
        return int(idx.item()), float(prio.item())


class PrioritizedNStepBuffer:
    def __init__(self, config: Config, state_dim: int, num_envs: int = 1, device=None):
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.device, self.capacity, self.batch_size = dev, int(config.memory_capacity), int(config.batch_size)
        self.n_steps, self.gamma, self.alpha = config.n_steps, config.gamma, config.alpha
        self.beta, self.beta_init = config.beta_init, config.beta_init
        self.N = num_envs
        self.sum_tree = SumTree(self.capacity, dev)
        self.ring = off.ReplayRing(self.capacity, state_dim, 1, True, dev)     # ring.done holds `terminal`
        self.window = off.NStepWindow(num_envs, state_dim, self.n_steps, dev)
        self.beta_t = torch.zeros(1, device=dev, dtype=f64)
        self.batch_index = torch.zeros(self.batch_size, device=dev, dtype=i32)
        self.is_weight = torch.zeros(self.batch_size, device=dev, dtype=f32)

    def store_lockstep(self, obs, action, reward, next_obs, terminal_u8, done_u8, truncated_u8=None):
        """terminal_u8 = the reference's `terminal`; or pass the env's terminated flags plus truncated_u8 and the kernel forms
        terminated & ~truncated itself."""
        full = self.window.push(obs, action, reward, next_obs, terminal_u8, done_u8, self.gamma, self.ring, self.ring.done,
                                trunc=truncated_u8)
        if full:
            self.sum_tree.store_new(self.N, self.ring.state)
            self.ring.advance(self.N)

    def store_transition(self, state, action, reward, next_state, terminal, done):
        """Single-env facade with the reference signature (ref :179-205)."""
        dev = self.device
        t = lambda x, dt: torch.as_tensor(np.asarray(x), device=dev).to(dt)
        self.store_lockstep(t(state, f32).reshape(1, -1), t([action], i32), t([reward], f32), t(next_state, f32).reshape(1, -1),
                            t([bool(terminal)], u8), t([bool(done)], u8))

    def set_beta(self, total_steps: int, max_train_steps: int):
        """IS exponent schedule (ref :222): host value -> device scalar (outside any captured graph)."""
        self.beta = self.beta_init + (1 - self.beta_init) * (total_steps / max_train_steps)
        self.beta_t.fill_(self.beta)

    def sample(self, total_steps: int, max_train_steps: int, uniforms=None, seed=0, draw=0, draw_base=None, set_beta=True):
        if set_beta:
            self.set_beta(total_steps, max_train_steps)
        self.sum_tree.sample(self.batch_size, self.ring.state, self.beta_t, uniforms=uniforms, out_idx=self.batch_index,
                             out_w=self.is_weight, seed=seed, draw=draw, draw_base=draw_base)
        return self.batch_index, self.is_weight

    def update_priorities(self, batch_index, td_errors):
        self.sum_tree.update(batch_index, td_error=td_errors, eps=0.01, alpha=self.alpha)

    def __len__(self) -> int:
        return len(self.ring)


class RainbowDQNTrainer:
    def __init__(self, config: Config):
        _ffi.require_cuda()
        self.cfg = cfg = config
        self.device = dev = torch.device("cuda", torch.cuda.current_device())
        self.N = N = int(cfg.num_envs)
        self.seed = int(cfg.seed) if cfg.seed is not None else int(time.time_ns() & 0x7FFFFFFF)
        self.env = ops.VecEnv(cfg.env_name, N, seed=self.seed)
        self.state_dim, self.action_dim = D, A = self.env.obs_dim, self.env.n_actions
        self.max_steps_per_episode = self.env.max_episode_steps
        self.max_train_steps = self.max_steps_per_episode * cfg.max_episodes
        H, B = cfg.hidden_dim, int(cfg.batch_size)
        self.policy_net = DuelingNoisyNetwork(D, A, H).to(dev)
        self.target_net = DuelingNoisyNetwork(D, A, H).to(dev)
        self.target_net.load_state_dict(self.policy_net.state_dict())
        self.target_net.eval()
        self.fp, self.fp_t = FlatParams(self.policy_net, device=dev), FlatParams(self.target_net, device=dev)
        self.optimizer = FusedAdam(self.fp, lr=cfg.lr)
        self.net_act = DuelingEngine(self.policy_net, self.fp, A, H, N, False, self.seed, 1)
        self.net_upd = DuelingEngine(self.policy_net, self.fp, A, H, B, True, self.seed, 2)
        self.net_nxt = DuelingEngine(self.policy_net, self.fp, A, H, B, False, self.seed, 3)
        self.net_tgt = DuelingEngine(self.target_net, self.fp_t, A, H, B, False, self.seed, 4)
        self.memory = PrioritizedNStepBuffer(cfg, D, N, dev)
        from ..graphs import Branches
        self.branches = Branches(2)
        self.td = torch.zeros(B, device=dev, dtype=f32)
        self.loss_acc = torch.zeros(2, device=dev, dtype=f32)
        self.action = torch.zeros(N, device=dev, dtype=i32)
        self.done = torch.zeros(N, device=dev, dtype=u8)
        self.total_steps = 0
        self.update_count = 0
        self.ctr_upd = torch.zeros(1, device=dev, dtype=i32)   # device mirror of update_count (PER sampling draw)
        self.cur = torch.zeros(N, D, device=dev, dtype=f32)
        self._g_lockstep = None
        self.graph_launches = 0
        self.episode_rewards = deque(maxlen=100)
        print(f"Device: {dev}")
        print(f"State dim: {D}, Action dim: {A}")

    def act(self, obs: torch.Tensor, deterministic: bool = False) -> torch.Tensor:
        if not deterministic:
            self.total_steps += 1
        out = self.net_act.forward(obs, self.N, noisy=not deterministic)
        # argmax(V + A - mean A) == argmax A : the greedy kernel runs on the advantage columns
        return ops.select_eps_greedy(out[:, :self.action_dim], 0.0, action=self.action)

    @torch.no_grad()
    def select_action(self, state: np.ndarray, deterministic: bool = False) -> int:
        obs = torch.as_tensor(np.asarray(state, np.float32), device=self.device).reshape(1, -1)
        eng = getattr(self, "_net_one", None) or DuelingEngine(self.policy_net, self.fp, self.action_dim, self.cfg.hidden_dim, 1, False,
                                                               self.seed, 5)
        self._net_one = eng
        if not deterministic:
            self.total_steps += 1
        out = eng.forward(obs, 1, noisy=not deterministic)
        return int(ops.select_eps_greedy(out[:, :self.action_dim], 0.0).item())

    def update(self, uniforms=None, xi_next=None, xi_cur=None):
        cfg, B, mem = self.cfg, int(self.cfg.batch_size), self.memory
        if len(mem) < B:
            return 0.0
        self.update_count += 1
        mem.set_beta(self.total_steps, self.max_train_steps)
        self.optimizer.sync_lr()
        self._update_device(uniforms, xi_next, xi_cur)
        self._schedule_lr()
        return self.loss_acc[0]

    def _schedule_lr(self):
        cfg = self.cfg
        lr_now = 0.9 * cfg.lr * (1 - self.total_steps / self.max_train_steps) + 0.1 * cfg.lr
        for g in self.optimizer.param_groups:
            g["lr"] = lr_now

    def _update_device(self, uniforms=None, xi_next=None, xi_cur=None):
        """The device side of update() (ref :311-361): capture-safe — beta, the learning rate and the draw counter are device
        scalars set outside."""
        cfg, B, A, mem = self.cfg, int(self.cfg.batch_size), self.action_dim, self.memory
        ring = mem.ring
        idx, w = mem.sample(self.total_steps, self.max_train_steps, uniforms=uniforms, seed=self.seed, draw=1, draw_base=self.ctr_upd,
                            set_beta=False)
        # three independent forwards (online net on s' with fresh noise (q8), target net in eval mode: mu only, online net on s)
        # as parallel branches
        qo, qt, q = self.branches.run(lambda: self.net_nxt.forward(ring.next_obs, B, row_index=idx, noisy=True, xi=xi_next),
                                      lambda: self.net_tgt.forward(ring.next_obs, B, row_index=idx, noisy=False),
                                      lambda: self.net_upd.forward(ring.obs, B, row_index=idx, noisy=True, xi=xi_cur))
        self.loss_acc.zero_()
        off.dqn_loss(q[:, :A], qt[:, :A], ring.action, ring.reward, ring.done, cfg.gamma ** cfg.n_steps, v=q[:, A:], vnext_target=qt[:, A:],
                     qnext_online=qo[:, :A], vnext_online=qo[:, A:], row_index=idx, is_weight=w, dq=self.net_upd.dout[:, :A],
                     dv=self.net_upd.dout[:, A:], td_error=self.td, loss_acc=self.loss_acc)
        # priorities before backward (ref :340); the tree write-back and the backward pass touch disjoint data: two branches
        self.branches.run(lambda: self.net_upd.backward(ring.obs, B, row_index=idx), lambda: mem.update_priorities(idx, self.td))
        self.optimizer.launch(max_norm=cfg.grad_clip)
        ops.polyak(self._target_params(), self._policy_params(), cfg.tau)
        ops.counter_add(self.ctr_upd, 1)

    # ---------------------------------------------------------------- one lockstep (ref train() loop body :371-388)
    def _lockstep_body(self):
        """act -> env.step -> n-step window / PER store -> update -> carry the observation.  Capture-safe once the n-step
        window is full and the ring holds a batch."""
        env, mem, cur, A = self.env, self.memory, self.cur, self.action_dim
        out = self.net_act.forward(cur, self.N, noisy=True)
        a = ops.select_eps_greedy(out[:, :A], 0.0, action=self.action)
        obs, r, te, tr, nobs = env.step(a, done=self.done)
        # time-limit truncation is not terminal (ref :376): the push kernel stores te & ~tr
        mem.store_lockstep(cur, a, r, nobs, te, self.done, truncated_u8=tr)
        self._update_device()
        cur.copy_(obs)

    def lockstep(self):
        """One lockstep of all N envs: eager while the n-step window / the ring are filling, then one CUDA-graph replay per
        lockstep with the host-scheduled scalars (PER beta, learning rate) written to their device slots first."""
        from ..graphs import capture
        cfg, mem, B = self.cfg, self.memory, int(self.cfg.batch_size)
        ready = len(mem) >= B and mem.window.pushed_host >= mem.n_steps
        if not getattr(cfg, "use_cuda_graph", True) or not ready:
            cur = self.cur
            a = self.act(cur)
            obs, r, te, tr, nobs = self.env.step(a, done=self.done)
            mem.store_lockstep(cur, a, r, nobs, te, self.done, truncated_u8=tr)
            self.update()
            cur.copy_(obs)
            return
        self.total_steps += 1
        self.update_count += 1
        mem.set_beta(self.total_steps, self.max_train_steps)
        self.optimizer.sync_lr()
        if self._g_lockstep is None:
            mirrors = (mem.window.pushed_host, mem.ring._size_host, [e.draws for e in self._engines()])
            self._g_lockstep = capture(self._lockstep_body)   # the capture's eager pass is this lockstep
            mem.window.pushed_host, mem.ring._size_host = mirrors[0], mirrors[1]
            for e, d in zip(self._engines(), mirrors[2]):
                e.draws = d
        else:
            self._g_lockstep.replay()
            self.graph_launches += self._g_lockstep.n_kernels
        # host mirrors of the counters the device body advanced
        mem.window.pushed_host += 1
        mem.ring._size_host = min(mem.ring.capacity, mem.ring._size_host + self.N)
        for e in (self.net_act, self.net_nxt, self.net_upd):
            e.draws += 1
        self._schedule_lr()

    def _engines(self):
        return (self.net_act, self.net_upd, self.net_nxt, self.net_tgt)

    # Polyak touches parameters only; both nets share the flat layout so one kernel covers them all (buffers are not in `flat`)
    def _policy_params(self):
        return self.fp.flat

    def _target_params(self):
        return self.fp_t.flat

    def train(self):
        print("Starting training...")
        cfg, env, mem = self.cfg, self.env, self.memory
        env.reset(out=self.cur)
        max_lock = cfg.max_locksteps or int(cfg.max_episodes * cfg.max_steps / self.N)
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
        print(f"\nEvaluating for {num_episodes} episodes...")
        env = ops.VecEnv(self.cfg.env_name, num_episodes, seed=self.seed + 999, first_env_id=1 << 32)
        eng = DuelingEngine(self.policy_net, self.fp, self.action_dim, self.cfg.hidden_dim, num_episodes, False, self.seed, 6)
        obs = env.reset()
        ret = torch.zeros(num_episodes, device=self.device, dtype=f64)
        alive = torch.ones(num_episodes, device=self.device, dtype=torch.bool)
        for _ in range(env.max_episode_steps):
            out = eng.forward(obs, num_episodes, noisy=False)
            a = ops.select_eps_greedy(out[:, :self.action_dim], 0.0)
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
    config.num_envs = 8192
    config.memory_capacity = 1 << 21
    config.batch_size = 8192
    trainer = RainbowDQNTrainer(config)
    trainer.train()
    trainer.test()
