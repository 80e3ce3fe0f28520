"""PPO for LunarLander-v3 on the B200 engine — same surface as the reference script
``algorithms/ppo_lunarlander.py`` (Config, ActorCritic, RolloutBuffer, PPOTrainer with
train()/eval()/test()/update()/collect_rollout()/compute_gae()), but every hot loop is a CUDA kernel
behind the C ABI:

    collect_rollout  (ref :198-231)  ->  T lockstep iterations over N env copies, captured as ONE CUDA graph:
                                         5 dense-layer kernels -> categorical sample -> env step (auto-reset)
    compute_gae      (ref :179-196)  ->  gymrl_gae (chunked fp64 scan over [T][N])
    update           (ref :233-330)  ->  per minibatch (one captured graph, replayed epochs x minibatches times):
                                         window of the device permutation -> forward (row gather fused) ->
                                         fused loss/grad -> backward -> global-norm clip + Adam
    LR anneal        (ref :337-341)  ->  host writes one float64 device scalar per update

Vectorisation knobs are *extra* Config attributes whose defaults reproduce the reference at
``num_envs = 1`` (T = update_freq = 2048 steps, 32 minibatches of 64, re-reset at every rollout):
    num_envs, num_steps (= update_freq // num_envs), num_minibatches (= N*T // batch_size),
    reset_each_rollout (= num_envs == 1), use_cuda_graph.
Multi-GPU (one process per GPU): env shards by global env id, one NCCL sum-all-reduce of
the flat gradient per optimizer step and a 3-scalar all-reduce for the advantage normalisation.
"""
from __future__ import annotations

import signal
import sys
import os
import time
from collections import deque
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import _ffi, dist as gdist, ops
from ..nn import FlatParams, FusedAdam, layer_init

f32, i32, u8, f64 = torch.float32, torch.int32, torch.uint8, torch.float64


class Config:
    def __init__(self):
        self.env_name = "LunarLander-v3"
        self.seed = None

        self.max_train_steps = 1_000_000
        self.update_freq = 2048
        self.num_epochs = 10
        self.batch_size = 64

        self.gamma = 0.99
        self.gae_lambda = 0.95
        self.clip_eps = 0.2
        self.dual_clip = 3.0
        self.entropy_coef = 0.01
        self.value_coef = 0.5
        self.max_grad_norm = 0.5

        self.lr = 3e-4
        self.anneal_lr = True

        self.hidden_dim = 256

        self.device = "cuda"  # there is no CPU path: the engine *is* the CUDA library

        # ---- engine extras (defaults reproduce the reference at num_envs = 1) ----
        self.num_envs = 1
        self.num_steps = None
        self.num_minibatches = None
        self.reset_each_rollout = None
        self.use_cuda_graph = True
        self.dist_epoch_graph = True   # multi-GPU: capture the per-minibatch gradient reduction inside the epoch graph
        self.peer_reduce = True        # multi-GPU: one-shot NVLink peer-memory reduction (csrc/comm.cu) instead of ncclAllReduce
        self.fused_heads = True   # heads + loss + heads backward as one kernel where the shape allows (H in {128,256}, 4 actions)


class ActorCritic(nn.Module):
    """Same module tree / state_dict keys as the reference ActorCritic (ref :63-90)."""

    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int = 256):
        super().__init__()
        self.state_dim, self.action_dim, self.hidden_dim = state_dim, action_dim, hidden_dim
        self.shared = nn.Sequential(
            layer_init(nn.Linear(state_dim, hidden_dim)), nn.Tanh(),
            layer_init(nn.Linear(hidden_dim, hidden_dim)), nn.Tanh())
        self.actor = nn.Sequential(
            layer_init(nn.Linear(hidden_dim, hidden_dim)), nn.Tanh(),
            layer_init(nn.Linear(hidden_dim, action_dim), std=0.01))
        self.critic = nn.Sequential(
            layer_init(nn.Linear(hidden_dim, hidden_dim)), nn.Tanh(),
            layer_init(nn.Linear(hidden_dim, 1), std=1.0))

    # flat layout: the two head-trunk layers are adjacent so they run as one [2H, H] GEMM
    PARAM_ORDER = ["shared.0.weight", "shared.0.bias", "shared.2.weight", "shared.2.bias",
                   "actor.0.weight", "critic.0.weight", "actor.0.bias", "critic.0.bias",
                   "actor.2.weight", "actor.2.bias", "critic.2.weight", "critic.2.bias"]

    def to_engine(self, device) -> "ActorCriticEngine":
        return ActorCriticEngine(self, device)


class _Acts:
    """Activation / gradient scratch for a fixed maximum batch M."""

    def __init__(self, M: int, H: int, A: int, device, backward: bool):
        self.M = M
        self.h1 = torch.empty(M, H, device=device, dtype=f32)
        self.h2 = torch.empty(M, H, device=device, dtype=f32)
        self.ac = torch.empty(M, 2 * H, device=device, dtype=f32)
        self.lv = torch.empty(M, 8, device=device, dtype=f32)  # [:, :A] logits, [:, A] value (A <= 7)
        if backward:
            self.dlv = torch.zeros(M, 8, device=device, dtype=f32)
            self.dac = torch.empty(M, 2 * H, device=device, dtype=f32)
            self.dh2 = torch.empty(M, H, device=device, dtype=f32)
            self.dh1 = torch.empty(M, H, device=device, dtype=f32)


class ActorCriticEngine:
    """Forward / backward of ActorCritic through the C ABI dense-layer kernels."""

    def __init__(self, model: ActorCritic, device):
        self.model = model.to(device)
        self.fp = FlatParams(self.model, ActorCritic.PARAM_ORDER, device)
        self.H, self.A, self.D = model.hidden_dim, model.action_dim, model.state_dim
        assert self.A <= 7
        fp, H = self.fp, self.H
        self.W1, self.b1 = fp.p("shared.0.weight"), fp.p("shared.0.bias")
        self.W2, self.b2 = fp.p("shared.2.weight"), fp.p("shared.2.bias")
        self.Wac = fp.span("actor.0.weight", "critic.0.weight", 2 * H, H)
        self.bac = fp.span("actor.0.bias", "critic.0.bias", 1, 2 * H).view(2 * H)
        self.Wa2, self.ba2 = fp.p("actor.2.weight"), fp.p("actor.2.bias")
        self.Wc2, self.bc2 = fp.p("critic.2.weight"), fp.p("critic.2.bias")
        self.gW1, self.gb1 = fp.g("shared.0.weight"), fp.g("shared.0.bias")
        self.gW2, self.gb2 = fp.g("shared.2.weight"), fp.g("shared.2.bias")
        self.gWac = fp.span("actor.0.weight", "critic.0.weight", 2 * H, H, grad=True)
        self.gbac = fp.span("actor.0.bias", "critic.0.bias", 1, 2 * H, grad=True).view(2 * H)
        self.gWa2, self.gba2 = fp.g("actor.2.weight"), fp.g("actor.2.bias")
        self.gWc2, self.gbc2 = fp.g("critic.2.weight"), fp.g("critic.2.bias")
        # the two 256-wide trunk layers are TMA-fed from pre-split tf32 weight images (csrc/wimages.cu, linear_tc.cu ws kernel)
        if H % 256 == 0:
            fp.enable_weight_images([("shared.2.weight", "shared.2.weight", H, H), ("actor.0.weight", "critic.0.weight", 2 * H, H)])
        self.workspace = None
        self.ws_layers = None
        self.deferred_reduce = True   # forward_trunk/backward may run inside an ops.reduce_defer_begin() scope
        # heads + loss + heads backward as one sweep (gymrl_ppo_heads_fused) where the kernel is built for the shape
        self.can_fuse_heads = self.A == 4 and self.H in (128, 256)
        self.can_fuse_rollout_tail = self.A <= 7 and self.H % 4 == 0     # gymrl_policy_heads_sample (PPO rollout tail)
        # GYMRL_PPO_BRANCHES=1: dW kernels of the trunk backward on a side stream / graph branch (see backward_trunk)
        self.branches = None
        if os.environ.get("GYMRL_PPO_BRANCHES", "0") == "1":
            from ..graphs import Branches
            self.branches = Branches(1)

    def alloc_workspace(self, M: int):
        need = max(ops.backward_weight_workspace(M, 2 * self.H, self.H), ops.backward_weight_workspace(M, self.H, self.H),
                   ops.backward_weight_workspace(M, self.H, self.D), ops.backward_weight_workspace(M, self.A, self.H))
        if self.can_fuse_heads:
            need = max(need, ops.ppo_heads_workspace_bytes(self.H, self.A))
        self.workspace = torch.empty(need, device=self.fp.flat.device, dtype=torch.uint8)
        # one workspace per backward call of a minibatch, so their partial sums can be folded together at the end
        # (ops.reduce_defer_begin / reduce_flush): [heads (or actor head), critic head, head trunks, shared.2, shared.0]
        dev = self.fp.flat.device
        heads = max(ops.backward_weight_workspace(M, self.A, self.H), ops.ppo_heads_workspace_bytes(self.H, self.A) if self.can_fuse_heads else 0)
        sizes = [heads, ops.backward_weight_workspace(M, 1, self.H), ops.backward_weight_workspace(M, 2 * self.H, self.H),
                 ops.backward_weight_workspace(M, self.H, self.H), ops.backward_weight_workspace(M, self.H, self.D)]
        self.ws_layers = [torch.empty(s, device=dev, dtype=torch.uint8) for s in sizes]

    def forward_trunk(self, x: torch.Tensor, acts: _Acts, M: int, row_index: Optional[torch.Tensor] = None):
        """acts.ac[:M] = (actor | critic) head-trunk activations."""
        T = _ffi.ACT_TANH
        ops.linear_forward(x, self.W1, self.b1, T, row_index=row_index, out=acts.h1, M=M)
        ops.linear_forward(acts.h1, self.W2, self.b2, T, out=acts.h2, M=M)
        ops.linear_forward(acts.h2, self.Wac, self.bac, T, out=acts.ac, M=M)

    def forward(self, x: torch.Tensor, acts: _Acts, M: int, row_index: Optional[torch.Tensor] = None):
        """logits = acts.lv[:M, :A], value = acts.lv[:M, A]."""
        H, A = self.H, self.A
        self.forward_trunk(x, acts, M, row_index)
        ops.linear_forward(acts.ac[:, :H], self.Wa2, self.ba2, _ffi.ACT_NONE, out=acts.lv[:, :A], M=M)
        ops.linear_forward(acts.ac[:, H:], self.Wc2, self.bc2, _ffi.ACT_NONE, out=acts.lv[:, A:A + 1], M=M)
        return acts.lv

    def backward(self, x: torch.Tensor, acts: _Acts, M: int, row_index: Optional[torch.Tensor] = None):
        """Given acts.dlv[:M] = dL/d(logits, value), fill the flat gradient buffer."""
        T, H, A, ws = _ffi.ACT_TANH, self.H, self.A, self.ws_layers
        dl, dv = acts.dlv[:, :A], acts.dlv[:, A:A + 1]
        ops.linear_backward(dl, acts.ac[:, :H], self.Wa2, self.gWa2, self.gba2, dx=acts.dac[:, :H], act_in=T, workspace=ws[0], M=M)
        ops.linear_backward(dv, acts.ac[:, H:], self.Wc2, self.gWc2, self.gbc2, dx=acts.dac[:, H:], act_in=T, workspace=ws[1], M=M)
        self.backward_trunk(x, acts, M, row_index)

    def backward_trunk(self, x: torch.Tensor, acts: _Acts, M: int, row_index: Optional[torch.Tensor] = None):
        """Given acts.dac[:M] = dL/d(pre-activation of the head trunks), fill the gradients of the three trunk layers."""
        T, ws = _ffi.ACT_TANH, self.ws_layers
        br = getattr(self, "branches", None)
        if br is not None:
            # dW of a layer is independent of its dX chain: two graph branches (the kernels fill the chip on their own, so what
            # overlaps is one kernel's tail / epilogue with the next one's ramp-up; opt-in, see GYMRL_PPO_BRANCHES)
            br.run(lambda: ops.linear_backward_input(acts.dac[:M], self.Wac, acts.h2[:M], T, out=acts.dh2),
                   lambda: ops.linear_backward_weight(acts.dac, acts.h2, self.gWac, self.gbac, workspace=ws[2], M=M))

            def rest():
                ops.linear_backward_input(acts.dh2[:M], self.W2, acts.h1[:M], T, out=acts.dh1)
                ops.linear_backward(acts.dh1, x, self.W1, self.gW1, self.gb1, row_index=row_index, workspace=ws[4], M=M)
            br.run(rest, lambda: ops.linear_backward_weight(acts.dh2, acts.h1, self.gW2, self.gb2, workspace=ws[3], M=M))
            return
        ops.linear_backward(acts.dac, acts.h2, self.Wac, self.gWac, self.gbac, dx=acts.dh2, act_in=T, workspace=ws[2], M=M)
        ops.linear_backward(acts.dh2, acts.h1, self.W2, self.gW2, self.gb2, dx=acts.dh1, act_in=T, workspace=ws[3], M=M)
        ops.linear_backward(acts.dh1, x, self.W1, self.gW1, self.gb1, row_index=row_index, workspace=ws[4], M=M)

    def heads_loss_backward(self, acts: _Acts, M: int, action, logp_old, adv, ret, loss_cfg, *, row_index=None, metrics=None,
                            lv_out=None):
        """Output heads, PPO loss and heads backward in one sweep over acts.ac: fills acts.dac and the head gradients."""
        ops.ppo_heads_fused(acts.ac, self.Wa2, self.ba2, self.Wc2, self.bc2, action, logp_old, adv, ret, loss_cfg,
                            dh=acts.dac, dWa=self.gWa2, dba=self.gba2, dWc=self.gWc2, dbc=self.gbc2, workspace=self.ws_layers[0],
                            M=M, row_index=row_index, act_in=_ffi.ACT_TANH, lv_out=lv_out, metrics=metrics)


class RolloutBuffer:
    """Device-resident [T][N] SoA rollout store (ref RolloutBuffer :120-154 keeps Python lists).
    obs rows are 32 B ([T+1][N][8]) so the env kernel writes coalesced lines and the minibatch gather
    touches exactly one sector per sample."""

    def __init__(self, T: int, N: int, D: int, device):
        self.T, self.N, self.D = T, N, D
        self.obs = torch.zeros(T + 1, N, D, device=device, dtype=f32)
        self.action = torch.zeros(T, N, device=device, dtype=i32)
        self.log_prob = torch.zeros(T, N, device=device, dtype=f32)
        self.value = torch.zeros(T, N, device=device, dtype=f32)
        self.reward = torch.zeros(T, N, device=device, dtype=f32)
        self.done = torch.zeros(T, N, device=device, dtype=u8)
        self.v_last = torch.zeros(N, device=device, dtype=f32)
        self.adv = torch.zeros(T, N, device=device, dtype=f32)
        self.ret = torch.zeros(T, N, device=device, dtype=f32)
        self.filled = 0

    # list-like views for code written against the reference buffer
    @property
    def states(self): return self.obs[:self.filled].reshape(-1, self.D)
    @property
    def actions(self): return self.action[:self.filled].reshape(-1)
    @property
    def log_probs(self): return self.log_prob[:self.filled].reshape(-1)
    @property
    def values(self): return self.value[:self.filled].reshape(-1)
    @property
    def rewards(self): return self.reward[:self.filled].reshape(-1)
    @property
    def dones(self): return self.done[:self.filled].reshape(-1)

    def clear(self):
        self.filled = 0

    def __len__(self) -> int:
        return self.filled * self.N


class PPOTrainer:
    MAX_EPOCH_GRAPH_MINIBATCHES = 64   # beyond this one minibatch graph is replayed with a device-side window counter

    def __init__(self, config: Config):
        _ffi.require_cuda()
        self.cfg = cfg = config
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.rank, self.world = gdist.info()
        N = int(cfg.num_envs)
        T = int(cfg.num_steps) if cfg.num_steps else max(1, int(cfg.update_freq) // N)
        self.N, self.T = N, T
        self.n_mb = int(cfg.num_minibatches) if cfg.num_minibatches else max(1, (N * T) // int(cfg.batch_size))
        assert (N * T) % self.n_mb == 0, "rollout size must be divisible by the number of minibatches"
        self.mb = (N * T) // self.n_mb
        self.reset_each_rollout = (N == 1) if cfg.reset_each_rollout is None else bool(cfg.reset_each_rollout)
        self.seed = int(cfg.seed) if cfg.seed is not None else int(time.time_ns() & 0x7FFFFFFF)

        self.env = ops.VecEnv(cfg.env_name, N, seed=self.seed, first_env_id=gdist.shard(N, self.rank)[0])
        state_dim, action_dim = self.env.obs_dim, self.env.n_actions
        self.model = self._make_model(state_dim, action_dim)
        gdist.broadcast_module_(self.model, self.device)   # identical replicas: rank 0's initialisation everywhere
        self.net = self.model.to_engine(self.device)
        self.optimizer = FusedAdam(self.net.fp, lr=cfg.lr, eps=1e-5)
        # multi-GPU: the gradient sum is a one-shot NVLink peer-memory reduction fused with the clip's sum of squares
        # (csrc/comm.cu); None on one GPU, with GYMRL_COMM=nccl, or when peer mappings are unavailable (then ncclAllReduce)
        self.comm = gdist.make_peer_reducer(self.net.fp.grad.numel()) if getattr(cfg, "peer_reduce", True) else None
        self.grad_reduced = torch.zeros_like(self.net.fp.grad) if self.comm is not None else None
        self.buffer = RolloutBuffer(T, N, state_dim, self.device)
        self.acts_roll = self._make_acts(N, backward=False)
        self.acts_mb = self._make_acts(self.mb, backward=True)
        self.net.alloc_workspace(self.mb)
        self.perm = torch.zeros(N * T, device=self.device, dtype=i32)
        self.idx_mb = torch.zeros(self.mb, device=self.device, dtype=i32)
        self.metrics = torch.zeros(8, device=self.device, dtype=f32)
        self.adv_sums = torch.zeros(3, device=self.device, dtype=f64)  # sum, sumsq, count
        self.ctr_action = torch.zeros(1, device=self.device, dtype=i32)   # draw counter: action sampling
        self.ctr_perm = torch.zeros(1, device=self.device, dtype=i32)     # draw counter: permutations
        self.ctr_mb = torch.zeros(1, device=self.device, dtype=i32)       # minibatch window within the epoch
        self.term = torch.zeros(N, device=self.device, dtype=u8)
        self.trunc = torch.zeros(N, device=self.device, dtype=u8)
        self.loss_cfg = self._make_loss_cfg()
        self.step_count = 0
        self.episode_rewards = deque(maxlen=100)
        self._episodes_seen = 0
        self._g_roll = None
        self._g_mb = None
        self._g_epoch = None
        self._n_sumsq = 0
        self._idx_cur = self.idx_mb
        self._g_mb_bwd = None
        self._g_opt = None
        self._started = False
        self.graph_launches = 0   # kernels launched through graph replays (native launch_count() sees eager calls only)
        if self.rank == 0:
            print(f"Device: {self.device} ({torch.cuda.get_device_name(self.device)})")
            print(f"State dim: {state_dim}, Action dim: {action_dim}")
            print(f"Model parameters: {sum(p.numel() for p in self.model.parameters()):,}")
            print(f"Envs: {N} x {self.world} GPU(s), rollout {T} steps, {self.n_mb} minibatches of {self.mb}")

    # ---------------------------------------------------------------- construction hooks (ppo_full overrides them)
    def _make_model(self, state_dim: int, action_dim: int) -> nn.Module:
        return ActorCritic(state_dim, action_dim, self.cfg.hidden_dim)

    def _make_acts(self, M: int, backward: bool, n_actions: Optional[int] = None):
        return _Acts(M, self.cfg.hidden_dim, n_actions or self.env.n_actions, self.device, backward)

    def _make_loss_cfg(self):
        cfg = self.cfg
        return _ffi.PPOCfg(mode=_ffi.PPO_DUALCLIP, clip_eps_min=cfg.clip_eps, clip_eps_max=cfg.clip_eps,
                           dual_clip=cfg.dual_clip, value_coef=cfg.value_coef, entropy_coef=cfg.entropy_coef)

    def _sample_step(self, acts, t: int):
        buf, N, A = self.buffer, self.N, self.env.n_actions
        ops.sample_categorical(acts.lv[:, :A], seed=self.seed, first_id=self.rank * N, draw=t, draw_base=self.ctr_action,
                               action=buf.action[t], logp=buf.log_prob[t], value_in=acts.lv[:, A:A + 1],
                               value_out=buf.value[t])

    def _loss_step(self, acts):
        buf, A = self.buffer, self.env.n_actions
        ops.ppo_loss(acts.lv[:, :A], acts.lv[:, A:A + 1], buf.action.view(-1), buf.log_prob.view(-1), buf.adv.view(-1),
                     buf.ret.view(-1), self.loss_cfg, row_index=self._idx_cur, dlogits=acts.dlv[:, :A],
                     dvalue=acts.dlv[:, A:A + 1], metrics=self.metrics)

    # ---------------------------------------------------------------- rollout
    def _rollout_body(self):
        buf, env, net, acts = self.buffer, self.env, self.net, self.acts_roll
        N, A = self.N, self.env.n_actions
        fused_tail = getattr(net, "can_fuse_rollout_tail", False) and getattr(self.cfg, "fused_rollout_tail", True)
        for t in range(self.T):
            # draw index = device counter + t: the counter advances once per rollout, not per step
            if fused_tail:
                # both heads + the categorical sample in one launch (bit-identical to the three it replaces)
                net.forward_trunk(buf.obs[t], acts, N)
                ops.policy_heads_sample(acts.ac, net.Wa2, net.ba2, net.Wc2, net.bc2, n=N, seed=self.seed, first_id=self.rank * N, draw=t,
                                        draw_base=self.ctr_action, action=buf.action[t], logp=buf.log_prob[t], value=buf.value[t])
            else:
                net.forward(buf.obs[t], acts, N)
                self._sample_step(acts, t)
            env.step(buf.action[t], obs=buf.obs[t + 1], reward=buf.reward[t], terminated=self.term, truncated=self.trunc,
                     want_next_obs=False, done=buf.done[t])
        ops.counter_add(self.ctr_action, self.T)
        net.forward(buf.obs[self.T], acts, N)
        buf.v_last.copy_(acts.lv[:, A])

    def collect_rollout(self):
        """T lockstep steps of all N envs (ref collect_rollout :198-231). Returns the bootstrap values [N]."""
        buf = self.buffer
        buf.clear()
        if self.cfg.use_cuda_graph and self._g_roll is None:
            # graph capture needs one eager warm-up pass; it is a real rollout (env + RNG advance) whose
            # transitions are dropped, exactly like the in-flight episode the reference drops at :200
            if not self._started:
                self.env.reset(out=buf.obs[0])
                self._started = True
            self._g_roll = self._capture(self._rollout_body)
        if self.reset_each_rollout or not self._started:
            self.env.reset(out=buf.obs[0])
            self._started = True
        else:
            buf.obs[0].copy_(buf.obs[self.T])
        if self.cfg.use_cuda_graph:
            self._replay(self._g_roll)
        else:
            self._rollout_body()
        buf.filled = self.T
        self.step_count += self.N * self.T * self.world
        return buf.v_last

    # ---------------------------------------------------------------- GAE
    def compute_gae(self, next_value=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """gymrl_gae over the [T][N] buffer (ref compute_gae :179-196).  Returns (advantages, returns) [T, N]."""
        buf = self.buffer
        if next_value is not None:
            nv = torch.as_tensor(next_value, dtype=f32, device=self.device).reshape(-1)
            buf.v_last.copy_(nv.expand(self.N) if nv.numel() == 1 else nv)
        ops.gae(buf.reward, buf.value, buf.v_last, buf.done, self.cfg.gamma, self.cfg.gae_lambda, adv=buf.adv, ret=buf.ret)
        return buf.adv, buf.ret

    # ---------------------------------------------------------------- update
    def _fwd_bwd_body(self, window: Optional[int] = None):
        """Forward + loss + backward of one minibatch.  `window` = static index of the minibatch inside the epoch's
        permutation (the row indices are then a view of self.perm); None = the window the device counter ctr_mb points
        at, copied into idx_mb (lets ONE captured graph walk through the epoch)."""
        buf, net, acts = self.buffer, self.net, self.acts_mb
        A, M = self.env.n_actions, self.mb
        obs_flat = buf.obs[:self.T].view(self.T * self.N, -1)
        if window is None:
            ops.slice_i32(self.idx_mb, self.perm, self.ctr_mb)
            ops.counter_add(self.ctr_mb, 1)
            idx = self.idx_mb
        else:
            idx = self.perm[window * M:(window + 1) * M]
        self._idx_cur = idx
        defer = getattr(net, "deferred_reduce", False)
        if defer:
            ops.reduce_defer_begin()   # the layers' partial-gradient folds run as one launch at the end
        self._n_sumsq = 0
        try:
            if getattr(self.cfg, "fused_heads", True) and getattr(net, "can_fuse_heads", False):
                net.forward_trunk(obs_flat, acts, M, row_index=idx)
                net.heads_loss_backward(acts, M, buf.action.view(-1), buf.log_prob.view(-1), buf.adv.view(-1), buf.ret.view(-1),
                                        self.loss_cfg, row_index=idx, metrics=self.metrics)
                net.backward_trunk(obs_flat, acts, M, row_index=idx)
            else:
                net.forward(obs_flat, acts, M, row_index=idx)
                self._loss_step(acts)
                net.backward(obs_flat, acts, M, row_index=idx)
        finally:
            if defer:
                # single GPU: the fold also leaves the sums of squares for the clip (every gradient passes through it)
                fuse_norm = self.world == 1 and self.cfg.max_grad_norm > 0
                n, covered = ops.reduce_flush(self.optimizer.sumsq_partials if fuse_norm else None)
                if fuse_norm:
                    assert covered == self.net.fp.n_params(), "the merged fold must produce every gradient element for the fused clip"
                    self._n_sumsq = n

    def _opt_body(self):
        if self._n_sumsq > 0:
            self.optimizer.launch_clipped(self._n_sumsq, max_norm=self.cfg.max_grad_norm, grad_scale=1.0 / self.world)
        else:
            self.optimizer.launch(max_norm=self.cfg.max_grad_norm, grad_scale=1.0 / self.world)

    def _minibatch_body(self, window: Optional[int] = None):
        self._fwd_bwd_body(window)
        self._opt_body()

    def _epoch_body(self):
        for w in range(self.n_mb):
            self._minibatch_body(w)

    def _reduce_opt_body(self):
        """Multi-GPU: sum the flat gradient over the ranks, then clip + Adam with grad_scale = 1 / world."""
        cfg = self.cfg
        if getattr(cfg, "_profile_skip_reduce", False):     # developer probe (tools/dist_phase_times.py): ranks diverge
            self._opt_body()
        elif self.comm is not None:
            n = self.comm.allreduce_sumsq(self.net.fp.grad, self.grad_reduced, self.optimizer.sumsq_partials)
            if cfg.max_grad_norm > 0:
                self.optimizer.launch_clipped(n, max_norm=cfg.max_grad_norm, grad_scale=1.0 / self.world, grad=self.grad_reduced)
            else:
                self.optimizer.launch(grad_scale=1.0 / self.world, grad=self.grad_reduced)
        else:
            gdist.allreduce_sum_(self.net.fp.grad)  # ncclAllReduce (sum; Adam rescales by 1/world)
            self._opt_body()

    def _epoch_body_dist(self):
        """Multi-GPU epoch: per minibatch backward -> the gradient sum over ranks -> clip + Adam, all on one stream so the
        whole epoch (the peer-memory reduction kernels, or the NCCL kernels) can be one CUDA graph."""
        for w in range(self.n_mb):
            self._fwd_bwd_body(w)
            self._reduce_opt_body()

    def _capture(self, fn, error_mode: str = "global"):
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()  # warm-up outside capture (first-launch module loads)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        c0 = _ffi.launch_count()
        # with side branches (GYMRL_PPO_BRANCHES=1) the main chain is captured on a high-priority stream, so the dX kernels of the
        # critical path win the SMs over the dW kernels of the side branch
        cap_stream = torch.cuda.Stream(priority=-1) if getattr(self.net, "branches", None) is not None else None
        with torch.cuda.graph(g, stream=cap_stream, capture_error_mode=error_mode):
            fn()
        g.n_kernels = _ffi.launch_count() - c0   # our kernels recorded in this graph (one launch call each)
        return g

    def _replay(self, g):
        g.replay()
        self.graph_launches += g.n_kernels

    def _snapshot(self):
        return {"flat": self.net.fp.flat.clone(), "m": self.optimizer.exp_avg.clone(), "v": self.optimizer.exp_avg_sq.clone(),
                "step": self.optimizer.step_t.clone(), "ctr_mb": self.ctr_mb.clone(), "ctr_action": self.ctr_action.clone(),
                "metrics": self.metrics.clone(), "done_ctr": self.optimizer.done_ctr.clone(), "sumsq": self.optimizer.sumsq.clone()}

    def _restore(self, s):
        self.net.fp.flat.copy_(s["flat"]); self.optimizer.exp_avg.copy_(s["m"]); self.optimizer.exp_avg_sq.copy_(s["v"])
        self.optimizer.step_t.copy_(s["step"]); self.ctr_mb.copy_(s["ctr_mb"]); self.ctr_action.copy_(s["ctr_action"])
        self.metrics.copy_(s["metrics"]); self.optimizer.done_ctr.copy_(s["done_ctr"]); self.optimizer.sumsq.copy_(s["sumsq"])
        self.net.fp.refresh_weight_images()    # the parameters were written behind the library's back

    def _ensure_update_graphs(self):
        if not self.cfg.use_cuda_graph or self._g_mb is not None or self._g_mb_bwd is not None or self._g_epoch is not None:
            return
        snap = self._snapshot()  # capture warm-ups execute real steps: undo them
        if self.world == 1 and self.n_mb <= self.MAX_EPOCH_GRAPH_MINIBATCHES:
            # a whole epoch as one graph: every minibatch reads its window of the permutation through a static view
            self._g_epoch = self._capture(self._epoch_body)
        elif self.world == 1:
            self._g_mb = self._capture(self._minibatch_body)
        elif getattr(self.cfg, "dist_epoch_graph", True) and self.n_mb <= self.MAX_EPOCH_GRAPH_MINIBATCHES:
            # NCCL collectives are capturable: one graph per epoch instead of 2 replays + 1 eager all-reduce per minibatch.
            # thread_local: the NCCL watchdog thread may touch CUDA while this thread captures.
            self._g_epoch = self._capture(self._epoch_body_dist, error_mode="thread_local")
        else:
            self._g_mb_bwd = self._capture(self._fwd_bwd_body)
            self._g_opt = self._capture(self._opt_body) if self.comm is None else None
        self._restore(snap)

    def total_launches(self) -> int:
        return _ffi.launch_count() + self.graph_launches

    def _run_epochs(self):
        """num_epochs passes over the rollout in shuffled minibatches (ref :258-330): per epoch one device permutation, then
        either one graph replay for the whole epoch, one per minibatch, or (multi-GPU) backward / all-reduce / optimiser."""
        cfg = self.cfg
        self.optimizer.sync_lr()
        self._ensure_update_graphs()
        self.metrics.zero_()
        for _ in range(cfg.num_epochs):
            ops.random_permutation(self.N * self.T, seed=self.seed + 7919 * self.rank, draw_base=self.ctr_perm, out=self.perm)
            ops.counter_add(self.ctr_perm, 1)
            self.ctr_mb.zero_()
            if self._g_epoch is not None:
                self._replay(self._g_epoch)
                continue
            for _ in range(self.n_mb):
                if self.world == 1:
                    if self._g_mb is not None:
                        self._replay(self._g_mb)
                    else:
                        self._minibatch_body()
                else:
                    if self._g_mb_bwd is not None:
                        self._replay(self._g_mb_bwd)
                    else:
                        self._fwd_bwd_body()
                    if self.comm is not None:
                        self._reduce_opt_body()                 # the one collective of the path, fused with the clip's norm
                    else:
                        gdist.allreduce_sum_(self.net.fp.grad)  # ncclAllReduce (sum; Adam rescales by 1/world)
                        if self._g_opt is not None:
                            self._replay(self._g_opt)
                        else:
                            self._opt_body()

    def update(self, next_value=None, read_metrics: bool = True) -> dict:
        cfg, buf = self.cfg, self.buffer
        adv, _ = self.compute_gae(next_value)
        # advantage normalisation, numpy semantics (ddof = 0, ref :236); global over all shards
        self.adv_sums.zero_()
        ops.sum_sumsq(adv, self.adv_sums[:2])
        count = gdist.global_moments_(self.adv_sums, adv.numel())
        ops.normalize_inplace(adv, self.adv_sums, count, ddof=0, eps=1e-8)

        self._run_epochs()
        if not read_metrics:
            return {}
        m = self.metrics.tolist()  # the single D2H of the update (ref: 5 .item() per minibatch, :309-322)
        k = max(m[7], 1.0)
        return {"policy_loss": m[0] / k, "value_loss": m[1] / k, "entropy": m[2] / k, "clip_frac": m[3] / k,
                "approx_kl": m[4] / k}

    # ---------------------------------------------------------------- train / eval / test
    def _anneal(self):
        if self.cfg.anneal_lr:
            frac = 1.0 - self.step_count / self.cfg.max_train_steps
            lr = self.cfg.lr * frac
            for group in self.optimizer.param_groups:
                group["lr"] = lr

    def _refresh_episode_rewards(self):
        mean_ret, _, total = self.env.episode_stats(100)
        if total > self._episodes_seen:
            self._episodes_seen = total
            self.episode_rewards.clear()
            self.episode_rewards.extend([mean_ret] * int(min(total, 100)))
        return mean_ret, total

    def _global_episode_stats(self, avg_reward, total):
        """The stop criterion must be the same on every rank (a rank that broke out alone would leave the others blocked in the
        next gradient all-reduce, possibly inside a captured graph): average the ranks' last-100 windows, weighted by how many
        episodes each window holds, and sum the episode counts.  Identity on one GPU."""
        if self.world == 1:
            return avg_reward, total
        k = float(min(total, 100))
        s, k_all, tot = gdist.allreduce_scalars([avg_reward * k, k, float(total)], self.device)
        return (s / k_all if k_all > 0 else 0.0), int(tot)

    def train_iteration(self):
        """One pass of the public training loop body (ref train() :336-346): LR anneal (one 8-byte H2D when the
        value changes), rollout, update, metrics + episode statistics read back to the host."""
        self._anneal()
        self.collect_rollout()
        metrics = self.update(None)
        avg_reward, total = self._refresh_episode_rewards()
        return metrics, avg_reward, total

    def train(self):
        if self.rank == 0:
            print("Starting training...")
        update_count = 0
        t0 = time.time()
        while self.step_count < self.cfg.max_train_steps:
            self._anneal()
            next_value = self.collect_rollout()
            metrics = self.update(None)
            update_count += 1
            avg_reward, total = self._refresh_episode_rewards()
            avg_reward, total = self._global_episode_stats(avg_reward, total)
            if total > 0 and self.rank == 0:
                sps = self.step_count / max(time.time() - t0, 1e-9)
                print(f"Step: {self.step_count:,} | Updates: {update_count} | Avg Reward: {avg_reward:.1f} | "
                      f"Policy Loss: {metrics['policy_loss']:.4f} | Value Loss: {metrics['value_loss']:.4f} | "
                      f"Entropy: {metrics['entropy']:.4f} | KL: {metrics['approx_kl']:.4f} | "
                      f"Clip: {metrics['clip_frac']:.2%} | {sps:,.0f} steps/s")
            if total >= 100 and avg_reward >= 200.0:
                if self.rank == 0:
                    print(f"\nEnvironment solved at step {self.step_count:,}!")
                break
        if self.rank == 0:
            print("Training completed!")

    @torch.no_grad()
    def select_action(self, state, deterministic: bool = False):
        """Single-observation convenience path (ref get_action :92-104) through the same kernels."""
        x = torch.as_tensor(np.asarray(state, dtype=np.float32), device=self.device).reshape(1, -1)
        acts = getattr(self, "_acts_one", None)
        if acts is None:
            acts = self._acts_one = self._make_acts(1, backward=False)
        A = self.env.n_actions
        self.net.forward(x, acts, 1)
        a, lp, _ = ops.sample_categorical(acts.lv[:, :A], seed=self.seed, first_id=1 << 40, draw_base=self.ctr_action,
                                          deterministic=deterministic)
        ops.counter_add(self.ctr_action, 1)
        return int(a.item()), float(lp.item()), float(acts.lv[0, A].item())

    def eval(self, num_episodes: int = 10):
        """Deterministic (argmax) episodes, one env copy per episode, stepped in lockstep (ref eval :368-399)."""
        print(f"\nEvaluating for {num_episodes} episodes...")
        env = ops.VecEnv(self.cfg.env_name, num_episodes, seed=self.seed + 12345, first_env_id=1 << 32)
        acts = self._make_acts(num_episodes, backward=False, n_actions=env.n_actions)
        A = env.n_actions
        obs = env.reset()
        action = torch.zeros(num_episodes, device=self.device, dtype=i32)
        logp = torch.zeros(num_episodes, device=self.device, dtype=f32)
        ret = torch.zeros(num_episodes, device=self.device, dtype=f64)
        alive = torch.ones(num_episodes, device=self.device, dtype=torch.bool)
        for _ in range(env.max_episode_steps):
            self.net.forward(obs, acts, num_episodes)
            ops.sample_categorical(acts.lv[:, :A], deterministic=True, action=action, logp=logp)
            obs, r, term, trunc, _ = env.step(action, want_next_obs=False)
            ret += torch.where(alive, r.double(), torch.zeros_like(ret))  # logging glue only
            alive &= ~((term | trunc).bool())
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
    config.num_envs = 4096
    config.num_steps = 128
    config.num_minibatches = 32
    config.max_train_steps = 50_000_000
    trainer = PPOTrainer(config)

    def signal_handler(signum, frame):
        print("\n\nTraining interrupted. Starting test...")
        trainer.test()
        sys.exit(0)

    signal.signal(signal.SIGINT, signal_handler)
    try:
        trainer.train()
    except KeyboardInterrupt:
        print("\nTraining interrupted.")
    trainer.test()


if __name__ == "__main__":
    main()
