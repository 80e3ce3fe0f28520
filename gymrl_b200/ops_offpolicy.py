"""Tensor-level wrappers for the off-policy part of the C ABI: replay ring, n-step window, PER sum-tree,
TD targets/losses, NoisyLinear helpers.  Same conventions as gymrl_b200/ops.py."""
from __future__ import annotations

from typing import Optional

import torch

from ._ffi import check, load, ptr, stream_ptr
from .ops import _ld, f32, f64, i32, u8


# ------------------------------------------------------------------------------------------------ replay ring
class ReplayRing:
    """SoA transition store in HBM with device-resident {cursor, size} (ReplayBuffer, dqn_cartpole.py:68-88)."""

    def __init__(self, capacity: int, obs_dim: int, act_width: int, discrete: bool, device, done_name: str = "done"):
        self.capacity, self.obs_dim, self.act_width, self.discrete = int(capacity), obs_dim, act_width, discrete
        self.state = torch.zeros(2, device=device, dtype=i32)            # {cursor, size}
        self.obs = torch.zeros(capacity, obs_dim, device=device, dtype=f32)
        self.next_obs = torch.zeros(capacity, obs_dim, device=device, dtype=f32)
        self.action = torch.zeros(capacity, act_width, device=device, dtype=i32 if discrete else f32)
        self.reward = torch.zeros(capacity, device=device, dtype=f32)
        self.done = torch.zeros(capacity, device=device, dtype=f32)      # stored as float like the reference's tensors
        self._size_host = 0
        self._done_ctr = torch.zeros(1, device=device, dtype=i32)        # gymrl_replay_store_all: blocks finished (zero between calls)

    def store(self, obs, action, reward, next_obs, done_u8):
        """One launch: the five fields at rows (cursor + i) % capacity and the {cursor, size} advance."""
        n = obs.shape[0]
        assert action.dtype == self.action.dtype and action.is_contiguous() and obs.is_contiguous() and next_obs.is_contiguous()
        check(load().gymrl_replay_store_all(ptr(self.obs), ptr(self.next_obs), ptr(self.action), ptr(self.reward), ptr(self.done),
                                            ptr(obs, f32), ptr(next_obs, f32), ptr(action), ptr(reward, f32), ptr(done_u8, u8), n,
                                            self.obs_dim, self.act_width, self.capacity, ptr(self.state, i32), ptr(self._done_ctr, i32),
                                            stream_ptr()))
        self._size_host = min(self.capacity, self._size_host + int(n))

    def advance(self, n):
        check(load().gymrl_replay_advance(ptr(self.state, i32), int(n), self.capacity, stream_ptr()))
        self._size_host = min(self.capacity, self._size_host + int(n))

    def sample_indices(self, batch, *, seed=0, draw=0, draw_base=None, out=None):
        out = torch.empty(batch, device=self.state.device, dtype=i32) if out is None else out
        check(load().gymrl_replay_sample_indices(ptr(out, i32), int(batch), ptr(self.state, i32), seed, draw, ptr(draw_base, i32),
                                                 stream_ptr()))
        return out

    def __len__(self):
        return self._size_host


def gather_concat(a, idx_a=None, b=None, idx_b=None, *, out=None, n=None):
    n = (idx_a.numel() if idx_a is not None else a.shape[0]) if n is None else n
    wa = a.shape[1]
    wb = b.shape[1] if b is not None else 0
    out = torch.empty(n, wa + wb, device=a.device, dtype=a.dtype) if out is None else out
    check(load().gymrl_gather_concat(ptr(out), _ld(out), ptr(a), wa, _ld(a), ptr(idx_a, i32), ptr(b) if b is not None else None, wb,
                                     _ld(b) if b is not None else 0, ptr(idx_b, i32), n, stream_ptr()))
    return out


# ------------------------------------------------------------------------------------------------ n-step window
class NStepWindow:
    def __init__(self, n_envs, obs_dim, n_steps, device):
        self.N, self.D, self.n = n_envs, obs_dim, n_steps
        z = lambda *s, dt=f32: torch.zeros(*s, device=device, dtype=dt)
        self.obs, self.nobs = z(n_steps, n_envs, obs_dim), z(n_steps, n_envs, obs_dim)
        self.act, self.rew = z(n_steps, n_envs, dt=i32), z(n_steps, n_envs)
        self.term, self.done = z(n_steps, n_envs, dt=u8), z(n_steps, n_envs, dt=u8)
        self.pushed = torch.zeros(2, device=device, dtype=i32)          # {pushes so far, blocks-done scratch}
        self.pushed_host = 0

    def push(self, obs, act, rew, nobs, term, done, gamma, ring: "ReplayRing", r_term, trunc=None):
        """Returns True when an n-step transition per env was written at ring rows [cursor, cursor+N).  With `trunc` the
        stored terminal flag is term & ~trunc (a time-limit cut is not terminal, rainbow :376)."""
        check(load().gymrl_nstep_push(ptr(self.obs), ptr(self.act), ptr(self.rew), ptr(self.nobs), ptr(self.term), ptr(self.done),
                                      ptr(obs, f32), ptr(act, i32), ptr(rew, f32), ptr(nobs, f32), ptr(term, u8), ptr(trunc, u8), ptr(done, u8),
                                      self.N, self.D, self.n, float(gamma), ptr(self.pushed, i32), ptr(ring.obs), ptr(ring.action),
                                      ptr(ring.reward), ptr(ring.next_obs), ptr(r_term, f32), ring.capacity, ptr(ring.state, i32),
                                      stream_ptr()))
        self.pushed_host += 1
        return self.pushed_host >= self.n


# ------------------------------------------------------------------------------------------------ sum tree
class DeviceSumTree:
    """float64 binary-heap sum tree with the reference's exact layout (rainbow_dqn_cartpole.py:116-152)."""

    def __init__(self, capacity, device):
        self.capacity = int(capacity)
        self.tree = torch.zeros(2 * self.capacity - 1, device=device, dtype=f64)
        n_scratch = int(load().gymrl_sumtree_scratch_ints(self.capacity))
        self.winner = torch.zeros(n_scratch, device=device, dtype=i32)     # winner[capacity] = -1 | per-subtree counters = 0
        self.winner[:self.capacity] = -1
        self.max_scratch = torch.zeros(1, device=device, dtype=f64)
        self.u32_scratch = torch.zeros(2, device=device, dtype=i32)      # {max IS-weight bits, blocks done}: zero between calls

    def update(self, idx, priority=None, td_error=None, eps=0.01, alpha=0.6, clip_max=0.0):
        n = idx.numel()
        check(load().gymrl_sumtree_update(ptr(self.tree, f64), self.capacity, ptr(idx, i32), ptr(priority, f64), ptr(td_error, f32), n,
                                          float(eps), float(alpha), float(clip_max), ptr(self.winner, i32), stream_ptr()))

    def store_new(self, n, ring_state):
        check(load().gymrl_sumtree_store_new(ptr(self.tree, f64), self.capacity, int(n), ptr(ring_state, i32), ptr(self.max_scratch, f64),
                                             ptr(self.winner, i32), stream_ptr()))

    def sample(self, batch, ring_state, beta_t, *, uniforms=None, out_idx=None, out_w=None, out_prio=None, tree_index=False, seed=0,
               draw=0, draw_base=None, raw_values=False):
        dev = self.tree.device
        out_idx = torch.empty(batch, device=dev, dtype=i32) if out_idx is None else out_idx
        out_w = torch.empty(batch, device=dev, dtype=f32) if out_w is None else out_w
        check(load().gymrl_sumtree_sample(ptr(self.tree, f64), self.capacity, int(batch), ptr(uniforms, f64), ptr(ring_state, i32),
                                          ptr(beta_t, f64), ptr(out_idx, i32), ptr(out_w, f32), ptr(out_prio, f64), ptr(self.u32_scratch, i32),
                                          int(tree_index) | (2 if raw_values else 0), seed, draw, ptr(draw_base, i32), stream_ptr()))
        return out_idx, out_w

    @property
    def priority_sum(self):
        return float(self.tree[0].item())

    @property
    def priority_max(self):
        return float(self.tree[self.capacity - 1:].max().item())


# ------------------------------------------------------------------------------------------------ TD losses
def dqn_loss(q, qnext_target, action, reward, done, gamma_n, *, v=None, vnext_target=None, qnext_online=None, vnext_online=None,
             row_index=None, is_weight=None, dq=None, dv=None, td_error=None, loss_acc=None):
    B, A = q.shape
    dev = q.device
    dq = torch.empty(B, A, device=dev, dtype=f32) if dq is None else dq
    if v is not None and dv is None:
        dv = torch.empty(B, 1, device=dev, dtype=f32)
    L = lambda t: _ld(t) if t is not None else 0
    check(load().gymrl_dqn_loss(ptr(q, f32), _ld(q), ptr(v, f32), L(v), ptr(qnext_target, f32), _ld(qnext_target), ptr(vnext_target, f32),
                                L(vnext_target), ptr(qnext_online, f32), L(qnext_online), ptr(vnext_online, f32), L(vnext_online),
                                ptr(row_index, i32), ptr(action, i32), ptr(reward, f32), ptr(done, f32), ptr(is_weight, f32), ptr(dq, f32),
                                _ld(dq), ptr(dv, f32), L(dv), ptr(td_error, f32), ptr(loss_acc, f32), B, A, float(gamma_n), stream_ptr()))
    return dq, dv


def twin_q_target(reward, done, q1t, q2t, gamma, *, row_index=None, logp_next=None, log_alpha=None, out=None):
    B = q1t.shape[0]
    out = torch.empty(B, device=q1t.device, dtype=f32) if out is None else out
    check(load().gymrl_twin_q_target(ptr(reward, f32), ptr(done, f32), ptr(row_index, i32), ptr(q1t, f32), _ld(q1t), ptr(q2t, f32),
                                     _ld(q2t), ptr(logp_next, f32), ptr(log_alpha, f64), float(gamma), ptr(out, f32), B, stream_ptr()))
    return out


def twin_q_loss(q1, q2, y, dq1, dq2, loss_acc=None):
    B = q1.shape[0]
    check(load().gymrl_twin_q_loss(ptr(q1, f32), _ld(q1), ptr(q2, f32), _ld(q2), ptr(y, f32), ptr(dq1, f32), _ld(dq1), ptr(dq2, f32),
                                   _ld(dq2), ptr(loss_acc, f32), B, stream_ptr()))


def min_q_grad(q1, q2, dq1, dq2, q1_only=False, acc=None):
    B = dq1.shape[0]
    L = lambda t: _ld(t) if t is not None else 0
    check(load().gymrl_min_q_grad(ptr(q1, f32), L(q1), ptr(q2, f32), L(q2), ptr(dq1, f32), _ld(dq1), ptr(dq2, f32), L(dq2), B,
                                  int(q1_only), ptr(acc, f32), stream_ptr()))


def sac_actor_grad(pre_tanh, noise, log_std, dq_daction, log_alpha, bound, ls_min, ls_max, dmean, dlog_std, logp, acc):
    B, A = pre_tanh.shape
    check(load().gymrl_sac_actor_grad(ptr(pre_tanh, f32), ptr(noise, f32), ptr(log_std, f32), _ld(log_std), ptr(dq_daction, f32),
                                      _ld(dq_daction), ptr(log_alpha, f64), float(bound), float(ls_min), float(ls_max), ptr(dmean, f32),
                                      ptr(dlog_std, f32), _ld(dmean), ptr(logp, f32), ptr(acc, f32), B, A, stream_ptr()))


def sac_alpha_step(log_alpha, adam_state, acc, batch, target_entropy, lr, loss_out=None):
    check(load().gymrl_sac_alpha_step(ptr(log_alpha, f64), ptr(adam_state, f64), ptr(acc, f32), int(batch), float(target_entropy), float(lr),
                                      ptr(loss_out, f32), stream_ptr()))


# ------------------------------------------------------------------------------------------------ discrete SAC
def sac_discrete_target(logits_next, q1t, q2t, reward, done, log_alpha, gamma, *, row_index=None, out=None):
    B, A = logits_next.shape
    out = torch.empty(B, device=q1t.device, dtype=f32) if out is None else out
    check(load().gymrl_sac_discrete_target(ptr(logits_next, f32), _ld(logits_next), ptr(q1t, f32), _ld(q1t), ptr(q2t, f32), _ld(q2t),
                                           ptr(reward, f32), ptr(done, f32), ptr(row_index, i32), ptr(log_alpha, f32), float(gamma),
                                           ptr(out, f32), B, A, stream_ptr()))
    return out


def sac_discrete_critic_loss(q1, q2, action, y, dq1, dq2, *, row_index=None, loss_acc=None):
    B, A = q1.shape
    check(load().gymrl_sac_discrete_critic_loss(ptr(q1, f32), _ld(q1), ptr(q2, f32), _ld(q2), ptr(action, i32), ptr(row_index, i32),
                                                ptr(y, f32), ptr(dq1, f32), _ld(dq1), ptr(dq2, f32), _ld(dq2), ptr(loss_acc, f32), B, A,
                                                stream_ptr()))


def sac_discrete_actor_grad(logits, q1, q2, log_alpha, dlogits, acc=None):
    B, A = logits.shape
    check(load().gymrl_sac_discrete_actor_grad(ptr(logits, f32), _ld(logits), ptr(q1, f32), _ld(q1), ptr(q2, f32), _ld(q2),
                                               ptr(log_alpha, f32), ptr(dlogits, f32), _ld(dlogits), ptr(acc, f32), B, A, stream_ptr()))


def sac_discrete_alpha_step(log_alpha, adam_state, acc, batch, target_entropy, lr, loss_out=None):
    check(load().gymrl_sac_discrete_alpha_step(ptr(log_alpha, f32), ptr(adam_state, f32), ptr(acc, f32), int(batch), float(target_entropy),
                                               float(lr), ptr(loss_out, f32), stream_ptr()))


def tanh_bound(z, bound, out=None):
    B, A = z.shape
    out = torch.empty(B, A, device=z.device, dtype=f32) if out is None else out
    check(load().gymrl_tanh_bound(ptr(z, f32), _ld(z), ptr(out, f32), float(bound), B, A, stream_ptr()))
    return out


def tanh_bound_grad(action, dq_daction, dz, bound):
    B, A = action.shape
    check(load().gymrl_tanh_bound_grad(ptr(action, f32), ptr(dq_daction, f32), _ld(dq_daction), ptr(dz, f32), _ld(dz), float(bound), B, A,
                                       stream_ptr()))


def fill_normal(out, *, seed=0, entity0=0, draw=0, draw_base=None):
    check(load().gymrl_fill_normal(ptr(out, f32), out.numel(), seed, entity0, draw, ptr(draw_base, i32), stream_ptr()))
    return out


def noisy_sample(eps, xi=None, *, seed=0, entity=0, draw=0, draw_base=None):
    check(load().gymrl_noisy_sample(ptr(eps, f32), ptr(xi, f32), eps.numel(), seed, entity, draw, ptr(draw_base, i32), stream_ptr()))
    return eps


def noisy_compose(w_mu, w_sigma, eps_in, eps_out, b_mu, b_sigma, w, b):
    N, K = w_mu.shape
    check(load().gymrl_noisy_compose(ptr(w_mu, f32), ptr(w_sigma, f32), ptr(eps_in, f32), ptr(eps_out, f32), ptr(b_mu, f32),
                                     ptr(b_sigma, f32), ptr(w, f32), ptr(b, f32), N, K, stream_ptr()))


def noisy_backward(dw, db, eps_in, eps_out, dw_mu, dw_sigma, db_mu, db_sigma, accumulate=False):
    N, K = dw.shape
    check(load().gymrl_noisy_backward(ptr(dw, f32), ptr(db, f32), ptr(eps_in, f32), ptr(eps_out, f32), ptr(dw_mu, f32), ptr(dw_sigma, f32),
                                      ptr(db_mu, f32), ptr(db_sigma, f32), N, K, int(accumulate), stream_ptr()))


def noisy_refresh(layers, K, *, noisy=True, seed=0, entity_base=0, entity_stride=4096, draw=0, draw_base=None, counter_inc=0):
    """layers: one or two dicts {w_mu, w_sigma, b_mu, b_sigma, eps_in, eps_out, w, b} (NoisyLinear layers sharing the input
    width K).  Draws the factorised noise (or zeroes it), composes W / b and advances the draw counter in ONE launch."""
    def args(l):
        if l is None:
            return [None] * 8 + [0]
        return [ptr(l["w_mu"], f32), ptr(l["w_sigma"], f32), ptr(l["b_mu"], f32), ptr(l["b_sigma"], f32), ptr(l["eps_in"], f32),
                ptr(l["eps_out"], f32), ptr(l["w"], f32), ptr(l["b"], f32), int(l["w_mu"].shape[0])]
    l0, l1 = layers[0], (layers[1] if len(layers) > 1 else None)
    check(load().gymrl_noisy_refresh(*args(l0), *args(l1), int(K), int(bool(noisy)), seed, int(entity_base), int(entity_stride), int(draw),
                                     ptr(draw_base, i32), int(counter_inc), stream_ptr()))


def noisy_backward2(l0, l1, K, accumulate=False):
    """l: dict {dw, db, eps_in, eps_out, dw_mu, dw_sigma, db_mu, db_sigma}; both layers in one launch."""
    def args(l):
        return [ptr(l["dw"], f32), ptr(l["db"], f32), ptr(l["eps_in"], f32), ptr(l["eps_out"], f32), ptr(l["dw_mu"], f32),
                ptr(l["dw_sigma"], f32), ptr(l["db_mu"], f32), ptr(l["db_sigma"], f32), int(l["dw"].shape[0])]
    check(load().gymrl_noisy_backward2(*args(l0), *args(l1), int(K), int(accumulate), stream_ptr()))
