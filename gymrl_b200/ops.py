"""Tensor-level wrappers over the C ABI (one function per extern "C" entry point).

Every function takes/returns CUDA torch tensors, launches on the current stream, never synchronises
(except VecEnv.episode_stats) and never falls back to PyTorch math: torch is only the allocator.
Output tensors can be passed in (``out=``) so steady-state loops and CUDA graphs allocate nothing.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _ffi
from ._ffi import check, load, ptr, stream_ptr

f32, i32, u8, f64 = torch.float32, torch.int32, torch.uint8, torch.float64


def _ld(t: torch.Tensor) -> int:
    """Leading dimension (row stride in elements) of a 2-D row-major view."""
    if t.dim() == 1:
        return 1
    assert t.stride(-1) == 1 or t.shape[-1] == 1, "innermost dimension must be contiguous"
    return t.stride(0)


# ------------------------------------------------------------------------------------------------
# environments
# ------------------------------------------------------------------------------------------------
class VecEnv:
    """N lockstep env copies on the current CUDA device (gym.make(name) x N)."""

    def __init__(self, name_or_kind, num_envs: int, seed: int = 0, first_env_id: int = 0):
        _ffi.require_cuda()
        lib = load()
        self.kind = _ffi.ENV_KINDS[name_or_kind] if isinstance(name_or_kind, str) else int(name_or_kind)
        od, ad, na, ms, sd = (C.c_int() for _ in range(5))
        ab = C.c_float()
        check(lib.gymrl_env_info(self.kind, C.byref(od), C.byref(ad), C.byref(na), C.byref(ms), C.byref(ab), C.byref(sd)))
        self.obs_dim, self.act_dim, self.n_actions = od.value, ad.value, na.value
        self.max_episode_steps, self.action_bound, self.state_doubles = ms.value, ab.value, sd.value
        self.discrete = self.n_actions > 0
        self.num_envs, self.seed, self.first_env_id = int(num_envs), int(seed), int(first_env_id)
        self.device = torch.device("cuda", torch.cuda.current_device())
        h = C.c_void_p()
        check(lib.gymrl_env_create(C.byref(h), self.kind, self.num_envs, self.seed & (2**64 - 1), self.first_env_id))
        self._h = h
        n = self.num_envs
        self.obs = torch.empty(n, self.obs_dim, device=self.device, dtype=f32)
        self.next_obs = torch.empty_like(self.obs)
        self.reward = torch.empty(n, device=self.device, dtype=f32)
        self.terminated = torch.empty(n, device=self.device, dtype=u8)
        self.truncated = torch.empty(n, device=self.device, dtype=u8)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            load().gymrl_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        obs = self.obs if out is None else out
        check(load().gymrl_env_reset(self._h, ptr(mask, u8), ptr(obs, f32), stream_ptr()))
        return obs

    def step(self, actions: torch.Tensor, obs: Optional[torch.Tensor] = None, next_obs: Optional[torch.Tensor] = None,
             reward: Optional[torch.Tensor] = None, terminated: Optional[torch.Tensor] = None,
             truncated: Optional[torch.Tensor] = None, want_next_obs: bool = True, done: Optional[torch.Tensor] = None):
        """Returns (obs_after_autoreset, reward, terminated, truncated, true_next_obs)."""
        obs = self.obs if obs is None else obs
        nobs = (self.next_obs if next_obs is None else next_obs) if want_next_obs else None
        reward = self.reward if reward is None else reward
        terminated = self.terminated if terminated is None else terminated
        truncated = self.truncated if truncated is None else truncated
        a = ptr(actions, i32 if self.discrete else f32)
        check(load().gymrl_env_step(self._h, a, ptr(obs, f32), ptr(nobs, f32) if nobs is not None else None,
                                    ptr(reward, f32), ptr(terminated, u8), ptr(truncated, u8), ptr(done, u8), stream_ptr()))
        return obs, reward, terminated, truncated, nobs

    def get_state(self) -> torch.Tensor:
        s = torch.empty(self.num_envs, self.state_doubles, device=self.device, dtype=f64)
        check(load().gymrl_env_get_state(self._h, ptr(s, f64), stream_ptr()))
        return s

    def set_state(self, s: torch.Tensor) -> None:
        s = s.to(self.device, f64).contiguous()
        assert s.shape == (self.num_envs, self.state_doubles)
        check(load().gymrl_env_set_state(self._h, ptr(s, f64), stream_ptr()))

    def set_profile(self, buf: Optional[torch.Tensor]) -> None:
        """Diagnostic: int64 [N, 8] device buffer the LunarLander step kernel fills with per-env cycle counters."""
        self._prof = buf
        check(load().gymrl_env_set_profile(self._h, ptr(buf, torch.int64) if buf is not None else None))

    def set_solver(self, variant: int) -> None:
        """LunarLander: arrangement of the solver loops in the step kernel (0, 2 or 3: bit-identical results; see gymrl.h)."""
        check(load().gymrl_env_set_solver(self._h, int(variant)))

    def get_solver(self) -> int:
        v = C.c_int()
        check(load().gymrl_env_get_solver(self._h, C.byref(v)))
        return int(v.value)

    def overflow_count(self) -> int:
        """LunarLander: touching manifolds dropped because all 8 contact slots of an env copy were taken (see gymrl.h)."""
        c = C.c_uint64()
        check(load().gymrl_env_overflow_count(self._h, C.byref(c), stream_ptr()))
        return int(c.value)

    def episode_stats(self, last_k: int = 100) -> Tuple[float, float, int]:
        mr, ml, te = C.c_double(), C.c_double(), C.c_uint64()
        check(load().gymrl_env_episode_stats(self._h, int(last_k), C.byref(mr), C.byref(ml), C.byref(te), stream_ptr()))
        return mr.value, ml.value, int(te.value)


# ------------------------------------------------------------------------------------------------
# action selection
# ------------------------------------------------------------------------------------------------
def sample_categorical(logits, noise=None, *, seed=0, first_id=0, draw=0, draw_base=None, deterministic=False,
                       action=None, logp=None, entropy=None, want_entropy=False, value_in=None, value_out=None):
    n, a = logits.shape
    dev = logits.device
    action = torch.empty(n, device=dev, dtype=i32) if action is None else action
    logp = torch.empty(n, device=dev, dtype=f32) if logp is None else logp
    if entropy is None and want_entropy:
        entropy = torch.empty(n, device=dev, dtype=f32)
    check(load().gymrl_sample_categorical(ptr(logits, f32), _ld(logits), ptr(noise, f32), ptr(action, i32), ptr(logp, f32),
                                          ptr(entropy, f32), ptr(value_in, f32), _ld(value_in) if value_in is not None else 0,
                                          ptr(value_out, f32), n, a, seed, first_id, draw, ptr(draw_base, i32), int(deterministic),
                                          stream_ptr()))
    return action, logp, entropy


def policy_heads_sample(h, Wa, ba, Wc, bc, *, n=None, seed=0, first_id=0, draw=0, draw_base=None, deterministic=False, action=None,
                        logp=None, entropy=None, value=None, lv_out=None):
    """Actor head + critic head + Categorical sample over h [n, 2H] = (actor | critic) in one launch (the PPO rollout tail)."""
    n = h.shape[0] if n is None else n
    A, H = Wa.shape
    action = torch.empty(n, device=h.device, dtype=i32) if action is None else action
    check(load().gymrl_policy_heads_sample(ptr(h, f32), _ld(h), ptr(Wa, f32), ptr(ba, f32), ptr(Wc, f32), ptr(bc, f32), int(H), int(A),
                                           ptr(action, i32), ptr(logp, f32), ptr(entropy, f32), ptr(value, f32), ptr(lv_out, f32), int(n),
                                           seed, first_id, draw, ptr(draw_base, i32), int(deterministic), stream_ptr()))
    return action


def select_eps_greedy(q, eps, *, seed=0, first_id=0, draw=0, draw_base=None, action=None):
    """eps: a Python float, or a device float32 tensor of one element (read by the kernel: capture-safe schedules)."""
    n, a = q.shape
    action = torch.empty(n, device=q.device, dtype=i32) if action is None else action
    if torch.is_tensor(eps):
        check(load().gymrl_select_eps_greedy_dev(ptr(q, f32), _ld(q), ptr(action, i32), n, a, ptr(eps, f32), seed, first_id, draw,
                                                 ptr(draw_base, i32), stream_ptr()))
        return action
    check(load().gymrl_select_eps_greedy(ptr(q, f32), _ld(q), ptr(action, i32), n, a, float(eps), seed, first_id, draw,
                                         ptr(draw_base, i32), stream_ptr()))
    return action


def sample_tanh_gaussian(mean, log_std, bound, log_std_min=-20.0, log_std_max=2.0, noise=None, *, seed=0, first_id=0, draw=0,
                         draw_base=None, deterministic=False, action=None, logp=None, pre_tanh=None, want_logp=True):
    n, a = mean.shape
    dev = mean.device
    action = torch.empty(n, a, device=dev, dtype=f32) if action is None else action
    if logp is None and want_logp and not deterministic:
        logp = torch.empty(n, device=dev, dtype=f32)
    check(load().gymrl_sample_tanh_gaussian(ptr(mean, f32), ptr(log_std, f32), _ld(mean), ptr(noise, f32), ptr(action, f32),
                                            ptr(logp, f32), ptr(pre_tanh, f32), n, a, float(bound), float(log_std_min),
                                            float(log_std_max), seed, first_id, draw, ptr(draw_base, i32), int(deterministic),
                                            stream_ptr()))
    return action, logp


def add_gaussian_noise_clip(mu, sigma, bound, noise_clip=0.0, noise=None, *, seed=0, first_id=0, draw=0, draw_base=None,
                            action=None):
    n, a = mu.shape
    assert mu.is_contiguous()
    action = torch.empty_like(mu) if action is None else action
    check(load().gymrl_add_gaussian_noise_clip(ptr(mu, f32), ptr(noise, f32), ptr(action, f32), n, a, float(sigma),
                                               float(noise_clip), float(bound), seed, first_id, draw, ptr(draw_base, i32),
                                               stream_ptr()))
    return action


# ------------------------------------------------------------------------------------------------
# GAE / normalisation
# ------------------------------------------------------------------------------------------------
def gae(reward, value, v_last_or_next, done, gamma, lam_actor, lam_critic=None, dw=None, dialect=0, adv=None, ret=None):
    """reward/value/done: [T, N] time-major.  dialect 0: v_last [N]; dialect 1: v_next [T, N] (+ dw)."""
    T, N = reward.shape
    lam_critic = lam_actor if lam_critic is None else lam_critic
    adv = torch.empty_like(reward) if adv is None else adv
    ret = torch.empty_like(reward) if ret is None else ret
    for t in (reward, value, v_last_or_next, done, adv, ret):
        assert t.is_contiguous()
    check(load().gymrl_gae(ptr(reward, f32), ptr(value, f32), ptr(v_last_or_next, f32), ptr(done, u8), ptr(dw, u8),
                           ptr(adv, f32), ptr(ret, f32), T, N, float(gamma), float(lam_actor), float(lam_critic), int(dialect),
                           stream_ptr()))
    return adv, ret


def sum_sumsq(x, sums=None):
    sums = torch.zeros(2, device=x.device, dtype=f64) if sums is None else sums
    assert x.is_contiguous()
    check(load().gymrl_sum_sumsq(ptr(x, f32), x.numel(), ptr(sums, f64), stream_ptr()))
    return sums


def normalize_inplace(x, sums, count, ddof=0, eps=1e-8):
    assert x.is_contiguous()
    check(load().gymrl_normalize_inplace(ptr(x, f32), x.numel(), ptr(sums, f64), float(count), int(ddof), float(eps), stream_ptr()))
    return x


# ------------------------------------------------------------------------------------------------
# PPO loss
# ------------------------------------------------------------------------------------------------
def ppo_loss(logits, value, action, logp_old, adv, ret, cfg: _ffi.PPOCfg, *, row_index=None, entropy_old=None, value_old=None,
             dlogits=None, dvalue=None, metrics=None):
    B, A = logits.shape
    dev = logits.device
    dlogits = torch.empty(B, A, device=dev, dtype=f32) if dlogits is None else dlogits
    dvalue = torch.empty(B, device=dev, dtype=f32) if dvalue is None else dvalue
    metrics = torch.zeros(8, device=dev, dtype=f32) if metrics is None else metrics
    check(load().gymrl_ppo_loss(ptr(logits, f32), _ld(logits), ptr(value, f32), _ld(value) if value.dim() > 1 else 1,
                                ptr(row_index, i32), ptr(action, i32), ptr(logp_old, f32), ptr(adv, f32), ptr(ret, f32),
                                ptr(entropy_old, f32), ptr(value_old, f32), ptr(dlogits, f32), _ld(dlogits), ptr(dvalue, f32),
                                _ld(dvalue) if dvalue.dim() > 1 else 1, ptr(metrics, f32), B, A, C.byref(cfg), stream_ptr()))
    return dlogits, dvalue, metrics


def weight_images_register(flat, mats):
    """Allocate and register pre-split tf32 images for the weight matrices `mats` = [(offset, rows, cols), ...] of the flat
    parameter buffer (csrc/wimages.cu): dense layers on those matrices then run the TMA-fed warp-specialised tensor-core GEMM.
    Returns the image tensor (keep it alive as long as the registration)."""
    import ctypes as C_
    n = flat.numel()
    images = torch.empty(4 * n, device=flat.device, dtype=f32)
    arr = (C_.c_int * (3 * len(mats)))(*[int(v) for m in mats for v in m])
    check(load().gymrl_weight_images_register(ptr(flat, f32), n, ptr(images, f32), arr, len(mats)))
    check(load().gymrl_weight_images_refresh(ptr(flat, f32), stream_ptr()))
    return images


def weight_images_refresh(flat):
    check(load().gymrl_weight_images_refresh(ptr(flat, f32), stream_ptr()))


def weight_images_unregister(flat):
    load().gymrl_weight_images_unregister(ptr(flat, f32))


def seq_gather(src, seq_index, seq_len, out=None):
    """out[b, t] = src.view(S, L, -1)[seq_index[b], t]  (ppo_lstm_lunarlander.py:682-707).  src: [S * L, D] rows."""
    src2 = src.reshape(src.shape[0], -1)
    nb, D = seq_index.numel(), src2.shape[1]
    out = torch.empty(nb * seq_len, D, device=src.device, dtype=f32) if out is None else out
    check(load().gymrl_seq_gather(ptr(src2, f32), _ld(src2), ptr(seq_index, i32), nb, int(seq_len), D, ptr(out, f32), _ld(out),
                                  stream_ptr()))
    return out


def gru_cell_forward(gi, gh, h, h_out=None, gates=None, save_gates=True):
    B, Hd = h.shape
    h_out = torch.empty(B, Hd, device=h.device, dtype=f32) if h_out is None else h_out
    if gates is None and save_gates:
        gates = torch.empty(B, 3 * Hd, device=h.device, dtype=f32)
    check(load().gymrl_gru_cell_forward(ptr(gi, f32), _ld(gi), ptr(gh, f32), _ld(gh), ptr(h, f32), _ld(h), ptr(h_out, f32), _ld(h_out),
                                        ptr(gates, f32), B, Hd, stream_ptr()))
    return h_out, gates


def gru_cell_backward(dh_out, gates, gh, h, dgi=None, dgh=None, dh=None, accumulate_dh=False):
    B, Hd = h.shape
    dev = h.device
    dgi = torch.empty(B, 3 * Hd, device=dev, dtype=f32) if dgi is None else dgi
    dgh = torch.empty(B, 3 * Hd, device=dev, dtype=f32) if dgh is None else dgh
    dh = torch.empty(B, Hd, device=dev, dtype=f32) if dh is None else dh
    assert gates.is_contiguous()
    check(load().gymrl_gru_cell_backward(ptr(dh_out, f32), _ld(dh_out), ptr(gates, f32), ptr(gh, f32), _ld(gh), ptr(h, f32), _ld(h),
                                         ptr(dgi, f32), _ld(dgi), ptr(dgh, f32), _ld(dgh), ptr(dh, f32), _ld(dh), int(accumulate_dh),
                                         B, Hd, stream_ptr()))
    return dgi, dgh, dh


def ppo_heads_workspace_bytes(H, A=4) -> int:
    return int(load().gymrl_ppo_heads_workspace_bytes(int(H), int(A)))


def ppo_heads_workspace(H, A=4, device="cuda"):
    return torch.empty(ppo_heads_workspace_bytes(H, A), device=device, dtype=torch.uint8)


def ppo_heads_fused(h, Wa, ba, Wc, bc, action, logp_old, adv, ret, cfg: _ffi.PPOCfg, *, dh, dWa, dba, dWc, dbc, workspace, M,
                    row_index=None, entropy_old=None, value_old=None, act_in=_ffi.ACT_TANH, lv_out=None, metrics=None,
                    accumulate=False):
    """Output heads + PPO loss + heads backward in one sweep over h [M, 2H] (csrc/ppo_loss.cu)."""
    A, H = Wa.shape
    check(load().gymrl_ppo_heads_fused(ptr(h, f32), _ld(h), ptr(Wa, f32), ptr(ba, f32), ptr(Wc, f32), ptr(bc, f32), ptr(row_index, i32),
                                       ptr(action, i32), ptr(logp_old, f32), ptr(adv, f32), ptr(ret, f32), ptr(entropy_old, f32),
                                       ptr(value_old, f32), ptr(dh, f32), _ld(dh), int(act_in), ptr(dWa, f32), ptr(dba, f32),
                                       ptr(dWc, f32), ptr(dbc, f32), ptr(lv_out, f32), ptr(metrics, f32), workspace.data_ptr(),
                                       workspace.numel() * workspace.element_size(), int(accumulate), int(M), int(H), int(A),
                                       C.byref(cfg), stream_ptr()))


# ------------------------------------------------------------------------------------------------
# dense layers
# ------------------------------------------------------------------------------------------------
def linear_forward(x, w, b=None, act=_ffi.ACT_NONE, *, row_index=None, out=None, M=None):
    """y = act(x[row_index] @ w.T + b);  w is [N, K] (torch.nn.Linear layout)."""
    N, K = w.shape
    M = (row_index.numel() if row_index is not None else x.shape[0]) if M is None else M
    assert w.is_contiguous() and x.shape[1] >= K
    out = torch.empty(M, N, device=x.device, dtype=f32) if out is None else out
    check(load().gymrl_linear_forward(ptr(x, f32), _ld(x), ptr(row_index, i32), ptr(w, f32), ptr(b, f32), ptr(out, f32), _ld(out),
                                      M, N, K, int(act), stream_ptr()))
    return out


def linear_backward_input(dy, w, h_in=None, act_in=_ffi.ACT_NONE, *, out=None, accumulate=False):
    """dx = (dy @ w) * act'(h_in)."""
    M = dy.shape[0]
    N, K = w.shape
    out = torch.empty(M, K, device=dy.device, dtype=f32) if out is None else out
    check(load().gymrl_linear_backward_input(ptr(dy, f32), _ld(dy), ptr(w, f32), ptr(h_in, f32), _ld(h_in) if h_in is not None else 0,
                                             ptr(out, f32), _ld(out), M, N, K, int(act_in), int(accumulate), stream_ptr()))
    return out


def backward_weight_workspace(M, N, K) -> int:
    return int(load().gymrl_linear_backward_weight_workspace(M, N, K))


def linear_backward_weight(dy, x, dw, db=None, *, row_index=None, workspace=None, accumulate=False, M=None):
    """dw = dy.T @ x[row_index];  db = dy.sum(0)."""
    N, K = dw.shape
    M = dy.shape[0] if M is None else M
    need = backward_weight_workspace(M, N, K)
    if workspace is None:
        workspace = torch.empty(need, device=dy.device, dtype=torch.uint8)
    assert workspace.numel() * workspace.element_size() >= need, "workspace too small"
    check(load().gymrl_linear_backward_weight(ptr(dy, f32), _ld(dy), ptr(x, f32), _ld(x), ptr(row_index, i32), ptr(dw, f32),
                                              ptr(db, f32), M, N, K, int(accumulate), workspace.data_ptr(),
                                              workspace.numel() * workspace.element_size(), stream_ptr()))
    return dw, db


def linear_backward(dy, x, w, dw, db=None, *, dx=None, act_in=_ffi.ACT_NONE, row_index=None, workspace=None, accumulate=False, M=None):
    """Whole layer backward: dw = dy.T @ x[row_index], db = dy.sum(0), dx = (dy @ w) * act'(x) (when dx is given)."""
    N, K = dw.shape
    M = dy.shape[0] if M is None else M
    need = backward_weight_workspace(M, N, K)
    if workspace is None:
        workspace = torch.empty(need, device=dy.device, dtype=torch.uint8)
    assert workspace.numel() * workspace.element_size() >= need, "workspace too small"
    check(load().gymrl_linear_backward(ptr(dy, f32), _ld(dy), ptr(x, f32), _ld(x), ptr(row_index, i32), ptr(w, f32), ptr(dw, f32),
                                       ptr(db, f32), ptr(dx, f32), _ld(dx) if dx is not None else 0, M, N, K, int(act_in),
                                       int(accumulate), workspace.data_ptr(), workspace.numel() * workspace.element_size(), stream_ptr()))
    return dw, db, dx


# ------------------------------------------------------------------------------------------------
# optimiser / target sync / permutation
# ------------------------------------------------------------------------------------------------
def grad_sumsq(grad, out=None):
    out = torch.zeros(1, device=grad.device, dtype=f64) if out is None else out
    check(load().gymrl_grad_sumsq(ptr(grad, f32), grad.numel(), ptr(out, f64), stream_ptr()))
    return out


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, *, beta1=0.9, beta2=0.999, eps=1e-8, sumsq=None, max_norm=0.0,
              clamp=0.0, grad_scale=1.0):
    """lr: float64 device scalar tensor; step: int32 device scalar tensor (incremented by the call)."""
    check(load().gymrl_adam_step(ptr(param, f32), ptr(grad, f32), ptr(exp_avg, f32), ptr(exp_avg_sq, f32), param.numel(),
                                 ptr(lr, f64), float(beta1), float(beta2), float(eps), ptr(step, i32), ptr(sumsq, f64),
                                 float(max_norm), float(clamp), float(grad_scale), stream_ptr()))


def clip_adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, *, sumsq_partials, n_partials, done_counter, max_norm, beta1=0.9,
                   beta2=0.999, eps=1e-8, grad_scale=1.0):
    """Clip-by-global-norm + Adam in one launch; the norm comes from reduce_flush's per-block sums of squares."""
    check(load().gymrl_clip_adam_step(ptr(param, f32), ptr(grad, f32), ptr(exp_avg, f32), ptr(exp_avg_sq, f32), param.numel(),
                                      ptr(lr, f64), float(beta1), float(beta2), float(eps), ptr(step, i32), ptr(sumsq_partials, f64),
                                      int(n_partials), float(max_norm), float(grad_scale), ptr(done_counter, i32), stream_ptr()))


def reduce_defer_begin():
    """Open a deferral scope: backward calls record their partial-gradient folds until reduce_flush()."""
    check(load().gymrl_reduce_defer_begin())


def reduce_flush(sumsq_partials=None):
    """Fold everything recorded since reduce_defer_begin() in one launch.  Returns (number of blocks = entries written to
    `sumsq_partials`, float64 [capacity], when given; number of gradient elements produced)."""
    n, cnt = C.c_int(0), C.c_longlong(0)
    cap = sumsq_partials.numel() if sumsq_partials is not None else 0
    check(load().gymrl_reduce_flush(ptr(sumsq_partials, f64), int(cap), C.byref(n), C.byref(cnt), stream_ptr()))
    return int(n.value), int(cnt.value)


def polyak(target, source, tau):
    check(load().gymrl_polyak(ptr(target, f32), ptr(source, f32), target.numel(), float(tau), stream_ptr()))


def random_permutation(n, *, seed=0, draw=0, draw_base=None, out=None, device=None):
    out = torch.empty(n, device=device or "cuda", dtype=i32) if out is None else out
    check(load().gymrl_random_permutation(ptr(out, i32), int(n), seed, draw, ptr(draw_base, i32), stream_ptr()))
    return out


def counter_add(counter, inc=1):
    check(load().gymrl_counter_add(ptr(counter, i32), int(inc), stream_ptr()))


def slice_i32(dst, src, block_index):
    """dst[:] = src[block*n:(block+1)*n] with `block` a device int32 scalar."""
    check(load().gymrl_slice_i32(ptr(dst, i32), ptr(src, i32), dst.numel(), ptr(block_index, i32), stream_ptr()))
    return dst


# ------------------------------------------------------------------------------------------------
# ppo_full network glue: mHC stages, RMSNorm (+SiLU)   (csrc/mhc.cu)
# ------------------------------------------------------------------------------------------------
def mhc_workspace_bytes(D, head_width=256, head_groups=2) -> int:
    return int(load().gymrl_mhc_workspace_bytes(int(D), int(head_width), int(head_groups)))


def mhc_stage_forward(h_prev, *, D, row_stride, branch_stride, M, z_prev=None, coef_prev=None, h_cur=None, params=None,
                      coef_cur=None, h_pre=None, final_weight=None, feat=None, sk_iters=10, eps=1e-6):
    """params = (g [2D], w [2D, 8], alpha [3], beta [8]) of the NEXT stage (None: no next stage)."""
    g = w = al = be = None
    if params is not None:
        g, w, al, be = params
    check(load().gymrl_mhc_stage_forward(ptr(h_prev, f32), int(row_stride), int(branch_stride), ptr(z_prev, f32), ptr(coef_prev, f32),
                                         ptr(h_cur, f32), ptr(g, f32), ptr(w, f32), ptr(al, f32), ptr(be, f32), ptr(coef_cur, f32),
                                         ptr(h_pre, f32), ptr(final_weight, f32), ptr(feat, f32), int(M), int(D), int(sk_iters),
                                         float(eps), stream_ptr()))


def mhc_stage_backward_a(h, z, dh_next, *, D, row_stride, branch_stride, M, dz, dh_partial, coef):
    """coef: the stage's [M, 24] coefficient rows written by mhc_stage_forward (in/out: gains dpost / dP)."""
    check(load().gymrl_mhc_stage_backward_a(ptr(h, f32), int(row_stride), int(branch_stride), ptr(z, f32), ptr(dh_next, f32),
                                            ptr(dz, f32), ptr(dh_partial, f32), ptr(coef, f32), int(M), int(D), stream_ptr()))


def mhc_stage_backward_b(h, dh_pre, scratch, dh_partial, params, grads, *, D, row_stride, branch_stride, M, workspace, dh=None,
                         dx0=None, accumulate=False):
    """grads = (dg, dw, dalpha, dbeta) views into the flat gradient buffer."""
    g, w, al, _ = params
    dg, dw, dal, dbe = grads
    check(load().gymrl_mhc_stage_backward_b(ptr(h, f32), int(row_stride), int(branch_stride), ptr(dh_pre, f32), ptr(scratch, f32),
                                            ptr(dh_partial, f32), ptr(dh, f32), ptr(dx0, f32), ptr(g, f32), ptr(w, f32), ptr(al, f32),
                                            ptr(dg, f32), ptr(dw, f32), ptr(dal, f32), ptr(dbe, f32), workspace.data_ptr(),
                                            workspace.numel() * workspace.element_size(), int(accumulate), int(M), int(D), stream_ptr()))


def rmsnorm_forward(x, weight, y, *, M, W, groups=1, sum2=False, silu=False, eps=1e-6):
    check(load().gymrl_rmsnorm_forward(ptr(x, f32), _ld(x), int(sum2), int(silu), ptr(weight, f32), ptr(y, f32), _ld(y), int(M), int(W),
                                       int(groups), float(eps), stream_ptr()))
    return y


def rmsnorm_backward(x, weight, dy, dx, dweight, *, M, W, workspace, groups=1, sum2=False, silu=False, eps=1e-6, accumulate=False):
    check(load().gymrl_rmsnorm_backward(ptr(x, f32), _ld(x), int(sum2), int(silu), ptr(weight, f32), ptr(dy, f32), _ld(dy), ptr(dx, f32),
                                        _ld(dx), ptr(dweight, f32), workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                                        int(accumulate), int(M), int(W), int(groups), float(eps), stream_ptr()))
    return dx, dweight


# ------------------------------------------------------------------------------------------------
# running normalisation (utils path, csrc/normalize.cu)
# ------------------------------------------------------------------------------------------------
def running_stats_update(x, state):
    """x [N, D] float32, state float64 [1 + 3 D]."""
    assert x.is_contiguous()
    N, D = x.shape
    check(load().gymrl_running_stats_update(ptr(x, f32), N, D, ptr(state, f64), stream_ptr()))


def running_normalize(x, state, center=True, out=None):
    assert x.is_contiguous()
    N, D = x.shape
    out = torch.empty_like(x) if out is None else out
    check(load().gymrl_running_normalize(ptr(x, f32), ptr(out, f32), N, D, ptr(state, f64), int(center), stream_ptr()))
    return out


def reward_scaling(r, R, state, gamma, reset=None, out=None):
    N = r.numel()
    out = torch.empty_like(r) if out is None else out
    check(load().gymrl_reward_scaling(ptr(r, f32), ptr(out, f32), ptr(R, f64), ptr(reset, u8), float(gamma), ptr(state, f64), N,
                                      stream_ptr()))
    return out
