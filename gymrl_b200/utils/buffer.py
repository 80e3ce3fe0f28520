"""utils.buffer on the device (reference utils/buffer.py).

Class names are the reference's — utils.runner.train picks its on/off-policy branch by isinstance on them (q17).
Storage is [T][N] SoA on the device (N = 1 when driven by the reference's single-env loop); `store` accepts the
reference's tuples of Python/NumPy scalars as well as tuples of device tensors [N, ...] from lockstep env copies.
  ReplayBuffer_on_policy.sample():  GAE with the utils dialect (bootstrap masked by dw, trace by done, per-step next
      values) = gymrl_gae dialect 1, then (adv - mean) / (std + 1e-8) with torch's ddof = 1 (:33) = gymrl_normalize_inplace.
  ReplayBuffer_off_policy.sample(): uniform draw WITHOUT replacement (np.random.choice(replace=False), :124) = a device
      permutation prefix, rows gathered on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _ffi, ops

f32, i32, i64, u8, f64 = torch.float32, torch.int32, torch.int64, torch.uint8, torch.float64


def _dev(cfg):
    d = getattr(cfg, "device", "cuda")
    d = torch.device(d)
    if d.type != "cuda":
        raise RuntimeError("gymrl_b200.utils.buffer keeps its storage on a CUDA device (cfg.device must be cuda)")
    return d


def _row(x, device, dtype=f32):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), device=device).to(dtype)


class ReplayBuffer_on_policy:
    FIELDS = ("s", "a", "r", "d", "dw", "logp", "v", "v_next")

    def __init__(self, cfg):
        _ffi.require_cuda()
        self.cfg = cfg
        self.device = _dev(cfg)
        self.clear()

    def clear(self):
        self.buffer = []          # list of per-step tuples of device tensors ([N, ...] each)
        self.samples = None
        self._n = 0

    def store(self, transitions):
        assert self.samples is None, 'Need to clear the buffer before storing new transitions.'
        s, a, r, d, dw, logp, v, v_next = transitions
        dev = self.device
        s = _row(s, dev)
        n = 1 if s.dim() == 1 else s.shape[0]
        row = (s.reshape(n, -1), _row(a, dev, i64).reshape(n), _row(r, dev).reshape(n), _row(d, dev, u8).reshape(n),
               _row(dw, dev, u8).reshape(n), _row(logp, dev).reshape(n), _row(v, dev).reshape(n), _row(v_next, dev).reshape(n))
        self.buffer.append(row)
        self._n += n

    def size(self):
        return self._n

    def compute_advantage(self, rewards, dones, dw, values, next_values):
        """[T, N] (or the reference's [T, 1]) float tensors -> (normalised adv, v_target), both shaped like `values`."""
        shape = values.shape
        T = rewards.shape[0]
        r = rewards.reshape(T, -1).to(self.device, f32).contiguous()
        v = values.reshape(T, -1).to(self.device, f32).contiguous()
        vn = next_values.reshape(T, -1).to(self.device, f32).contiguous()
        dn = dones.reshape(T, -1).to(self.device).to(u8).contiguous()
        dwb = dw.reshape(T, -1).to(self.device).to(u8).contiguous()
        adv, v_target = ops.gae(r, v, vn, dn, self.cfg.gamma, self.cfg.lamda, dw=dwb, dialect=1)
        sums = ops.sum_sumsq(adv)
        ops.normalize_inplace(adv, sums, adv.numel(), ddof=1, eps=1e-8)
        return adv.reshape(shape), v_target.reshape(shape)

    def sample(self):
        if self.samples is None:
            cols = list(zip(*self.buffer))
            s = torch.stack(cols[0])                                   # [T, N, D]
            a, r, d, dw, logp, v, vn = (torch.stack(c) for c in cols[1:])   # [T, N]
            adv, v_target = self.compute_advantage(r, d, dw, v, vn)
            T, N = r.shape
            flat = lambda x: x.reshape(T * N, 1)
            self.samples = (s.reshape(T * N, -1), flat(a), flat(logp), flat(adv), flat(v_target))
        return self.samples


class ReplayBuffer_on_policy_v2:
    """Padded per-episode store of the recurrent scripts (reference :53-102).  Host-side container only — the recurrent
    trainers are outside the B200 path — kept so that utils.runner's isinstance dispatch has the name."""
    KEYS = ("s", "a", "a_logprob", "r", "d", "dw", "v", "v_", "active")

    def __init__(self, cfg):
        self.cfg = cfg
        self.clear()

    def clear(self):
        B, L = self.cfg.batch_size, self.cfg.max_steps
        z = lambda *shape, dtype=np.float32: np.zeros(shape, dtype=dtype)
        self.buffer = {"s": z(B, L, *self.cfg.state_shape), "a": z(B, L, dtype=np.int64), "a_logprob": z(B, L), "r": z(B, L),
                       "d": z(B, L), "dw": np.ones((B, L), np.float32), "v": z(B, L), "v_": z(B, L), "active": z(B, L, dtype=np.int8)}
        self.size = np.zeros(B, dtype=int)
        self.episode_num = 0

    def store(self, transitions):
        s, a, r, d, dw, a_logprob, v, v_ = transitions
        e, t = self.episode_num, self.size[self.episode_num]
        for k, val in zip(("s", "a", "r", "d", "dw", "a_logprob", "v", "v_"), (s, a, r, d, dw, a_logprob, v, v_)):
            self.buffer[k][e, t] = val
        self.buffer["active"][e, t] = 1
        self.size[e] += 1

    def next_episode(self):
        self.episode_num += 1

    def sample(self):
        L = self.size.max()
        dt = {"a": torch.long}
        return tuple(torch.tensor(self.buffer[k][:, :L], dtype=dt.get(k, torch.float32), device=self.cfg.device)
                     for k in ("s", "a", "a_logprob", "r", "d", "dw", "v", "v_", "active"))


class ReplayBuffer_off_policy:
    """Uniform replay (reference :105-135: object ring + np.random.choice(size, B, replace=False)) as a device SoA ring.

    store(transitions) takes the reference's tuple (s, a, r, s', done ...) for ONE env copy, or the same fields with a leading
    [N] axis for N lockstep copies (N = cfg.num_envs when set, else inferred from the first store's state against
    cfg.state_shape).  Every field keeps its per-item shape: sample() returns [B] for scalars, [B, 1] for 1-element vectors
    (e.g. Pendulum actions, like torch.tensor(np.array(...)) in the reference), [B, C, H, W] for image states.
    Writes go through gymrl_replay_store with the ring cursor in device memory; sampling is the O(B) keyed-bijection prefix
    of gymrl_replay_sample_indices (without replacement) into a preallocated index buffer."""

    def __init__(self, cfg):
        _ffi.require_cuda()
        self.cfg = cfg
        self.capacity, self.batch_size = int(cfg.memory_capacity), int(cfg.batch_size)
        self.device = _dev(cfg)
        self.seed = int(getattr(cfg, "seed", 0) or 0)
        self.clear()

    def clear(self):
        self.fields = None
        self.item_shapes = None
        self.pointer, self.is_full = 0, False
        self._draw = 0
        self.state = torch.zeros(2, device=self.device, dtype=torch.int32)        # {cursor, size} on the device
        self._idx = torch.zeros(self.batch_size, device=self.device, dtype=torch.int32)

    def _infer_n(self, first):
        """(n, batched): how many transitions this store() carries and whether the fields have a leading [n] axis."""
        n = int(getattr(self.cfg, "num_envs", 0) or 0)
        shape = tuple(first.shape) if torch.is_tensor(first) else tuple(np.shape(first))
        if n > 1:
            assert shape[:1] == (n,), f"expected a leading [{n}] axis (cfg.num_envs), got {shape}"
            return n, True
        st = getattr(self.cfg, "state_shape", None)
        if st is not None and shape != tuple(st) and shape[1:] == tuple(st):
            return shape[0], True
        return 1, False

    def store(self, transitions):
        n, batched = self._infer_n(transitions[0])
        rows = [x.to(self.device, f32) if torch.is_tensor(x) else torch.as_tensor(np.asarray(x, dtype=np.float32), device=self.device)
                for x in transitions]
        if self.fields is None:
            self.item_shapes = [tuple(t.shape[1:]) if batched else tuple(t.shape) for t in rows]
            self.fields = [torch.zeros(self.capacity, max(1, int(np.prod(sh, dtype=np.int64))), device=self.device, dtype=f32)
                           for sh in self.item_shapes]
        L, sp = _ffi.load(), _ffi.stream_ptr()
        for f, t in zip(self.fields, rows):
            src = t.reshape(n, -1).contiguous()
            assert src.shape[1] == f.shape[1], "field width changed between stores"
            _ffi.check(L.gymrl_replay_store(_ffi.ptr(f, f32), _ffi.ptr(src, f32), n, f.shape[1], 0, self.capacity,
                                            _ffi.ptr(self.state, torch.int32), sp))
        _ffi.check(L.gymrl_replay_advance(_ffi.ptr(self.state, torch.int32), n, self.capacity, sp))
        if self.pointer + n >= self.capacity:
            self.is_full = True
        self.pointer = (self.pointer + n) % self.capacity

    def size(self):
        return self.capacity if self.is_full else self.pointer

    def sample(self):
        size = self.size()
        b = min(self.batch_size, size)
        _ffi.check(_ffi.load().gymrl_replay_sample_indices(_ffi.ptr(self._idx, torch.int32), b, _ffi.ptr(self.state, torch.int32),
                                                           self.seed, self._draw, None, _ffi.stream_ptr()))
        self._draw += 1
        idx = self._idx[:b].long()
        return iter([torch.index_select(f, 0, idx).reshape((b,) + sh) for f, sh in zip(self.fields, self.item_shapes)])


class Queue:
    """Fixed-size ring of Python objects with uniform sampling (reference :139-169; host-side helper of StateManager)."""

    def __init__(self, buffer_size):
        self.buffer_size = buffer_size
        self.buffer = np.empty(buffer_size, dtype=object)
        self.index, self.filled = 0, False

    def put(self, item):
        self.buffer[self.index] = item
        self.index = (self.index + 1) % self.buffer_size
        self.filled = self.filled or self.index == 0

    def size(self):
        return self.buffer_size if self.filled else self.index

    def sample(self):
        if self.is_empty():
            raise ValueError('Queue is empty!')
        return self.buffer[np.random.randint(0, self.size())]

    def is_empty(self):
        return self.size() == 0

    def is_full(self):
        return self.filled

    def capacity(self):
        return self.buffer_size
