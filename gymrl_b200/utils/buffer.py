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
    def __init__(self, cfg):
        _ffi.require_cuda()
        self.capacity, self.batch_size = int(cfg.memory_capacity), int(cfg.batch_size)
        self.device = _dev(cfg)
        self.seed = int(getattr(cfg, "seed", 0) or 0)
        self.clear()

    def clear(self):
        self.fields = None
        self.pointer, self.is_full = 0, False
        self._draw = 0

    def _alloc(self, row):
        self.fields = [torch.zeros((self.capacity,) + tuple(x.shape[1:]), device=self.device, dtype=f32) for x in row]

    def store(self, transitions):
        row = [_row(x, self.device) for x in transitions]
        n = row[0].shape[0] if row[0].dim() >= 2 else 1     # state [D] = one env copy, [N, D] = N lockstep copies
        row = [x.reshape(n, -1) for x in row]
        if self.fields is None:
            self._alloc(row)
        idx = (self.pointer + torch.arange(n, device=self.device)) % self.capacity
        for f, x in zip(self.fields, row):
            f[idx] = x
        if self.pointer + n >= self.capacity:
            self.is_full = True
        self.pointer = (self.pointer + n) % self.capacity

    def size(self):
        return self.capacity if self.is_full else self.pointer

    def sample(self):
        size = self.size()
        b = min(self.batch_size, size)
        perm = ops.random_permutation(size, seed=self.seed, draw=self._draw, device=self.device)   # without replacement
        self._draw += 1
        idx = perm[:b].long()
        out = []
        for f in self.fields:
            x = f[idx]
            out.append(x.squeeze(-1) if x.shape[-1] == 1 else x)
        return iter(out)


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
