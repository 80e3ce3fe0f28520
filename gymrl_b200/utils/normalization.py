"""utils.normalization on the device (reference utils/normalization.py:4-52).

Same classes and call signatures.  A call takes one observation (NumPy [D], the reference's usage in
utils/runner.py:112,125-126 — returned as NumPy) or a device batch [N, D] from N lockstep env copies (returned as a
device tensor).  The statistic lives in one float64 device vector updated by gymrl_running_stats_update; with one
observation per call it reproduces the reference's arithmetic exactly, first-sample quirk (mean = std = x) included.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _ffi, ops

f32, f64, u8 = torch.float32, torch.float64, torch.uint8


def _dim(shape) -> int:
    if isinstance(shape, int):
        return int(shape)
    d = 1
    for s in shape:
        d *= int(s)
    return d


class RunningMeanStd:
    def __init__(self, shape, device=None):
        _ffi.require_cuda()
        self.shape, self.D = shape, _dim(shape)
        self.device = torch.device(device or "cuda")
        self.state = torch.zeros(1 + 3 * self.D, device=self.device, dtype=f64)   # n, mean[D], S[D], std[D]

    def _as_batch(self, x):
        t = torch.as_tensor(np.asarray(x, dtype=np.float32) if not torch.is_tensor(x) else x, dtype=f32, device=self.device)
        return t.reshape(-1, self.D).contiguous()

    def update(self, x):
        ops.running_stats_update(self._as_batch(x), self.state)

    # read-back views with the reference attribute names (one small D2H each; log / checkpoint time only)
    @property
    def n(self): return int(self.state[0].item())
    @property
    def mean(self): return self.state[1:1 + self.D].cpu().numpy().astype(np.float32 if self.n >= 1 else np.float64)
    @property
    def S(self): return self.state[1 + self.D:1 + 2 * self.D].cpu().numpy()
    @property
    def std(self): return self.state[1 + 2 * self.D:].cpu().numpy()

    def state_dict(self):
        return {"state": self.state.clone()}

    def load_state_dict(self, sd):
        self.state.copy_(sd["state"])

    # Pickled in the REFERENCE's shape (plain n / mean / S / std NumPy fields, ref :6-10): the reference's ModelLoader stores
    # `state_norm` / `reward_scaler` as pickled attributes (utils/model.py:343-345), so a checkpoint written here unpickles
    # into the reference's own classes and one written by the reference unpickles into these.
    def __getstate__(self):
        n = self.n
        return {"n": n, "mean": self.mean, "S": self.S, "std": self.std}

    def __setstate__(self, st):
        _ffi.require_cuda()
        mean = np.asarray(st["mean"], dtype=np.float64).reshape(-1)
        self.shape, self.D = (mean.shape[0] if mean.shape[0] != 1 else 1), mean.shape[0]
        self.device = torch.device("cuda", torch.cuda.current_device())
        flat = np.concatenate([[float(st["n"])], mean, np.asarray(st["S"], np.float64).reshape(-1), np.asarray(st["std"], np.float64).reshape(-1)])
        self.state = torch.as_tensor(flat, dtype=f64).to(self.device)


class Normalization:
    def __init__(self, shape, device=None):
        self.running_ms = RunningMeanStd(shape, device)

    def __call__(self, x, update=True):
        rm = self.running_ms
        xb = rm._as_batch(x)
        if update:
            ops.running_stats_update(xb, rm.state)
        y = ops.running_normalize(xb, rm.state, center=True)
        if torch.is_tensor(x):
            return y.view(x.shape)
        return y.cpu().numpy().reshape(np.shape(x))

    def state_dict(self): return self.running_ms.state_dict()
    def load_state_dict(self, sd): self.running_ms.load_state_dict(sd)


class RewardScaling:
    """x / std(discounted return), one accumulator R per env copy (reference: shape = 1, a single env)."""

    def __init__(self, shape, gamma, device=None, num_envs: int = 1):
        self.shape, self.gamma = shape, gamma
        self.running_ms = RunningMeanStd(shape=1, device=device)
        self.device = self.running_ms.device
        self.R = torch.zeros(num_envs, device=self.device, dtype=f64)
        self._reset = torch.zeros(num_envs, device=self.device, dtype=u8)

    def __call__(self, x):
        if torch.is_tensor(x):
            r = x.to(self.device, f32).reshape(-1).contiguous()
        else:
            r = torch.as_tensor(np.asarray(x, dtype=np.float32).reshape(-1), device=self.device)
        if r.numel() != self.R.numel():
            self.R = torch.zeros(r.numel(), device=self.device, dtype=f64)
            self._reset = torch.zeros(r.numel(), device=self.device, dtype=u8)
        out = ops.reward_scaling(r, self.R, self.running_ms.state, self.gamma, reset=self._reset)
        self._reset.zero_()
        if torch.is_tensor(x):
            return out.view(x.shape)
        return out.cpu().numpy().astype(np.float64).reshape(-1)   # the runner takes [0] (utils/runner.py:125)

    def reset(self, mask=None):
        """Start of an episode: R = 0 (reference :51-52); `mask` selects env copies in the batched use."""
        if mask is None:
            self._reset.fill_(1)
        else:
            self._reset.copy_(torch.as_tensor(mask, device=self.device).to(u8))

    def state_dict(self): return {"state": self.running_ms.state.clone(), "R": self.R.clone()}

    def __getstate__(self):     # the reference's fields (ref :39-43): shape, gamma, running_ms, R (NumPy)
        return {"shape": self.shape, "gamma": self.gamma, "running_ms": self.running_ms, "R": self.R.cpu().numpy().astype(np.float64)}

    def __setstate__(self, st):
        self.shape, self.gamma, self.running_ms = st["shape"], st["gamma"], st["running_ms"]
        self.device = self.running_ms.device
        self.R = torch.as_tensor(np.asarray(st["R"], np.float64).reshape(-1), dtype=f64).to(self.device)
        self._reset = torch.zeros(self.R.numel(), device=self.device, dtype=u8)

    def load_state_dict(self, sd):
        self.running_ms.state.copy_(sd["state"])
        self.R = sd["R"].clone().to(self.device)


# pickle.dump stores classes by (module, qualname): under gymrl_b200.utils.install() these ARE `utils.normalization.*`
for _c in (RunningMeanStd, Normalization, RewardScaling):
    _c.__module__ = "utils.normalization"
