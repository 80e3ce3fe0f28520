"""PPO "with all tricks" for LunarLander-v3 on the B200 engine — same surface as the reference script
``algorithms/ppo_full_lunarlander.py`` (Config, RMSNorm / ManifoldHyperConnectionFuse / MHCBlock / MHCBackbone / MLP /
ActorCritic with the reference's state_dict keys, RolloutBuffer, PPOTrainer with collect_experience() /
compute_advantages() / update_model() / train() / eval() / test()).

    ActorCritic.forward (ref :391-393)   ->  input_proj GEMM -> 4 x [mHC stage kernel -> 128x128 GEMM] -> final
                                             branch-sum + RMSNorm -> merged head GEMM [512,128] -> SiLU+RMSNorm kernel ->
                                             the two skinny output layers            (csrc/mhc.cu + the dense-layer kernels)
    collect_experience (ref :462-505)    ->  T lockstep iterations over N env copies in one CUDA graph; the sampler also
                                             stores the policy entropy (old_entropies, ref :488)
    compute_advantages (ref :507-535)    ->  gymrl_gae dialect 2 (decoupled lam_actor / lam_critic, float32 gamma*V product)
    update_model (ref :537-679)          ->  per minibatch: forward, GYMRL_PPO_FULL loss (clip-higher, ratio clamp [0, dual_clip],
                                             ERC mask, plain .mean()), backward, global-norm clip + Adam; LR and entropy
                                             coefficient annealed AFTER the update (ref :659-666) through device scalars
Not carried over: the PSCN backbone (`use_mhc = False`, ref :321-360) and the covariance clip (`clip_cov_ratio > 0`,
ref :608-616; 0 by default) raise NotImplementedError.
"""
from __future__ import annotations

import signal
import sys
import time
from collections import deque
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _ffi, dist as gdist, ops
from ..nn import FlatParams, FusedAdam, layer_init
from . import ppo_lunarlander as base

f32, i32, u8, f64 = torch.float32, torch.int32, torch.uint8, torch.float64


class Config:
    def __init__(self):
        self.env_name = "LunarLander-v3"
        self.seed = None
        # mHC parameters (ref :25-30)
        self.use_mhc = True
        self.mhc_dim = 128
        self.mhc_rate = 2
        self.mhc_layers = 2
        self.mhc_sk_it = 10
        # training parameters (ref :32-52)
        self.max_train_steps = 5e6
        self.update_freq = 4096
        self.num_epochs = 4
        self.batch_size = 1024
        self.gamma = 0.995
        self.lam_actor = 0.95
        self.lam_critic = 0.95
        self.clip_eps_min = 0.2
        self.clip_eps_max = 0.28
        self.clip_cov_ratio = 0.0
        self.clip_cov_min = 1.0
        self.clip_cov_max = 5.0
        self.dual_clip = 3.0
        self.entropy_coef = 0.01
        self.erc_beta_low = 0.06
        self.erc_beta_high = 0.06
        self.lr = 3e-4
        self.max_grad_norm = 0.5
        self.anneal = True
        self.device = "cuda"  # the engine is the CUDA library (the reference hard-codes "cpu", :52)
        # ---- engine extras (defaults reproduce the reference at num_envs = 1) ----
        self.num_envs = 1
        self.num_steps = None
        self.num_minibatches = None
        self.reset_each_rollout = None
        self.use_cuda_graph = True


# ------------------------------------------------------------------------------------------------ modules (host side:
# parameter containers with the reference's names; the arithmetic runs in ActorCriticEngine)
class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))


class ManifoldHyperConnectionFuse(nn.Module):
    """Parameters of one mHC fusion layer (ref :112-142): norm.weight [n*dim], w [n*dim, n^2+2n], alpha [3], beta [n^2+2n]."""

    def __init__(self, dim: int, rate: int, max_sk_it: int):
        super().__init__()
        self.n, self.dim, self.max_sk_it = rate, dim, max_sk_it
        self.nc, self.n2 = rate * dim, rate * rate
        self.norm = RMSNorm(dim * rate)
        self.w = nn.Parameter(torch.zeros(self.nc, self.n2 + 2 * rate))
        self.alpha = nn.Parameter(torch.ones(3) * 0.01)
        beta_init = torch.zeros(self.n2 + 2 * rate)
        beta_init[:2 * rate] = 0.01
        res_beta = torch.full((rate, rate), -2.0)
        res_beta.fill_diagonal_(2.0)
        beta_init[2 * rate:] = res_beta.flatten()
        self.beta = nn.Parameter(beta_init)


class MHCBlock(nn.Module):
    def __init__(self, dim: int, rate: int, max_sk_it: int):
        super().__init__()
        self.linear1 = nn.Linear(dim, dim)
        self.mhc1 = ManifoldHyperConnectionFuse(dim, rate, max_sk_it)
        self.linear2 = nn.Linear(dim, dim)
        self.mhc2 = ManifoldHyperConnectionFuse(dim, rate, max_sk_it)


class MHCBackbone(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, rate: int, num_layers: int, max_sk_it: int):
        super().__init__()
        self.rate, self.output_dim = rate, output_dim
        self.input_proj = nn.Linear(input_dim, output_dim)
        self.layers = nn.ModuleList([MHCBlock(output_dim, rate, max_sk_it) for _ in range(num_layers)])
        self.final_norm = RMSNorm(output_dim)


class MLP(nn.Module):
    """[in, hidden, out]: Linear -> SiLU -> RMSNorm -> Linear, module indices as in the reference Sequential (ref :298-315)."""

    def __init__(self, dim_list: List[int], last_std: Optional[float] = None):
        super().__init__()
        assert len(dim_list) == 3, "the engine implements the reference's two-layer heads"
        self.mlp = nn.Sequential(layer_init(nn.Linear(dim_list[0], dim_list[1])), nn.SiLU(), RMSNorm(dim_list[1]),
                                 layer_init(nn.Linear(dim_list[1], dim_list[2]), std=last_std if last_std else 2 ** 0.5))


class ActorCritic(nn.Module):
    """Same module tree / state_dict keys as the reference ActorCritic with use_mhc = True (ref :364-393)."""

    HEAD = 256

    def __init__(self, state_dim: int, action_dim: int, config=None):
        super().__init__()
        cfg = config or Config()
        if not getattr(cfg, "use_mhc", True):
            raise NotImplementedError("the PSCN backbone (use_mhc = False) is not on the B200 path")
        if cfg.mhc_rate != 2 or cfg.mhc_dim not in (128, 256):
            raise NotImplementedError("mHC kernels are built for mhc_rate = 2 and mhc_dim in {128, 256}")
        self.state_dim, self.action_dim = state_dim, action_dim
        self.dim, self.rate, self.n_layers, self.sk_it = cfg.mhc_dim, cfg.mhc_rate, cfg.mhc_layers, cfg.mhc_sk_it
        self.shared = MHCBackbone(state_dim, cfg.mhc_dim, cfg.mhc_rate, cfg.mhc_layers, cfg.mhc_sk_it)
        self.actor = MLP([cfg.mhc_dim, self.HEAD, action_dim], last_std=0.001)
        self.critic = MLP([cfg.mhc_dim, self.HEAD, 1], last_std=1.0)

    def param_order(self) -> List[str]:
        """Flat layout: the two head trunks adjacent (one [512, D] GEMM, one [512] bias, one [512] RMSNorm weight)."""
        names = ["shared.input_proj.weight", "shared.input_proj.bias"]
        for l in range(self.n_layers):
            for k in (1, 2):
                p = f"shared.layers.{l}"
                names += [f"{p}.mhc{k}.norm.weight", f"{p}.mhc{k}.w", f"{p}.mhc{k}.alpha", f"{p}.mhc{k}.beta",
                          f"{p}.linear{k}.weight", f"{p}.linear{k}.bias"]
        names += ["shared.final_norm.weight", "actor.mlp.0.weight", "critic.mlp.0.weight", "actor.mlp.0.bias", "critic.mlp.0.bias",
                  "actor.mlp.2.weight", "critic.mlp.2.weight", "actor.mlp.3.weight", "actor.mlp.3.bias", "critic.mlp.3.weight",
                  "critic.mlp.3.bias"]
        return names

    def to_engine(self, device) -> "ActorCriticEngine":
        return ActorCriticEngine(self, device)


class _Acts:
    """Activation / gradient scratch of the mHC ActorCritic for a fixed maximum batch M."""

    def __init__(self, M: int, D: int, S: int, HW: int, device, backward: bool):
        e = lambda *shape: torch.empty(*shape, device=device, dtype=f32)
        self.M = M
        self.x0 = e(M, D)                                   # input_proj output = both branches of the stage-0 input
        self.h = [None] + [e(M, 2 * D) for _ in range(S)]   # h[s] = input of stage s (s >= 1), h[S] = backbone output
        self.coef = [e(M, 24) for _ in range(S)]     # per stage: pre/post | P | r_ s dpost | dP | H (forward writes, backward completes)
        self.hpre = [e(M, D) for _ in range(S)]
        self.z = [e(M, D) for _ in range(S)]
        self.feat = e(M, D)
        self.zh = e(M, 2 * HW)                              # head trunks, pre-activation (actor | critic)
        self.yh = e(M, 2 * HW)                              # SiLU -> RMSNorm
        self.lv = e(M, 8)                                   # [:, :A] logits, [:, A] value
        if backward:
            self.dlv = torch.zeros(M, 8, device=device, dtype=f32)
            self.dyh, self.dzh = e(M, 2 * HW), e(M, 2 * HW)
            self.dfeat = e(M, D)
            self.dh_a, self.dh_b = e(M, 2 * D), e(M, 2 * D)  # ping-pong: dL/dh[s+1] and the partial / full dL/dh[s]
            self.dz, self.dhpre = e(M, D), e(M, D)
            self.dx0 = e(M, D)


class ActorCriticEngine:
    """Forward / backward of the mHC ActorCritic through the C ABI (csrc/mhc.cu + dense-layer kernels)."""

    def __init__(self, model: ActorCritic, device):
        self.model = model.to(device)
        self.fp = fp = FlatParams(self.model, model.param_order(), device)
        self.D, self.A, self.S, self.HW = model.dim, model.action_dim, 2 * model.n_layers, model.HEAD
        self.Din, self.sk = model.state_dim, model.sk_it
        assert self.A <= 7
        P, G = fp.p, fp.g
        self.Win, self.bin, self.gWin, self.gbin = P("shared.input_proj.weight"), P("shared.input_proj.bias"), \
            G("shared.input_proj.weight"), G("shared.input_proj.bias")
        self.stage, self.gstage, self.lin, self.glin = [], [], [], []
        for l in range(model.n_layers):
            for k in (1, 2):
                m, ln = f"shared.layers.{l}.mhc{k}", f"shared.layers.{l}.linear{k}"
                self.stage.append((P(m + ".norm.weight"), P(m + ".w"), P(m + ".alpha"), P(m + ".beta")))
                self.gstage.append((G(m + ".norm.weight"), G(m + ".w"), G(m + ".alpha"), G(m + ".beta")))
                self.lin.append((P(ln + ".weight"), P(ln + ".bias")))
                self.glin.append((G(ln + ".weight"), G(ln + ".bias")))
        self.gf, self.ggf = P("shared.final_norm.weight"), G("shared.final_norm.weight")
        HW, D = self.HW, self.D
        self.Wh = fp.span("actor.mlp.0.weight", "critic.mlp.0.weight", 2 * HW, D)
        self.bh = fp.span("actor.mlp.0.bias", "critic.mlp.0.bias", 1, 2 * HW).view(2 * HW)
        self.gh = fp.span("actor.mlp.2.weight", "critic.mlp.2.weight", 1, 2 * HW).view(2 * HW)
        self.gWh = fp.span("actor.mlp.0.weight", "critic.mlp.0.weight", 2 * HW, D, grad=True)
        self.gbh = fp.span("actor.mlp.0.bias", "critic.mlp.0.bias", 1, 2 * HW, grad=True).view(2 * HW)
        self.ggh = fp.span("actor.mlp.2.weight", "critic.mlp.2.weight", 1, 2 * HW, grad=True).view(2 * HW)
        self.Wa, self.ba, self.gWa, self.gba = P("actor.mlp.3.weight"), P("actor.mlp.3.bias"), G("actor.mlp.3.weight"), G("actor.mlp.3.bias")
        self.Wc, self.bc, self.gWc, self.gbc = P("critic.mlp.3.weight"), P("critic.mlp.3.bias"), G("critic.mlp.3.weight"), G("critic.mlp.3.bias")
        # the D x D layers and the head trunk are TMA-fed from pre-split tf32 weight images (csrc/wimages.cu): at the update's
        # 131072-row minibatches they run on the persistent warp-specialised GEMM (128-wide pair tiles for the D = 128 layers)
        if D % 128 == 0 and (2 * HW) % 128 == 0:
            mats = [(f"shared.layers.{l}.linear{k}.weight", f"shared.layers.{l}.linear{k}.weight", D, D)
                    for l in range(model.n_layers) for k in (1, 2)]
            mats.append(("actor.mlp.0.weight", "critic.mlp.0.weight", 2 * HW, D))
            fp.enable_weight_images(mats)
        self.workspace = None
        self.ws_rows = None

    def make_acts(self, M: int, backward: bool) -> _Acts:
        return _Acts(M, self.D, self.S, self.HW, self.fp.flat.device, backward)

    def alloc_workspace(self, M: int):
        D, HW = self.D, self.HW
        need = max(ops.backward_weight_workspace(M, 2 * HW, D), ops.backward_weight_workspace(M, D, D),
                   ops.backward_weight_workspace(M, D, self.Din), ops.backward_weight_workspace(M, self.A, HW))
        self.workspace = torch.empty(need, device=self.fp.flat.device, dtype=torch.uint8)
        self.ws_rows = torch.empty(ops.mhc_workspace_bytes(D, HW, 2), device=self.fp.flat.device, dtype=torch.uint8)

    def forward(self, x: torch.Tensor, acts: _Acts, M: int, row_index: Optional[torch.Tensor] = None):
        """logits = acts.lv[:M, :A], value = acts.lv[:M, A]."""
        D, S, HW, A, N_ = self.D, self.S, self.HW, self.A, _ffi.ACT_NONE
        ops.linear_forward(x, self.Win, self.bin, N_, row_index=row_index, out=acts.x0, M=M)
        # stage 0 reads x0 as both branches (h.repeat, ref :256-258): row stride D, branch stride 0
        ops.mhc_stage_forward(acts.x0, D=D, row_stride=D, branch_stride=0, M=M, params=self.stage[0], coef_cur=acts.coef[0],
                              h_pre=acts.hpre[0], sk_iters=self.sk)
        for s in range(S):
            W, b = self.lin[s]
            ops.linear_forward(acts.hpre[s], W, b, N_, out=acts.z[s], M=M)
            src, rs, bs = (acts.x0, D, 0) if s == 0 else (acts.h[s], 2 * D, D)
            last = s == S - 1
            ops.mhc_stage_forward(src, D=D, row_stride=rs, branch_stride=bs, M=M, z_prev=acts.z[s], coef_prev=acts.coef[s],
                                  h_cur=acts.h[s + 1], params=None if last else self.stage[s + 1],
                                  coef_cur=None if last else acts.coef[s + 1], h_pre=None if last else acts.hpre[s + 1],
                                  final_weight=self.gf if last else None, feat=acts.feat if last else None, sk_iters=self.sk)
        ops.linear_forward(acts.feat, self.Wh, self.bh, N_, out=acts.zh, M=M)
        ops.rmsnorm_forward(acts.zh, self.gh, acts.yh, M=M, W=HW, groups=2, silu=True)
        ops.linear_forward(acts.yh[:, :HW], self.Wa, self.ba, N_, out=acts.lv[:, :A], M=M)
        ops.linear_forward(acts.yh[:, HW:], self.Wc, self.bc, N_, out=acts.lv[:, A:A + 1], M=M)
        return acts.lv

    def backward(self, x: torch.Tensor, acts: _Acts, M: int, row_index: Optional[torch.Tensor] = None):
        """Given acts.dlv[:M] = dL/d(logits, value), fill the flat gradient buffer."""
        D, S, HW, A, N_, ws, wr = self.D, self.S, self.HW, self.A, _ffi.ACT_NONE, self.workspace, self.ws_rows
        dl, dv = acts.dlv[:, :A], acts.dlv[:, A:A + 1]
        ops.linear_backward(dl, acts.yh[:, :HW], self.Wa, self.gWa, self.gba, dx=acts.dyh[:, :HW], act_in=N_, workspace=ws, M=M)
        ops.linear_backward(dv, acts.yh[:, HW:], self.Wc, self.gWc, self.gbc, dx=acts.dyh[:, HW:], act_in=N_, workspace=ws, M=M)
        ops.rmsnorm_backward(acts.zh, self.gh, acts.dyh, acts.dzh, self.ggh, M=M, W=HW, groups=2, silu=True, workspace=wr)
        ops.linear_backward(acts.dzh, acts.feat, self.Wh, self.gWh, self.gbh, dx=acts.dfeat, act_in=N_, workspace=ws, M=M)
        # final branch-sum + RMSNorm: the same gradient for both branches of h[S]
        ops.rmsnorm_backward(acts.h[S], self.gf, acts.dfeat, acts.dh_a, self.ggf, M=M, W=D, groups=1, sum2=True, workspace=wr)
        for s in range(S - 1, -1, -1):
            src, rs, bs = (acts.x0, D, 0) if s == 0 else (acts.h[s], 2 * D, D)
            ops.mhc_stage_backward_a(src, acts.z[s], acts.dh_a, D=D, row_stride=rs, branch_stride=bs, M=M, dz=acts.dz,
                                     dh_partial=acts.dh_b, coef=acts.coef[s])
            W, _ = self.lin[s]
            gW, gb = self.glin[s]
            ops.linear_backward(acts.dz, acts.hpre[s], W, gW, gb, dx=acts.dhpre, act_in=N_, workspace=ws, M=M)
            ops.mhc_stage_backward_b(src, acts.dhpre, acts.coef[s], acts.dh_b, self.stage[s], self.gstage[s], D=D, row_stride=rs,
                                     branch_stride=bs, M=M, workspace=wr, dh=None if s == 0 else acts.dh_a,
                                     dx0=acts.dx0 if s == 0 else None)
        ops.linear_backward(acts.dx0, x, self.Win, self.gWin, self.gbin, row_index=row_index, workspace=ws, M=M)


class RolloutBuffer(base.RolloutBuffer):
    """The reference buffer additionally keeps old_entropies and next_value (ref :416-436)."""

    def __init__(self, T: int, N: int, D: int, device):
        super().__init__(T, N, D, device)
        self.entropy = torch.zeros(T, N, device=device, dtype=f32)

    @property
    def old_entropies(self): return self.entropy[:self.filled].reshape(-1)
    @property
    def next_value(self): return self.v_last


class PPOTrainer(base.PPOTrainer):
    def __init__(self, config: Config):
        if getattr(config, "clip_cov_ratio", 0.0) > 0:
            raise NotImplementedError("covariance clipping (clip_cov_ratio > 0) is not on the B200 path; the reference default is 0")
        config.hidden_dim = getattr(config, "hidden_dim", 256)
        config.anneal_lr = False   # base-class knob; this trainer anneals after the update (ref :659-666)
        self._ent_coef_t = None
        super().__init__(config)
        cfg, dev = self.cfg, self.device
        self.buffer = RolloutBuffer(self.T, self.N, self.env.obs_dim, dev)
        self.episode_rewards = deque(maxlen=10)
        self.lr = cfg.lr
        self.ent_coef = cfg.entropy_coef
        if self.rank == 0:
            print(f"Model device: {next(self.model.parameters()).device}")

    # ---- construction hooks
    def _make_model(self, state_dim: int, action_dim: int) -> nn.Module:
        return ActorCritic(state_dim, action_dim, config=self.cfg)

    def _make_acts(self, M: int, backward: bool, n_actions: Optional[int] = None):
        return self.net.make_acts(M, backward)

    def _make_loss_cfg(self):
        cfg = self.cfg
        self._ent_coef_t = torch.full((1,), float(cfg.entropy_coef), device=self.device, dtype=f32)
        return _ffi.PPOCfg(mode=_ffi.PPO_FULL, clip_eps_min=cfg.clip_eps_min, clip_eps_max=cfg.clip_eps_max, dual_clip=cfg.dual_clip,
                           value_coef=0.5, entropy_coef=cfg.entropy_coef, erc_low=cfg.erc_beta_low, erc_high=cfg.erc_beta_high,
                           d_entropy_coef=self._ent_coef_t.data_ptr())

    def _sample_step(self, acts, t: int):
        buf, N, A = self.buffer, self.N, self.env.n_actions
        ops.sample_categorical(acts.lv[:, :A], seed=self.seed, first_id=self.rank * N, draw=t, draw_base=self.ctr_action,
                               action=buf.action[t], logp=buf.log_prob[t], entropy=buf.entropy[t], value_in=acts.lv[:, A:A + 1],
                               value_out=buf.value[t])

    def _loss_step(self, acts):
        buf, A = self.buffer, self.env.n_actions
        ops.ppo_loss(acts.lv[:, :A], acts.lv[:, A:A + 1], buf.action.view(-1), buf.log_prob.view(-1), buf.adv.view(-1),
                     buf.ret.view(-1), self.loss_cfg, row_index=self._idx_cur, entropy_old=buf.entropy.view(-1),
                     dlogits=acts.dlv[:, :A], dvalue=acts.dlv[:, A:A + 1], metrics=self.metrics)

    # ---- the reference's method names
    def collect_experience(self):
        return self.collect_rollout()

    def compute_advantages(self):
        """gymrl_gae dialect 2 (ref :507-535).  Returns (adv_actor, returns) [T, N]."""
        buf, cfg = self.buffer, self.cfg
        ops.gae(buf.reward, buf.value, buf.v_last, buf.done, cfg.gamma, cfg.lam_actor, cfg.lam_critic, dialect=2, adv=buf.adv, ret=buf.ret)
        return buf.adv, buf.ret

    def update_model(self, advantages=None, returns=None, read_metrics: bool = True) -> dict:
        """ref :537-679.  advantages / returns default to the buffers compute_advantages() filled (no normalisation here)."""
        cfg, buf = self.cfg, self.buffer
        if advantages is not None and advantages.data_ptr() != buf.adv.data_ptr():
            buf.adv.copy_(torch.as_tensor(advantages, dtype=f32, device=self.device).view_as(buf.adv))
        if returns is not None and returns.data_ptr() != buf.ret.data_ptr():
            buf.ret.copy_(torch.as_tensor(returns, dtype=f32, device=self.device).view_as(buf.ret))
        self._run_epochs()
        if cfg.anneal:   # after the update, from the post-rollout step count
            frac = 1 - self.step_count / cfg.max_train_steps
            self.lr = cfg.lr * frac
            for group in self.optimizer.param_groups:
                group["lr"] = self.lr
            self.ent_coef = cfg.entropy_coef * frac
            self._ent_coef_t.fill_(float(self.ent_coef))
        if not read_metrics:
            return {}
        m = self.metrics.tolist()
        k = max(m[7], 1.0)
        return {"policy_loss": m[0] / k, "value_loss": m[1] / k, "entropy": m[2] / k, "clip_frac": m[3] / k, "approx_kl": m[4] / k,
                "erc_clip_frac": m[5] / k}

    def update(self, next_value=None, read_metrics: bool = True) -> dict:
        if next_value is not None:
            nv = torch.as_tensor(next_value, dtype=f32, device=self.device).reshape(-1)
            self.buffer.v_last.copy_(nv.expand(self.N) if nv.numel() == 1 else nv)
        self.compute_advantages()
        return self.update_model(read_metrics=read_metrics)

    def train_iteration(self):
        self.collect_experience()
        metrics = self.update(None)
        avg_reward, total = self._refresh_episode_rewards()
        return metrics, avg_reward, total

    def _refresh_episode_rewards(self):
        mean_ret, _, total = self.env.episode_stats(10)
        if total > self._episodes_seen:
            self._episodes_seen = total
            self.episode_rewards.clear()
            self.episode_rewards.extend([mean_ret] * int(min(total, 10)))
        return mean_ret, total

    def train(self):
        while self.step_count < self.cfg.max_train_steps:
            self.collect_experience()
            advantages, returns = self.compute_advantages()
            m = self.update_model(advantages, returns)
            self._refresh_episode_rewards()
            if self.rank == 0:
                print(f"Policy Loss: {m['policy_loss']:.4f} | Value Loss: {m['value_loss']:.4f} | Entropy: {m['entropy']:.4f} | "
                      f"KL: {m['approx_kl']:.4f} | Clip Frac: {m['clip_frac']:.2%} | ERC Clip: {m['erc_clip_frac']:.2%} | "
                      f"LR: {self.optimizer.param_groups[0]['lr']:.6f} | Ent Coef: {self.ent_coef:.4f}")
                if len(self.episode_rewards) > 0:
                    print(f"Step: {self.step_count}, Avg Reward: {np.mean(self.episode_rewards):.2f}")

    def eval(self, num_episodes: int = 10):
        rewards = super().eval(num_episodes)
        print(f"Test Results: Mean Reward {np.mean(rewards):.2f} +/- {np.std(rewards):.2f}")
        return None   # the reference's eval() returns nothing (ref :693-719)

    def test(self):
        self.eval(num_episodes=10)
        print("(visual test skipped: the device env has no renderer)")


def main():
    config = Config()
    config.num_envs, config.num_steps, config.num_minibatches = 4096, 128, 4
    ppo = PPOTrainer(config)

    def signal_handler(signum, frame):
        print("\nCtrl+C detected, stopping training and starting test...")
        ppo.test()
        sys.exit(0)

    signal.signal(signal.SIGINT, signal_handler)
    try:
        ppo.train()
    except KeyboardInterrupt:
        print("\nTraining interrupted, starting test...")
        ppo.test()
    else:
        ppo.test()


if __name__ == "__main__":
    main()
