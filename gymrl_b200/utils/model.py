"""utils.model names used by the in-scope legacy scripts (reference utils/model.py): initialize_weights, MLP,
NoisyLinear, PSCN, ModelLoader, StateManager.  These are host-side torch module *definitions* — a script that builds its
network from them and writes its own autograd update() keeps running through torch; the B200 kernels sit under the
re-hosted trainers in gymrl_b200/algorithms/ and under utils.buffer / utils.normalization / the env.  Conv, attention
and GRU blocks of the reference file (:112-324) belong to the image / recurrent scripts and are out of scope.
"""
from __future__ import annotations

import math
import os

import torch
from torch import nn
from torch.nn import functional as F

from .buffer import Queue


def initialize_weights(layer, init_type='kaiming', nonlinearity='leaky_relu'):
    if isinstance(layer, (nn.Linear, nn.Conv2d)):
        init = {'kaiming': lambda w: nn.init.kaiming_uniform_(w, nonlinearity=nonlinearity),
                'xavier': nn.init.xavier_uniform_,
                'orthogonal': lambda w: nn.init.orthogonal_(w, gain=math.sqrt(2))}.get(init_type)
        if init is None:
            raise ValueError(f"Unknown initialization type: {init_type}")
        init(layer.weight)
        if layer.bias is not None:
            nn.init.zeros_(layer.bias)
    return layer


class MLP(nn.Module):
    """Linear [-> LayerNorm] -> activation stack; the last layer is bare unless last_act (ref :26-52)."""

    def __init__(self, dim_list, activation=nn.PReLU(), last_act=False, use_norm=False, linear=nn.Linear, *args, **kwargs):
        super().__init__()
        assert dim_list, "Dim list can't be empty!"
        mods, n = [], len(dim_list) - 1
        for i, (a, b) in enumerate(zip(dim_list[:-1], dim_list[1:])):
            mods.append(initialize_weights(linear(a, b, *args, **kwargs)))
            if i < n - 1 or last_act:
                if use_norm:
                    mods.append(nn.LayerNorm(b))
                mods.append(activation)
        self.mlp = nn.Sequential(*mods)

    def forward(self, x):
        return self.mlp(x)


class NoisyLinear(nn.Module):
    """Factorised-noise linear layer, noise redrawn on every training-mode forward (ref :56-108).  The device
    kernels for the same arithmetic are gymrl_noisy_sample / gymrl_noisy_compose / gymrl_noisy_backward."""

    def __init__(self, in_features, out_features, sigma_init=0.5):
        super().__init__()
        self.in_features, self.out_features, self.sigma_init = in_features, out_features, sigma_init
        self.weight_mu = nn.Parameter(torch.empty(out_features, in_features))
        self.weight_sigma = nn.Parameter(torch.empty(out_features, in_features))
        self.bias_mu = nn.Parameter(torch.empty(out_features))
        self.bias_sigma = nn.Parameter(torch.empty(out_features))
        self.register_buffer('weight_epsilon', torch.zeros(out_features, in_features), persistent=False)
        self.register_buffer('bias_epsilon', torch.zeros(out_features), persistent=False)
        self.reset_parameters()
        self.reset_noise()

    def reset_parameters(self):
        bound = 1 / math.sqrt(self.in_features)
        nn.init.uniform_(self.weight_mu, -bound, bound)
        nn.init.uniform_(self.bias_mu, -bound, bound)
        nn.init.constant_(self.weight_sigma, self.sigma_init / math.sqrt(self.in_features))
        nn.init.constant_(self.bias_sigma, self.sigma_init / math.sqrt(self.out_features))

    def scale_noise(self, size: int):
        x = torch.randn(size)
        return x.sign() * x.abs().sqrt()

    def reset_noise(self):
        eps_in, eps_out = self.scale_noise(self.in_features), self.scale_noise(self.out_features)
        self.weight_epsilon.copy_(torch.outer(eps_out, eps_in))
        self.bias_epsilon.copy_(eps_out)

    def forward(self, x):
        if not self.training:
            return F.linear(x, self.weight_mu, self.bias_mu)
        self.reset_noise()
        return F.linear(x, self.weight_mu + self.weight_sigma * self.weight_epsilon, self.bias_mu + self.bias_sigma * self.bias_epsilon)

    def __repr__(self):
        return f"{type(self).__name__}(in_features={self.in_features}, out_features={self.out_features}, sigma_init={self.sigma_init})"


class PSCN(nn.Module):
    """Split-and-concat MLP: every level keeps half of its activations as output and feeds the other half on (ref :256-286)."""

    def __init__(self, input_dim, output_dim, depth=4, linear=nn.Linear):
        super().__init__()
        min_dim = 2 ** (depth - 1)
        assert depth >= 1, "depth must be at least 1"
        assert output_dim >= min_dim and output_dim % min_dim == 0, f"output_dim must be a multiple of {min_dim} for depth {depth}"
        self.output_dim = output_dim
        widths = [output_dim >> i for i in range(depth)]
        ins = [input_dim] + [w // 2 for w in widths[:-1]]
        self.layers = nn.ModuleList(MLP([i, w], last_act=True, linear=linear) for i, w in zip(ins, widths))

    def forward(self, x):
        keep = []
        for level, layer in enumerate(self.layers):
            x = layer(x)
            if level + 1 < len(self.layers):
                half = x.shape[-1] // 2
                keep.append(x[..., :half])
                x = x[..., half:]
        keep.append(x)
        return torch.cat(keep, dim=-1)


class ModelLoader:
    """Checkpoint mixin (ref :330-375): one torch.save file ./checkpoints/<algo>_<env>.pth holding `<attr>_state_dict` for every
    attribute with a state_dict (networks, optimizers, the device Normalization / RewardScaling) and the plain attributes
    (learn_step ...), except cfg / memory / state_buffer."""
    _SKIP = ('state_buffer', 'cfg', 'memory')

    def __init__(self, cfg):
        cfg.save_path = f'./checkpoints/{cfg.algo_name}_{cfg.env_name.replace("/", "-")}.pth'
        self.cfg = cfg
        os.makedirs(os.path.dirname(cfg.save_path), exist_ok=True)

    def save_model(self):
        from . import install, normalization as _nz
        install()   # the normalisers pickle as `utils.normalization.*` (the reference's class paths)
        state = {}
        for key, value in self.__dict__.items():
            if key in self._SKIP:
                continue
            if isinstance(value, (_nz.Normalization, _nz.RewardScaling, _nz.RunningMeanStd)):
                state[key] = value      # a pickled plain attribute, like the reference (its classes have no state_dict)
            elif hasattr(value, 'state_dict'):
                state[f'{key}_state_dict'] = value.state_dict()
            else:
                state[key] = value
        torch.save(state, self.cfg.save_path)

    def load_model(self):
        from . import install
        install()   # a reference-written checkpoint names `utils.normalization.Normalization`: resolve it to the device class
        checkpoint = torch.load(self.cfg.save_path, map_location=self.cfg.device, weights_only=False)
        for key, value in checkpoint.items():
            if key in self._SKIP:
                continue
            if key.endswith('_state_dict'):
                target = getattr(self, key[:-len('_state_dict')], None)
                if target is not None:
                    target.load_state_dict(value)
            else:
                setattr(self, key, value)


class StateManager:
    def __init__(self, buffer_size=100):
        self.state_buffer = Queue(buffer_size)

    def save_state(self, *args):
        self.state_buffer.put(args)

    def load_state(self):
        return self.state_buffer.sample()
