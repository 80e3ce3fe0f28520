"""Generate tests/golden/mhc_actor_critic.npz from the UNMODIFIED reference module
(/root/reference/algorithms/ppo_full_lunarlander.py: ActorCritic with the MHCBackbone, :76-412).

TEST INFRASTRUCTURE.  Run in the build container only:   python -m oracle.make_golden_mhc

The reference initialises the mHC projection `w` to zeros (:125), which makes the mapping input-independent; to
pin every term of the forward and backward pass the fixture perturbs all parameters with seeded noise first, then
records logits / value and the torch-autograd gradients of  sum(logits * Gl) + sum(value * Gv).
"""
from pathlib import Path

import numpy as np
import torch

from . import ref_loader as rl

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def main():
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(8, n_actions=4, max_steps=1000))
    m = rl.load("algorithms/ppo_full_lunarlander.py")
    cfg = m.Config()
    torch.manual_seed(0)
    net = m.ActorCritic(8, 4, config=cfg)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith(".w"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif name.endswith(".alpha"):
                p.copy_(0.5 + 0.5 * torch.rand(p.shape, generator=g))
            elif name.endswith(".beta") or name.endswith("norm.weight") or name.endswith("mlp.2.weight"):
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
            elif name.endswith("bias"):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
        net.actor.mlp[3].weight.mul_(100.0)   # last_std = 0.001 would hide the actor path in the gradients
    B = 48
    x = torch.randn(B, 8, generator=g)
    Gl = torch.randn(B, 4, generator=g)
    Gv = torch.randn(B, 1, generator=g)
    logits, value = net(x)
    loss = (logits * Gl).sum() + (value * Gv).sum()
    loss.backward()
    out = {"x": x.numpy(), "Gl": Gl.numpy(), "Gv": Gv.numpy(), "logits": logits.detach().numpy(), "value": value.detach().numpy(),
           "rate": cfg.mhc_rate, "layers": cfg.mhc_layers, "sk_it": cfg.mhc_sk_it, "dim": cfg.mhc_dim,
           "source": "algorithms/ppo_full_lunarlander.py:76-412 (ActorCritic.forward + autograd)"}
    for k, v in net.state_dict().items():
        out["p:" + k] = v.numpy()
    for k, p in net.named_parameters():
        out["g:" + k] = p.grad.numpy()
    # default initialisation (w = 0): outputs only, parameters re-derivable from the perturbed set is not possible, so
    # store the few tensors that differ in closed form: none needed - the perturbed case covers the arithmetic.
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / "mhc_actor_critic.npz", **out)
    print("wrote", OUT / "mhc_actor_critic.npz", sum(v.size for v in out.values() if hasattr(v, "size")), "values")
    print("param count", sum(p.numel() for p in net.parameters()))


if __name__ == "__main__":
    main()
