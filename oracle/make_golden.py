"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes from /root/reference.

TEST INFRASTRUCTURE.  Run in the build container only (the reference tree does not exist on the GPU box):

    python -m oracle.make_golden

The reference has no tests / golden vectors of its own, so these fixtures — outputs
of the reference's own code on seeded inputs — are what pins the oracle and the CUDA path.  Every
fixture records the reference file:line that produced it.
"""
from __future__ import annotations

import random
from collections import deque
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

from . import ref_loader as rl

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _ppo_module():
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(8, n_actions=4, max_steps=1000))
    return rl.load("algorithms/ppo_lunarlander.py")


def _bare_ppo_trainer(m, cfg=None):
    t = m.PPOTrainer.__new__(m.PPOTrainer)
    t.cfg = cfg or m.Config()
    t.cfg.device = "cpu"
    t.buffer = m.RolloutBuffer()
    return t


def gen_gae():
    """PPOTrainer.compute_gae (algorithms/ppo_lunarlander.py:179-196) and compute_advantages
    (algorithms/ppo_full_lunarlander.py:507-535) on seeded inputs; fp64 outputs as the reference produces them."""
    m = _ppo_module()
    rng = np.random.default_rng(0)
    out = {}
    # (a) the reference's own shape: one env, T = 2048
    T = 2048
    r = rng.standard_normal(T).astype(np.float32)
    v = rng.standard_normal(T).astype(np.float32)
    d = rng.random(T) < 0.01
    nv = np.float32(rng.standard_normal())
    t = _bare_ppo_trainer(m)
    t.buffer.rewards = [float(x) for x in r]
    t.buffer.values = [float(x) for x in v]
    t.buffer.dones = [bool(x) for x in d]
    adv, ret = t.compute_gae(float(nv))
    assert adv.dtype == np.float64
    out.update(a_reward=r, a_value=v, a_done=d.astype(np.uint8), a_next_value=nv, a_adv=adv, a_ret=ret,
               a_gamma=t.cfg.gamma, a_lam=t.cfg.gae_lambda)
    # (b) [T=96][N=40] lockstep batch: the reference function applied per env column
    T, N = 96, 40
    r = rng.standard_normal((T, N)).astype(np.float32)
    v = rng.standard_normal((T, N)).astype(np.float32)
    d = rng.random((T, N)) < 0.05
    nv = rng.standard_normal(N).astype(np.float32)
    adv = np.zeros((T, N)); ret = np.zeros((T, N))
    for n in range(N):
        t = _bare_ppo_trainer(m)
        t.buffer.rewards = [float(x) for x in r[:, n]]
        t.buffer.values = [float(x) for x in v[:, n]]
        t.buffer.dones = [bool(x) for x in d[:, n]]
        adv[:, n], ret[:, n] = t.compute_gae(float(nv[n]))
    out.update(b_reward=r, b_value=v, b_done=d.astype(np.uint8), b_next_value=nv, b_adv=adv, b_ret=ret)
    # (c) decoupled lambdas (ppo_full)
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(8, n_actions=4, max_steps=1000))
    mf = rl.load("algorithms/ppo_full_lunarlander.py")
    T = 300
    r = rng.standard_normal(T).astype(np.float32)
    v = rng.standard_normal(T).astype(np.float32)
    d = rng.random(T) < 0.02
    nv = np.float32(rng.standard_normal())
    tf = mf.PPOTrainer.__new__(mf.PPOTrainer)
    tf.cfg = mf.Config()
    tf.cfg.lam_actor, tf.cfg.lam_critic = 0.95, 0.9
    tf.buffer = mf.RolloutBuffer()
    tf.buffer.rewards = [float(x) for x in r]
    tf.buffer.values = [torch.tensor(float(x)) for x in v]  # ppo_full stores 0-dim fp32 tensors
    tf.buffer.dones = [bool(x) for x in d]
    tf.buffer.next_value = torch.tensor(float(nv))
    adv_a, ret = tf.compute_advantages()
    out.update(c_reward=r, c_value=v, c_done=d.astype(np.uint8), c_next_value=nv, c_adv=adv_a, c_ret=ret,
               c_gamma=tf.cfg.gamma, c_lam_actor=0.95, c_lam_critic=0.9)
    np.savez(OUT / "gae_algorithms.npz", **out)

    # utils dialect: ReplayBuffer_on_policy.compute_advantage (utils/buffer.py:21-35)
    ub = rl.load("utils/buffer.py")
    import types
    cfg = types.SimpleNamespace(gamma=0.99, lamda=0.95, device="cpu")
    T = 500
    r = torch.tensor(rng.standard_normal((T, 1)).astype(np.float32))
    v = torch.tensor(rng.standard_normal((T, 1)).astype(np.float32))
    v2 = torch.tensor(rng.standard_normal((T, 1)).astype(np.float32))
    done = torch.tensor((rng.random((T, 1)) < 0.03).astype(np.float32))
    dw = done * torch.tensor((rng.random((T, 1)) < 0.5).astype(np.float32))
    buf = ub.ReplayBuffer_on_policy(cfg)
    adv_n, v_target = buf.compute_advantage(r, done, dw, v, v2)
    raw_adv = (v_target - v)  # exact: v_target = adv + values in fp32
    np.savez(OUT / "gae_utils.npz", reward=r.numpy(), value=v.numpy(), next_value=v2.numpy(), done=done.numpy().astype(np.uint8),
             dw=dw.numpy().astype(np.uint8), adv_normalized=adv_n.numpy(), v_target=v_target.numpy(), gamma=0.99, lamda=0.95)


def gen_categorical():
    """Categorical(logits).sample()/log_prob/entropy as used by ActorCritic.get_action
    (algorithms/ppo_lunarlander.py:92-104) with the Exp(1) noise torch draws."""
    from torch.distributions import Categorical
    g = torch.Generator().manual_seed(1)
    N, A = 512, 4
    logits = torch.randn(N, A, generator=g) * 2.0
    logits[:8] = 0.0  # exact ties in p: the noise decides
    dist = Categorical(logits=logits)
    torch.manual_seed(1234)
    action = dist.sample()
    torch.manual_seed(1234)
    q = torch.empty_like(dist.probs).exponential_(1)
    assert torch.equal(torch.argmax(dist.probs / q, dim=-1), action)
    vals = (dist.probs / q).sort(dim=-1, descending=True).values
    margin = (vals[:, 0] - vals[:, 1]) / vals[:, 0]
    np.savez(OUT / "categorical.npz", logits=logits.numpy(), noise=q.numpy(), action=action.numpy().astype(np.int32),
             log_prob=dist.log_prob(action).numpy(), entropy=dist.entropy().numpy(), margin=margin.numpy(),
             greedy=logits.argmax(dim=-1).numpy().astype(np.int32))


def gen_ppo_loss():
    """Loss-level fixtures: d loss / d(logits, V) from the reference's own update loops, obtained by
    swapping the network for a lookup table of leaf (logits, value) parameters indexed by state[:, 0].
      - PPOTrainer.update, algorithms/ppo_lunarlander.py:261-322 (dual-clip form)
      - PPOTrainer.update_model, algorithms/ppo_full_lunarlander.py:573-657 (ERC mask, clip-higher)"""
    m = _ppo_module()
    rng = np.random.default_rng(5)
    B, A = 384, 4
    logits0 = (rng.standard_normal((B, A)) * 1.5).astype(np.float32)
    value0 = rng.standard_normal(B).astype(np.float32)
    actions = rng.integers(0, A, B)
    ln = torch.log_softmax(torch.tensor(logits0), -1)
    old_lp = (ln[torch.arange(B), torch.tensor(actions)] + torch.tensor(rng.standard_normal(B).astype(np.float32)) * 0.3).numpy()
    rewards = rng.standard_normal(B).astype(np.float32)
    vals_old = rng.standard_normal(B).astype(np.float32)
    dones = rng.random(B) < 0.05

    class Table(m.ActorCritic):
        def __init__(self):
            nn.Module.__init__(self)
            self.l = nn.Parameter(torch.tensor(logits0))
            self.v = nn.Parameter(torch.tensor(value0))

        def forward(self, x):
            idx = x[:, 0].long()
            return self.l[idx], self.v[idx].unsqueeze(-1)

    t = _bare_ppo_trainer(m)
    t.cfg.num_epochs, t.cfg.batch_size, t.cfg.max_grad_norm = 1, B, 1e9
    t.model = Table()
    t.optimizer = torch.optim.SGD(t.model.parameters(), lr=0.0)
    for i in range(B):
        s = np.zeros(8, np.float32); s[0] = i
        t.buffer.add(s, int(actions[i]), float(old_lp[i]), float(vals_old[i]), float(rewards[i]), bool(dones[i]))
    np.random.seed(0)
    adv, ret = t.compute_gae(0.37)
    adv_n = (adv - adv.mean()) / (adv.std() + 1e-8)
    metrics = t.update(0.37)
    out = dict(logits=logits0, value=value0, action=actions.astype(np.int32), logp_old=old_lp.astype(np.float32),
               adv=adv_n.astype(np.float32), ret=ret.astype(np.float32), dlogits=t.model.l.grad.numpy(), dvalue=t.model.v.grad.numpy(),
               **{"m_" + k: np.float64(v) for k, v in metrics.items()},
               clip_eps=t.cfg.clip_eps, dual_clip=t.cfg.dual_clip, value_coef=t.cfg.value_coef, entropy_coef=t.cfg.entropy_coef)
    np.savez(OUT / "ppo_loss_dualclip.npz", **out)

    # ---- ppo_full ----
    mf = rl.load("algorithms/ppo_full_lunarlander.py")
    tf = mf.PPOTrainer.__new__(mf.PPOTrainer)
    tf.cfg = mf.Config()
    tf.cfg.num_epochs, tf.cfg.batch_size, tf.cfg.max_grad_norm, tf.cfg.anneal = 1, B, 1e9, False
    tf.ent_coef = tf.cfg.entropy_coef
    tf.lr = tf.cfg.lr
    tf.step_count = 0
    p = torch.softmax(torch.tensor(logits0), -1)
    H_new = -(p * torch.log(p)).sum(-1).numpy()
    old_ent = (H_new * (1.0 + rng.standard_normal(B).astype(np.float32) * 0.05)).astype(np.float32)  # ~1/4 outside the ERC band

    class TableF(nn.Module):
        def __init__(self):
            super().__init__()
            self.l = nn.Parameter(torch.tensor(logits0))
            self.v = nn.Parameter(torch.tensor(value0))

        def forward(self, x):
            idx = x[:, 0].long()
            return self.l[idx], self.v[idx].unsqueeze(-1)

    tf.model = TableF()
    tf.optimizer = torch.optim.SGD(tf.model.parameters(), lr=0.0)
    tf.buffer = mf.RolloutBuffer()
    tf.buffer.states = [[float(i)] + [0.0] * 7 for i in range(B)]
    tf.buffer.actions = [int(a) for a in actions]
    tf.buffer.log_probs = [float(x) for x in old_lp]
    tf.buffer.old_entropies = [float(x) for x in old_ent]
    advf = rng.standard_normal(B).astype(np.float32)
    retf = rng.standard_normal(B).astype(np.float32)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        tf.update_model(advf, retf)
    np.savez(OUT / "ppo_loss_full.npz", logits=logits0, value=value0, action=actions.astype(np.int32),
             logp_old=old_lp.astype(np.float32), entropy_old=old_ent, adv=advf, ret=retf, dlogits=tf.model.l.grad.numpy(),
             dvalue=tf.model.v.grad.numpy(), clip_eps_min=tf.cfg.clip_eps_min, clip_eps_max=tf.cfg.clip_eps_max,
             dual_clip=tf.cfg.dual_clip, entropy_coef=tf.cfg.entropy_coef, erc_low=tf.cfg.erc_beta_low, erc_high=tf.cfg.erc_beta_high)


def gen_ppo_update():
    """End-to-end PPOTrainer.update (algorithms/ppo_lunarlander.py:233-330) with the real ActorCritic:
    (1) one full-batch minibatch, unclipped -> per-parameter gradients; (2) 1 epoch x 4 minibatches with
    Adam(lr 3e-4, eps 1e-5) + clip_grad_norm_(0.5) -> parameters after the update."""
    m = _ppo_module()
    rng = np.random.default_rng(11)
    torch.manual_seed(3)
    B = 512
    states = rng.standard_normal((B, 8)).astype(np.float32)
    model = m.ActorCritic(8, 4, 256)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        logits, values = model(torch.tensor(states))
    from torch.distributions import Categorical
    dist = Categorical(logits=logits)
    actions = dist.sample()
    old_lp = dist.log_prob(actions) + 0.1 * torch.randn(B)
    rewards = rng.standard_normal(B).astype(np.float32)
    dones = rng.random(B) < 0.03

    def fill(t):
        for i in range(B):
            t.buffer.add(states[i], int(actions[i]), float(old_lp[i]), float(values[i, 0]), float(rewards[i]), bool(dones[i]))

    out = dict(states=states, action=actions.numpy().astype(np.int32), logp_old=old_lp.numpy(), value_old=values[:, 0].numpy(),
               reward=rewards, done=dones.astype(np.uint8), next_value=np.float32(0.21))
    for k, v in sd0.items():
        out["w0_" + k] = v.numpy()
    # (1) gradients
    t = _bare_ppo_trainer(m)
    t.cfg.num_epochs, t.cfg.batch_size, t.cfg.max_grad_norm = 1, B, 1e9
    t.model = m.ActorCritic(8, 4, 256); t.model.load_state_dict(sd0)
    t.optimizer = torch.optim.SGD(t.model.parameters(), lr=0.0)
    fill(t)
    np.random.seed(0)
    met = t.update(0.21)
    for k, p in t.model.named_parameters():
        out["g_" + k] = p.grad.numpy().copy()
    out.update({"m1_" + k: np.float64(v) for k, v in met.items()})
    # (2) one epoch of 4 minibatches with the real optimiser
    t = _bare_ppo_trainer(m)
    t.cfg.num_epochs, t.cfg.batch_size = 1, B // 4
    t.model = m.ActorCritic(8, 4, 256); t.model.load_state_dict(sd0)
    t.optimizer = torch.optim.Adam(t.model.parameters(), lr=t.cfg.lr, eps=1e-5)
    fill(t)
    np.random.seed(42)
    perm = np.arange(B); np.random.shuffle(perm)
    np.random.seed(42)
    met = t.update(0.21)
    out["perm"] = perm.astype(np.int32)
    for k, v in t.model.state_dict().items():
        out["w1_" + k] = v.numpy().copy()
    out.update({"m2_" + k: np.float64(v) for k, v in met.items()})
    out.update(gamma=t.cfg.gamma, lam=t.cfg.gae_lambda, lr=t.cfg.lr, max_grad_norm=t.cfg.max_grad_norm)
    np.savez(OUT / "ppo_update.npz", **out)


def main():
    assert rl.available(), "run in the build container (needs /root/reference)"
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(1)
    gen_gae()
    gen_categorical()
    gen_ppo_loss()
    gen_ppo_update()
    from . import make_golden_offpolicy as off
    off.main()
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
