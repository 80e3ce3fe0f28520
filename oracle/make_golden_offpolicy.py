"""Off-policy golden fixtures from the UNMODIFIED reference classes (SumTree / PrioritizedNStepBuffer /
DQN / Rainbow / SAC / TD3 update()).  TEST INFRASTRUCTURE; run via `python -m oracle.make_golden`.

Randomness is pinned the way SURVEY §8c describes: NumPy / torch generators are seeded, the draws the
reference consumes are re-generated from the same seed and stored next to the outputs, and replay
sampling is replaced by a recorded index batch.
"""
from __future__ import annotations

import random
from collections import deque
from pathlib import Path

import numpy as np
import torch

from . import ref_loader as rl

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _sd(module, prefix):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


# ------------------------------------------------------------------------------------------------ SumTree / PER
def gen_sumtree():
    """SumTree (algorithms/rainbow_dqn_cartpole.py:116-152): rotation for non-power-of-two capacity, tie rule, drift."""
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(4, n_actions=2, max_steps=500))
    m = rl.load("algorithms/rainbow_dqn_cartpole.py")
    out = {}
    rng = np.random.default_rng(0)
    for cap in (37, 64, 20000):
        t = m.SumTree(cap)
        n_upd = 3 * cap if cap < 1000 else 30000
        idx = rng.integers(0, cap, n_upd)
        pr = rng.random(n_upd) ** 0.6
        if cap == 64:
            pr = np.round(pr * 8) + 1.0          # integer priorities: exact ties in the descent
        for i, p in zip(idx, pr):
            t.update(int(i), float(p))
        vs = np.concatenate([rng.random(200) * t.priority_sum, [0.0, t.priority_sum, t.tree[2 * 0 + 1]]])
        if cap == 64:
            vs = np.concatenate([vs, np.cumsum(t.tree[cap - 1:])[:40]])  # v exactly on prefix boundaries
        res = [t.get_index(float(v)) for v in vs]
        out.update({f"c{cap}_idx": idx.astype(np.int32), f"c{cap}_prio": pr, f"c{cap}_tree": t.tree.copy(), f"c{cap}_v": vs,
                    f"c{cap}_leaf": np.array([r[0] for r in res], np.int32), f"c{cap}_leafp": np.array([r[1] for r in res]),
                    f"c{cap}_max": t.priority_max})
    np.savez(OUT / "sumtree.npz", **out)


def gen_per_nstep():
    """PrioritizedNStepBuffer store/sample/update_priorities (rainbow :155-264) on a synthetic 2-env-free stream."""
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(4, n_actions=2, max_steps=500))
    m = rl.load("algorithms/rainbow_dqn_cartpole.py")
    cfg = m.Config(); cfg.device = "cpu"; cfg.memory_capacity = 300; cfg.batch_size = 64
    buf = m.PrioritizedNStepBuffer(cfg, 4)
    rng = np.random.default_rng(1)
    T = 420   # wraps the 300-slot ring
    S = rng.standard_normal((T, 4)).astype(np.float32); S2 = rng.standard_normal((T, 4)).astype(np.float32)
    A = rng.integers(0, 2, T); R = rng.standard_normal(T).astype(np.float32)
    done = rng.random(T) < 0.15
    timelimit = done & (rng.random(T) < 0.3)            # done but not terminal (ref :376)
    terminal = done & ~timelimit
    for t in range(T):
        buf.store_transition(S[t], int(A[t]), float(R[t]), S2[t], bool(terminal[t]), bool(done[t]))
    out = dict(S=S, S2=S2, A=A.astype(np.int32), R=R, done=done.astype(np.uint8), terminal=terminal.astype(np.uint8),
               b_state=buf.buffer["state"].copy(), b_action=buf.buffer["action"].copy(), b_reward=buf.buffer["reward"].copy(),
               b_next_state=buf.buffer["next_state"].copy(), b_terminal=buf.buffer["terminal"].copy(),
               tree_after_store=buf.sum_tree.tree.copy(), count=buf.count, size=buf.current_size)
    np.random.seed(7)
    u = np.random.random_sample(cfg.batch_size)
    np.random.seed(7)
    batch, bidx, w = buf.sample(1234, 250000)
    td = rng.standard_normal(cfg.batch_size).astype(np.float32)
    bidx_dup = bidx.copy(); bidx_dup[10:20] = bidx_dup[0:10]    # duplicates: last writer wins (q6)
    buf.update_priorities(bidx_dup, td)
    out.update(u=u, batch_index=bidx.astype(np.int32), is_weight=w.numpy(), beta=buf.beta, td=td, batch_index_dup=bidx_dup.astype(np.int32),
               tree_after_update=buf.sum_tree.tree.copy(), s_state=batch["state"].numpy(), s_reward=batch["reward"].numpy(),
               gamma=cfg.gamma, n_steps=cfg.n_steps, alpha=cfg.alpha, capacity=cfg.memory_capacity)
    np.savez(OUT / "per_nstep.npz", **out)


# ------------------------------------------------------------------------------------------------ DQN
def gen_dqn():
    """DQNTrainer.update (algorithms/dqn_cartpole.py:135-168) x2 on a fixed batch: MSE, clamp(+-1), Adam(1e-3)."""
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(4, n_actions=2, max_steps=500))
    m = rl.load("algorithms/dqn_cartpole.py")
    torch.manual_seed(0)
    cfg = m.Config(); cfg.device = "cpu"; cfg.batch_size = 256; cfg.hidden_dim = 64
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        t = m.DQNTrainer(cfg)
    # make the target differ from the online net and the gradients large enough to hit the clamp
    with torch.no_grad():
        for p in t.target_net.parameters():
            p.add_(0.05 * torch.randn_like(p))
        t.policy_net.net[4].weight.mul_(30.0)
    rng = np.random.default_rng(2)
    B = cfg.batch_size
    batch = (rng.standard_normal((B, 4)).astype(np.float32), rng.integers(0, 2, B), (rng.standard_normal(B) * 3).astype(np.float32),
             rng.standard_normal((B, 4)).astype(np.float32), rng.random(B) < 0.1)
    out = dict(states=batch[0], action=batch[1].astype(np.int32), reward=batch[2], next_states=batch[3], done=batch[4].astype(np.uint8),
               gamma=cfg.gamma, lr=cfg.lr)
    out.update(_sd(t.policy_net, "p0_")); out.update(_sd(t.target_net, "t0_"))
    t.memory.sample = lambda bs: batch
    t.memory.buffer = deque([0] * B)
    losses = [t.update(), t.update()]
    out.update(_sd(t.policy_net, "p2_"), losses=np.array(losses))
    np.savez(OUT / "dqn_update.npz", **out)


# ------------------------------------------------------------------------------------------------ Rainbow
def gen_rainbow():
    """RainbowDQNTrainer.update (rainbow :311-361): PER sample, NoisyNet forwards, double-Q n-step target, IS-weighted loss,
    priority write-back, clip 10, Adam, Polyak, LR decay."""
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(4, n_actions=2, max_steps=500))
    m = rl.load("algorithms/rainbow_dqn_cartpole.py")
    torch.manual_seed(1)
    cfg = m.Config(); cfg.device = "cpu"; cfg.memory_capacity = 1024; cfg.batch_size = 128; cfg.hidden_dim = 64
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        t = m.RainbowDQNTrainer(cfg)
    with torch.no_grad():
        for p in t.target_net.parameters():
            p.add_(0.02 * torch.randn_like(p))
    rng = np.random.default_rng(3)
    T = 900
    S = rng.standard_normal((T, 4)).astype(np.float32); S2 = rng.standard_normal((T, 4)).astype(np.float32)
    A = rng.integers(0, 2, T); R = np.ones(T, np.float32)
    done = rng.random(T) < 0.05; terminal = done & (rng.random(T) < 0.8)
    for k in range(T):
        t.memory.store_transition(S[k], int(A[k]), float(R[k]), S2[k], bool(terminal[k]), bool(done[k]))
    # give the tree non-uniform priorities
    pr = rng.random(t.memory.current_size) + 0.05
    for i, p in enumerate(pr):
        t.memory.sum_tree.update(i, float(p))
    t.total_steps = 5000
    out = dict(S=S, S2=S2, A=A.astype(np.int32), R=R, done=done.astype(np.uint8), terminal=terminal.astype(np.uint8), prio=pr,
               total_steps=t.total_steps, max_train_steps=t.max_train_steps, tree0=t.memory.sum_tree.tree.copy(),
               gamma=cfg.gamma, n_steps=cfg.n_steps, tau=cfg.tau, lr=cfg.lr, grad_clip=cfg.grad_clip, alpha=cfg.alpha,
               capacity=cfg.memory_capacity, batch_size=cfg.batch_size)
    out.update(_sd(t.policy_net, "p0_")); out.update(_sd(t.target_net, "t0_"))
    np.random.seed(11)
    out["u"] = np.random.random_sample(cfg.batch_size)
    torch.manual_seed(99)
    H, Aa = cfg.hidden_dim, 2
    for tag in ("next", "cur"):       # the two policy_net forwards of update(): next_state first, then state (q8)
        out[f"xi_{tag}_in_a"] = torch.randn(H).numpy(); out[f"xi_{tag}_out_a"] = torch.randn(Aa).numpy()
        out[f"xi_{tag}_in_v"] = torch.randn(H).numpy(); out[f"xi_{tag}_out_v"] = torch.randn(1).numpy()
    np.random.seed(11)
    torch.manual_seed(99)
    loss = t.update()
    out.update(_sd(t.policy_net, "p1_")); out.update(_sd(t.target_net, "t1_"))
    out.update(loss=loss, tree1=t.memory.sum_tree.tree.copy(), lr_after=t.optimizer.param_groups[0]["lr"], beta=t.memory.beta)
    np.savez(OUT / "rainbow_update.npz", **out)


# ------------------------------------------------------------------------------------------------ SAC / TD3
def _patch_rsample(eps_list):
    """Normal.rsample() == loc + eps * scale with eps ~ N(0,1): feed recorded eps."""
    it = iter(eps_list)
    orig = torch.distributions.Normal.rsample

    def rsample(self, sample_shape=torch.Size()):
        return self.loc + next(it) * self.scale
    torch.distributions.Normal.rsample = rsample
    return orig


def gen_sac():
    """SACTrainer.update (algorithms/sac_pendulum.py:213-267) x2 on fixed batches and fixed N(0,1) draws."""
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(3, act_dim=1, bound=2.0, max_steps=200))
    m = rl.load("algorithms/sac_pendulum.py")
    torch.manual_seed(5)
    cfg = m.Config(); cfg.device = "cpu"; cfg.batch_size = 256; cfg.hidden_dim = 64
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        t = m.SACTrainer(cfg)
    with torch.no_grad():
        for p in t.critic_target.parameters():
            p.add_(0.02 * torch.randn_like(p))
        t.actor.log_std.bias.fill_(-1.0)
        t.actor.log_std.weight.mul_(40.0)      # push some log_std values past the clamp limits
    rng = np.random.default_rng(6)
    B = cfg.batch_size
    out = dict(gamma=cfg.gamma, tau=cfg.tau, lr=cfg.lr_actor, init_alpha=cfg.init_alpha)
    out.update(_sd(t.actor, "a0_")); out.update(_sd(t.critic, "c0_")); out.update(_sd(t.critic_target, "ct0_"))
    batches, eps = [], []
    for k in range(2):
        th = rng.uniform(-np.pi, np.pi, B)
        s = np.stack([np.cos(th), np.sin(th), rng.uniform(-8, 8, B)], 1).astype(np.float32)
        th2 = rng.uniform(-np.pi, np.pi, B)
        s2 = np.stack([np.cos(th2), np.sin(th2), rng.uniform(-8, 8, B)], 1).astype(np.float32)
        a = rng.uniform(-2, 2, (B, 1)).astype(np.float32)
        r = (-rng.random(B) * 16).astype(np.float64)     # env rewards are float64 in the reference's buffer
        d = rng.random(B) < 0.05
        batches.append((s, a, r, s2, d))
        eps += [torch.randn(B, 1), torch.randn(B, 1)]
        out.update({f"b{k}_s": s, f"b{k}_a": a, f"b{k}_r": r.astype(np.float32), f"b{k}_s2": s2, f"b{k}_d": d.astype(np.uint8),
                    f"b{k}_eps_next": eps[-2].numpy(), f"b{k}_eps_new": eps[-1].numpy()})
    bi = iter(batches)
    t.memory.sample = lambda bs: next(bi)
    t.memory.buffer = deque([0] * B)
    orig = _patch_rsample(eps)
    try:
        l0 = t.update(); l1 = t.update()
    finally:
        torch.distributions.Normal.rsample = orig
    out.update(_sd(t.actor, "a2_")); out.update(_sd(t.critic, "c2_")); out.update(_sd(t.critic_target, "ct2_"))
    out.update(losses=np.array([l0, l1]), log_alpha2=t.log_alpha.item())
    np.savez(OUT / "sac_update.npz", **out)


def gen_td3():
    """TD3Trainer.update (algorithms/td3_pendulum.py:172-228) x2 (the second one runs the delayed actor step)."""
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(3, act_dim=1, bound=2.0, max_steps=200))
    m = rl.load("algorithms/td3_pendulum.py")
    torch.manual_seed(8)
    cfg = m.Config(); cfg.device = "cpu"; cfg.batch_size = 256; cfg.hidden_dim = 64
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        t = m.TD3Trainer(cfg)
    with torch.no_grad():
        for p in list(t.critic_target.parameters()) + list(t.actor_target.parameters()):
            p.add_(0.02 * torch.randn_like(p))
    rng = np.random.default_rng(9)
    B = cfg.batch_size
    out = dict(gamma=cfg.gamma, tau=cfg.tau)
    out.update(_sd(t.actor, "a0_")); out.update(_sd(t.actor_target, "at0_")); out.update(_sd(t.critic, "c0_")); out.update(_sd(t.critic_target, "ct0_"))
    batches, noises = [], []
    for k in range(2):
        s = rng.standard_normal((B, 3)).astype(np.float32); s2 = rng.standard_normal((B, 3)).astype(np.float32)
        a = rng.uniform(-2, 2, (B, 1)).astype(np.float32)
        r = (-rng.random(B) * 16).astype(np.float64); d = rng.random(B) < 0.05
        batches.append((s, a, r, s2, d))
        noises.append(torch.randn(B, 1))
        out.update({f"b{k}_s": s, f"b{k}_a": a, f"b{k}_r": r.astype(np.float32), f"b{k}_s2": s2, f"b{k}_d": d.astype(np.uint8),
                    f"b{k}_noise": noises[-1].numpy()})
    bi, ni = iter(batches), iter(noises)
    t.memory.sample = lambda bs: next(bi)
    t.memory.buffer = deque([0] * B)
    orig = torch.randn_like
    torch.randn_like = lambda x, **kw: next(ni)
    try:
        l0 = t.update(); l1 = t.update()
    finally:
        torch.randn_like = orig
    out.update(_sd(t.actor, "a2_")); out.update(_sd(t.actor_target, "at2_")); out.update(_sd(t.critic, "c2_")); out.update(_sd(t.critic_target, "ct2_"))
    out.update(losses=np.array([l0[0], l0[1], l1[0], l1[1]], dtype=np.float64))
    np.savez(OUT / "td3_update.npz", **out)


def gen_ddpg():
    """DDPGTrainer.update (algorithms/ddpg_pendulum.py:154-195) x2."""
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(3, act_dim=1, bound=2.0, max_steps=200))
    m = rl.load("algorithms/ddpg_pendulum.py")
    torch.manual_seed(18)
    cfg = m.Config(); cfg.device = "cpu"; cfg.batch_size = 256; cfg.hidden_dim = 64
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        t = m.DDPGTrainer(cfg)
    with torch.no_grad():
        for p in list(t.critic_target.parameters()) + list(t.actor_target.parameters()):
            p.add_(0.02 * torch.randn_like(p))
    rng = np.random.default_rng(19)
    B = cfg.batch_size
    out = dict(gamma=cfg.gamma, tau=cfg.tau)
    out.update(_sd(t.actor, "a0_")); out.update(_sd(t.actor_target, "at0_")); out.update(_sd(t.critic, "c0_")); out.update(_sd(t.critic_target, "ct0_"))
    batches = []
    for k in range(2):
        s = rng.standard_normal((B, 3)).astype(np.float32); s2 = rng.standard_normal((B, 3)).astype(np.float32)
        a = rng.uniform(-2, 2, (B, 1)).astype(np.float32)
        r = (-rng.random(B) * 16).astype(np.float64); d = rng.random(B) < 0.05
        batches.append((s, a, r, s2, d))
        out.update({f"b{k}_s": s, f"b{k}_a": a, f"b{k}_r": r.astype(np.float32), f"b{k}_s2": s2, f"b{k}_d": d.astype(np.uint8)})
    bi = iter(batches)
    t.memory.sample = lambda bs: next(bi)
    t.memory.buffer = deque([0] * B)
    l0 = t.update(); l1 = t.update()
    out.update(_sd(t.actor, "a2_")); out.update(_sd(t.actor_target, "at2_")); out.update(_sd(t.critic, "c2_")); out.update(_sd(t.critic_target, "ct2_"))
    out.update(losses=np.array([l0[0], l0[1], l1[0], l1[1]], dtype=np.float64))
    np.savez(OUT / "ddpg_update.npz", **out)



# ------------------------------------------------------------------------------------------------ NoisyNet DQN
def gen_noisy_dqn():
    """NoisyDQNTrainer.update (algorithms/noisy_dqn_cartpole.py:206-253) x2 on a fixed batch: four NoisyLinear layers with fresh
    factorised noise per forward (online net on s, then on s'), eval-mode target net, double-Q, MSE, Adam, hard sync on the 2nd.

    NOTE — the unmodified script cannot run its own update(): the second forward (on s', under no_grad, :232) re-draws the noise
    with in-place `copy_` into the epsilon buffers (:86-87) that the first forward's autograd graph saved, and loss.backward()
    (:239) raises "one of the variables needed for gradient computation has been modified by an inplace operation".  The golden
    is therefore generated with NoisyLinear.reset_noise rebinding the two buffers to new tensors (same draws, same order, not in
    place) — the only reading under which the script's update() executes: the gradient uses the first forward's noise."""
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(4, n_actions=2, max_steps=500))
    m = rl.load("algorithms/noisy_dqn_cartpole.py")

    def reset_noise(self):
        eps_i = self._scale_noise(self.in_features)
        eps_j = self._scale_noise(self.out_features)
        self.weight_epsilon = torch.outer(eps_j, eps_i)
        self.bias_epsilon = eps_j.clone()
    m.NoisyLinear.reset_noise = reset_noise
    torch.manual_seed(4)
    cfg = m.Config(); cfg.device = "cpu"; cfg.batch_size = 256; cfg.hidden_dim = 64; cfg.target_update_freq = 2
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        t = m.NoisyDQNTrainer(cfg)
    with torch.no_grad():
        for p in t.target_net.parameters():
            p.add_(0.05 * torch.randn_like(p))
    rng = np.random.default_rng(6)
    B, D, H, A = cfg.batch_size, 4, cfg.hidden_dim, 2
    batch = (rng.standard_normal((B, D)).astype(np.float32), rng.integers(0, A, B), (rng.standard_normal(B) * 2).astype(np.float32),
             rng.standard_normal((B, D)).astype(np.float32), rng.random(B) < 0.1)
    out = dict(states=batch[0], action=batch[1].astype(np.int32), reward=batch[2], next_states=batch[3], done=batch[4].astype(np.uint8),
               gamma=cfg.gamma, lr=cfg.lr)
    out.update(_sd(t.policy_net, "p0_")); out.update(_sd(t.target_net, "t0_"))
    t.memory.sample = lambda bs: batch
    t.memory.buffer = deque([0] * B)
    dims = {"fc1": (D, H), "fc2": (H, H), "value_stream": (H, 1), "advantage_stream": (H, A)}
    torch.manual_seed(123)
    for u in (1, 2):
        for tag in ("cur", "next"):       # update(): policy_net(states) first, then policy_net(next_states) (ref :229-232)
            for name in ("fc1", "fc2", "value_stream", "advantage_stream"):      # forward order (ref :126-133)
                out[f"xi{u}_{tag}_{name}_in"] = torch.randn(dims[name][0]).numpy()
                out[f"xi{u}_{tag}_{name}_out"] = torch.randn(dims[name][1]).numpy()
    torch.manual_seed(123)
    m1 = t.update()
    out.update(_sd(t.policy_net, "p1_"))
    m2 = t.update()
    out.update(_sd(t.policy_net, "p2_")); out.update(_sd(t.target_net, "t2_"))
    out.update(losses=np.array([m1["loss"], m2["loss"]]))
    np.savez(OUT / "noisy_dqn_update.npz", **out)


# ------------------------------------------------------------------------------------------------ DDQN + PER (dialect B)
def gen_ddqn_per(duel: bool):
    """DDQNPER(Duel)Trainer.update x2 (algorithms/ddqn_per_cartpole.py:206-247 / ddqn_per_duel_cartpole.py): stratified SumTree
    sample with beta += 0.001, double-Q target, IS-weighted loss, priorities min(|td| + 1e-4, 1)^0.6, grad clamp(+-1), Adam."""
    import random
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(4, n_actions=2, max_steps=500))
    m = rl.load("algorithms/ddqn_per_duel_cartpole.py" if duel else "algorithms/ddqn_per_cartpole.py")
    torch.manual_seed(7 + duel)
    cfg = m.Config(); cfg.device = "cpu"; cfg.memory_capacity = 1024; cfg.batch_size = 128; cfg.hidden_dim = 64
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        t = (m.DDQNPERDuelTrainer if duel else m.DDQNPERTrainer)(cfg)
    with torch.no_grad():
        for p in t.target_net.parameters():
            p.add_(0.05 * torch.randn_like(p))
    rng = np.random.default_rng(8 + duel)
    T = 700
    S = rng.standard_normal((T, 4)).astype(np.float32); S2 = rng.standard_normal((T, 4)).astype(np.float32)
    A = rng.integers(0, 2, T); R = (rng.standard_normal(T)).astype(np.float32); done = rng.random(T) < 0.08
    for k in range(T):
        t.memory.push((S[k], int(A[k]), float(R[k]), S2[k], bool(done[k])))
    pr = rng.random(T) + 0.05                      # non-uniform priorities
    for i, p in enumerate(pr):
        t.memory.tree.update(i + cfg.memory_capacity - 1, float(p))
    out = dict(S=S, S2=S2, A=A.astype(np.int32), R=R, done=done.astype(np.uint8), prio=pr, tree0=t.memory.tree.tree.copy(),
               gamma=cfg.gamma, lr=cfg.lr, capacity=cfg.memory_capacity, batch_size=cfg.batch_size, beta0=cfg.beta)
    out.update(_sd(t.policy_net, "p0_")); out.update(_sd(t.target_net, "t0_"))
    random.seed(21)
    out["u1"] = np.array([random.random() for _ in range(cfg.batch_size)])      # random.uniform(a, b) = a + (b - a) * random()
    out["u2"] = np.array([random.random() for _ in range(cfg.batch_size)])
    random.seed(21)
    l1 = t.update()
    out.update(_sd(t.policy_net, "p1_")); out["tree1"] = t.memory.tree.tree.copy()
    l2 = t.update()
    out.update(_sd(t.policy_net, "p2_")); out["tree2"] = t.memory.tree.tree.copy()
    out.update(losses=np.array([l1, l2]), beta2=cfg.beta)
    np.savez(OUT / ("ddqn_per_duel_update.npz" if duel else "ddqn_per_update.npz"), **out)


# ------------------------------------------------------------------------------------------------ discrete SAC
def gen_sac_discrete():
    """SACTrainer.update x2 (algorithms/sac_cartpole.py:155-221) on a fixed batch: softmax actor, twin critics with separate
    optimisers, entropy-regularised target, float32 log_alpha with its own Adam, Polyak of both targets."""
    rl.install_gymnasium_stub(make=lambda name, **k: rl.FakeEnv(4, n_actions=2, max_steps=500))
    m = rl.load("algorithms/sac_cartpole.py")
    torch.manual_seed(12)
    cfg = m.Config(); cfg.device = "cpu"; cfg.batch_size = 256; cfg.hidden_dim = 64
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        t = m.SACTrainer(cfg)
    with torch.no_grad():
        for net in (t.critic1_target, t.critic2_target):
            for p in net.parameters():
                p.add_(0.05 * torch.randn_like(p))
        t.actor.fc3.weight.mul_(3.0)               # a policy that is not uniform
        t.log_alpha.fill_(float(np.log(0.2)))
    rng = np.random.default_rng(13)
    B = cfg.batch_size
    batch = (rng.standard_normal((B, 4)).astype(np.float32), rng.integers(0, 2, B), (rng.standard_normal(B) * 2).astype(np.float32),
             rng.standard_normal((B, 4)).astype(np.float32), rng.random(B) < 0.1)
    out = dict(states=batch[0], action=batch[1].astype(np.int32), reward=batch[2], next_states=batch[3], done=batch[4].astype(np.uint8),
               gamma=cfg.gamma, tau=cfg.tau, log_alpha0=float(t.log_alpha.item()))
    nets = dict(a=t.actor, c1=t.critic1, c2=t.critic2, c1t=t.critic1_target, c2t=t.critic2_target)
    for k, net in nets.items():
        out.update(_sd(net, f"{k}0_"))
    t.memory.sample = lambda bs: batch
    t.memory.buffer = deque([0] * B)
    losses = []
    for u in (1, 2):
        losses.append(t.update())
        for k, net in nets.items():
            out.update(_sd(net, f"{k}{u}_"))
        out[f"log_alpha{u}"] = float(t.log_alpha.item())
    out["losses"] = np.array(losses)               # rows: (actor, critic1, critic2, alpha)
    np.savez(OUT / "sac_discrete_update.npz", **out)

def main():
    torch.set_num_threads(1)
    gen_sumtree()
    gen_per_nstep()
    gen_dqn()
    gen_rainbow()
    gen_noisy_dqn()
    gen_ddqn_per(False)
    gen_ddqn_per(True)
    gen_sac_discrete()
    gen_sac()
    gen_td3()
    gen_ddpg()
