"""GPU parity of the off-policy path (replay ring, n-step fold, PER sum-tree, DQN / Rainbow / SAC / TD3 updates)
against golden vectors produced by the reference's own classes (oracle/make_golden_offpolicy.py).

Index work (sampled leaves, ring rows, last-writer-wins) is compared exactly; float64 tree sums to 1e-12
relative (different but equally valid summation orders); network parameters after update() to fp32 tolerances
stated per test.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

f32, f64, i32, u8 = torch.float32, torch.float64, torch.int32, torch.uint8


def cu(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


# ------------------------------------------------------------------------------------------------ sum tree
@pytest.mark.parametrize("cap", [37, 64, 20000])
def test_sumtree_matches_reference(golden, cap):
    from gymrl_b200.algorithms.rainbow_dqn_cartpole import SumTree
    g = golden("sumtree.npz")
    t = SumTree(cap)
    idx, pr = g[f"c{cap}_idx"], g[f"c{cap}_prio"]
    # the reference applies the updates one by one; batches of 13 exercise last-writer-wins inside a batch
    for s in range(0, len(idx), 13):
        t.update(cu(idx[s:s + 13]), cu(pr[s:s + 13], f64))
    tree = t.tree.cpu().numpy()
    assert np.array_equal(tree[cap - 1:], g[f"c{cap}_tree"][cap - 1:])                    # leaves: exact
    np.testing.assert_allclose(tree, g[f"c{cap}_tree"], rtol=1e-12, atol=1e-12)           # sums: order-of-addition only
    assert abs(t.priority_max - float(g[f"c{cap}_max"])) == 0.0
    # descents on the reference's own tree contents -> exact leaf indices incl. the rotated order (q4) and the tie rule
    t.tree.copy_(cu(g[f"c{cap}_tree"]))
    v = cu(g[f"c{cap}_v"], f64)
    prio = torch.zeros(len(v), device="cuda", dtype=f64)
    leaf, _ = t.sample(len(v), t._ring1, t._beta0, uniforms=v, out_prio=prio, raw_values=True)
    assert np.array_equal(leaf.cpu().numpy(), g[f"c{cap}_leaf"])
    assert np.array_equal(prio.cpu().numpy(), g[f"c{cap}_leafp"])
    i0, p0 = t.get_index(float(g[f"c{cap}_v"][3]))
    assert i0 == int(g[f"c{cap}_leaf"][3]) and p0 == float(g[f"c{cap}_leafp"][3])


def test_per_nstep_buffer_matches_reference(golden):
    from gymrl_b200.algorithms import rainbow_dqn_cartpole as R
    g = golden("per_nstep.npz")
    cfg = R.Config()
    cfg.memory_capacity, cfg.batch_size = int(g["capacity"]), 64
    buf = R.PrioritizedNStepBuffer(cfg, 4, num_envs=1)
    for t in range(len(g["A"])):
        buf.store_transition(g["S"][t], int(g["A"][t]), float(g["R"][t]), g["S2"][t], bool(g["terminal"][t]), bool(g["done"][t]))
    ring = buf.ring
    st = ring.state.cpu().numpy()
    assert st[0] == int(g["count"]) and st[1] == int(g["size"])
    assert np.array_equal(ring.obs.cpu().numpy(), g["b_state"].astype(np.float32))
    assert np.array_equal(ring.next_obs.cpu().numpy(), g["b_next_state"].astype(np.float32))
    assert np.array_equal(ring.action.cpu().numpy()[:, 0], g["b_action"][:, 0].astype(np.int32))
    assert np.array_equal(ring.done.cpu().numpy(), g["b_terminal"].astype(np.float32))
    np.testing.assert_allclose(ring.reward.cpu().numpy(), g["b_reward"].astype(np.float32), rtol=0, atol=0)  # fp64 fold -> fp32
    np.testing.assert_allclose(buf.sum_tree.tree.cpu().numpy(), g["tree_after_store"], rtol=1e-12, atol=1e-12)
    # sample with the reference's uniforms
    buf.sum_tree.tree.copy_(cu(g["tree_after_store"]))
    idx, w = buf.sample(1234, 250000, uniforms=cu(g["u"], f64))
    assert abs(buf.beta - float(g["beta"])) < 1e-15
    assert np.array_equal(idx.cpu().numpy(), g["batch_index"])
    np.testing.assert_allclose(w.cpu().numpy(), g["is_weight"], rtol=2e-6, atol=0)
    np.testing.assert_array_equal(ring.obs[idx.long()].cpu().numpy(), g["s_state"])
    # priority write-back with duplicate indices: last writer (in batch order) wins
    buf.update_priorities(cu(g["batch_index_dup"]), cu(g["td"]))
    tree = buf.sum_tree.tree.cpu().numpy()
    cap = cfg.memory_capacity
    np.testing.assert_allclose(tree[cap - 1:], g["tree_after_update"][cap - 1:], rtol=1e-6, atol=0)   # float32 powf vs NumPy pow
    np.testing.assert_allclose(tree[0], g["tree_after_update"][0], rtol=1e-6)


def test_nstep_window_lockstep_equals_per_env_streams():
    """N envs pushed in lockstep == N independent single-env buffers (the reference semantics per env)."""
    from gymrl_b200.algorithms import rainbow_dqn_cartpole as R
    rng = np.random.default_rng(0)
    N, T, D = 5, 40, 4
    cfg = R.Config(); cfg.memory_capacity = 512
    vec = R.PrioritizedNStepBuffer(cfg, D, num_envs=N)
    singles = [R.PrioritizedNStepBuffer(cfg, D, num_envs=1) for _ in range(N)]
    for t in range(T):
        s, s2 = rng.standard_normal((N, D)).astype(np.float32), rng.standard_normal((N, D)).astype(np.float32)
        a, r = rng.integers(0, 2, N).astype(np.int32), rng.standard_normal(N).astype(np.float32)
        d = (rng.random(N) < 0.2); te = d & (rng.random(N) < 0.7)
        vec.store_lockstep(cu(s), cu(a), cu(r), cu(s2), cu(te.astype(np.uint8)), cu(d.astype(np.uint8)))
        for n in range(N):
            singles[n].store_transition(s[n], int(a[n]), float(r[n]), s2[n], bool(te[n]), bool(d[n]))
    rows = T - cfg.n_steps + 1
    for n in range(N):
        assert torch.equal(vec.ring.reward[n:rows * N:N], singles[n].ring.reward[:rows])
        assert torch.equal(vec.ring.next_obs[n:rows * N:N], singles[n].ring.next_obs[:rows])
        assert torch.equal(vec.ring.done[n:rows * N:N], singles[n].ring.done[:rows])


def test_uniform_replay_sample_without_replacement():
    from gymrl_b200 import ops_offpolicy as off
    ring = off.ReplayRing(1000, 3, 1, False, torch.device("cuda"))
    for k in range(7):
        n = 128
        ring.store(torch.full((n, 3), float(k), device="cuda"), torch.zeros(n, 1, device="cuda"), torch.arange(n, device="cuda", dtype=f32),
                   torch.zeros(n, 3, device="cuda"), torch.zeros(n, device="cuda", dtype=u8))
    assert len(ring) == 896 and ring.state.tolist() == [896, 896]
    idx = ring.sample_indices(896, seed=1, draw=3)
    assert torch.equal(torch.sort(idx.long()).values, torch.arange(896, device="cuda"))           # a full draw is a permutation
    idx = ring.sample_indices(256, seed=1, draw=4)
    assert idx.unique().numel() == 256 and int(idx.max()) < 896                                    # no repeats (random.sample)
    ring.store(torch.ones(128, 3, device="cuda"), torch.zeros(128, 1, device="cuda"), torch.zeros(128, device="cuda"),
               torch.zeros(128, 3, device="cuda"), torch.zeros(128, device="cuda", dtype=u8))
    assert ring.state.tolist() == [24, 1000]                                                       # wrapped like deque(maxlen)


# ------------------------------------------------------------------------------------------------ helpers
def _load(module, g, prefix):
    sd = {k[len(prefix):]: torch.as_tensor(g[k]) for k in g.files if k.startswith(prefix)}
    module.load_state_dict(sd)


def _cmp(module, g, prefix, rtol, atol):
    for k, v in module.state_dict().items():
        if prefix + k in g.files and "epsilon" not in k:
            np.testing.assert_allclose(v.cpu().numpy(), g[prefix + k], rtol=rtol, atol=atol, err_msg=prefix + k)


def _fill_ring(ring, s, a, r, s2, d):
    n = len(r)
    ring.store(cu(s), cu(a).reshape(n, -1), cu(r), cu(s2), cu(d.astype(np.uint8)))


# ------------------------------------------------------------------------------------------------ DQN
def test_dqn_update_matches_reference(golden):
    from gymrl_b200.algorithms import dqn_cartpole as D
    g = golden("dqn_update.npz")
    cfg = D.Config(); cfg.batch_size, cfg.hidden_dim, cfg.seed, cfg.memory_capacity = 256, 64, 0, 1024
    t = D.DQNTrainer(cfg)
    _load(t.policy_net, g, "p0_"); t.fp.refresh_views()
    _load(t.target_net, g, "t0_"); t.fp_t.refresh_views()
    _fill_ring(t.memory, g["states"], g["action"], g["reward"], g["next_states"], g["done"])
    idx = torch.arange(256, device="cuda", dtype=i32)
    losses = [float(t.update(idx)), float(t.update(idx))]
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-5)
    _cmp(t.policy_net, g, "p2_", rtol=2e-4, atol=2e-6)


# ------------------------------------------------------------------------------------------------ Rainbow
def test_rainbow_update_matches_reference(golden):
    from gymrl_b200.algorithms import rainbow_dqn_cartpole as R
    g = golden("rainbow_update.npz")
    cfg = R.Config()
    cfg.memory_capacity, cfg.batch_size, cfg.hidden_dim, cfg.seed = int(g["capacity"]), int(g["batch_size"]), 64, 0
    t = R.RainbowDQNTrainer(cfg)
    _load(t.policy_net, g, "p0_"); t.fp.refresh_views()
    _load(t.target_net, g, "t0_"); t.fp_t.refresh_views()
    for k in range(len(g["A"])):
        t.memory.store_transition(g["S"][k], int(g["A"][k]), float(g["R"][k]), g["S2"][k], bool(g["terminal"][k]), bool(g["done"][k]))
    t.memory.sum_tree.tree.copy_(cu(g["tree0"]))
    t.total_steps = int(g["total_steps"])
    assert t.max_train_steps == int(g["max_train_steps"])
    xi = lambda tag: {k: cu(g[f"xi_{tag}_{k}"]) for k in ("in_a", "out_a", "in_v", "out_v")}
    loss = float(t.update(uniforms=cu(g["u"], f64), xi_next=xi("next"), xi_cur=xi("cur")))
    np.testing.assert_allclose(loss, float(g["loss"]), rtol=2e-5)
    assert abs(t.memory.beta - float(g["beta"])) < 1e-15
    assert abs(t.optimizer.param_groups[0]["lr"] - float(g["lr_after"])) < 1e-15
    cap = cfg.memory_capacity
    np.testing.assert_allclose(t.memory.sum_tree.tree.cpu().numpy()[cap - 1:], g["tree1"][cap - 1:], rtol=2e-4, atol=1e-7)
    _cmp(t.policy_net, g, "p1_", rtol=2e-4, atol=2e-6)
    _cmp(t.target_net, g, "t1_", rtol=2e-5, atol=2e-7)



# ------------------------------------------------------------------------------------------------ NoisyNet DQN
def test_noisy_dqn_update_matches_reference(golden):
    """Two updates of algorithms/noisy_dqn_cartpole.py (fresh factorised noise on all four layers per forward, double-Q with an
    eval-mode target, MSE, Adam, hard target sync on the second) with the reference's own noise draws fed in."""
    from gymrl_b200.algorithms import noisy_dqn_cartpole as Nd
    g = golden("noisy_dqn_update.npz")
    cfg = Nd.Config(); cfg.batch_size, cfg.hidden_dim, cfg.seed, cfg.memory_capacity, cfg.target_update_freq = 256, 64, 0, 1024, 2
    t = Nd.NoisyDQNTrainer(cfg)
    _load(t.policy_net, g, "p0_"); t.fp.refresh_views()
    _load(t.target_net, g, "t0_"); t.fp_t.refresh_views()
    _fill_ring(t.memory, g["states"], g["action"], g["reward"], g["next_states"], g["done"])
    idx = torch.arange(256, device="cuda", dtype=i32)
    xi = lambda u, tag: {n: (cu(g[f"xi{u}_{tag}_{n}_in"]), cu(g[f"xi{u}_{tag}_{n}_out"])) for n in Nd.NoisyDuelingQNetwork.LAYERS}
    losses = []
    for u in (1, 2):
        losses.append(float(t.update(idx, xi_cur=xi(u, "cur"), xi_next=xi(u, "next"))["loss"]))
        if u == 1:
            _cmp(t.policy_net, g, "p1_", rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(losses, g["losses"], rtol=3e-5)
    _cmp(t.policy_net, g, "p2_", rtol=3e-4, atol=3e-6)
    _cmp(t.target_net, g, "t2_", rtol=3e-4, atol=3e-6)           # hard sync happened on the second update
    torch.testing.assert_close(t.fp_t.flat, t.fp.flat, rtol=0, atol=0)


# ------------------------------------------------------------------------------------------------ DDQN + PER (dialect B)
@pytest.mark.parametrize("duel", [False, True])
def test_ddqn_per_update_matches_reference(golden, duel):
    """Two updates of algorithms/ddqn_per(_duel)_cartpole.py with the reference's random.uniform draws fed in: sampled leaves /
    IS weights (beta += 0.001 per sample), double-Q target, IS-weighted loss, priorities min(|td| + 1e-4, 1)^0.6 written to the
    tree, grad clamp, Adam."""
    import importlib
    M = importlib.import_module("gymrl_b200.algorithms." + ("ddqn_per_duel_cartpole" if duel else "ddqn_per_cartpole"))
    g = golden("ddqn_per_duel_update.npz" if duel else "ddqn_per_update.npz")
    cfg = M.Config()
    cfg.memory_capacity, cfg.batch_size, cfg.hidden_dim, cfg.seed = int(g["capacity"]), int(g["batch_size"]), 64, 0
    t = (M.DDQNPERDuelTrainer if duel else M.DDQNPERTrainer)(cfg)
    _load(t.policy_net, g, "p0_"); t.fp.refresh_views()
    _load(t.target_net, g, "t0_"); t.fp_t.refresh_views()
    n = len(g["A"])
    t.memory.store(cu(g["S"]), cu(g["A"]).reshape(n, 1), cu(g["R"]), cu(g["S2"]), cu(g["done"]))
    assert len(t.memory) == n
    t.memory.tree.tree.copy_(cu(g["tree0"], f64))
    cap = cfg.memory_capacity
    losses = []
    for u, tree_key, p_key in (("u1", "tree1", "p1_"), ("u2", "tree2", "p2_")):
        losses.append(float(t.update(uniforms=cu(g[u], f64))))
        np.testing.assert_allclose(t.memory.tree.tree.cpu().numpy()[cap - 1:], g[tree_key][cap - 1:], rtol=3e-4, atol=1e-7, err_msg=tree_key)
        _cmp(t.policy_net, g, p_key, rtol=3e-4, atol=3e-6)
    np.testing.assert_allclose(losses, g["losses"], rtol=5e-5)
    assert abs(cfg.beta - float(g["beta2"])) < 1e-15
    # facade: leaves are addressed by tree index in this dialect (ref :75-107)
    leaf, prio = t.memory.tree.get_leaf(0.5 * t.memory.tree.total_priority())
    assert cap - 1 <= leaf < 2 * cap - 1 and prio > 0


# ------------------------------------------------------------------------------------------------ discrete SAC
def test_sac_discrete_update_matches_reference(golden):
    """Two updates of algorithms/sac_cartpole.py: every network, log_alpha and the four losses after each update."""
    from gymrl_b200.algorithms import sac_cartpole as Sd
    g = golden("sac_discrete_update.npz")
    cfg = Sd.Config(); cfg.batch_size, cfg.hidden_dim, cfg.seed, cfg.memory_capacity = 256, 64, 0, 1024
    t = Sd.SACTrainer(cfg)
    nets = dict(a=(t.actor, t.fp_a), c1=(t.critic1, t.fp_c1), c2=(t.critic2, t.fp_c2), c1t=(t.critic1_target, t.fp_c1t),
                c2t=(t.critic2_target, t.fp_c2t))
    for k, (net, fp) in nets.items():
        _load(net, g, f"{k}0_"); fp.refresh_views()
    t.log_alpha.fill_(float(g["log_alpha0"]))
    _fill_ring(t.memory, g["states"], g["action"], g["reward"], g["next_states"], g["done"])
    idx = torch.arange(256, device="cuda", dtype=i32)
    for u in (1, 2):
        t.update(idx)
        np.testing.assert_allclose(t.losses(), g["losses"][u - 1], rtol=2e-4, atol=2e-6)
        for k, (net, _) in nets.items():
            _cmp(net, g, f"{k}{u}_", rtol=3e-4, atol=3e-6)
        np.testing.assert_allclose(float(t.log_alpha.item()), float(g[f"log_alpha{u}"]), rtol=1e-5)


@pytest.mark.parametrize("B,A_", [(1, 2), (5, 4), (1777, 2), (4096, 3)])
def test_sac_discrete_kernels_vs_oracle(B, A_):
    """gymrl_sac_discrete_target / critic_loss / actor_grad against oracle.algos_np at ragged sizes, with the row gather and
    padded leading dimensions the trainers use."""
    from oracle import algos_np as Anp
    from gymrl_b200 import ops_offpolicy as off
    rng = np.random.default_rng(B + A_)
    Btot = B + 300
    idx = rng.permutation(Btot)[:B].astype(np.int32)
    z, zn = (rng.standard_normal((B, A_)) * 2).astype(np.float32), (rng.standard_normal((B, A_)) * 2).astype(np.float32)
    q1, q2, q1t, q2t = (rng.standard_normal((B, A_)).astype(np.float32) for _ in range(4))
    r, d = rng.standard_normal(Btot).astype(np.float32), (rng.random(Btot) < 0.2).astype(np.float32)
    act = rng.integers(0, A_, Btot).astype(np.int32)
    la = torch.tensor([np.log(0.3)], device="cuda", dtype=torch.float32)
    pad = lambda x: torch.nn.functional.pad(cu(x), (0, 8 - A_))[:, :A_]          # row stride 8: non-trivial leading dimension
    y = off.sac_discrete_target(pad(zn), pad(q1t), pad(q2t), cu(r), cu(d), la, 0.9, row_index=cu(idx))
    y_ref = Anp.sac_discrete_target(zn, q1t, q2t, r[idx], d[idx], float(la.item()), 0.9)
    np.testing.assert_allclose(y.cpu().numpy(), y_ref, rtol=2e-5, atol=2e-6)
    dq1, dq2 = torch.full((B, 8), float("nan"), device="cuda")[:, :A_], torch.full((B, 8), float("nan"), device="cuda")[:, :A_]
    acc = torch.zeros(2, device="cuda")
    off.sac_discrete_critic_loss(pad(q1), pad(q2), cu(act), y, dq1, dq2, row_index=cu(idx), loss_acc=acc)
    c = Anp.sac_discrete_critic(q1, q2, act[idx], y.cpu().numpy())
    np.testing.assert_allclose(acc.cpu().numpy(), [c["loss1"], c["loss2"]], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(dq1.cpu().numpy(), c["dq1"], rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(dq2.cpu().numpy(), c["dq2"], rtol=1e-4, atol=1e-8)
    dz = torch.full((B, 8), float("nan"), device="cuda")[:, :A_]
    acc2 = torch.zeros(2, device="cuda")
    off.sac_discrete_actor_grad(pad(z), pad(q1), pad(q2), la, dz, acc2)
    a = Anp.sac_discrete_actor(z, q1, q2, float(la.item()))
    np.testing.assert_allclose(dz.cpu().numpy(), a["dlogits"], rtol=2e-4, atol=2e-7 / B)
    np.testing.assert_allclose(acc2.cpu().numpy(), [a["loss"], a["sum_entropy"]], rtol=1e-4, atol=1e-5)

# ------------------------------------------------------------------------------------------------ SAC
def test_sac_update_matches_reference(golden):
    from gymrl_b200.algorithms import sac_pendulum as S
    g = golden("sac_update.npz")
    cfg = S.Config(); cfg.batch_size, cfg.hidden_dim, cfg.seed, cfg.memory_capacity = 256, 64, 0, 1024
    t = S.SACTrainer(cfg)
    _load(t.actor, g, "a0_"); t.fp_a.refresh_views()
    _load(t.critic, g, "c0_"); t.fp_c.refresh_views()
    _load(t.critic_target, g, "ct0_"); t.fp_ct.refresh_views()
    losses = []
    for k in range(2):
        _fill_ring(t.memory, g[f"b{k}_s"], g[f"b{k}_a"], g[f"b{k}_r"], g[f"b{k}_s2"], g[f"b{k}_d"])
        idx = torch.arange(256 * k, 256 * (k + 1), device="cuda", dtype=i32)
        t.update(idx, noise_next=cu(g[f"b{k}_eps_next"]), noise_new=cu(g[f"b{k}_eps_new"]))
        losses.append(t.losses())
    np.testing.assert_allclose(np.array(losses), g["losses"], rtol=5e-4, atol=5e-5)
    np.testing.assert_allclose(t.log_alpha.item(), float(g["log_alpha2"]), rtol=1e-9)
    _cmp(t.critic, g, "c2_", rtol=3e-4, atol=3e-6)
    _cmp(t.critic_target, g, "ct2_", rtol=3e-5, atol=3e-7)
    _cmp(t.actor, g, "a2_", rtol=3e-4, atol=3e-6)


# ------------------------------------------------------------------------------------------------ TD3
def test_td3_update_matches_reference(golden):
    from gymrl_b200.algorithms import td3_pendulum as T
    g = golden("td3_update.npz")
    cfg = T.Config(); cfg.batch_size, cfg.hidden_dim, cfg.seed, cfg.memory_capacity = 256, 64, 0, 1024
    t = T.TD3Trainer(cfg)
    for mod, fp, pre in ((t.actor, t.fp_a, "a0_"), (t.actor_target, t.fp_at, "at0_"), (t.critic, t.fp_c, "c0_"), (t.critic_target, t.fp_ct, "ct0_")):
        _load(mod, g, pre); fp.refresh_views()
    out = []
    for k in range(2):
        _fill_ring(t.memory, g[f"b{k}_s"], g[f"b{k}_a"], g[f"b{k}_r"], g[f"b{k}_s2"], g[f"b{k}_d"])
        idx = torch.arange(256 * k, 256 * (k + 1), device="cuda", dtype=i32)
        al, cl = t.update(idx, noise=cu(g[f"b{k}_noise"]))
        out += [float(al[0]) if k == 1 else 0.0, float(cl[0])]
    np.testing.assert_allclose(out, g["losses"], rtol=5e-4, atol=5e-5)
    _cmp(t.critic, g, "c2_", rtol=3e-4, atol=3e-6)
    _cmp(t.actor, g, "a2_", rtol=3e-4, atol=3e-6)
    _cmp(t.critic_target, g, "ct2_", rtol=3e-5, atol=3e-7)
    _cmp(t.actor_target, g, "at2_", rtol=3e-5, atol=3e-7)


# ------------------------------------------------------------------------------------------------ DDPG
def test_ddpg_update_matches_reference(golden):
    """Two DDPGTrainer.update() calls of the reference (algorithms/ddpg_pendulum.py:154-195) on the same batches."""
    from gymrl_b200.algorithms import ddpg_pendulum as Dm
    g = golden("ddpg_update.npz")
    cfg = Dm.Config(); cfg.batch_size, cfg.hidden_dim, cfg.seed, cfg.memory_capacity = 256, 64, 0, 1024
    t = Dm.DDPGTrainer(cfg)
    for mod, fp, pre in ((t.actor, t.fp_a, "a0_"), (t.actor_target, t.fp_at, "at0_"), (t.critic, t.fp_c, "c0_"), (t.critic_target, t.fp_ct, "ct0_")):
        _load(mod, g, pre); fp.refresh_views()
    out = []
    for k in range(2):
        _fill_ring(t.memory, g[f"b{k}_s"], g[f"b{k}_a"], g[f"b{k}_r"], g[f"b{k}_s2"], g[f"b{k}_d"])
        idx = torch.arange(256 * k, 256 * (k + 1), device="cuda", dtype=i32)
        al, cl = t.update(idx)
        out += [float(al[0]), 0.5 * float(cl[0])]
    np.testing.assert_allclose(out, g["losses"], rtol=5e-4, atol=5e-5)
    _cmp(t.critic, g, "c2_", rtol=3e-4, atol=3e-6)
    _cmp(t.actor, g, "a2_", rtol=3e-4, atol=3e-6)
    _cmp(t.critic_target, g, "ct2_", rtol=3e-5, atol=3e-7)
    _cmp(t.actor_target, g, "at2_", rtol=3e-5, atol=3e-7)


# ------------------------------------------------------------------------------------------------ smoke at BASELINE sizes
@pytest.mark.parametrize("algo", ["dqn", "rainbow", "sac", "td3", "ddpg", "noisy_dqn", "ddqn_per", "ddqn_per_duel", "sac_discrete"])
def test_offpolicy_trainers_run_vectorised(algo):
    import importlib
    name = {"dqn": "dqn_cartpole", "rainbow": "rainbow_dqn_cartpole", "sac": "sac_pendulum", "td3": "td3_pendulum", "ddpg": "ddpg_pendulum",
            "noisy_dqn": "noisy_dqn_cartpole", "ddqn_per": "ddqn_per_cartpole", "ddqn_per_duel": "ddqn_per_duel_cartpole",
            "sac_discrete": "sac_cartpole"}[algo]
    M = importlib.import_module(f"gymrl_b200.algorithms.{name}")
    cfg = M.Config()
    cfg.num_envs, cfg.seed, cfg.max_locksteps = 1024, 3, 30
    cfg.batch_size, cfg.memory_capacity = 1024, 1 << 16
    cls = [getattr(M, k) for k in dir(M) if k.endswith("Trainer")][0]
    t = cls(cfg)
    t.train()
    flat = getattr(t, "fp", None) or getattr(t, "fp_a")
    assert torch.isfinite(flat.flat).all()
    r = t.eval(4)
    assert len(r) == 4 and all(np.isfinite(r))


@pytest.mark.parametrize("algo", ["rainbow", "sac", "td3", "ddpg", "sac_discrete", "dqn", "noisy_dqn", "ddqn_per", "ddqn_per_duel"])
def test_graph_lockstep_equals_eager_lockstep(algo):
    """train()'s captured lockstep (act -> env step -> store -> update as one CUDA graph, RNG draw counters / PER beta /
    learning rate in device scalars) leaves the same parameters, replay contents and env stream as the eager lockstep."""
    import importlib
    name = {"rainbow": "rainbow_dqn_cartpole", "sac": "sac_pendulum", "td3": "td3_pendulum", "ddpg": "ddpg_pendulum",
            "sac_discrete": "sac_cartpole", "dqn": "dqn_cartpole", "noisy_dqn": "noisy_dqn_cartpole", "ddqn_per": "ddqn_per_cartpole",
            "ddqn_per_duel": "ddqn_per_duel_cartpole"}[algo]
    M = importlib.import_module(f"gymrl_b200.algorithms.{name}")
    cls = [getattr(M, k) for k in dir(M) if k.endswith("Trainer") and getattr(M, k).__module__ == M.__name__][0]
    res = []
    for use_graph in (False, True):
        cfg = M.Config()
        cfg.num_envs, cfg.seed, cfg.batch_size, cfg.memory_capacity, cfg.use_cuda_graph = 256, 5, 256, 1 << 14, use_graph
        torch.manual_seed(0)
        t = cls(cfg)
        t.env.reset(out=t.cur)
        for _ in range(40):
            t.lockstep()
        torch.cuda.synchronize()
        assert bool(getattr(t, "_g_lockstep", None)) == use_graph
        if use_graph:
            assert t.graph_launches > 0
        if algo == "td3" and use_graph:
            assert sorted(t._g_lockstep) == [0, 1]          # one graph per phase: critic + actor / critic only
        fps = [getattr(t, k) for k in ("fp", "fp_t", "fp_a", "fp_at", "fp_c", "fp_ct", "fp_c1", "fp_c2", "fp_c1t", "fp_c2t") if hasattr(t, k)]
        ring = t.memory.ring if hasattr(t.memory, "ring") else t.memory
        res.append(([f.flat.clone() for f in fps], t.cur.clone(), ring.obs.clone(), ring.reward.clone(), ring.state.clone(), len(t.memory)))
    (pa, ca, oa, ra, sa, la), (pb, cb, ob, rb, sb, lb) = res
    assert la == lb and torch.equal(sa, sb)
    assert torch.equal(ca, cb) and torch.equal(oa, ob) and torch.equal(ra, rb)
    for a, b in zip(pa, pb):
        torch.testing.assert_close(a, b, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("cap", [37, 4096, 20000, 1 << 21])
def test_sumtree_invariant_and_determinism_full_size(cap):
    """Properties of the batched tree writes that hold at any size (C3's capacity 2^21 included): after every batch each internal
    node is EXACTLY fl(left + right) (so the root is the tree-order pairwise sum of the leaves), duplicates resolve last-writer-
    wins in batch order, the scratch is left clean, and the same batches give bitwise the same tree."""
    from gymrl_b200 import ops_offpolicy as off
    g = torch.Generator().manual_seed(cap)
    n = min(8192, 4 * cap)

    def run():
        t = off.DeviceSumTree(cap, torch.device("cuda"))
        ring = torch.tensor([0, 0], device="cuda", dtype=i32)
        leaves = np.zeros(cap)
        gg = torch.Generator().manual_seed(cap + 1)
        for b in range(4):
            idx = torch.randint(0, cap, (n,), generator=gg, dtype=torch.int32)
            idx[1::7] = idx[0]                                     # heavy duplication of one leaf
            pr = torch.rand(n, generator=gg, dtype=torch.float64) + 0.01
            t.update(idx.cuda(), pr.cuda())
            for i, p in zip(idx.tolist(), pr.tolist()):           # the reference's loop: last writer wins
                leaves[i] = p
        m = min(cap, 8192)
        ring[0] = cap - m // 2                                     # a wrapped store range, priority = max(leaves)
        ring[1] = cap
        t.store_new(m, ring)
        pos = (np.arange(m) + cap - m // 2) % cap
        leaves[pos] = leaves.max()
        return t, leaves

    t, leaves = run()
    tree = t.tree.cpu().numpy()
    assert np.array_equal(tree[cap - 1:], leaves)
    internal = np.arange(cap - 1)
    assert np.array_equal(tree[internal], tree[2 * internal + 1] + tree[2 * internal + 2])       # exact, every node
    np.testing.assert_allclose(tree[0], leaves.sum(), rtol=1e-12)
    w = t.winner.cpu().numpy()
    assert (w[:cap] == -1).all() and (w[cap:] == 0).all()                                        # scratch restored
    assert float(t.max_scratch.item()) == 0.0 and t.u32_scratch.tolist() == [0, 0]
    t2, _ = run()
    assert torch.equal(t.tree, t2.tree)                                                          # bitwise reproducible


def test_replay_store_all_wraps_like_a_deque():
    """gymrl_replay_store_all: five fields + the ring advance in one launch, wrapping at the capacity (deque(maxlen) semantics)."""
    from gymrl_b200 import ops_offpolicy as off
    cap, D, A = 1000, 3, 2
    ring = off.ReplayRing(cap, D, A, False, torch.device("cuda"))
    ref = {k: np.zeros(s, np.float32) for k, s in (("obs", (cap, D)), ("nobs", (cap, D)), ("act", (cap, A)), ("rew", (cap,)), ("done", (cap,)))}
    rng = np.random.default_rng(0)
    cursor = size = 0
    for n in (400, 400, 400, 7, 1000):
        o, o2 = rng.standard_normal((n, D)).astype(np.float32), rng.standard_normal((n, D)).astype(np.float32)
        a, r = rng.standard_normal((n, A)).astype(np.float32), rng.standard_normal(n).astype(np.float32)
        d = (rng.random(n) < 0.3).astype(np.uint8)
        ring.store(cu(o), cu(a), cu(r), cu(o2), cu(d))
        pos = (cursor + np.arange(n)) % cap
        ref["obs"][pos], ref["nobs"][pos], ref["act"][pos], ref["rew"][pos], ref["done"][pos] = o, o2, a, r, d.astype(np.float32)
        cursor, size = (cursor + n) % cap, min(cap, size + n)
        assert ring.state.tolist() == [cursor, size] and len(ring) == size
    assert np.array_equal(ring.obs.cpu().numpy(), ref["obs"]) and np.array_equal(ring.next_obs.cpu().numpy(), ref["nobs"])
    assert np.array_equal(ring.action.cpu().numpy(), ref["act"]) and np.array_equal(ring.reward.cpu().numpy(), ref["rew"])
    assert np.array_equal(ring.done.cpu().numpy(), ref["done"]) and int(ring._done_ctr.item()) == 0
