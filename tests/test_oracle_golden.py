"""Pin the CPU oracle (oracle/algos_np.py, oracle/philox.py) against the golden vectors produced by the
reference's own classes (oracle/make_golden.py).  No GPU needed."""
import numpy as np
import pytest

from oracle import algos_np as A
from oracle import philox as px


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def kat(ctr, key):
        out = px.philox((key[1] << 32) | key[0], np.uint64((ctr[3] << 32) | ctr[0]), ctr[1], ctr[2])
        return [int(x) for x in out]
    assert kat([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_gae_algorithms_dialect(golden):
    g = golden("gae_algorithms.npz")
    adv, ret = A.gae_algorithms(g["a_reward"][:, None], g["a_value"][:, None], g["a_next_value"][None], g["a_done"][:, None],
                                float(g["a_gamma"]), float(g["a_lam"]))
    np.testing.assert_allclose(adv[:, 0], g["a_adv"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ret[:, 0], g["a_ret"], rtol=1e-12, atol=1e-12)
    adv, ret = A.gae_algorithms(g["b_reward"], g["b_value"], g["b_next_value"], g["b_done"], float(g["a_gamma"]), float(g["a_lam"]))
    np.testing.assert_allclose(adv, g["b_adv"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ret, g["b_ret"], rtol=1e-12, atol=1e-12)
    adv, ret = A.gae_algorithms(g["c_reward"][:, None], g["c_value"][:, None], g["c_next_value"][None], g["c_done"][:, None],
                                float(g["c_gamma"]), float(g["c_lam_actor"]), float(g["c_lam_critic"]), coef_f32=False, boot_f32=True)
    np.testing.assert_allclose(adv[:, 0], g["c_adv"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ret[:, 0], g["c_ret"], rtol=1e-12, atol=1e-12)


def test_gae_utils_dialect_bit_exact(golden):
    g = golden("gae_utils.npz")
    adv, vt = A.gae_utils(g["reward"], g["value"], g["next_value"], g["done"], g["dw"], float(g["gamma"]), float(g["lamda"]))
    assert np.array_equal(vt, g["v_target"])  # float32 recurrence reproduced bit for bit
    mean, std = adv.mean(dtype=np.float64), adv.std(ddof=1, dtype=np.float64)  # torch .std() => ddof 1
    np.testing.assert_allclose((adv - mean) / (std + 1e-8), g["adv_normalized"], rtol=2e-5, atol=2e-6)


def test_categorical(golden):
    g = golden("categorical.npz")
    ln, p, action, ent = A.categorical(g["logits"], g["noise"])
    safe = g["margin"] > 1e-5  # near ties are flagged, not counted
    assert safe.mean() > 0.97
    assert np.array_equal(action[safe], g["action"][safe])
    np.testing.assert_allclose(ln[np.arange(len(action)), g["action"]], g["log_prob"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ent, g["entropy"], rtol=1e-5, atol=1e-6)


def test_ppo_loss_dualclip(golden):
    g = golden("ppo_loss_dualclip.npz")
    r = A.ppo_loss_grad(g["logits"], g["value"], g["action"], g["logp_old"], g["adv"], g["ret"], mode="dualclip",
                        clip_eps_min=float(g["clip_eps"]), clip_eps_max=float(g["clip_eps"]), dual_clip=float(g["dual_clip"]),
                        value_coef=float(g["value_coef"]), entropy_coef=float(g["entropy_coef"]))
    np.testing.assert_allclose(r["dlogits"], g["dlogits"], rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(r["dvalue"], g["dvalue"], rtol=2e-4, atol=2e-7)
    for k in ("policy_loss", "value_loss", "entropy", "clip_frac", "approx_kl"):
        np.testing.assert_allclose(r[k], g["m_" + k], rtol=1e-4, atol=1e-6)


def test_ppo_loss_full(golden):
    g = golden("ppo_loss_full.npz")
    r = A.ppo_loss_grad(g["logits"], g["value"], g["action"], g["logp_old"], g["adv"], g["ret"], mode="full",
                        clip_eps_min=float(g["clip_eps_min"]), clip_eps_max=float(g["clip_eps_max"]),
                        dual_clip=float(g["dual_clip"]), value_coef=0.5, entropy_coef=float(g["entropy_coef"]),
                        entropy_old=g["entropy_old"], erc_low=float(g["erc_low"]), erc_high=float(g["erc_high"]))
    assert 0.02 < r["erc_frac"] < 0.9  # the fixture exercises both sides of the ERC mask
    np.testing.assert_allclose(r["dlogits"], g["dlogits"], rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(r["dvalue"], g["dvalue"], rtol=2e-4, atol=2e-7)


def test_ppo_update_gradients_via_oracle_forward(golden):
    """The e2e fixture is self-consistent: oracle forward + oracle loss reproduce the reference's metrics."""
    g = golden("ppo_update.npz")
    sd = {k[3:]: g[k] for k in g.files if k.startswith("w0_")}
    logits, value = A.mlp_actor_critic_forward(sd, g["states"])
    adv, ret = A.gae_algorithms(g["reward"][:, None], g["value_old"][:, None], g["next_value"][None], g["done"][:, None],
                                float(g["gamma"]), float(g["lam"]))
    adv = adv[:, 0]; ret = ret[:, 0]
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    r = A.ppo_loss_grad(logits, value, g["action"], g["logp_old"], adv, ret)
    for k in ("policy_loss", "value_loss", "entropy", "clip_frac", "approx_kl"):
        np.testing.assert_allclose(r[k], g["m1_" + k], rtol=2e-4, atol=2e-6)


def test_mhc_oracle_matches_reference_module(golden):
    """oracle/mhc_np.py (float64) vs the reference ActorCritic's own forward and torch-autograd gradients
    (algorithms/ppo_full_lunarlander.py:76-412, fixture from oracle/make_golden_mhc.py)."""
    from oracle.mhc_np import ActorCriticMHC
    g = golden("mhc_actor_critic.npz")
    sd = {k[2:]: g[k] for k in g.files if k.startswith("p:")}
    assert sum(v.size for v in sd.values()) == 144433        # ppo_full parameter count
    net = ActorCriticMHC(sd, int(g["rate"]), int(g["layers"]), int(g["sk_it"]))
    logits, value = net.forward(g["x"])
    np.testing.assert_allclose(logits, g["logits"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(value, g["value"], rtol=0, atol=2e-6)
    grads = net.backward(g["Gl"], g["Gv"])
    for k in sd:
        ref = g["g:" + k]
        assert np.abs(grads[k].reshape(ref.shape) - ref).max() <= 2e-5 * max(np.abs(ref).max(), 1e-6), k


def test_normalization_oracle_matches_reference(golden):
    """oracle restatement of utils/normalization.py vs the reference classes' own outputs (oracle/make_golden_utils.py)."""
    g = golden("normalization.npz")
    y, rm = A.normalize_stream(g["x"])
    np.testing.assert_allclose(y, g["y"].astype(np.float32), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rm.mean, g["mean"], rtol=0, atol=0)
    np.testing.assert_allclose(rm.std, g["std"], rtol=1e-15)
    rs = A.reward_scaling_stream(g["r"], g["reset_at"], float(g["gamma"]))
    np.testing.assert_allclose(rs, g["r_scaled"].astype(np.float32), rtol=2e-6)


def test_sqrt_threshold():
    """csrc/env_lunar.cu replaces `sqrtf(x) <= B2_LINEAR_SLOP` in the joint position solve by `x <= 0x1.a36e3p-16f`:
    exact for every float32 x (checked on the 2^17 floats around the boundary and on random magnitudes)."""
    slop, T = np.float32(0.005), np.float32(float.fromhex("0x1.a36e3p-16"))
    base = T.view(np.uint32)
    xs = np.arange(int(base) - 65536, int(base) + 65536, dtype=np.uint32).view(np.float32)
    assert np.array_equal(np.sqrt(xs) <= slop, xs <= T)
    rng = np.random.default_rng(0)
    xr = np.exp(rng.uniform(-30, 5, 200000)).astype(np.float32)
    assert np.array_equal(np.sqrt(xr) <= slop, xr <= T)


# ------------------------------------------------------------------------------------------------ golden provenance
@pytest.mark.parametrize("gen,name", [("gen_noisy_dqn", "noisy_dqn_update.npz"), ("gen_sac_discrete", "sac_discrete_update.npz"),
                                      ("gen_ddqn_per", "ddqn_per_update.npz"), ("gen_ddqn_per_duel", "ddqn_per_duel_update.npz")])
def test_offpolicy_goldens_regenerate_from_the_reference(gen, name, golden, tmp_path, monkeypatch):
    """The committed fixtures of the §8f rank-2 trainers are what oracle/make_golden_offpolicy.py produces from the reference
    tree (run here, in the build container; skipped on the GPU box where /root/reference does not exist)."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    import oracle.make_golden_offpolicy as mg
    monkeypatch.setattr(mg, "OUT", tmp_path)
    if gen == "gen_ddqn_per":
        mg.gen_ddqn_per(False)
    elif gen == "gen_ddqn_per_duel":
        mg.gen_ddqn_per(True)
    else:
        getattr(mg, gen)()
    new, old = np.load(tmp_path / name), golden(name)
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        np.testing.assert_allclose(new[k], old[k], rtol=1e-6, atol=1e-7, err_msg=k)


def test_sac_discrete_oracle_matches_torch_autograd():
    """oracle.algos_np.sac_discrete_* against the reference's own expressions (algorithms/sac_cartpole.py:164-204) evaluated
    with torch autograd in float64 on the CPU."""
    import torch
    rng = np.random.default_rng(3)
    B, A_ = 257, 3
    z, zn = rng.standard_normal((B, A_)) * 2, rng.standard_normal((B, A_)) * 2
    q1, q2, q1t, q2t = (rng.standard_normal((B, A_)) for _ in range(4))
    r, d, act = rng.standard_normal(B), (rng.random(B) < 0.2).astype(np.float64), rng.integers(0, A_, B)
    log_alpha, gamma = np.log(0.3), 0.9
    T = lambda x: torch.tensor(x, dtype=torch.float64)
    alpha = torch.tensor(log_alpha, dtype=torch.float64).exp()
    # target (ref :164-176)
    npb = torch.softmax(T(zn), dim=-1)
    nlp = torch.log(npb + 1e-8)
    nH = -(npb * nlp).sum(1, keepdim=True)
    y_ref = T(r)[:, None] + gamma * (1 - T(d))[:, None] * ((npb * torch.min(T(q1t), T(q2t))).sum(1, keepdim=True) + alpha * nH)
    y = A.sac_discrete_target(zn, q1t, q2t, r, d, log_alpha, gamma)
    np.testing.assert_allclose(y, y_ref[:, 0].numpy(), rtol=1e-12, atol=1e-12)
    # critics (ref :178-189)
    tq1, tq2 = T(q1).requires_grad_(), T(q2).requires_grad_()
    l1 = torch.nn.functional.mse_loss(tq1.gather(1, torch.tensor(act)[:, None]), y_ref)
    l2 = torch.nn.functional.mse_loss(tq2.gather(1, torch.tensor(act)[:, None]), y_ref)
    (l1 + l2).backward()
    c = A.sac_discrete_critic(q1, q2, act, y)
    np.testing.assert_allclose([c["loss1"], c["loss2"]], [l1.item(), l2.item()], rtol=1e-12)
    np.testing.assert_allclose(c["dq1"], tq1.grad.numpy(), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(c["dq2"], tq2.grad.numpy(), rtol=1e-12, atol=1e-15)
    # actor (ref :191-200)
    tz = T(z).requires_grad_()
    pb = torch.softmax(tz, dim=-1)
    lp = torch.log(pb + 1e-8)
    H = -(pb * lp).sum(1, keepdim=True)
    la = torch.mean(-alpha * H - (pb * torch.min(T(q1), T(q2))).sum(1, keepdim=True))
    la.backward()
    a = A.sac_discrete_actor(z, q1, q2, log_alpha)
    np.testing.assert_allclose(a["loss"], la.item(), rtol=1e-12)
    np.testing.assert_allclose(a["dlogits"], tz.grad.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(a["sum_entropy"], H.sum().item(), rtol=1e-12)


# ---------------------------------------------------------------- §8f rank 3: masked_mean / value-clip loss, GRU cell
@pytest.mark.parametrize("tag", ["script_window", "tight_window"])
def test_oracle_ppo_lstm_masked_mean_value_clip_golden(golden, tag):
    """oracle ppo_loss_grad(mode=full, masked_mean, value-clip) == autograd of the reference's own update_model expressions
    (ppo_lstm_lunarlander.py:646-655 masked_mean, :757-771 losses; golden from oracle/make_golden_f3.py)."""
    g = golden("ppo_lstm_loss.npz")
    k = lambda n: g[f"{tag}__{n}"]
    r = A.ppo_loss_grad(k("logits"), k("values"), k("action"), k("logp_old"), k("adv"), k("ret"), mode="full",
                        clip_eps_min=float(k("clip_eps_min")), clip_eps_max=float(k("clip_eps_max")), dual_clip=float(k("dual_clip")),
                        value_coef=0.5, entropy_coef=float(k("entropy_coef")), entropy_old=k("entropy_old"),
                        erc_low=float(k("erc_low")), erc_high=float(k("erc_high")), value_old=k("value_old"),
                        vclip_eps_min=float(k("clip_eps_min")), vclip_eps_max=float(k("clip_eps_max")), masked_mean=True)
    np.testing.assert_allclose(r["dlogits"], k("dlogits"), rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(r["dvalue"], k("dvalues"), rtol=2e-4, atol=1e-6)
    if tag == "tight_window":
        assert 0.2 < r["erc_frac"] < 0.5          # the denominator of masked_mean is not B here
    # the sequence minibatch of the reference is a gather of whole sequences
    S, L = int(k("n_seq")), int(k("seq_len"))
    np.testing.assert_array_equal(k("states").reshape(S, L, -1)[k("perm")], k("s_batch"))


def test_oracle_gru_sequence_golden(golden):
    """oracle gru_sequence (cell forward + BPTT) == torch.nn.GRU of the reference's MLPRNN (ppo_rnn_lunarlander.py:124-139)."""
    g = golden("gru_mlprnn.npz")
    r = A.gru_sequence(g["x"], g["h0"], g["w_ih"], g["w_hh"], g["b_ih"], g["b_hh"], dout=g["dout"], dhT=g["dhT"])
    np.testing.assert_allclose(r["out"], g["out"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r["hT"], g["hT"], rtol=1e-5, atol=1e-6)
    for name in ("dx", "dh0", "dw_ih", "dw_hh", "db_ih", "db_hh"):
        np.testing.assert_allclose(r[name], g[name], rtol=2e-4, atol=2e-5, err_msg=name)
