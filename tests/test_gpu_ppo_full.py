"""GPU parity of the ppo_full path: the mHC ActorCritic forward and backward through the C ABI
  (1) against the golden fixture the UNMODIFIED reference module produced (forward + torch-autograd gradients,
      tests/golden/mhc_actor_critic.npz, oracle/make_golden_mhc.py),
  (2) against the float64 NumPy oracle (oracle/mhc_np.py) on another seed / a ragged batch / the bench minibatch size,
and the trainer built on it (rollout -> decoupled-lambda GAE -> ERC-masked update).
Tolerances: float32 network, so 2e-5 absolute on O(0.1..1) outputs and 2e-4 of the largest entry per gradient tensor.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
from oracle import algos_np as A  # noqa: E402
from oracle.mhc_np import ActorCriticMHC  # noqa: E402


def cu(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _engine(sd, cfg=None):
    from gymrl_b200.algorithms import ppo_full_lunarlander as F
    cfg = cfg or F.Config()
    model = F.ActorCritic(8, 4, config=cfg)
    missing = model.load_state_dict({k: torch.as_tensor(v) for k, v in sd.items()})
    assert not missing.missing_keys and not missing.unexpected_keys   # the reference's state_dict keys, exactly
    return model.to_engine("cuda")


def _run(eng, x, dl, dv):
    M = x.shape[0]
    acts = eng.make_acts(M, backward=True)
    eng.alloc_workspace(M)
    xd = cu(x, torch.float32)
    eng.forward(xd, acts, M)
    logits, value = acts.lv[:, :4].cpu().numpy().copy(), acts.lv[:, 4].cpu().numpy().copy()
    acts.dlv.zero_()
    acts.dlv[:, :4].copy_(cu(dl, torch.float32)); acts.dlv[:, 4].copy_(cu(dv, torch.float32).reshape(-1))
    eng.fp.grad.zero_()
    eng.backward(xd, acts, M)
    torch.cuda.synchronize()
    grads = {n: eng.fp.g(n).cpu().numpy().copy() for n in eng.fp.views}
    return logits, value, grads


def _check_grads(grads, ref, tol):
    for k, g in grads.items():
        r = np.asarray(ref[k]).reshape(g.shape)
        scale = max(np.abs(r).max(), 1e-6)
        assert np.abs(g - r).max() <= tol * scale, (k, np.abs(g - r).max(), scale)


def test_mhc_actor_critic_golden(golden):
    g = golden("mhc_actor_critic.npz")
    sd = {k[2:]: g[k] for k in g.files if k.startswith("p:")}
    eng = _engine(sd)
    assert eng.fp.numel() >= 144433 and sum(p.numel() for p in eng.model.parameters()) == 144433
    logits, value, grads = _run(eng, g["x"], g["Gl"], g["Gv"])
    np.testing.assert_allclose(logits, g["logits"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(value, g["value"].reshape(-1), rtol=0, atol=2e-5)
    _check_grads(grads, {k[2:]: g[k] for k in g.files if k.startswith("g:")}, 2e-4)


@pytest.mark.parametrize("M,seed", [(1, 3), (333, 4), (16384, 5)])
def test_mhc_actor_critic_vs_oracle(golden, M, seed):
    g = golden("mhc_actor_critic.npz")
    rng = np.random.default_rng(seed)
    sd = {k[2:]: (g[k] + 0.02 * rng.standard_normal(g[k].shape)).astype(np.float32) for k in g.files if k.startswith("p:")}
    x = rng.standard_normal((M, 8)).astype(np.float32) * np.array([0.5, 0.7, 1.0, 1.0, 0.5, 1.0, 0.5, 0.5], np.float32)
    dl = (rng.standard_normal((M, 4)) / M).astype(np.float32)
    dv = (rng.standard_normal((M, 1)) / M).astype(np.float32)
    ora = ActorCriticMHC(sd, 2, 2, 10)
    ol, ov = ora.forward(x)
    og = ora.backward(dl, dv)
    logits, value, grads = _run(_engine(sd), x, dl, dv)
    np.testing.assert_allclose(logits, ol, rtol=0, atol=3e-5)
    np.testing.assert_allclose(value, ov.reshape(-1), rtol=0, atol=3e-5)
    _check_grads(grads, og, 3e-4)


def test_mhc_default_init_is_identity_like():
    """With the reference initialisation (w = 0) the mapping is input independent: pre = sigmoid(0.01), post =
    2 sigmoid(0.01), P = Sinkhorn(exp([[2,-2],[-2,2]])) for every row (ref :125-139) - checked through the oracle."""
    from gymrl_b200.algorithms import ppo_full_lunarlander as F
    torch.manual_seed(0)
    model = F.ActorCritic(8, 4, config=F.Config())
    sd = {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    eng = model.to_engine("cuda")
    x = np.random.default_rng(0).standard_normal((64, 8)).astype(np.float32)
    acts = eng.make_acts(64, backward=False)
    eng.forward(cu(x), acts, 64)
    c = acts.coef[0].cpu().numpy()
    assert np.allclose(c[:, 0], 1 / (1 + np.exp(-0.01)), atol=1e-6) and np.allclose(c[:, 2], 2 / (1 + np.exp(-0.01)), atol=1e-6)
    assert np.allclose(c[:, 4] + c[:, 5], 1.0, atol=1e-5) and np.allclose(c[:, 4] + c[:, 6], 1.0, atol=1e-5)   # doubly stochastic
    ol, ov = ActorCriticMHC(sd, 2, 2, 10).forward(x)
    np.testing.assert_allclose(acts.lv[:, :4].cpu().numpy(), ol, atol=2e-5)
    np.testing.assert_allclose(acts.lv[:, 4].cpu().numpy(), ov.reshape(-1), atol=2e-5)


@pytest.mark.parametrize("graph", [False, True])
def test_ppo_full_trainer_iteration(graph):
    from gymrl_b200.algorithms import ppo_full_lunarlander as F
    cfg = F.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.num_epochs, cfg.seed, cfg.use_cuda_graph = 128, 32, 4, 2, 11, graph
    tr = F.PPOTrainer(cfg)
    tr.collect_experience()
    buf = tr.buffer
    # the stored entropies / log-probs / values are those of the sampled policy (ref :476-488)
    acts = tr.net.make_acts(128, backward=False)
    tr.net.forward(buf.obs[5], acts, 128)
    lp_all, _, _, ent = A.categorical(acts.lv[:, :4].cpu().numpy())
    a = buf.action[5].cpu().numpy()
    np.testing.assert_allclose(buf.log_prob[5].cpu().numpy(), lp_all[np.arange(128), a], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(buf.entropy[5].cpu().numpy(), ent, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(buf.value[5].cpu().numpy(), acts.lv[:, 4].cpu().numpy(), rtol=0, atol=0)
    adv, ret = tr.compute_advantages()
    ea, er = A.gae_algorithms(buf.reward.cpu().numpy(), buf.value.cpu().numpy(), buf.v_last.cpu().numpy(), buf.done.cpu().numpy(),
                              cfg.gamma, cfg.lam_actor, cfg.lam_critic, coef_f32=False, boot_f32=True)
    np.testing.assert_allclose(adv.cpu().numpy(), ea, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ret.cpu().numpy(), er, rtol=1e-5, atol=1e-5)
    before = tr.net.fp.flat.clone()
    m = tr.update_model(adv, ret)
    assert all(np.isfinite(v) for v in m.values()), m
    assert not torch.equal(before, tr.net.fp.flat)
    # anneal after the update (ref :659-666)
    frac = 1 - tr.step_count / cfg.max_train_steps
    assert abs(tr.optimizer.param_groups[0]["lr"] - cfg.lr * frac) < 1e-12 and abs(tr.ent_coef - cfg.entropy_coef * frac) < 1e-12
    assert abs(float(tr._ent_coef_t.item()) - tr.ent_coef) < 1e-9
    m2, _, _ = tr.train_iteration()
    assert all(np.isfinite(v) for v in m2.values()), m2


def test_ppo_full_update_matches_oracle_step():
    """One minibatch = the whole rollout, one epoch: the parameters after update_model equal an oracle step built from
    the oracle network gradient + oracle loss gradient + oracle Adam (global-norm clip 0.5, eps 1e-5)."""
    from gymrl_b200.algorithms import ppo_full_lunarlander as F
    cfg = F.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.num_epochs, cfg.seed, cfg.use_cuda_graph = 64, 8, 1, 1, 5, False
    cfg.anneal = False
    tr = F.PPOTrainer(cfg)
    # make the mapping input dependent
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for n, p in tr.model.named_parameters():
            if n.endswith(".w"):
                p.copy_((torch.randn(p.shape, generator=g) * 0.05).cuda())
            if n == "actor.mlp.3.weight":
                p.mul_(100.0)
    tr.collect_experience()
    adv, ret = tr.compute_advantages()
    buf = tr.buffer
    sd = {k: v.detach().cpu().numpy().copy() for k, v in tr.model.state_dict().items()}
    ora = ActorCriticMHC(sd, 2, 2, 10)
    x = buf.obs[:8].reshape(-1, 8).cpu().numpy()
    logits, value = ora.forward(x)
    out = A.ppo_loss_grad(logits.astype(np.float32), value.reshape(-1).astype(np.float32), buf.action.view(-1).cpu().numpy(),
                     buf.log_prob.view(-1).cpu().numpy(), adv.view(-1).cpu().numpy(), ret.view(-1).cpu().numpy(), mode="full",
                     clip_eps_min=cfg.clip_eps_min, clip_eps_max=cfg.clip_eps_max, dual_clip=cfg.dual_clip, value_coef=0.5,
                     entropy_coef=cfg.entropy_coef, entropy_old=buf.entropy.view(-1).cpu().numpy(), erc_low=cfg.erc_beta_low,
                     erc_high=cfg.erc_beta_high)
    grads = ora.backward(out["dlogits"], out["dvalue"])
    names = list(tr.net.fp.views)
    flat_g = np.zeros(tr.net.fp.numel(), np.float64)
    flat_p = tr.net.fp.flat.cpu().numpy().astype(np.float64)
    for n in names:
        o, shape = tr.net.fp.views[n]
        flat_g[o:o + int(np.prod(shape))] = grads[n].reshape(-1)
    norm = np.sqrt((flat_g ** 2).sum())
    flat_g *= min(1.0, cfg.max_grad_norm / (norm + 1e-6))
    mhat, vhat = flat_g, flat_g ** 2          # first Adam step: m/(1-b1) = g, v/(1-b2) = g^2
    expect = flat_p - cfg.lr * mhat / (np.sqrt(vhat) + 1e-5)
    tr.update_model(adv, ret)
    got = tr.net.fp.flat.cpu().numpy()
    # Adam's first step moves every touched parameter by ~lr; compare the step itself
    step_err = np.abs((got - flat_p) - (expect - flat_p))
    big = np.abs(flat_g) > 1e-6 * np.abs(flat_g).max()    # sign-stable entries (|g| >> fp32 noise)
    assert step_err[big].max() < 0.05 * cfg.lr, step_err[big].max()
