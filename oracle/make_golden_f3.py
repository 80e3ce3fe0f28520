"""Generate tests/golden/ppo_lstm_loss.npz and gru_mlprnn.npz from the UNMODIFIED reference.
TEST INFRASTRUCTURE; build container only:  python -m oracle.make_golden_f3

ppo_lstm_loss.npz — the real `algorithms/ppo_lstm_lunarlander.py` PPOTrainer on oracle/gymnasium_shim: collect_experience
    (256 steps), compute_advantages, then ONE minibatch of update_model (:657-816) holding every sequence, with the model's
    outputs intercepted so that autograd's d loss / d logits and d loss / d values of the reference's own loss expressions
    (ERC mask, masked_mean :646-655, value-clip :763-771) are recorded, together with the permuted sequence minibatch
    (states.view(S, L, -1)[perm], :682-707) for the sequence-gather kernel.  Two variants: the script's ERC window (most rows
    unmasked) and a tight window (a third of the rows masked), so masked_mean's denominator matters.
gru_mlprnn.npz — `algorithms/ppo_rnn_lunarlander.py` MLPRNN(256, 256).rnn (torch.nn.GRU, hidden 64, :124-139) unrolled over
    T = 6 steps for B = 5 rows with autograd gradients of a weighted sum of the outputs.
"""
from pathlib import Path

import numpy as np
import torch

from . import gymnasium_shim, ref_loader as rl

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def gen_ppo_lstm(tag, erc_low, erc_high, seed):
    import contextlib
    import io
    import random
    gymnasium_shim.install()
    mod = rl.load("algorithms/ppo_lstm_lunarlander.py", name=f"ref_f3_ppo_lstm_{tag}")
    torch.manual_seed(seed); np.random.seed(seed); random.seed(seed)
    cfg = mod.Config()
    cfg.update_freq, cfg.num_epochs, cfg.seed = 256, 1, seed
    cfg.batch_size = cfg.update_freq // cfg.seq_len          # one minibatch holding all 32 sequences
    if erc_low is not None:
        cfg.erc_beta_low, cfg.erc_beta_high = erc_low, erc_high
    with contextlib.redirect_stdout(io.StringIO()):
        tr = mod.PPOTrainer(cfg)
        tr.collect_experience()
        adv, ret = tr.compute_advantages()
    # a few policy-changing steps so that ratio != 1 and the ERC mask / clips have something to do
    opt = torch.optim.SGD(tr.model.parameters(), lr=0.03 if erc_low is None else 0.01)
    s = torch.tensor(np.array(tr.buffer.states), dtype=torch.float32).view(-1, cfg.seq_len, len(tr.buffer.states[0]))
    h0 = torch.tensor(np.array(tr.buffer.hidden_states), dtype=torch.float32)[:: cfg.seq_len]
    for _ in range(2):
        out = tr.model(s, h0)
        opt.zero_grad(); (out[0].pow(2).mean() * 3 + out[1].mean()).backward(); opt.step()
    if erc_low is None:     # tight variant: pick the window from the data so that about a third of the rows are masked
        with torch.no_grad():
            lgt = tr.model(s, h0)[0].reshape(-1, 4)
            Hn = torch.distributions.Categorical(logits=lgt).entropy()
            er = (Hn / (torch.tensor(tr.buffer.old_entropies, dtype=torch.float32) + 1e-8) - 1).abs()
            srt = torch.sort(er).values          # a window half-way between two samples: no row sits on the boundary
            kq = int(0.65 * srt.numel())
            erc_low = erc_high = float(0.5 * (srt[kq] + srt[kq + 1]))
        tr.cfg.erc_beta_low, tr.cfg.erc_beta_high = erc_low, erc_high
    captured = {}
    real_forward = tr.model.forward

    def forward(x, hidden):
        out = real_forward(x, hidden)
        if torch.is_grad_enabled() and "logits" not in captured and x.dim() == 3:
            out[0].retain_grad(); out[1].retain_grad()
            captured.update(logits=out[0], values=out[1], s_batch=x.detach().clone(), hidden=hidden.detach().clone())
        return out

    tr.model.forward = forward
    real_step = tr.optimizer.step
    tr.optimizer.step = lambda *a, **k: None                  # keep the parameters: the golden is the loss, not Adam
    rp = torch.randperm
    perm_used = {}

    def randperm(n, **k):
        p = rp(n, **k); perm_used["perm"] = p.clone(); return p

    torch.randperm = randperm
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            tr.update_model(adv, ret)
    finally:
        torch.randperm = rp
        tr.optimizer.step = real_step
    perm = perm_used["perm"].numpy()
    L, S = cfg.seq_len, cfg.update_freq // cfg.seq_len
    lg, vl = captured["logits"], captured["values"]
    f = lambda a: np.asarray(a, np.float32)
    g = lambda a: f(a).reshape(S, L)[perm].reshape(-1)        # per-step field in minibatch (permuted sequence) order
    return dict(
        tag=tag, seq_len=L, n_seq=S, perm=perm.astype(np.int32),
        states=f(np.array(tr.buffer.states)), s_batch=captured["s_batch"].numpy(),
        logits=lg.detach().reshape(-1, lg.shape[-1]).numpy(), values=vl.detach().reshape(-1).numpy(),
        dlogits=lg.grad.reshape(-1, lg.shape[-1]).numpy(), dvalues=vl.grad.reshape(-1).numpy(),
        action=np.asarray(tr.buffer.actions, np.int32).reshape(S, L)[perm].reshape(-1),
        logp_old=g(tr.buffer.log_probs), entropy_old=g(tr.buffer.old_entropies), value_old=g(tr.buffer.values),
        adv=g(adv), ret=g(ret),
        flat_action=np.asarray(tr.buffer.actions, np.int32), flat_logp_old=f(tr.buffer.log_probs),
        flat_entropy_old=f(tr.buffer.old_entropies), flat_value_old=f(tr.buffer.values), flat_adv=f(adv), flat_ret=f(ret),
        clip_eps_min=cfg.clip_eps_min, clip_eps_max=cfg.clip_eps_max, dual_clip=cfg.dual_clip, entropy_coef=float(tr.ent_coef),
        erc_low=erc_low, erc_high=erc_high)


def gen_gru():
    mod = rl.load("algorithms/ppo_rnn_lunarlander.py", name="ref_f3_ppo_rnn")
    torch.manual_seed(3)
    m = mod.MLPRNN(256, 256, batch_first=True)
    B, T = 5, 6
    x = torch.randn(B, T, 256, requires_grad=True)
    h0 = (torch.randn(1, B, 64) * 0.5).requires_grad_(True)
    out, hT = m.rnn(x, h0)
    w = torch.randn(B, T, 64)
    wT = torch.randn(1, B, 64)
    ((out * w).sum() + (hT * wT).sum()).backward()
    p = {k: v.detach().numpy() for k, v in m.rnn.named_parameters()}
    return dict(x=x.detach().numpy(), h0=h0.detach().numpy()[0], out=out.detach().numpy(), hT=hT.detach().numpy()[0], dout=w.numpy(),
                dhT=wT.numpy()[0], dx=x.grad.numpy(), dh0=h0.grad.numpy()[0], w_ih=p["weight_ih_l0"], w_hh=p["weight_hh_l0"],
                b_ih=p["bias_ih_l0"], b_hh=p["bias_hh_l0"],
                dw_ih=m.rnn.weight_ih_l0.grad.numpy(), dw_hh=m.rnn.weight_hh_l0.grad.numpy(), db_ih=m.rnn.bias_ih_l0.grad.numpy(),
                db_hh=m.rnn.bias_hh_l0.grad.numpy(), source="ppo_rnn_lunarlander.py:124-139 MLPRNN.rnn = nn.GRU(256, 64, batch_first=True)")


def main():
    a = gen_ppo_lstm("script_window", 0.06, 0.06, 0)
    b = gen_ppo_lstm("tight_window", None, None, 1)
    out = {}
    for d in (a, b):
        for k, v in d.items():
            out[f"{d['tag']}__{k}"] = v
    np.savez_compressed(OUT / "ppo_lstm_loss.npz", source="ppo_lstm_lunarlander.py:657-816 update_model (one minibatch, all sequences)", **out)
    np.savez_compressed(OUT / "gru_mlprnn.npz", **gen_gru())
    for d in (a, b):
        H = d["logits"]
        print(d["tag"], "erc window", d["erc_low"], "rows", H.shape[0], "|dlogits| max", np.abs(d["dlogits"]).max(), "nonzero rows", int((np.abs(d["dlogits"]).sum(1) > 0).sum()))
    print("wrote ppo_lstm_loss.npz, gru_mlprnn.npz")


if __name__ == "__main__":
    main()
