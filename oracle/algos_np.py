"""NumPy restatements of the learner-side arithmetic of the reference hot path.  TEST INFRASTRUCTURE.

Pinned against tests/golden/*.npz (outputs of the reference's own classes, see oracle/make_golden.py)
by the `not gpu` tests; the GPU parity tests then compare the CUDA kernels with these functions on
other seeds / sizes.  Each function cites the reference lines it restates.
"""
import numpy as np


# ---- GAE (algorithms/ppo_lunarlander.py:179-196; ppo_full_lunarlander.py:507-535) --------------------------
def gae_algorithms(reward, value, v_last, done, gamma, lam_actor, lam_critic=None, coef_f32=True, boot_f32=False):
    """[T, N] arrays; float64 recurrence exactly as the reference loop; returns (adv_actor, returns) float64.

    coef_f32: ppo_lunarlander.py builds `dones` as a float32 array (:181), so under NumPy >= 2 (NEP 50, the
    version in this image) the trace coefficient `gamma * lam * (1 - dones[t])` is a *float32* scalar before it
    multiplies the float64 `last_gae`.  ppo_full keeps `dones` boolean (:510), so its coefficient stays float64
    (coef_f32=False) - but it stores `values` as 0-dim float32 tensors, so np.array(values) is float32 and
    `gamma * values[t + 1]` is a float32 product (boot_f32=True)."""
    lam_critic = lam_actor if lam_critic is None else lam_critic
    reward = np.asarray(reward, np.float64); value = np.asarray(value, np.float64)
    T = reward.shape[0]
    vals = np.concatenate([value, np.asarray(v_last, np.float64)[None]], axis=0)
    nd = 1.0 - np.asarray(done, np.float64)
    adv_a = np.zeros_like(reward); adv_c = np.zeros_like(reward)
    la = np.zeros_like(reward[0]); lc = np.zeros_like(reward[0])
    for t in reversed(range(T)):
        boot = gamma * vals[t + 1]
        if boot_f32:
            boot = (np.float32(gamma) * vals[t + 1].astype(np.float32)).astype(np.float64)
        delta = reward[t] + boot * nd[t] - vals[t]
        ca, cc = gamma * lam_actor * nd[t], gamma * lam_critic * nd[t]
        if coef_f32:
            ca, cc = ca.astype(np.float32).astype(np.float64), cc.astype(np.float32).astype(np.float64)
        adv_a[t] = la = delta + ca * la
        adv_c[t] = lc = delta + cc * lc
    return adv_a, adv_c + vals[:-1]


# ---- GAE, utils dialect (utils/buffer.py:21-35): float32 recurrence --------------------------------------
def gae_utils(reward, value, next_value, done, dw, gamma, lamda):
    f = np.float32
    reward, value, next_value = (np.asarray(x, f) for x in (reward, value, next_value))
    done, dw = np.asarray(done, f), np.asarray(dw, f)
    td = (reward + (f(gamma) * next_value) * (f(1) - dw)) - value
    T = reward.shape[0]
    adv = np.zeros_like(reward)
    gl = f(gamma * lamda)
    gae = None
    for t in reversed(range(T)):
        gae = td[t] if gae is None else (gl * gae) * (f(1) - done[t]) + td[t]
        adv[t] = gae
    return adv, adv + value


# ---- Categorical (algorithms/ppo_lunarlander.py:92-117; torch.distributions.Categorical) --------------------
def categorical(logits, noise=None):
    """Returns (normalised logits, probs, action or None, entropy) in float32, following torch's op sequence."""
    z = np.asarray(logits, np.float32)
    m = z.max(-1, keepdims=True)
    lse = np.log(np.exp(z - m).sum(-1, keepdims=True, dtype=np.float32)) + m
    ln = z - lse
    e = np.exp(ln - ln.max(-1, keepdims=True))
    p = e / e.sum(-1, keepdims=True, dtype=np.float32)
    ent = -(np.maximum(ln, np.finfo(np.float32).min) * p).sum(-1, dtype=np.float32)
    action = None if noise is None else np.argmax(p / np.asarray(noise, np.float32), axis=-1).astype(np.int32)
    return ln, p, action, ent


# ---- PPO losses + analytic gradients (ppo_lunarlander.py:278-300; ppo_full_lunarlander.py:586-633;
#      ppo_lstm_lunarlander.py:763-771) -----------------------------------------------------------------------
def ppo_loss_grad(logits, value, action, logp_old, adv, ret, *, mode="dualclip", clip_eps_min=0.2, clip_eps_max=0.2,
                  dual_clip=3.0, value_coef=0.5, entropy_coef=0.01, entropy_old=None, erc_low=0.06, erc_high=0.06,
                  value_old=None, vclip_eps_min=0.2, vclip_eps_max=0.2, masked_mean=False):
    """float64 restatement with torch's tie conventions (min/max split ties evenly; clamp passes grads on the
    closed interval).  Returns dict(dlogits, dvalue, policy_loss, value_loss, entropy, clip_frac, approx_kl, erc_frac)."""
    z = np.asarray(logits, np.float64); V = np.asarray(value, np.float64)
    B, A = z.shape
    a = np.asarray(action)
    ln = z - (np.log(np.exp(z - z.max(-1, keepdims=True)).sum(-1, keepdims=True)) + z.max(-1, keepdims=True))
    p = np.exp(ln)
    H = -(p * ln).sum(-1)
    lp = ln[np.arange(B), a]
    ratio = np.exp(lp - logp_old)
    Adv = np.asarray(adv, np.float64); R = np.asarray(ret, np.float64)
    lo, hi = 1 - clip_eps_min, 1 + clip_eps_max
    in_range = (ratio >= lo) & (ratio <= hi)
    surr2 = np.clip(ratio, lo, hi) * Adv
    g2 = np.where(in_range, Adv, 0.0)
    mask = np.ones(B)
    if mode == "full":
        er = H / (np.asarray(entropy_old, np.float64) + 1e-8)
        mask = ((er > 1 - erc_low) & (er < 1 + erc_high)).astype(np.float64)
        surr1 = np.clip(ratio, 0.0, dual_clip) * Adv
        g1 = np.where((ratio >= 0) & (ratio <= dual_clip), Adv, 0.0)
        obj = np.minimum(surr1, surr2)
        g = np.where(surr1 < surr2, g1, np.where(surr1 == surr2, 0.5 * (g1 + g2), g2))
    else:
        surr1 = ratio * Adv
        min_surr = np.minimum(surr1, surr2)
        gm = np.where(surr1 < surr2, Adv, np.where(surr1 == surr2, 0.5 * (Adv + g2), g2))
        dc = dual_clip * Adv
        neg = Adv < 0
        obj = np.where(neg, np.maximum(min_surr, dc), min_surr)
        g = np.where(neg, np.where(min_surr > dc, gm, np.where(min_surr == dc, 0.5 * gm, 0.0)), gm)
    e1 = V - R
    vterm, dv = e1 * e1, 2 * e1
    if value_old is not None:
        vo = np.asarray(value_old, np.float64)
        dlt = V - vo
        vcl = vo + np.clip(dlt, -vclip_eps_min, vclip_eps_max)
        e2 = vcl - R
        l2 = e2 * e2
        d2 = 2 * e2 * ((dlt >= -vclip_eps_min) & (dlt <= vclip_eps_max))
        dv = np.where(l2 > vterm, d2, np.where(l2 == vterm, 0.5 * (dv + d2), dv))
        vterm = np.maximum(vterm, l2)
    # masked_mean (ppo_lstm_lunarlander.py:646-655): sum(x * mask) / mask.sum(), 0 when mask.sum() == 0; ppo_full: plain mean
    Nn = B
    if masked_mean:
        Nn = mask.sum() if mask.sum() > 0 else np.inf
    dL_dlp = -(g * ratio) * mask / Nn
    dL_dH = -entropy_coef * mask / Nn
    onehot = np.zeros((B, A)); onehot[np.arange(B), a] = 1.0
    dlogits = dL_dlp[:, None] * (onehot - p) + dL_dH[:, None] * (-p * (ln + H[:, None]))
    dvalue = value_coef * mask * dv / Nn
    return dict(dlogits=dlogits, dvalue=dvalue, policy_loss=(-obj * mask).sum() / Nn, value_loss=(value_coef * mask * vterm).sum() / Nn,
                entropy=(H * mask).sum() / Nn, clip_frac=(((ratio < lo) | (ratio > hi)) * mask).sum() / Nn,
                approx_kl=(np.asarray(logp_old, np.float64) - lp).mean(), erc_frac=1.0 - mask.mean())


# ---- torch.optim.Adam single-tensor update + clip_grad_norm_ (ppo_lunarlander.py:169,304-306) ---------------
def adam_step(param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, max_norm=0.0, clamp=0.0):
    """float32 arrays; returns (param, m, v) after one step (step = 1-based count after increment)."""
    f = np.float32
    g = np.asarray(grad, f).copy()
    if max_norm > 0:
        tn = f(np.sqrt((g.astype(np.float64) ** 2).sum()))
        coef = min(f(max_norm) / (tn + f(1e-6)), f(1.0))
        g = g * f(coef)
    if clamp > 0:
        g = np.clip(g, -f(clamp), f(clamp))
    m = m + f(1 - beta1) * (g - m)
    v = v * f(beta2) + (f(1 - beta2) * g) * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / f(np.sqrt(bc2)) + f(eps)
    param = param + (f(-lr / bc1) * m) / denom
    return param.astype(f), m.astype(f), v.astype(f)


def mlp_actor_critic_forward(sd, x):
    """ActorCritic.forward (algorithms/ppo_lunarlander.py:86-90) in float64 NumPy from a state_dict of arrays."""
    x = np.asarray(x, np.float64)
    h = np.tanh(x @ sd["shared.0.weight"].T.astype(np.float64) + sd["shared.0.bias"])
    h = np.tanh(h @ sd["shared.2.weight"].T.astype(np.float64) + sd["shared.2.bias"])
    a = np.tanh(h @ sd["actor.0.weight"].T.astype(np.float64) + sd["actor.0.bias"])
    c = np.tanh(h @ sd["critic.0.weight"].T.astype(np.float64) + sd["critic.0.bias"])
    return a @ sd["actor.2.weight"].T.astype(np.float64) + sd["actor.2.bias"], (c @ sd["critic.2.weight"].T.astype(np.float64) + sd["critic.2.bias"])[:, 0]


# ---- running normalisation (utils/normalization.py:4-52) -------------------------------------------------------------
class RunningMeanStdNP:
    """Same update rule and precisions as the reference: mean float32, S float64, std float64 (float32 = x at n == 1)."""

    def __init__(self, D):
        self.n, self.mean, self.S, self.std = 0, np.zeros(D), np.zeros(D), np.zeros(D)

    def update(self, x):
        x = np.array(x, dtype=np.float32)
        self.n += 1
        if self.n == 1:
            self.mean, self.std = x, x
        else:
            old = self.mean.copy()
            self.mean = (old + (x - old) / np.float32(self.n)).astype(np.float32)
            self.S = self.S + ((x - old) * (x - self.mean)).astype(np.float64)
            self.std = np.sqrt(self.S / self.n)


def normalize_stream(x):
    """Normalization.__call__(x_t) for t = 0.. (update=True) -> float32 rows."""
    rm = RunningMeanStdNP(x.shape[1])
    out = np.zeros_like(x, dtype=np.float32)
    for t in range(len(x)):
        rm.update(x[t])
        out[t] = ((x[t] - rm.mean).astype(np.float64) / (np.asarray(rm.std, np.float64) + 1e-8)).astype(np.float32)
    return out, rm


def reward_scaling_stream(r, reset_at, gamma):
    rm, R, out = RunningMeanStdNP(1), np.zeros(1), np.zeros(len(r), np.float32)
    for t in range(len(r)):
        if reset_at[t]:
            R = np.zeros(1)
        R = gamma * R + r[t]
        rm.update(R)
        out[t] = np.float32(r[t] / (float(np.asarray(rm.std, np.float64)[0]) + 1e-8))
    return out


# ---- discrete SAC (algorithms/sac_cartpole.py:155-221) ------------------------------------------------------
def _softmax_entropy(logits):
    """p = softmax(logits), lp = log(p + 1e-8), H = -sum p lp  (Actor.forward :96-99 and the update :166-168, :191-193)."""
    z = np.asarray(logits, np.float64)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    lp = np.log(p + 1e-8)
    return p, lp, -(p * lp).sum(axis=1)


def sac_discrete_target(logits_next, q1t, q2t, reward, done, log_alpha, gamma):
    """target_q = r + gamma (1 - d) (sum_a p' min(Q1t, Q2t) + alpha H(p'))  (ref :164-176); float64 [B]."""
    p, _, H = _softmax_entropy(logits_next)
    mq = (p * np.minimum(np.asarray(q1t, np.float64), np.asarray(q2t, np.float64))).sum(axis=1)
    return np.asarray(reward, np.float64) + gamma * (1.0 - np.asarray(done, np.float64)) * (mq + np.exp(float(log_alpha)) * H)


def sac_discrete_critic(q1, q2, action, y):
    """mse(Q1(s)[a], y), mse(Q2(s)[a], y) and their gradients wrt the full [B, A] outputs (ref :178-189)."""
    q1, q2, y = np.asarray(q1, np.float64), np.asarray(q2, np.float64), np.asarray(y, np.float64)
    B = len(y)
    rows = np.arange(B)
    e1, e2 = q1[rows, action] - y, q2[rows, action] - y
    d1, d2 = np.zeros_like(q1), np.zeros_like(q2)
    d1[rows, action] = 2.0 * e1 / B
    d2[rows, action] = 2.0 * e2 / B
    return dict(loss1=(e1 ** 2).mean(), loss2=(e2 ** 2).mean(), dq1=d1, dq2=d2)


def sac_discrete_actor(logits, q1, q2, log_alpha):
    """actor_loss = mean(-alpha H - sum_a p min(Q1, Q2)) and d/dlogits through softmax and log(p + 1e-8) (ref :191-200);
    sum_entropy feeds the alpha loss mean(exp(log_alpha) (H - target_entropy)) (ref :202-204)."""
    p, lp, H = _softmax_entropy(logits)
    alpha = np.exp(float(log_alpha))
    m = np.minimum(np.asarray(q1, np.float64), np.asarray(q2, np.float64))
    B = p.shape[0]
    g = alpha * (lp + p / (p + 1e-8)) - m                  # dL_i / dp_j
    dz = p * (g - (p * g).sum(axis=1, keepdims=True)) / B
    return dict(loss=(-alpha * H - (p * m).sum(axis=1)).mean(), dlogits=dz, sum_entropy=H.sum())


# ---- torch.nn.GRU cell (ppo_rnn_lunarlander.py:124-139 MLPRNN.rnn; gate order r | z | n) -----------------------
def gru_cell_forward(gi, gh, h):
    """float64: returns (h_new, r, z, n) for gi = x W_ih^T + b_ih, gh = h W_hh^T + b_hh of shape [B, 3H]."""
    gi, gh, h = (np.asarray(a, np.float64) for a in (gi, gh, h))
    Hd = h.shape[-1]
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    r = sig(gi[:, :Hd] + gh[:, :Hd])
    z = sig(gi[:, Hd:2 * Hd] + gh[:, Hd:2 * Hd])
    n = np.tanh(gi[:, 2 * Hd:] + r * gh[:, 2 * Hd:])
    return (1 - z) * n + z * h, r, z, n


def gru_cell_backward(dh_new, r, z, n, gh, h):
    """Returns (dgi, dgh, dh_direct) given dL/dh_new."""
    Hd = h.shape[-1]
    dn = dh_new * (1 - z)
    dz = dh_new * (h - n)
    dan = dn * (1 - n * n)
    dr = dan * gh[:, 2 * Hd:]
    daz, dar = dz * z * (1 - z), dr * r * (1 - r)
    dgi = np.concatenate([dar, daz, dan], axis=1)
    dgh = np.concatenate([dar, daz, dan * r], axis=1)
    return dgi, dgh, dh_new * z


def gru_sequence(x, h0, w_ih, w_hh, b_ih, b_hh, dout=None, dhT=None):
    """x [B, T, I], h0 [B, H]: unrolled GRU forward and (when dout [B, T, H] is given) BPTT, float64.
    Returns dict(out, hT[, dx, dh0, dw_ih, dw_hh, db_ih, db_hh])."""
    x, h0, w_ih, w_hh, b_ih, b_hh = (np.asarray(a, np.float64) for a in (x, h0, w_ih, w_hh, b_ih, b_hh))
    B, T, _ = x.shape
    hs, saved, h = [], [], h0
    for t in range(T):
        gi, gh = x[:, t] @ w_ih.T + b_ih, h @ w_hh.T + b_hh
        hn, r, z, n = gru_cell_forward(gi, gh, h)
        saved.append((r, z, n, gh, h))
        hs.append(hn)
        h = hn
    res = {"out": np.stack(hs, axis=1), "hT": h}
    if dout is None:
        return res
    dout = np.asarray(dout, np.float64)
    dh = np.zeros_like(h0) if dhT is None else np.asarray(dhT, np.float64)
    dx = np.zeros_like(x)
    dw_ih, dw_hh, db_ih, db_hh = np.zeros_like(w_ih), np.zeros_like(w_hh), np.zeros_like(b_ih), np.zeros_like(b_hh)
    for t in reversed(range(T)):
        r, z, n, gh, hp = saved[t]
        dgi, dgh, dh_dir = gru_cell_backward(dout[:, t] + dh, r, z, n, gh, hp)
        dx[:, t] = dgi @ w_ih
        dw_ih += dgi.T @ x[:, t]; db_ih += dgi.sum(0)
        dw_hh += dgh.T @ hp; db_hh += dgh.sum(0)
        dh = dh_dir + dgh @ w_hh
    res.update(dx=dx, dh0=dh, dw_ih=dw_ih, dw_hh=dw_hh, db_ih=db_ih, db_hh=db_hh)
    return res
