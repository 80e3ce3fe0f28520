"""NumPy restatement (float64) of ppo_full's ActorCritic: MHCBackbone + RMSNorm/SiLU heads, forward AND backward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against tests/golden/mhc_actor_critic.npz, which
oracle/make_golden_mhc.py produces by running the UNMODIFIED reference module
(/root/reference/algorithms/ppo_full_lunarlander.py) forward and through torch autograd.

Restates, with the reference's state_dict keys:
    sinkhorn_knopp_batched                 ppo_full_lunarlander.py:76-103   (u, v only; the caller rebuilds P, :175-178)
    ManifoldHyperConnectionFuse.mapping    :144-180   (u, v are detached: no gradient flows through the iterations)
                               .process    :182-188
                               .depth_connection :190-194
    MHCBlock.forward                       :210-229
    MHCBackbone.forward                    :251-267
    RMSNorm                                :273-284
    MLP (Linear -> SiLU -> RMSNorm -> Linear) :287-318, ActorCritic.forward :391-393
"""
import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _silu(x):
    return x * _sigmoid(x)


def _silu_grad(x):
    s = _sigmoid(x)
    return s * (1.0 + x * (1.0 - s))


def sinkhorn_uv(E, iters, eps=1e-8):
    """E [B, n, n] positive.  Returns u, v [B, n] after `iters` alternating normalisations (ref :87-97)."""
    B, n, _ = E.shape
    u = np.ones((B, n), E.dtype)
    v = np.ones((B, n), E.dtype)
    for _ in range(iters):
        u = 1.0 / (np.einsum("bij,bj->bi", E, v) + eps)
        v = 1.0 / (np.einsum("bij,bi->bj", E, u) + eps)
    return u, v


def rmsnorm_fwd(x, weight, eps=1e-6):
    inv = 1.0 / np.sqrt((x * x).mean(-1, keepdims=True) + eps)
    return x * inv * weight, inv


def rmsnorm_bwd(dy, x, inv, weight):
    nh = x * inv
    dyw = dy * weight
    dx = inv * (dyw - nh * (dyw * nh).mean(-1, keepdims=True))
    return dx, (dy * nh).sum(0)


class MHCStage:
    """One ManifoldHyperConnectionFuse + its Linear + SiLU (half of an MHCBlock)."""

    def __init__(self, sd, prefix_mhc, prefix_lin, n, sk_it, dt=np.float64):
        g = lambda k: np.asarray(sd[k], dt)
        self.g, self.w, self.alpha, self.beta = g(prefix_mhc + ".norm.weight"), g(prefix_mhc + ".w"), g(prefix_mhc + ".alpha"), g(prefix_mhc + ".beta")
        self.W, self.b = g(prefix_lin + ".weight"), g(prefix_lin + ".bias")
        self.n, self.sk_it = n, sk_it
        self.names = dict(g=prefix_mhc + ".norm.weight", w=prefix_mhc + ".w", alpha=prefix_mhc + ".alpha", beta=prefix_mhc + ".beta",
                          W=prefix_lin + ".weight", b=prefix_lin + ".bias")

    def coefficients(self, h):
        B, n, D = h.shape
        hv = h.reshape(B, n * D)
        H = (self.g * hv) @ self.w                                         # :151-154
        r = np.sqrt((hv * hv).sum(-1, keepdims=True)) / np.sqrt(n * D)     # :157
        r_ = 1.0 / (r + 1e-6)
        a = np.concatenate([np.full(n, self.alpha[0]), np.full(n, self.alpha[1]), np.full(n * n, self.alpha[2])])
        t = r_ * H * a + self.beta                                          # :162-164
        pre = _sigmoid(t[:, :n])
        post = 2.0 * _sigmoid(t[:, n:2 * n])
        E = np.exp(t[:, 2 * n:]).reshape(B, n, n)
        u, v = sinkhorn_uv(E, self.sk_it)
        P = u[:, :, None] * E * v[:, None, :]                               # :177
        return dict(hv=hv, H=H, r_=r_, a=a, t=t, pre=pre, post=post, E=E, u=u, v=v, P=P)

    def forward(self, h):
        c = self.coefficients(h)
        h_pre = np.einsum("bi,bid->bd", c["pre"], h)                         # :184
        h_res = np.einsum("bij,bjd->bid", c["P"], h)                         # :187
        z = h_pre @ self.W.T + self.b
        h_out = _silu(z)
        out = c["post"][:, :, None] * h_out[:, None, :] + h_res              # :192-193
        self.cache = (h, c, h_pre, z, h_out)
        return out

    def backward(self, dout, grads):
        h, c, h_pre, z, h_out = self.cache
        B, n, D = h.shape
        dh_out = np.einsum("bi,bid->bd", c["post"], dout)
        dpost = np.einsum("bid,bd->bi", dout, h_out)
        dP = np.einsum("bid,bjd->bij", dout, h)
        dh = np.einsum("bij,bid->bjd", c["P"], dout)
        dz = dh_out * _silu_grad(z)
        grads[self.names["W"]] = dz.T @ h_pre
        grads[self.names["b"]] = dz.sum(0)
        dh_pre = dz @ self.W
        dh = dh + c["pre"][:, :, None] * dh_pre[:, None, :]
        dpre = np.einsum("bd,bid->bi", dh_pre, h)
        dt = np.concatenate([dpre * c["pre"] * (1.0 - c["pre"]), dpost * c["post"] * (1.0 - 0.5 * c["post"]),
                             (dP * c["P"]).reshape(B, n * n)], axis=1)        # dE = u v dP (u, v detached), dt = dE * E
        grads[self.names["beta"]] = dt.sum(0)
        rH = c["r_"] * c["H"]
        dal = dt * rH
        grads[self.names["alpha"]] = np.array([dal[:, :n].sum(), dal[:, n:2 * n].sum(), dal[:, 2 * n:].sum()])
        dH = dt * c["r_"] * c["a"]
        dr_ = (dt * c["H"] * c["a"]).sum(-1, keepdims=True)
        hv = c["hv"]
        grads[self.names["w"]] = (self.g * hv).T @ dH
        wdH = dH @ self.w.T
        grads[self.names["g"]] = (hv * wdH).sum(0)
        dhv = self.g * wdH
        s = (hv * hv).sum(-1, keepdims=True)
        r = np.sqrt(s) / np.sqrt(n * D)
        dr = -c["r_"] ** 2 * dr_
        dhv = dhv + hv * dr / (np.sqrt(s) * np.sqrt(n * D))                  # d r / d hv = hv / (sqrt(s) sqrt(nD))
        del r
        return dh + dhv.reshape(B, n, D)


class ActorCriticMHC:
    def __init__(self, sd, rate=2, layers=2, sk_it=10, dt=np.float64):
        self.sd = {k: np.asarray(v, dt) for k, v in sd.items()}
        self.n, self.dt = rate, dt
        self.stages = []
        for l in range(layers):
            self.stages.append(MHCStage(self.sd, f"shared.layers.{l}.mhc1", f"shared.layers.{l}.linear1", rate, sk_it, dt))
            self.stages.append(MHCStage(self.sd, f"shared.layers.{l}.mhc2", f"shared.layers.{l}.linear2", rate, sk_it, dt))

    def forward(self, x):
        sd = self.sd
        x = np.asarray(x, self.dt)
        x0 = x @ sd["shared.input_proj.weight"].T + sd["shared.input_proj.bias"]
        h = np.repeat(x0[:, None, :], self.n, axis=1)
        for st in self.stages:
            h = st.forward(h)
        hs = h.sum(1)
        feat, inv_f = rmsnorm_fwd(hs, sd["shared.final_norm.weight"])
        heads = {}
        for name in ("actor", "critic"):
            z = feat @ sd[f"{name}.mlp.0.weight"].T + sd[f"{name}.mlp.0.bias"]
            a = _silu(z)
            y, inv = rmsnorm_fwd(a, sd[f"{name}.mlp.2.weight"])
            out = y @ sd[f"{name}.mlp.3.weight"].T + sd[f"{name}.mlp.3.bias"]
            heads[name] = (z, a, inv, y, out)
        self.cache = (x, x0, hs, inv_f, feat, heads)
        return heads["actor"][4], heads["critic"][4]

    def backward(self, dlogits, dvalue):
        """Returns {state_dict key: gradient} for every parameter."""
        sd = self.sd
        x, x0, hs, inv_f, feat, heads = self.cache
        grads = {}
        dfeat = np.zeros_like(feat)
        for name, dout in (("actor", np.asarray(dlogits, self.dt)), ("critic", np.asarray(dvalue, self.dt).reshape(len(x), -1))):
            z, a, inv, y, _ = heads[name]
            grads[f"{name}.mlp.3.weight"] = dout.T @ y
            grads[f"{name}.mlp.3.bias"] = dout.sum(0)
            dy = dout @ sd[f"{name}.mlp.3.weight"]
            da, grads[f"{name}.mlp.2.weight"] = rmsnorm_bwd(dy, a, inv, sd[f"{name}.mlp.2.weight"])
            dz = da * _silu_grad(z)
            grads[f"{name}.mlp.0.weight"] = dz.T @ feat
            grads[f"{name}.mlp.0.bias"] = dz.sum(0)
            dfeat += dz @ sd[f"{name}.mlp.0.weight"]
        dhs, grads["shared.final_norm.weight"] = rmsnorm_bwd(dfeat, hs, inv_f, sd["shared.final_norm.weight"])
        dh = np.repeat(dhs[:, None, :], self.n, axis=1)
        for st in reversed(self.stages):
            dh = st.backward(dh, grads)
        dx0 = dh.sum(1)
        grads["shared.input_proj.weight"] = dx0.T @ x
        grads["shared.input_proj.bias"] = dx0.sum(0)
        return grads
