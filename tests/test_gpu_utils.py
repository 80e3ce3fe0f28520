"""GPU parity of the utils path: device Normalization / RewardScaling vs the
reference classes' outputs (tests/golden/normalization.npz), utils.buffer vs the reference's compute_advantage golden,
the Gymnasium-shaped env view vs the CPU oracle env, and the reference's runner loop driving an agent end to end.
"""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
from oracle import algos_np as A  # noqa: E402
from oracle import envs_np  # noqa: E402


def test_normalization_matches_reference_stream(golden):
    from gymrl_b200.utils.normalization import Normalization
    g = golden("normalization.npz")
    norm = Normalization(shape=(8,))
    ys = np.stack([norm(g["x"][t]) for t in range(len(g["x"]))])          # one observation per call, like utils/runner.py:112
    np.testing.assert_array_equal(ys, g["y"].astype(np.float32))          # same operations, same precisions: bit-exact
    np.testing.assert_array_equal(norm.running_ms.mean.astype(np.float64), g["mean"])
    np.testing.assert_allclose(norm.running_ms.std, g["std"], rtol=1e-15)
    np.testing.assert_allclose(norm.running_ms.S, g["S"], rtol=1e-15)
    assert norm.running_ms.n == int(g["n"])
    np.testing.assert_array_equal(norm(g["x"][0], update=False), g["y_eval"].astype(np.float32))
    assert norm.running_ms.n == int(g["n"])                               # update=False leaves the statistic alone


def test_reward_scaling_matches_reference_stream(golden):
    from gymrl_b200.utils.normalization import RewardScaling
    g = golden("normalization.npz")
    rs = RewardScaling(shape=1, gamma=float(g["gamma"]))
    out = []
    for t in range(len(g["r"])):
        if g["reset_at"][t]:
            rs.reset()
        out.append(rs(float(g["r"][t]))[0])
    np.testing.assert_allclose(np.array(out), g["r_scaled"], rtol=2e-6)   # reward is float64 in the reference, float32 on the device


def test_normalization_batched_merge():
    """N > 32 rows per call: merged batch moments equal the float64 moments of everything seen so far."""
    from gymrl_b200.utils.normalization import Normalization
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((5, 4096, 8)) * 3 + 1).astype(np.float32)
    norm = Normalization(shape=(8,))
    for t in range(5):
        y = norm(torch.as_tensor(x[t]).cuda())
    allx = x.reshape(-1, 8).astype(np.float64)
    np.testing.assert_allclose(norm.running_ms.mean, allx.mean(0), rtol=1e-6)
    np.testing.assert_allclose(norm.running_ms.std, allx.std(0), rtol=1e-9)
    np.testing.assert_allclose(y.cpu().numpy(), (x[4] - allx.mean(0)) / (allx.std(0) + 1e-8), rtol=1e-4, atol=1e-5)


def _cfg(**kw):
    c = types.SimpleNamespace(gamma=0.99, lamda=0.95, device=torch.device("cuda"), batch_size=64, memory_capacity=1000, seed=3)
    c.__dict__.update(kw)
    return c


def test_on_policy_buffer_matches_reference(golden):
    """store() the reference's 8-tuples one env step at a time, sample(): normalised advantage (ddof = 1) and v_target equal the
    reference ReplayBuffer_on_policy.compute_advantage outputs (utils/buffer.py:21-35)."""
    from gymrl_b200.utils.buffer import ReplayBuffer_on_policy
    g = golden("gae_utils.npz")
    T = len(g["reward"])
    buf = ReplayBuffer_on_policy(_cfg(gamma=float(g["gamma"]), lamda=float(g["lamda"])))
    rng = np.random.default_rng(0)
    states = rng.standard_normal((T, 8)).astype(np.float32)
    for t in range(T):
        buf.store((states[t], int(t % 4), float(g["reward"][t, 0]), bool(g["done"][t, 0]), bool(g["dw"][t, 0]), -0.5,
                   float(g["value"][t, 0]), float(g["next_value"][t, 0])))
    assert buf.size() == T
    s, a, logp, adv, v_target = buf.sample()
    assert s.shape == (T, 8) and a.shape == (T, 1) and a.dtype == torch.long and adv.shape == (T, 1)
    np.testing.assert_array_equal(v_target.cpu().numpy(), g["v_target"])                 # float32 recurrence: bit-exact
    np.testing.assert_allclose(adv.cpu().numpy(), g["adv_normalized"], rtol=2e-5, atol=2e-6)
    with pytest.raises(AssertionError):
        buf.store((states[0], 0, 0.0, False, False, 0.0, 0.0, 0.0))                      # store after sample (ref :11)
    buf.clear()
    assert buf.size() == 0


def test_off_policy_buffer_semantics():
    from gymrl_b200.utils.buffer import ReplayBuffer_off_policy
    buf = ReplayBuffer_off_policy(_cfg(memory_capacity=50, batch_size=16))
    for i in range(10):
        buf.store((np.full(4, i, np.float32), i % 2, float(i), np.full(4, i + 1, np.float32), i == 9))
    assert buf.size() == 10
    s, a, r, s2, d = buf.sample()
    assert s.shape == (10, 4) and r.shape == (10,)                      # min(batch_size, size) (ref :122)
    assert sorted(r.cpu().tolist()) == list(map(float, range(10)))      # without replacement (ref :124)
    np.testing.assert_array_equal(s2.cpu().numpy()[:, 0], r.cpu().numpy() + 1)
    for i in range(10, 120):
        buf.store((np.full(4, i, np.float32), 0, float(i), np.full(4, i + 1, np.float32), False))
    assert buf.size() == 50 and buf.is_full
    s, a, r, s2, d = buf.sample()
    rr = r.cpu().numpy()
    assert len(rr) == 16 and len(set(rr.tolist())) == 16 and rr.min() >= 70   # the ring kept the newest 50


def test_off_policy_buffer_keeps_item_shapes_and_vector_stores():
    """1-element action vectors come back as [B, 1] like the reference's torch.tensor(np.array(...)), image-shaped
    states keep their shape (a [C, H, W] state is ONE transition, not C env copies), and N lockstep copies store in one call."""
    from gymrl_b200.utils.buffer import ReplayBuffer_off_policy
    cfg = _cfg(memory_capacity=64, batch_size=8)
    cfg.state_shape = (3,)
    buf = ReplayBuffer_off_policy(cfg)
    for i in range(12):                                   # Pendulum-like: action shape (1,)
        buf.store((np.full(3, i, np.float32), np.array([0.5 * i], np.float32), -float(i), np.full(3, i + 1, np.float32), False))
    s, a, r, s2, d = buf.sample()
    assert s.shape == (8, 3) and a.shape == (8, 1) and r.shape == (8,) and d.shape == (8,)
    np.testing.assert_allclose(torch.cat([s, a], 1)[:, 3].cpu().numpy(), -0.5 * r.cpu().numpy())      # the reference idiom works
    cfg2 = _cfg(memory_capacity=16, batch_size=4)
    cfg2.state_shape = (2, 4, 4)
    img = ReplayBuffer_off_policy(cfg2)
    for i in range(5):
        img.store((np.full((2, 4, 4), i, np.float32), i, 1.0, np.full((2, 4, 4), i + 1, np.float32), False))
    s, a, r, s2, d = img.sample()
    assert img.size() == 5 and s.shape == (4, 2, 4, 4) and a.shape == (4,)
    assert (s.reshape(4, -1).std(dim=1) == 0).all()       # every sampled state is one stored image, not a mix
    cfg3 = _cfg(memory_capacity=100, batch_size=32)
    cfg3.state_shape, cfg3.num_envs = (4,), 16
    vec = ReplayBuffer_off_policy(cfg3)
    for t in range(5):                                    # 16 env copies per store, device tensors accepted
        st = torch.full((16, 4), float(t), device="cuda") + torch.arange(16, device="cuda")[:, None] * 0.01
        vec.store((st, torch.zeros(16, device="cuda"), torch.full((16,), float(t), device="cuda"), st + 1, torch.zeros(16, device="cuda")))
    assert vec.size() == 80
    s, a, r, s2, d = vec.sample()
    assert s.shape == (32, 4) and r.shape == (32,) and len({(float(x[0]), float(y)) for x, y in zip(s.cpu(), r.cpu())}) == 32
    np.testing.assert_allclose(np.floor(s[:, 0].cpu().numpy() + 1e-6), r.cpu().numpy())


def test_gym_view_matches_oracle_env():
    from gymrl_b200.utils import env as E
    env = E.make("CartPole-v1")
    obs, info = env.reset(seed=123)
    ora = envs_np.CartPoleVec(1, seed=123)
    o0 = ora.reset()
    np.testing.assert_allclose(obs, o0[0], rtol=1e-6, atol=1e-7)
    assert env.observation_space.shape == (4,) and env.action_space.n == 2 and env.spec.max_episode_steps == 500
    rng = np.random.default_rng(0)
    steps = 0
    for _ in range(200):
        a = int(rng.integers(2))
        obs, r, term, trunc, _ = env.step(a)
        cur, nobs, rr, te, tr = ora.step(np.array([a]))
        np.testing.assert_allclose(obs, nobs[0], rtol=1e-5, atol=1e-6)     # gymnasium returns the terminal observation itself
        assert r == float(rr[0]) and term == bool(te[0]) and trunc == bool(tr[0])
        steps += 1
        if term or trunc:
            obs, _ = env.reset()                                             # unseeded: the next episode of the same stream
            np.testing.assert_allclose(obs, cur[0], rtol=1e-5, atol=1e-6)
    env.close()
    p = E.make("Pendulum-v1")
    o, _ = p.reset(seed=1)
    assert o.shape == (3,) and p.action_space.shape == (1,) and float(p.action_space.high[0]) == 2.0
    o, r, te, tr, _ = p.step(np.array([0.5], np.float32))
    assert o.shape == (3,) and r <= 0.0 and not te
    p.close()


def test_runner_train_loop_drives_an_agent(tmp_path, monkeypatch):
    """utils.runner.BenchMark.train with a minimal on-policy agent (the agent protocol of SURVEY §1): the loop attaches
    the device normalisers, infers on_policy from the buffer class (q17), calls update() whenever the buffer holds
    batch_size transitions, evaluates and saves a checkpoint in the reference's layout."""
    monkeypatch.chdir(tmp_path)
    from gymrl_b200.utils import runner as R
    from gymrl_b200.utils.buffer import ReplayBuffer_on_policy
    from gymrl_b200.utils.model import MLP, ModelLoader

    class Config(R.BasicConfig):
        def __init__(self):
            super().__init__()
            self.env_name, self.algo_name = "CartPole-v1", "TestPPO"
            self.train_eps, self.eval_freq, self.save_freq, self.batch_size = 6, 3, 3, 32

    class Agent(ModelLoader):
        def __init__(self, cfg):
            super().__init__(cfg)
            self.net = MLP([cfg.n_states, 16, cfg.n_actions + 1]).to(cfg.device)
            self.memory = ReplayBuffer_on_policy(cfg)
            self.learn_step, self.updates = 0, []

        @torch.no_grad()
        def choose_action(self, state):
            out = self.net(torch.as_tensor(state, dtype=torch.float32, device=self.cfg.device).unsqueeze(0))[0]
            dist = torch.distributions.Categorical(logits=out[:-1])
            a = dist.sample()
            return int(a), float(dist.log_prob(a)), float(out[-1])

        @torch.no_grad()
        def evaluate(self, state):
            out = self.net(torch.as_tensor(state, dtype=torch.float32, device=self.cfg.device).unsqueeze(0))[0]
            return int(out[:-1].argmax())

        def update(self):
            s, a, logp, adv, v_target = self.memory.sample()
            assert s.shape[0] == a.shape[0] == adv.shape[0] >= self.cfg.batch_size and torch.isfinite(adv).all()
            self.updates.append(s.shape[0])
            self.memory.clear()
            self.learn_step += 1
            return {"adv_mean": float(adv.mean()), "nan_metric": float("nan")}

    agents = []
    R.BenchMark.train(lambda cfg: agents.append(Agent(cfg)) or agents[-1], Config)
    ag = agents[0]
    assert ag.cfg.on_policy is True and ag.cfg.use_rnn is False and ag.cfg.n_states == 4 and ag.cfg.n_actions == 2
    assert len(ag.updates) >= 1 and ag.state_norm.running_ms.n > 32
    ckpt = torch.load(ag.cfg.save_path, weights_only=False)
    # the normalisers are pickled plain attributes like in the reference (utils/model.py:343-345), not *_state_dict entries
    assert "net_state_dict" in ckpt and "state_norm" in ckpt and "state_norm_state_dict" not in ckpt and ckpt["learn_step"] == ag.learn_step
    ag.learn_step = -1
    ag.load_model()
    assert ag.learn_step == ckpt["learn_step"]


# ----------------------------------------------------------------------------------------------- §8f rank 4: checkpoint wire format
class _CkptCfg:
    algo_name, env_name = "PPO", "LunarLander-v3"
    device = "cuda"


def _ckpt_agent(tmp_path, monkeypatch, fused_adam):
    from gymrl_b200.nn import FlatParams, FusedAdam
    from gymrl_b200.utils.model import MLP, ModelLoader
    from gymrl_b200.utils.normalization import Normalization, RewardScaling
    monkeypatch.chdir(tmp_path)

    class Agent(ModelLoader):
        def __init__(self, cfg):
            super().__init__(cfg)
            self.net = MLP([8, 32, 4]).to("cuda")
            if fused_adam:
                self.optimizer = FusedAdam(FlatParams(self.net, device=torch.device("cuda")), lr=3e-4, eps=1e-5)
            else:
                self.optimizer = torch.optim.Adam(self.net.parameters(), lr=3e-4, eps=1e-5)
            self.state_norm = Normalization(shape=(8,))
            self.reward_scaler = RewardScaling(shape=1, gamma=0.99)
            self.learn_step = 0

    return Agent(_CkptCfg())


@pytest.mark.parametrize("fused_adam", [False, True])
def test_load_checkpoint_written_by_reference_modelloader(tmp_path, monkeypatch, golden, fused_adam):
    """tests/golden/ref_checkpoint.pth was written by the UNMODIFIED reference ModelLoader.save_model (oracle/make_golden_checkpoint.py):
    `state_norm` / `reward_scaler` are pickled reference objects, `optimizer_state_dict` is torch.optim.Adam's layout."""
    import shutil
    from pathlib import Path
    ag = _ckpt_agent(tmp_path, monkeypatch, fused_adam)
    g = golden("ref_checkpoint_expect.npz")
    shutil.copy(Path(__file__).parent / "golden" / "ref_checkpoint.pth", ag.cfg.save_path)
    ag.load_model()
    assert ag.learn_step == int(g["learn_step"])
    for k, v in ag.net.state_dict().items():
        np.testing.assert_array_equal(v.cpu().numpy(), g["w_" + k.replace(".", "_")])
    np.testing.assert_allclose(ag.net(torch.as_tensor(g["x4"], device="cuda")).detach().cpu().numpy(), g["net_out"], rtol=1e-5, atol=1e-6)
    # the pickled reference normalisers became device objects with the same statistic, and keep working
    rm = ag.state_norm.running_ms
    assert rm.n == int(g["norm_n"])
    np.testing.assert_array_equal(rm.mean.astype(np.float64), g["norm_mean"])
    np.testing.assert_allclose(rm.std, g["norm_std"], rtol=1e-15)
    np.testing.assert_allclose(rm.S, g["norm_S"], rtol=1e-15)
    np.testing.assert_array_equal(ag.state_norm(g["probe"], update=False), g["probe_normalized"].astype(np.float32))
    assert ag.reward_scaler.running_ms.n == int(g["rs_n"]) and ag.reward_scaler.gamma == 0.99
    np.testing.assert_allclose(ag.reward_scaler.R.cpu().numpy(), g["rs_R"].reshape(-1), rtol=1e-15)
    np.testing.assert_allclose(ag.reward_scaler.running_ms.std, g["rs_std"].reshape(-1), rtol=1e-15)
    # optimizer state in torch's layout, for torch.optim.Adam and for the flat FusedAdam alike
    sd = ag.optimizer.state_dict()
    assert float(sd["state"][0]["step"]) == float(g["adam_step"])
    np.testing.assert_array_equal(sd["state"][0]["exp_avg"].cpu().numpy(), g["adam_exp_avg0"])


def test_checkpoint_written_here_unpickles_as_reference_objects(tmp_path, monkeypatch):
    """The other direction: a checkpoint saved by gymrl_b200's ModelLoader names only the classes the reference's own file does,
    and — unpickled WITHOUT this package's classes (stand-ins with no __setstate__, like the reference's) — yields the reference's
    field layout: state_norm.running_ms.{n, mean, S, std} NumPy arrays, reward_scaler.{shape, gamma, running_ms, R}."""
    import pickle
    import zipfile
    ag = _ckpt_agent(tmp_path, monkeypatch, True)
    rng = np.random.default_rng(1)
    for _ in range(17):
        ag.state_norm(rng.standard_normal(8).astype(np.float32))
        ag.reward_scaler(float(rng.standard_normal()))
    ag.learn_step = 9
    ag.save_model()
    z = zipfile.ZipFile(ag.cfg.save_path)
    raw = z.read([n for n in z.namelist() if n.endswith("data.pkl")][0])
    assert b"gymrl_b200" not in raw                       # no class path of this package inside the file

    class RunningMeanStd: pass
    class Normalization: pass
    class RewardScaling: pass
    standins = {"RunningMeanStd": RunningMeanStd, "Normalization": Normalization, "RewardScaling": RewardScaling}

    class RefSideUnpickler(pickle.Unpickler):
        def find_class(self, module, name):
            if module == "utils.normalization":
                return standins[name]
            assert module.split(".")[0] in ("torch", "collections", "numpy", "_codecs"), f"unexpected global {module}.{name}"
            return super().find_class(module, name)

    class _P:   # torch.load's pickle_module protocol
        Unpickler = RefSideUnpickler
        load = staticmethod(pickle.load)
        __name__ = "pickle"

    ck = torch.load(ag.cfg.save_path, map_location="cpu", weights_only=False, pickle_module=_P)
    assert set(ck) == {"net_state_dict", "optimizer_state_dict", "state_norm", "reward_scaler", "learn_step"}
    rm = ck["state_norm"].running_ms
    assert isinstance(ck["state_norm"], Normalization) and isinstance(rm, RunningMeanStd)
    assert rm.n == 17 and rm.mean.shape == (8,) and rm.mean.dtype == np.float32 and rm.S.dtype == np.float64 and rm.std.shape == (8,)
    np.testing.assert_array_equal(rm.mean, ag.state_norm.running_ms.mean)
    rs = ck["reward_scaler"]
    assert rs.gamma == 0.99 and rs.shape == 1 and rs.running_ms.n == 17 and isinstance(rs.R, np.ndarray)
    assert set(ck["optimizer_state_dict"]) == {"state", "param_groups"} and ck["learn_step"] == 9
    # and torch.optim.Adam (what the reference agent holds) accepts the optimizer entry
    from gymrl_b200.utils.model import MLP
    ref_net = MLP([8, 32, 4])
    ref_net.load_state_dict(ck["net_state_dict"])
    torch.optim.Adam(ref_net.parameters(), lr=1e-3).load_state_dict(ck["optimizer_state_dict"])


def test_runner_train_vectorised(tmp_path, monkeypatch):
    """utils.runner.train with cfg.num_envs = 64: the reference's loop over N lockstep copies on the device — device normalisers
    (one RewardScaling accumulator per copy), [N]-row buffer stores, one update() per filled buffer, episode accounting."""
    monkeypatch.chdir(tmp_path)
    from gymrl_b200.utils import runner as R
    from gymrl_b200.utils.buffer import ReplayBuffer_on_policy
    from gymrl_b200.utils.model import MLP, ModelLoader

    class Config(R.BasicConfig):
        def __init__(self):
            super().__init__()
            self.env_name, self.algo_name = "CartPole-v1", "TestVecPPO"
            self.num_envs, self.train_eps, self.batch_size, self.seed = 64, 200, 2048, 1
            self.n_states, self.n_actions = 4, 2          # make_env is not needed for the device env

    class Agent(ModelLoader):
        def __init__(self, cfg):
            super().__init__(cfg)
            self.net = MLP([cfg.n_states, 16, cfg.n_actions + 1]).to("cuda")
            self.memory = ReplayBuffer_on_policy(cfg)
            self.learn_step, self.updates = 0, []

        @torch.no_grad()
        def choose_action(self, state):
            out = self.net(state)
            dist = torch.distributions.Categorical(logits=out[:, :-1])
            a = dist.sample()
            return a, dist.log_prob(a), out[:, -1]

        def update(self):
            s, a, logp, adv, v_target = self.memory.sample()
            assert s.shape[0] == a.shape[0] == adv.shape[0] >= self.cfg.batch_size and torch.isfinite(adv).all()
            assert s.shape[0] % 64 == 0
            self.updates.append(s.shape[0])
            self.memory.clear()
            self.learn_step += 1
            return {"adv_mean": float(adv.mean())}

    cfg = Config()
    ag = Agent(cfg)
    env = R.train(None, ag, cfg)
    assert cfg.on_policy is True and len(ag.updates) >= 1 and all(u == 2048 for u in ag.updates)
    avg, _, total = env.episode_stats(100)
    assert total >= 200 and 8.0 < avg < 200.0                       # a random CartPole policy lasts ~20 steps
    assert ag.state_norm.running_ms.n >= 2048 and ag.reward_scaler.R.numel() == 64
    ckpt = torch.load(cfg.save_path, weights_only=False)
    assert "net_state_dict" in ckpt and "state_norm" in ckpt
