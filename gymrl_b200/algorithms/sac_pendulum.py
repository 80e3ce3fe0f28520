"""SAC for Pendulum-v1 on the B200 engine — same surface as the reference ``algorithms/sac_pendulum.py``
(Config, Actor, Critic, ReplayBuffer, SACTrainer.train/eval/test/update/select_action/soft_update, .alpha).

    select_action (ref :201-211) -> actor trunk GEMMs + fused [mean | log_std] head + tanh-Gaussian sample kernel, all N envs
    update        (ref :213-267) -> on-device sample (without replacement) -> target: actor(s') sample, twin target critic,
                                    y kernel -> critic fwd/loss/bwd/Adam -> actor fwd + sample -> critic fwd (input-grad only)
                                    -> tanh-Gaussian reparameterisation gradient kernel -> actor bwd/Adam -> float64 alpha
                                    Adam step on device -> Polyak kernel over the flat critic buffer.
The reference back-propagates the actor loss into the critic's parameter gradients too and discards them;
here only the critic's *input* gradient is computed for that pass.  One update of ``batch_size`` per lockstep.
"""
from __future__ import annotations

import time
from collections import deque

import numpy as np
import torch
import torch.nn as nn

from .. import _ffi, ops, ops_offpolicy as off
from ..mlp import Chain
from ..nn import FlatParams, FusedAdam

f32, i32, u8, f64 = torch.float32, torch.int32, torch.uint8, torch.float64
RELU, NONE = _ffi.ACT_RELU, _ffi.ACT_NONE


class Config:
    def __init__(self):
        self.env_name = "Pendulum-v1"
        self.seed = None
        self.max_episodes = 500
        self.max_steps = 200
        self.batch_size = 128
        self.gamma = 0.99
        self.lr_actor = 3e-4
        self.lr_critic = 3e-4
        self.lr_alpha = 3e-4
        self.tau = 0.005
        self.init_alpha = 0.2
        self.memory_capacity = 100000
        self.hidden_dim = 256
        self.log_std_min = -20
        self.log_std_max = 2
        self.device = "cuda"
        # ---- engine extras ----
        self.num_envs = 1
        self.max_locksteps = None
        self.use_cuda_graph = True   # train(): act -> env step -> store -> update as ONE captured graph per lockstep


class Actor(nn.Module):
    def __init__(self, state_dim, action_dim, hidden_dim, action_bound, log_std_min, log_std_max):
        super().__init__()
        self.action_bound, self.log_std_min, self.log_std_max = action_bound, log_std_min, log_std_max
        self.fc1 = nn.Linear(state_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.mean = nn.Linear(hidden_dim, action_dim)
        self.log_std = nn.Linear(hidden_dim, action_dim)

    # mean and log_std heads are adjacent in the flat buffer -> one [2A, H] GEMM
    PARAM_ORDER = ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "mean.weight", "log_std.weight", "mean.bias", "log_std.bias"]


class Critic(nn.Module):
    def __init__(self, state_dim, action_dim, hidden_dim):
        super().__init__()
        self.fc1 = nn.Linear(state_dim + action_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)
        self.fc3 = nn.Linear(hidden_dim, 1)
        self.fc4 = nn.Linear(state_dim + action_dim, hidden_dim)
        self.fc5 = nn.Linear(hidden_dim, hidden_dim)
        self.fc6 = nn.Linear(hidden_dim, 1)

    Q1 = [("fc1.weight", "fc1.bias", RELU), ("fc2.weight", "fc2.bias", RELU), ("fc3.weight", "fc3.bias", NONE)]
    Q2 = [("fc4.weight", "fc4.bias", RELU), ("fc5.weight", "fc5.bias", RELU), ("fc6.weight", "fc6.bias", NONE)]


def _actor_chain(fp: FlatParams, A: int, H: int, M: int, backward: bool) -> Chain:
    head_w = fp.span("mean.weight", "log_std.weight", 2 * A, H)
    head_b = fp.span("mean.bias", "log_std.bias", 1, 2 * A).view(2 * A)
    head_gw = fp.span("mean.weight", "log_std.weight", 2 * A, H, grad=True)
    head_gb = fp.span("mean.bias", "log_std.bias", 1, 2 * A, grad=True).view(2 * A)
    return Chain(fp, [(fp.p("fc1.weight"), fp.p("fc1.bias"), fp.g("fc1.weight"), fp.g("fc1.bias"), RELU),
                      (fp.p("fc2.weight"), fp.p("fc2.bias"), fp.g("fc2.weight"), fp.g("fc2.bias"), RELU),
                      (head_w, head_b, head_gw, head_gb, NONE)], M, backward)


class ReplayBuffer(off.ReplayRing):
    def __init__(self, capacity: int, obs_dim: int = 3, act_dim: int = 1, device=None):
        super().__init__(capacity, obs_dim, act_dim, False, device or torch.device("cuda", torch.cuda.current_device()))


class SACTrainer:
    def __init__(self, config: Config):
        _ffi.require_cuda()
        self.cfg = cfg = config
        self.device = dev = torch.device("cuda", torch.cuda.current_device())
        self.N = N = int(cfg.num_envs)
        self.seed = int(cfg.seed) if cfg.seed is not None else int(time.time_ns() & 0x7FFFFFFF)
        self.env = ops.VecEnv(cfg.env_name, N, seed=self.seed)
        D, A, H, B = self.env.obs_dim, self.env.act_dim, cfg.hidden_dim, int(cfg.batch_size)
        self.D, self.A, self.B = D, A, B
        self.action_bound = float(self.env.action_bound)
        self.actor = Actor(D, A, H, self.action_bound, cfg.log_std_min, cfg.log_std_max).to(dev)
        self.critic = Critic(D, A, H).to(dev)
        self.critic_target = Critic(D, A, H).to(dev)
        self.critic_target.load_state_dict(self.critic.state_dict())
        self.fp_a = FlatParams(self.actor, Actor.PARAM_ORDER, dev)
        self.fp_c = FlatParams(self.critic, device=dev)
        self.fp_ct = FlatParams(self.critic_target, device=dev)
        self.actor_optimizer = FusedAdam(self.fp_a, lr=cfg.lr_actor)
        self.critic_optimizer = FusedAdam(self.fp_c, lr=cfg.lr_critic)
        self.target_entropy = -A
        self.log_alpha = torch.tensor([np.log(cfg.init_alpha)], device=dev, dtype=f64)   # float64 like the reference (q9)
        self.alpha_state = torch.zeros(3, device=dev, dtype=f64)
        self.pi_act = _actor_chain(self.fp_a, A, H, N, False)
        self.pi_upd = _actor_chain(self.fp_a, A, H, B, True)
        self.q1 = Chain.from_names(self.fp_c, Critic.Q1, B, True)
        self.q2 = Chain.from_names(self.fp_c, Critic.Q2, B, True)
        self.q1t = Chain.from_names(self.fp_ct, Critic.Q1, B, False)
        self.q2t = Chain.from_names(self.fp_ct, Critic.Q2, B, False)
        self.memory = ReplayBuffer(cfg.memory_capacity, D, A, dev)
        from ..graphs import Branches
        # update() is recorded as a DAG: an outer three-way fork and two nested two-way forks (the twin critics)
        self.br_outer, self.branches, self.branches_b = Branches(2), Branches(1), Branches(1)
        self.pi_nxt = _actor_chain(self.fp_a, A, H, B, False)      # the target branch's own forward scratch
        z = lambda *s, dt=f32: torch.zeros(*s, device=dev, dtype=dt)
        self.idx = z(B, dt=i32)
        self.sa, self.sa2 = z(B, D + A), z(B, D + A)
        self.act_b, self.logp_b, self.pre_b, self.noise_b = z(B, A), z(B), z(B, A), z(B, A)
        self.act_n, self.logp_n, self.noise_n, self.sa2n = z(B, A), z(B), z(B, A), z(B, D + A)   # the target branch's copies
        self.y = z(B)
        self.acc = z(4)            # [0] alpha*mean(logpi)  [1] sum(logpi)  [2] -mean(minQ)  [3] alpha loss
        self.closs = z(2)
        self.action = z(N, A)
        self.done = z(N, dt=u8)
        self.update_count = 0
        self.act_count = 0
        # RNG draw counters on the device (the kernels add them to their `draw` argument), so that a captured lockstep
        # draws fresh noise on every replay: act (+1 per act), sample (+1 per update), normals (+2 per update)
        self.ctr_act, self.ctr_upd, self.ctr_nz = z(1, dt=i32), z(1, dt=i32), z(1, dt=i32)
        self.cur = z(N, D)
        self._g_lockstep = None
        self.graph_launches = 0
        self.episode_rewards = deque(maxlen=100)
        print(f"Device: {dev}")
        print(f"State dim: {D}, Action dim: {A}")
        print(f"Action bound: [-{self.action_bound}, {self.action_bound}]")
        print(f"Target entropy: {self.target_entropy}")

    @property
    def alpha(self) -> torch.Tensor:
        return self.log_alpha.exp()

    def soft_update(self, target=None, source=None):
        ops.polyak(self.fp_ct.flat, self.fp_c.flat, self.cfg.tau)

    def act(self, obs: torch.Tensor, deterministic: bool = False) -> torch.Tensor:
        A = self.A
        out = self.pi_act.forward(obs, self.N)
        self.act_count += 1
        ops.sample_tanh_gaussian(out[:, :A], out[:, A:], self.action_bound, self.cfg.log_std_min, self.cfg.log_std_max,
                                 seed=self.seed, draw=1, draw_base=self.ctr_act, deterministic=deterministic, action=self.action,
                                 want_logp=False)
        ops.counter_add(self.ctr_act, 1)
        return self.action

    @torch.no_grad()
    def select_action(self, state: np.ndarray, deterministic: bool = False) -> np.ndarray:
        obs = torch.as_tensor(np.asarray(state, np.float32), device=self.device).reshape(1, -1)
        chain = getattr(self, "_pi_one", None) or _actor_chain(self.fp_a, self.A, self.cfg.hidden_dim, 1, False)
        self._pi_one = chain
        out = chain.forward(obs, 1)
        self.act_count += 1
        a, _ = ops.sample_tanh_gaussian(out[:, :self.A], out[:, self.A:], self.action_bound, self.cfg.log_std_min,
                                        self.cfg.log_std_max, seed=self.seed, first_id=1 << 40, draw=self.act_count,
                                        deterministic=deterministic, want_logp=False)
        return a[0].cpu().numpy()

    def update(self, idx: torch.Tensor = None, noise_next: torch.Tensor = None, noise_new: torch.Tensor = None):
        """One SAC update.  idx / noise_* let the parity tests feed the reference's own sample and N(0,1) draws."""
        cfg, B, A, D, mem = self.cfg, self.B, self.A, self.D, self.memory
        if len(mem) < B:
            return 0.0, 0.0, 0.0
        self.update_count += 1
        if idx is None:
            idx = mem.sample_indices(B, seed=self.seed, draw=1, draw_base=self.ctr_upd, out=self.idx)
        lsmin, lsmax, bound = cfg.log_std_min, cfg.log_std_max, self.action_bound
        br, tw_a, tw_b = self.br_outer, self.branches, self.branches_b
        # The update as a DAG (recorded as parallel branches of the lockstep graph; eagerly: concurrent streams).  Phase 1: the
        # target y, the critics' forward on (s, a) and the actor's forward on s with its fresh action are mutually independent
        # (each owns its scratch); inside a branch the twin critics fork again.

        def target_branch():        # ---- target (ref :233-237) ----
            out = self.pi_nxt.forward(mem.next_obs, B, row_index=idx)
            nz = noise_next if noise_next is not None else off.fill_normal(self.noise_n, seed=self.seed, entity0=0, draw=2, draw_base=self.ctr_nz)
            ops.sample_tanh_gaussian(out[:, :A], out[:, A:], bound, lsmin, lsmax, nz, action=self.act_n, logp=self.logp_n)
            off.gather_concat(mem.next_obs, idx, self.act_n, None, out=self.sa2n, n=B)
            q1t, q2t = tw_a.run(lambda: self.q1t.forward(self.sa2n, B), lambda: self.q2t.forward(self.sa2n, B))
            off.twin_q_target(mem.reward, mem.done, q1t, q2t, cfg.gamma, row_index=idx, logp_next=self.logp_n, log_alpha=self.log_alpha,
                              out=self.y)

        def critic_forward_branch():   # ---- critic forward (ref :239-241) ----
            off.gather_concat(mem.obs, idx, mem.action, idx, out=self.sa, n=B)
            return tw_b.run(lambda: self.q1.forward(self.sa, B), lambda: self.q2.forward(self.sa, B))

        def actor_prep_branch():    # ---- new action of the current policy (ref :248-249): does not depend on the critics ----
            out = self.pi_upd.forward(mem.obs, B, row_index=idx)
            nz = noise_new if noise_new is not None else off.fill_normal(self.noise_b, seed=self.seed, entity0=0, draw=3, draw_base=self.ctr_nz)
            ops.sample_tanh_gaussian(out[:, :A], out[:, A:], bound, lsmin, lsmax, nz, action=self.act_b, logp=self.logp_b,
                                     pre_tanh=self.pre_b)
            off.gather_concat(mem.obs, idx, self.act_b, None, out=self.sa2, n=B)
            return out, nz

        (q1, q2), _, (out, nz) = br.run(critic_forward_branch, target_branch, actor_prep_branch)
        # ---- critic loss / backward / step (ref :242-246) ----
        self.closs.zero_()
        off.twin_q_loss(q1, q2, self.y, self.q1.dout, self.q2.dout, self.closs)
        tw_a.run(lambda: self.q1.backward(self.sa, B), lambda: self.q2.backward(self.sa, B))   # the twins are independent
        self.critic_optimizer.step()

        # ---- actor (ref :250-255) through the UPDATED critics; the target sync (ref :265) only reads them: a side branch ----
        def actor_branch():
            q1n, q2n = tw_a.run(lambda: self.q1.forward(self.sa2, B), lambda: self.q2.forward(self.sa2, B))
            self.acc.zero_()
            off.min_q_grad(q1n, q2n, self.q1.dout, self.q2.dout, acc=self.acc[2:3])
            dx1, dx2 = tw_a.run(lambda: self.q1.backward(self.sa2, B, param_grads=False, input_grad=True),
                                lambda: self.q2.backward(self.sa2, B, param_grads=False, input_grad=True))
            dx1.add_(dx2)                                       # d(-minQ/B)/d[s, a]   (torch add: tiny [B, D+A] glue)
            dhead = self.pi_upd.dout
            off.sac_actor_grad(self.pre_b, nz, out[:, A:], dx1[:, D:], self.log_alpha, bound, lsmin, lsmax, dhead[:, :A], dhead[:, A:],
                               self.logp_b, self.acc)

            def actor_step():
                self.pi_upd.backward(mem.obs, B, row_index=idx)
                self.actor_optimizer.step()

            def alpha_and_counters():   # ---- alpha (ref :257-263): needs only the log-prob sums of sac_actor_grad ----
                off.sac_alpha_step(self.log_alpha, self.alpha_state, self.acc, B, self.target_entropy, cfg.lr_alpha, loss_out=self.acc[3:4])
                ops.counter_add(self.ctr_upd, 1)
                ops.counter_add(self.ctr_nz, 2)
            tw_b.run(actor_step, alpha_and_counters)

        br.run(actor_branch, self.soft_update)
        return self.acc, self.closs

    # ---------------------------------------------------------------- one lockstep (ref train() loop body :278-294)
    def _lockstep_body(self):
        """act -> env.step -> store -> update -> carry the observation; capture-safe once the replay holds a batch."""
        env, mem, cur = self.env, self.memory, self.cur
        a = self.act(cur)
        obs, r, te, tr, nobs = env.step(a, done=self.done)
        mem.store(cur, a, r, nobs, self.done)      # done = terminated | truncated (ref :281-283)
        self.update()
        cur.copy_(obs)

    def lockstep(self):
        """One lockstep of all N envs.  Eager until the replay buffer holds a batch (the update is a host-side no-op before
        that), then one CUDA-graph replay per lockstep (cfg.use_cuda_graph)."""
        from ..graphs import capture
        if not getattr(self.cfg, "use_cuda_graph", True) or len(self.memory) < self.B:
            return self._lockstep_body()
        if self._g_lockstep is None:
            mirrors = (self.act_count, self.update_count, self.memory._size_host)
            self._g_lockstep = capture(self._lockstep_body)   # the capture's eager pass is this lockstep ...
            self.act_count, self.update_count = mirrors[0] + 1, mirrors[1] + 1   # ... the recording pass is not
            self.memory._size_host = min(self.memory.capacity, mirrors[2] + self.N)
            return
        self._g_lockstep.replay()
        self.graph_launches += self._g_lockstep.n_kernels
        # host mirrors of what the replayed body would have counted
        self.act_count += 1
        self.update_count += 1
        self.memory._size_host = min(self.memory.capacity, self.memory._size_host + self.N)

    def losses(self):
        """(actor_loss, critic_loss, alpha_loss) as floats — one D2H, call at log time."""
        a, c = self.acc.tolist(), self.closs.tolist()
        return a[0] + a[2], c[0], a[3]

    def train(self):
        print("Starting training...")
        cfg, env, mem = self.cfg, self.env, self.memory
        env.reset(out=self.cur)
        max_lock = cfg.max_locksteps or int(cfg.max_episodes * cfg.max_steps / self.N)
        t0, last_total = time.time(), 0
        for step in range(max_lock):
            self.lockstep()
            if step % cfg.max_steps == cfg.max_steps - 1:
                avg, _, total = env.episode_stats(100)
                if total != last_total:
                    last_total = total
                    self.episode_rewards.extend([avg] * min(self.N, 100))
                    al, cl, el = self.losses()
                    sps = (step + 1) * self.N / max(time.time() - t0, 1e-9)
                    print(f"Episodes {total} | Avg(100): {avg:.1f} | Alpha: {float(self.alpha):.4f} | Actor: {al:.3f} | "
                          f"Critic: {cl:.3f} | {sps:,.0f} steps/s")
                    if avg >= -200.0 and total >= 100:
                        print(f"\nEnvironment solved in {total} episodes!")
                        break
        print("Training completed!")

    def eval(self, num_episodes: int = 10):
        print(f"\nEvaluating for {num_episodes} episodes...")
        env = ops.VecEnv(self.cfg.env_name, num_episodes, seed=self.seed + 999, first_env_id=1 << 32)
        chain = _actor_chain(self.fp_a, self.A, self.cfg.hidden_dim, num_episodes, False)
        obs = env.reset()
        ret = torch.zeros(num_episodes, device=self.device, dtype=f64)
        for _ in range(env.max_episode_steps):
            out = chain.forward(obs, num_episodes)
            a, _ = ops.sample_tanh_gaussian(out[:, :self.A], out[:, self.A:], self.action_bound, deterministic=True, want_logp=False)
            obs, r, te, tr, _ = env.step(a, want_next_obs=False)
            ret += r.double()
        rewards = ret.tolist()
        for i, r in enumerate(rewards):
            print(f"  Episode {i + 1}: Reward = {r:.1f}")
        print(f"Evaluation Results: Mean = {np.mean(rewards):.1f} +/- {np.std(rewards):.1f}")
        env.close()
        return rewards

    def test(self):
        self.eval(num_episodes=5)
        print("\n(visual test skipped: the device env has no renderer)")


if __name__ == "__main__":
    config = Config()
    config.num_envs = 4096
    config.batch_size = 4096
    config.memory_capacity = 1 << 20
    trainer = SACTrainer(config)
    trainer.train()
    trainer.test()
