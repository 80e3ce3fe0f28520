"""Device-time split of one PPO iteration on the bench workload (C2: 4096 envs x 128 steps, 10 epochs x 32 mb):
rollout graph vs GAE vs update, plus the env step alone and the policy forward alone.  CUDA events, 5 repeats.

    python tools/phase_times.py [--envs 4096] [--steps 128]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def timed(fn, reps=5):
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=128)
    a = ap.parse_args()
    from gymrl_b200 import ops
    from gymrl_b200.algorithms import ppo_lunarlander as P
    cfg = P.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.num_epochs = a.envs, a.steps, 32, 10
    tr = P.PPOTrainer(cfg)
    for _ in range(3):
        tr.collect_rollout()
        tr.update(None, read_metrics=False)
    out = {}
    out["rollout_ms"] = timed(tr.collect_rollout)
    out["update_ms"] = timed(lambda: tr.update(None, read_metrics=False))
    # env step alone, with the actions of the last rollout
    env = tr.env
    acts = tr.buffer.action
    obs = torch.empty_like(tr.buffer.obs[0])
    rew = torch.empty(a.envs, device="cuda")
    te = torch.empty(a.envs, dtype=torch.uint8, device="cuda")
    tu = torch.empty_like(te)
    dn = torch.empty_like(te)

    def env_only():
        for t in range(a.steps):
            env.step(acts[t], obs=obs, reward=rew, terminated=te, truncated=tu, want_next_obs=False, done=dn)
    out["env_steps_only_ms"] = timed(env_only)
    x = tr.buffer.obs[0]
    out["policy_forward_x_steps_ms"] = timed(lambda: [tr.net.forward(x, tr.acts_roll, a.envs) for _ in range(a.steps)])
    A = env.n_actions

    def sample_only():
        for t in range(a.steps):
            ops.sample_categorical(tr.acts_roll.lv[:, :A], seed=1, first_id=0, draw_base=tr.ctr_action, action=tr.buffer.action[t],
                                   logp=tr.buffer.log_prob[t], value_in=tr.acts_roll.lv[:, A:A + 1], value_out=tr.buffer.value[t])
            ops.counter_add(tr.ctr_action, 1)
    out["sample_x_steps_ms"] = timed(sample_only)
    # one minibatch: forward+loss+backward vs optimizer
    tr.ctr_mb.zero_()
    out["mb_fwd_bwd_ms"] = timed(lambda: (tr.ctr_mb.zero_(), tr._fwd_bwd_body()))
    out["mb_opt_ms"] = timed(tr._opt_body)
    M = tr.mb
    obs_flat = tr.buffer.obs[:tr.T].view(tr.T * tr.N, -1)
    out["mb_fwd_ms"] = timed(lambda: tr.net.forward(obs_flat, tr.acts_mb, M, row_index=tr.idx_mb))
    out["mb_bwd_ms"] = timed(lambda: tr.net.backward(obs_flat, tr.acts_mb, M, row_index=tr.idx_mb))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
