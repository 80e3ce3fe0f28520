"""Warm per-stage timing of one PPO minibatch (C2 shape): every stage captured `reps` times back to back in a CUDA graph and
replayed between two CUDA events (the ncu launch list is cold-cache and serialised; this is the steady-state view).

    python tools/mb_stage_times.py [reps]
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    from gymrl_b200 import _ffi, ops
    from gymrl_b200.algorithms import ppo_lunarlander as P
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    cfg = P.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.num_epochs, cfg.use_cuda_graph, cfg.seed = 4096, 128, 32, 1, False, 1
    tr = P.PPOTrainer(cfg)
    tr.collect_rollout()
    tr.update(None, read_metrics=False)
    net, acts, buf, M = tr.net, tr.acts_mb, tr.buffer, tr.mb
    T, H, A, ws = _ffi.ACT_TANH, net.H, net.A, net.ws_layers
    obs = buf.obs[:tr.T].view(tr.T * tr.N, -1)
    idx = tr.perm[:M]
    flat = buf.action.view(-1), buf.log_prob.view(-1), buf.adv.view(-1), buf.ret.view(-1)

    def timed(fn):
        # `reps` launches captured in one CUDA graph: no host launch cost between them (a ctypes call is ~10 us of Python)
        g = tr._capture(lambda: [fn() for _ in range(reps)])
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) * 1e3 / reps, 2)

    def deferred_nofold(fn):
        # reps calls inside ONE deferral scope would overflow the pending table: flush after each, fold is then 1 launch
        def run():
            ops.reduce_defer_begin()
            fn()
            ops.reduce_flush()
        return run

    def deferred(fn):
        def run():
            ops.reduce_defer_begin()
            fn()
            ops.reduce_flush()
        return run

    out = {}
    only_l1 = len(sys.argv) > 2 and sys.argv[2] == "l1"
    if only_l1:
        out["L1 fwd (gather, K=8)"] = timed(lambda: ops.linear_forward(obs, net.W1, net.b1, T, row_index=idx, out=acts.h1, M=M))
        out["L1 dW (gather) excl. fold"] = timed(deferred_nofold(lambda: ops.linear_backward(acts.dh1, obs, net.W1, net.gW1, net.gb1, row_index=idx, workspace=ws[4], M=M)))
        out["whole minibatch (graph replay)"] = timed(lambda: tr._minibatch_body(0))
        print(json.dumps(out))
        return
    out["L1 fwd (gather, K=8)"] = timed(lambda: ops.linear_forward(obs, net.W1, net.b1, T, row_index=idx, out=acts.h1, M=M))
    out["L2 fwd 256x256"] = timed(lambda: ops.linear_forward(acts.h1, net.W2, net.b2, T, out=acts.h2, M=M))
    out["Lac fwd 512x256"] = timed(lambda: ops.linear_forward(acts.h2, net.Wac, net.bac, T, out=acts.ac, M=M))
    out["heads+loss+bwd (fused) incl. fold"] = timed(lambda: net.heads_loss_backward(acts, M, *flat, tr.loss_cfg, row_index=idx, metrics=tr.metrics))
    out["Lac bwd (dW+dX) incl. fold"] = timed(lambda: ops.linear_backward(acts.dac, acts.h2, net.Wac, net.gWac, net.gbac, dx=acts.dh2, act_in=T, workspace=ws[2], M=M))
    out["Lac dW only incl. fold"] = timed(lambda: ops.linear_backward_weight(acts.dac, acts.h2, net.gWac, net.gbac, workspace=ws[2], M=M))
    out["Lac dX only"] = timed(lambda: ops.linear_backward_input(acts.dac, net.Wac, acts.h2, T, out=acts.dh2))
    out["L2 bwd (dW+dX) incl. fold"] = timed(lambda: ops.linear_backward(acts.dh2, acts.h1, net.W2, net.gW2, net.gb2, dx=acts.dh1, act_in=T, workspace=ws[3], M=M))
    out["L1 dW (gather) incl. fold"] = timed(lambda: ops.linear_backward(acts.dh1, obs, net.W1, net.gW1, net.gb1, row_index=idx, workspace=ws[4], M=M))
    out["grad_sumsq + adam + post"] = timed(tr.optimizer.launch if False else (lambda: tr.optimizer.launch(max_norm=0.5)))
    tr.ctr_mb.zero_()
    out["whole minibatch (eager, static window)"] = timed(lambda: tr._minibatch_body(0))
    out["whole minibatch (graph replay)"] = out["whole minibatch (eager, static window)"]
    del out["whole minibatch (eager, static window)"]
    out["sum of stages"] = round(sum(v for k, v in out.items() if "only" not in k and "whole" not in k), 2)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
