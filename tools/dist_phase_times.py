"""Where a multi-GPU PPO iteration goes (run under torch.distributed.run, one rank per GPU): rollout / update split per rank
with CUDA events, for the three gradient paths — peer-memory reduction (csrc/comm.cu), ncclAllReduce, and no reduction at all
(profiling only: isolates rank skew and the unfused optimizer from the collective itself) — plus %globaltimer stamps from
inside the peer-reduce kernel (entry -> published -> peers arrived -> reduced).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/dist_phase_times.py
"""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def run(mode, iters=5):
    from gymrl_b200 import _ffi
    from gymrl_b200.algorithms import ppo_lunarlander as P
    cfg = P.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.num_epochs, cfg.seed, cfg.max_train_steps = 4096, 128, 32, 10, 0, 10 ** 12
    cfg.peer_reduce = mode == "peer"
    cfg._profile_skip_reduce = mode == "none"
    torch.manual_seed(0)
    if mode == "solo":      # every rank runs the single-GPU code path on its own GPU (no collectives at all)
        import gymrl_b200.dist as gd
        real = gd.info
        gd.info = lambda: (0, 1)
        try:
            tr = P.PPOTrainer(cfg)
        finally:
            gd.info = real
    else:
        tr = P.PPOTrainer(cfg)
    stamps = None
    if tr.comm is not None:
        stamps = torch.zeros(8, device="cuda", dtype=torch.int64)
        lib = _ffi.load()
        lib.gymrl_debug_comm_stamps.argtypes = [C.c_void_p, C.c_void_p]
        lib.gymrl_debug_comm_stamps(tr.comm._h, stamps.data_ptr())
    for _ in range(3):
        tr.collect_rollout(); tr.update(None, read_metrics=False)
    dist.barrier(); torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(iters)]
    for k in range(iters):
        ev[k][0].record(); tr.collect_rollout(); ev[k][1].record(); tr.update(None, read_metrics=False); ev[k][2].record()
    dist.barrier(); torch.cuda.synchronize()
    ro = sorted(e[0].elapsed_time(e[1]) for e in ev)[iters // 2]
    up = sorted(e[1].elapsed_time(e[2]) for e in ev)[iters // 2]
    out = {"mode": mode, "rank": dist.get_rank(), "rollout_ms": round(ro, 2), "update_ms": round(up, 2), "per_minibatch_us": round(up * 1e3 / 320, 1)}
    if stamps is not None:
        s = stamps.tolist()
        out["peer_kernel_block0_us"] = {"publish": (s[1] - s[0]) / 1e3, "wait_peers": (s[2] - s[1]) / 1e3, "reduce": (s[3] - s[2]) / 1e3}
    return out


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for mode in (sys.argv[1:] or ["peer", "nccl", "none", "solo"]):
        r = run(mode)
        allr = [None] * dist.get_world_size()
        dist.all_gather_object(allr, r)
        if dist.get_rank() == 0:
            for x in allr:
                print(json.dumps(x), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
