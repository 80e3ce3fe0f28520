"""Warp-specialised TMA / 2-CTA GEMM (gemm3x_ws_kernel) vs the one-tile-per-CTA register-split kernel at the C2 update's shapes.
CUDA events on the launching stream, operand sets larger than L2 (6 x (16 + 32 + ...) MB), weights registered / unregistered.

    python tools/ws_gemm_bench.py [--M 16384]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gymrl_b200 import ops  # noqa: E402


def timeit(fn, sets, reps=5):
    """One CUDA graph holding a pass over all operand sets (as the trainer's epoch graph does: no host launch cost between the
    kernels — the ws path encodes a tensor map per launch on the host), replayed `reps` times between two events."""
    for i in range(sets):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(sets):
            fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(sets):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * sets)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=16384)
    a = ap.parse_args()
    M, sets = a.M, 6
    rows = []
    for (N, K) in [(512, 256), (256, 256)]:
        xs = [torch.randn(M, K, device="cuda") for _ in range(sets)]
        hs = [torch.tanh(torch.randn(M, K, device="cuda")) for _ in range(sets)]
        dys = [torch.randn(M, N, device="cuda") / M for _ in range(sets)]
        ys = [torch.empty(M, N, device="cuda") for _ in range(sets)]
        dxs = [torch.empty(M, K, device="cuda") for _ in range(sets)]
        w, b = torch.randn(N, K, device="cuda") / 16, torch.zeros(N, device="cuda")
        flat = torch.zeros(N * K + 8, device="cuda")
        flat[4:4 + N * K] = w.reshape(-1)
        wv = flat[4:4 + N * K].view(N, K)
        fl = 2.0 * M * N * K
        for engine in ("register-split (r1)", "ws TMA/2-CTA (r2)"):
            if engine.startswith("ws"):
                img = ops.weight_images_register(flat, [(4, N, K)])
                W = wv
            else:
                W = w
            t_f = timeit(lambda i: ops.linear_forward(xs[i], W, b, 1, out=ys[i]), sets)
            t_b = timeit(lambda i: ops.linear_backward_input(dys[i], W, hs[i], 1, out=dxs[i]), sets)
            if not engine.startswith("ws"):
                dw, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
                wsb = torch.empty(ops.backward_weight_workspace(M, N, K), device="cuda", dtype=torch.uint8)
                t_w = timeit(lambda i: ops.linear_backward_weight(dys[i], hs[i], dw, db, workspace=wsb), sets)
                print(json.dumps({"dW+db (TS kernel + fold)": True, "N": N, "K": K, "us": round(t_w, 2), "TFLOPs": round(fl / t_w / 1e6, 1)}), flush=True)
            if engine.startswith("ws"):
                ops.weight_images_unregister(flat)
            rows.append({"engine": engine, "M": M, "N": N, "K": K, "fwd+tanh_us": round(t_f, 2), "fwd_TFLOPs": round(fl / t_f / 1e6, 1),
                         "dX*tanh'_us (MxK out, reduce N)": round(t_b, 2), "dX_TFLOPs": round(fl / t_b / 1e6, 1)})
            print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
