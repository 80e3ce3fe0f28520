"""Micro-benchmark of the dense-layer kernels at the shapes of the C2 update (CUDA events, operand sets > L2).

    python tools/gemm_bench.py            # prints one line per (product, shape, engine)
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gymrl_b200 import _ffi, ops  # noqa: E402


def timeit(fn, sets, reps=4):
    for i in range(sets):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i in range(sets):
            fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * sets)


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default=None, help="N,K: only this shape")
    ap.add_argument("--product", default=None, help="only this product: fwd+tanh | fwd | dX*act | dW+db")
    ap.add_argument("--engine", default=None, help="ffma | tc")
    ap.add_argument("--M", type=int, default=16384)
    a = ap.parse_args()
    lib = _ffi.load()
    M = a.M
    sets = 5
    shapes = [(256, 256), (512, 256)]          # (N_out, K_in)
    if a.shape:
        shapes = [tuple(int(v) for v in a.shape.split(","))]
    want = lambda prod: a.product is None or a.product == prod
    print(f"{'product':10s} {'M':>6s} {'N':>4s} {'K':>4s} {'engine':>6s} {'ms':>8s} {'TFLOP/s':>8s}")
    for N, K in shapes:
        xs = [torch.randn(M, K, device="cuda") for _ in range(sets)]
        hs = [torch.tanh(torch.randn(M, K, device="cuda")) for _ in range(sets)]
        dys = [torch.randn(M, N, device="cuda") / M for _ in range(sets)]
        ys = [torch.empty(M, N, device="cuda") for _ in range(sets)]
        dxs = [torch.empty(M, K, device="cuda") for _ in range(sets)]
        w, b = torch.randn(N, K, device="cuda") / 16, torch.zeros(N, device="cuda")
        dw, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
        ws = torch.empty(ops.backward_weight_workspace(M, N, K), device="cuda", dtype=torch.uint8)
        fl = 2.0 * M * N * K
        for mode, name in ((0, "ffma"), (1, "tc")):
            if a.engine and a.engine != name:
                continue
            lib.gymrl_set_gemm_mode(mode)
            if want('fwd+tanh'):
                t = timeit(lambda i: ops.linear_forward(xs[i], w, b, 1, out=ys[i]), sets)
                print(f"{'fwd+tanh':10s} {M:6d} {N:4d} {K:4d} {name:>6s} {t:8.4f} {fl / t / 1e9:8.1f}")
            if want('fwd'):
                t = timeit(lambda i: ops.linear_forward(xs[i], w, b, 0, out=ys[i]), sets)
                print(f"{'fwd':10s} {M:6d} {N:4d} {K:4d} {name:>6s} {t:8.4f} {fl / t / 1e9:8.1f}")
            if want('dX*act'):
                t = timeit(lambda i: ops.linear_backward_input(dys[i], w, hs[i], 1, out=dxs[i]), sets)
                print(f"{'dX*act':10s} {M:6d} {N:4d} {K:4d} {name:>6s} {t:8.4f} {fl / t / 1e9:8.1f}")
            if want('dW+db'):
                t = timeit(lambda i: ops.linear_backward_weight(dys[i], hs[i], dw, db, workspace=ws), sets)
                print(f"{'dW+db':10s} {M:6d} {N:4d} {K:4d} {name:>6s} {t:8.4f} {fl / t / 1e9:8.1f}")
    lib.gymrl_set_gemm_mode(1)


if __name__ == "__main__":
    main()
