"""Wall-clock (%globaltimer) stamps of every CTA of one tcgen05 GEMM launch (developer probe, gymrl_debug_tc_cta_times):
when each CTA starts, finishes its main loop, gets its accumulator, finishes its epilogue and exits, and on which SM.

    python tools/tc_cta_times.py fwd|dx|dw [M] [N] [K]
"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gymrl_b200 import _ffi, ops  # noqa: E402


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    K = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    lib = _ffi.load()
    lib.gymrl_debug_tc_cta_times.argtypes = [ctypes.c_void_p]
    buf = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
    x, w, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / 16, torch.zeros(N, device="cuda")
    y = torch.empty(M, N, device="cuda")
    dy = torch.randn(M, N, device="cuda")
    dx = torch.empty(M, K, device="cuda")
    gw, gb = torch.empty(N, K, device="cuda"), torch.empty(N, device="cuda")
    ws = torch.empty(ops.backward_weight_workspace(M, N, K), dtype=torch.uint8, device="cuda")

    def run():
        if kind == "fwd":
            ops.linear_forward(x, w, b, _ffi.ACT_TANH, out=y)
        elif kind == "dx":
            ops.linear_backward_input(dy, w, x, _ffi.ACT_TANH, out=dx)
        else:
            ops.linear_backward_weight(dy, x, gw, gb, workspace=ws)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    print(f"{kind} M={M} N={N} K={K}: {e0.elapsed_time(e1) * 1e3:.1f} us (events, includes the reduction for dw)")
    lib.gymrl_debug_tc_cta_times(ctypes.c_void_p(buf.data_ptr()))
    run()
    torch.cuda.synchronize()
    lib.gymrl_debug_tc_cta_times(ctypes.c_void_p(0))
    t = buf.view(-1, 8).cpu().numpy()
    t = t[t[:, 0] != 0]
    t0 = t[:, 0].min()
    names = ["entry", "tmem_ready", "mainloop_end", "acc_ready", "epilogue_end", "exit"]
    rel = (t[:, :6] - t0) / 1e3
    print(f"CTAs {len(t)}, SMs used {len(np.unique(t[:, 6]))}, launch span {rel[:, 5].max():.1f} us")
    for i, n in enumerate(names):
        print(f"  {n:13s} min {rel[:, i].min():7.2f}  median {np.median(rel[:, i]):7.2f}  max {rel[:, i].max():7.2f} us")
    d = np.diff(rel, axis=1)
    for i, n in enumerate(["alloc", "mainloop", "acc wait", "epilogue", "teardown"]):
        print(f"  phase {n:9s} median {np.median(d[:, i]):7.2f}  p90 {np.percentile(d[:, i], 90):7.2f}  max {d[:, i].max():7.2f} us")
    first = rel[:, 0] < 2.0
    print(f"  first-wave CTAs {first.sum()}: exit median {np.median(rel[first, 5]):.2f}; later CTAs {len(t) - first.sum()}: "
          f"entry median {np.median(rel[~first, 0]) if (~first).any() else float('nan'):.2f}")
    # per-SM: how many CTAs and the gap between one CTA's exit and the next one's entry
    gaps = []
    for sm in np.unique(t[:, 6]):
        r = rel[t[:, 6] == sm]
        r = r[np.argsort(r[:, 0])]
        for a, b2 in zip(r[:-1], r[1:]):
            gaps.append(b2[0] - a[5])
    if gaps:
        print(f"  same-SM exit -> next entry gap: median {np.median(gaps):.2f} max {np.max(gaps):.2f} us ({len(gaps)} pairs)")


if __name__ == "__main__":
    main()
