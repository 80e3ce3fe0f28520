"""%globaltimer stamps from inside gemm3x_ws_kernel (developer probe: gymrl_debug_tc_cta_times, 32 slots per CTA): where a
persistent CTA pair spends its time — setup, first operands, main loop per tile, accumulator hand-over, epilogue.

    python tools/ws_cta_times.py fwd|dx [M] [N] [K]
"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gymrl_b200 import _ffi, ops  # noqa: E402

SLOTS = {0: "entry", 1: "setup done (barriers, TMEM, cluster sync)", 16: "TMA: first box issued", 8: "producer: first A operands in registers",
         9: "producer: slab 0 staged", 4: "MMA: slab 0 full", 17: "TMA: last box of tile 0 issued", 10: "producer: tile 0 staged", 5: "MMA: tile 0 issued",
         12: "epilogue: tile 0 accumulator ready", 6: "MMA: tile 1 slab 0 full", 13: "epilogue: tile 0 stored", 11: "producer: tile 1 staged",
         7: "MMA: tile 1 issued", 14: "epilogue: tile 1 accumulator ready", 15: "epilogue: tile 1 stored", 2: "exit (after the cluster sync)"}


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "fwd"
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    N = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    K = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    lib = _ffi.load()
    lib.gymrl_debug_tc_cta_times.argtypes = [ctypes.c_void_p]
    buf = torch.zeros(32 * 1024, dtype=torch.int64, device="cuda")
    x, b = torch.randn(M, K, device="cuda"), torch.zeros(N, device="cuda")
    flat = torch.zeros(N * K + 8, device="cuda")
    flat[4:4 + N * K] = (torch.randn(N, K, device="cuda") / 16).reshape(-1)
    w = flat[4:4 + N * K].view(N, K)
    img = ops.weight_images_register(flat, [(4, N, K)])
    y, dy, dx = torch.empty(M, N, device="cuda"), torch.randn(M, N, device="cuda"), torch.empty(M, K, device="cuda")
    flush = torch.empty(64 * 1024 * 1024, device="cuda")

    def run():
        if kind == "fwd":
            ops.linear_forward(x, w, b, _ffi.ACT_TANH, out=y)
        else:
            ops.linear_backward_input(dy, w, x, _ffi.ACT_TANH, out=dx)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    for cold in (False, True):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        print(f"{kind} M={M} N={N} K={K} ({'L2 flushed' if cold else 'warm L2'}): {e0.elapsed_time(e1) * 1e3:.1f} us (events, eager launch)")
        buf.zero_()
        if cold:
            flush.zero_()
        lib.gymrl_debug_tc_cta_times(ctypes.c_void_p(buf.data_ptr()))
        run()
        torch.cuda.synchronize()
        lib.gymrl_debug_tc_cta_times(ctypes.c_void_p(0))
        t = buf.view(-1, 32).cpu().numpy()
        t = t[t[:, 0] != 0]
        t0 = t[:, 0].min()
        print(f"  CTAs {len(t)}, SMs {len(np.unique(t[:, 3]))}")
        for slot, name in SLOTS.items():
            v = t[:, slot]
            v = v[v != 0]
            if len(v) == 0:
                continue
            r = (v - t0) / 1e3
            print(f"  {name:46s} n={len(v):3d}  min {r.min():7.2f}  median {np.median(r):7.2f}  max {r.max():7.2f} us")
    ops.weight_images_unregister(flat)


if __name__ == "__main__":
    main()
