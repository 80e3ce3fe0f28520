"""Per-phase clock64() timeline of CTA 0 of the tcgen05 GEMM (developer probe; needs the gymrl_debug_tc_timeline hook).

    python tools/tc_timeline.py [N] [K]
"""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from gymrl_b200 import _ffi, ops  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    M = 16384
    lib = _ffi.load()
    lib.gymrl_debug_tc_timeline.argtypes = [ctypes.c_void_p]
    buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
    x, w, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / 16, torch.zeros(N, device="cuda")
    y = torch.empty(M, N, device="cuda")
    for _ in range(3):
        ops.linear_forward(x, w, b, 0, out=y)
    torch.cuda.synchronize()
    lib.gymrl_debug_tc_timeline(ctypes.c_void_p(buf.data_ptr()))
    ops.linear_forward(x, w, b, 0, out=y)
    torch.cuda.synchronize()
    lib.gymrl_debug_tc_timeline(ctypes.c_void_p(0))
    for name, off in (("producer thread 0", 0), ("MMA thread", 512)):
        t = buf[off:off + 512].cpu().tolist()
        t0 = t[0]
        print(f"--- {name}: start 0; mainloop end {t[1]-t0}; acc ready {t[2]-t0}; epilogue end {t[3]-t0}; after sync {t[4]-t0}")
        nslab = K // 32
        print(" slab: enter  +wait  +work   (producer: wait stage free, convert+STS+arrive; MMA thread: wait stage full, issue 12 MMAs + commit)")
        for kb in range(nslab):
            s = t[8 + kb * 8: 8 + kb * 8 + 3]
            d = [s[0] - t0] + [s[i] - s[i - 1] for i in range(1, 3)]
            print(f"  {kb:2d}: {d[0]:7d} {d[1]:7d} {d[2]:7d}")
        if off == 0:
            print(" epilogue chunk (thread 0): tcgen05.ld done (since acc ready / previous chunk end), +STS+syncwarp, +LDS/math/STG issue")
            prev = t[2]
            for c in range(4):
                s = t[300 + c * 4: 300 + c * 4 + 3]
                print(f"  chunk {c}: {s[0] - prev:7d} {s[1] - s[0]:7d} {s[2] - s[1]:7d}")
                prev = s[2]


if __name__ == "__main__":
    main()
