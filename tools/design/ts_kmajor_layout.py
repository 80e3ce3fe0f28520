"""Design aid for DESIGN.md §8 item 1 (K-major A through tensor memory): checks, in NumPy, the index algebra of the planned
variant before any PTX is written.

Plan: per 32-k slab a warp stores its 16 rows x 32 k block of A (hi or lo image) with ONE `tcgen05.st.16x256b.x4`.
In that shape (CUTLASS copy_traits_sm100.hpp, SM100_TMEM_STORE_16dp256b1x: thread t of the warp holds, per repeat q,
  regs 2q*2+{0,1} -> (lane t/4,     columns 8q + 2(t%4) + {0,1})
  regs 2q*2+{2,3} -> (lane t/4 + 8, columns 8q + 2(t%4) + {0,1}) )
a quad of threads owns 8 consecutive TMEM columns of one row.  If TMEM column c of the slab held k = c, thread t would need the
k-values {8q + 2(t%4) + e}: four 8-byte pieces 32 B apart.  Permuting the slab's k axis at 8-byte granularity,
        TMEM / MMA column 8q + 2m + e   <->   k = 8m + 2q + e        (m = t % 4, q = 0..3, e = 0..1),
makes thread (row r, m) need k = 8m .. 8m + 7: two LDG.128 from one row, and a quad reads the row's whole 128-byte line.
The B operand must present the same permutation to the tensor core: the staging thread that holds B[n][4c .. 4c+3] writes
its two 8-byte halves to MMA columns pi(4c), pi(4c + 2) (two STS.64 instead of one STS.128).

This script verifies: (1) pi is a bijection on 0..31 that maps each MMA k-step (8 consecutive MMA columns) to a set of k's that
is the same for A and B; (2) every thread's register list is exactly k = 8m..8m+7 of rows t/4 and t/4 + 8; (3) the GEMM computed
slab by slab through the permuted operands equals A @ B.T.
"""
import numpy as np


def pi_k_of_col(c):
    """k index stored in TMEM / MMA column c of the slab."""
    q, m, e = c // 8, (c % 8) // 2, c % 2
    return 8 * m + 2 * q + e


def main():
    cols = np.arange(32)
    k_of_col = np.array([pi_k_of_col(c) for c in cols])
    assert sorted(k_of_col) == list(range(32)), "pi must be a bijection"
    # (2) register contents of thread t for the 16x256b.x4 store
    for t in range(32):
        m, r = t % 4, t // 4
        held = []
        for q in range(4):
            for e in range(2):
                held.append(int(k_of_col[8 * q + 2 * m + e]))
        assert sorted(held) == list(range(8 * m, 8 * m + 8)), (t, held)
    print("thread (row t/4 [+8], m = t%4) holds k = 8m .. 8m+7 in register order",
          [int(k_of_col[8 * q + 2 * 1 + e]) for q in range(4) for e in range(2)], "(example m = 1)")
    # (1)+(3) slab-wise product through permuted operands
    rng = np.random.default_rng(0)
    M, N, K = 16, 24, 64
    A, B = rng.standard_normal((M, K)), rng.standard_normal((N, K))
    C = np.zeros((M, N))
    for s in range(K // 32):
        a_tmem = A[:, 32 * s + k_of_col]          # TMEM columns of the slab (what tcgen05.st wrote)
        b_smem = B[:, 32 * s + k_of_col]          # B image columns in MMA order (what the permuted STS.64 pairs wrote)
        for j in range(4):                        # four k-steps of 8 MMA columns
            C += a_tmem[:, 8 * j:8 * j + 8] @ b_smem[:, 8 * j:8 * j + 8].T
    assert np.allclose(C, A @ B.T)
    # B staging: thread with original chunk c (k = 4c..4c+3) -> MMA columns of its two halves
    col_of_k = np.argsort(k_of_col)
    for c in range(8):
        lo, hi = col_of_k[4 * c], col_of_k[4 * c + 2]
        assert col_of_k[4 * c + 1] == lo + 1 and col_of_k[4 * c + 3] == hi + 1      # each half stays an aligned 8-byte pair
        print(f"B chunk {c} (k {4 * c}..{4 * c + 3}) -> MMA columns {lo},{lo + 1} and {hi},{hi + 1}")
    print("OK: permuted slab product equals A @ B.T")


if __name__ == "__main__":
    main()
