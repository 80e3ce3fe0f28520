// linear_tc.cu — tcgen05 (5th-gen tensor core) path of the dense layers: fp32-accurate GEMM by 3xTF32 splitting.
//
//   C[M][N] (+epilogue) = sum_k A(m,k) B(n,k),  A/B fp32 in global memory, accumulation in TMEM (fp32).
//
// Every fp32 operand x is split in registers into hi = tf32(x) and lo = tf32(x - hi); the product is
// accumulated as  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (the dropped lo*lo term is ~2^-22 relative), which keeps
// the reference's fp32 parity bar while running on the tensor pipe (kind::tf32, UMMA 128 x BN x 8).
//
// One CTA (8 producer / epilogue warps + 1 MMA-issue warp) owns a 128 x BN output tile (BN = 64/128/256 TMEM columns):
//   stage loop over K in slabs of 32 floats (= one 128-byte swizzle span), two smem stages:
//     producer warps:  LDG.128 the A/B slab one slab ahead (row gather fused for A), split hi/lo, st.shared.v4 into the
//                      canonical UMMA shared-memory layout (SWIZZLE_128B; K-major: 8-row x 128 B atoms, MN-major: 128B_BASE32B
//                      atoms), fence.proxy.async, mbarrier arrive (full)
//     MMA warp, one lane: 4 k-steps x 3 tcgen05.mma (descriptors advance 32 B inside the swizzle atom for K-major,
//                      one atom row-block per k-step for MN-major), tcgen05.commit -> mbarrier that frees the stage
//   so the tensor core works on stage s while the producers fill stage s^1 (operands never round-trip through HBM
//   in split form).  Epilogue: tcgen05.ld 32 lanes x 32 columns per warp -> per-warp smem transpose -> bias / tanh / relu /
//   act'(h) (one instantiation per variant) -> coalesced 128 B row segments.
//   Backward-weight runs split-K over blockIdx.z with deterministic partial tiles (folded by reduce.cu / linear_skinny.cu);
//   its MN-major A operand goes through tensor memory instead of shared memory (gemm_tf32x3_ts_kernel below).
// No TMA: the operands need the register pass for the split (and the minibatch row gather), so a bulk tensor
// copy cannot produce them; staging is plain coalesced 128 B row reads instead.
// All mbarrier waits are bounded (trap on timeout) so a descriptor bug cannot hang the GPU.
#include "common.cuh"
#include <cstdlib>

void gymrl_count_launch(int n = 1);

#include "linear_tc.cuh"

// developer timeline probe (tools/tc_timeline.py): CTA 0 / thread 0 and thread 255 stamp clock64() per pipeline phase
__device__ long long* g_tc_dbg = nullptr;
extern "C" int gymrl_debug_tc_timeline(long long* d_buf) {
    return cudaMemcpyToSymbol(g_tc_dbg, &d_buf, sizeof(d_buf)) == cudaSuccess ? 0 : -1;
}
#define TC_STAMP(slot) do { if (dbg) dbg[(slot)] = clock64(); } while (0)
// per-CTA wall-clock probe: thread 0 of every CTA stamps %globaltimer (ns) at entry / TMEM ready / mainloop end /
// accumulator ready / epilogue end / exit, plus its SM id -> 8 slots per CTA
__device__ long long* g_tc_cta_dbg = nullptr;
extern "C" int gymrl_debug_tc_cta_times(long long* d_buf) {
    return cudaMemcpyToSymbol(g_tc_cta_dbg, &d_buf, sizeof(d_buf)) == cudaSuccess ? 0 : -1;
}
__device__ __forceinline__ long long globaltimer_ns() {
    long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}
#define TC_GSTAMP(slot) do { if (cdbg) cdbg[(slot)] = globaltimer_ns(); } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Explicit shared-state-space accesses: the stage ring is reached through a pointer that was aligned by integer
// arithmetic, so the compiler no longer knows its address space and emits generic ST.E / LD.E for plain dereferences.
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        uint32_t done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();  // bounded wait: fail loudly instead of hanging the GPU
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100): SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;             // version = 1 (Blackwell)
    d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B (K-major), 1 = SWIZZLE_128B_BASE32B (tf32 MN-major)
    return d;
}
// MN-major tf32 slab [ROWS x 32 k]: block (k-atom jk of 4 rows, MN-atom i of 32 elements) at (jk*ROWS/32 + i)*512 B.
// One UMMA k-step (8 k) = k-atoms 2j, 2j+1:  LBO = 512 B between MN atoms, SBO = ROWS/32 * 512 B between k-atoms.
template <int ROWS>
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t base, int j) {
    return make_desc(base + (uint32_t)(2 * j * (ROWS / 32)) * 512u, 512u, (uint32_t)(ROWS / 32) * 512u, 1u);
}
// fp32 -> tf32, round to nearest (ties away), for finite inputs bit-identical to cvt.rna.tf32.f32; two integer ops
// instead of the four ptxas emits for the cvt (its Inf/NaN guard; non-finite inputs stay non-finite here too).
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split4(const float4 v, uint4& hi, uint4& lo) {
    hi.x = to_tf32(v.x); hi.y = to_tf32(v.y); hi.z = to_tf32(v.z); hi.w = to_tf32(v.w);
    lo.x = to_tf32(v.x - __uint_as_float(hi.x)); lo.y = to_tf32(v.y - __uint_as_float(hi.y));
    lo.z = to_tf32(v.z - __uint_as_float(hi.z)); lo.w = to_tf32(v.w - __uint_as_float(hi.w));
}

// Stage one [ROWS x 32] operand slab into smem (hi and lo images).
//  KMAJOR: global rows are the MN index (128 B = 32 k per row);  smem atom = 8 MN-rows x 128 B, atoms stacked along MN (1024 B).
//  !KMAJOR: global rows are the k index (contiguous MN);         smem block(j, i) = k-atom j (8 k-rows x 128 B) of MN-atom i
//           (32 MN elements), at ((j * ROWS/32) + i) * 1024.
#define TC_THREADS 256   // 8 warps: two warps per TMEM lane quarter, so staging and the epilogue have latency-hiding partners

// One operand's [ROWS x 32 k] slab pipeline: global -> registers (load) and registers -> hi/lo tf32 images in the
// canonical UMMA smem layout (store).  Loads are unconditional (out-of-range MN rows are clamped onto valid ones:
// they only feed output rows/columns the epilogue never writes) and addressed through per-thread pointers set up
// once, so the compiler keeps every LDG.128 of a slab in flight at once; the kernel holds two register sets so
// the loads of slab k+1 are issued before slab k is converted.
//  KMAJOR: global rows are the MN index (128 B = 32 k per row);  smem atom = 8 MN-rows x 128 B, atoms stacked along MN (1024 B).
//  !KMAJOR: global rows are the k index (contiguous MN);  tf32 MN-major operands must use SWIZZLE_128B_BASE32B
//           (cutlass sm100_common.inl:92): atoms of 4 k-rows x 128 B, Swizzle<2,5,2> = the 32-byte chunk index is XOR-ed
//           with the k-row index inside the atom; block (k-atom jk, MN-atom i) at (jk * ROWS/32 + i) * 512 B.
template <int ROWS, bool KMAJOR>
struct Stager {
    static constexpr int PASSES = ROWS * 8 / TC_THREADS;   // float4 per thread per slab
    static constexpr int CH = ROWS / 4;                    // MN-major: float4 per k-row
    static constexpr int KSTEP = TC_THREADS / CH;          // MN-major: k-rows covered per pass (multiple of 4)
    const float* ptr[KMAJOR ? PASSES : 1];
    const int32_t* rows;
    long long ld;
    int k;          // MN-major + gather: next k-row of this thread's first pass
    uint32_t soff;  // byte offset of pass 0 inside an image; pass i adds a compile-time stride

    __device__ __forceinline__ void init(const float* g, int ld_, const int32_t* rows_, int mn0, int mn_total, int kbeg) {
        const int t = threadIdx.x;
        ld = ld_;
        rows = rows_;
        if (KMAJOR) {
            const int r0 = t >> 3, c = t & 7;
#pragma unroll
            for (int i = 0; i < PASSES; ++i) {
                const int mn = min(mn0 + r0 + (TC_THREADS / 8) * i, mn_total - 1);
                const long long gr = rows_ ? (long long)rows_[mn] : (long long)mn;
                ptr[i] = g + gr * ld + kbeg + c * 4;
            }
            soff = (uint32_t)(r0 >> 3) * 1024u + (uint32_t)(r0 & 7) * 128u + (uint32_t)((c ^ (r0 & 7)) << 4);
        } else {
            const int kk = t / CH, c = t % CH;
            const int mn = min(mn0 + c * 4, mn_total - 4);
            k = kbeg + kk;
            ptr[0] = g + mn + (rows_ ? 0ll : (long long)k * ld);
            const int jk = kk >> 2, kr = kk & 3, ai = c >> 3, c16 = c & 7;
            soff = (uint32_t)(jk * (ROWS / 32) + ai) * 512u + (uint32_t)kr * 128u + (uint32_t)((((c16 >> 1) ^ kr) << 5) | ((c16 & 1) << 4));
        }
    }
    // issue the loads of the next slab and advance
    __device__ __forceinline__ void load(float4 (&v)[PASSES]) {
        if (KMAJOR) {
#pragma unroll
            for (int i = 0; i < PASSES; ++i) {
                v[i] = __ldg(reinterpret_cast<const float4*>(ptr[i]));
                ptr[i] += 32;
            }
        } else if (rows == nullptr) {
#pragma unroll
            for (int i = 0; i < PASSES; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(ptr[0] + (long long)(KSTEP * i) * ld));
            ptr[0] += 32 * ld;
        } else {
            int32_t ri[PASSES];
#pragma unroll
            for (int i = 0; i < PASSES; ++i) ri[i] = __ldg(rows + k + KSTEP * i);
#pragma unroll
            for (int i = 0; i < PASSES; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(ptr[0] + (long long)ri[i] * ld));
            k += 32;
        }
    }
    __device__ __forceinline__ void store(uint32_t s_hi, uint32_t s_lo, const float4 (&v)[PASSES]) const {
        constexpr uint32_t STRIDE = KMAJOR ? (uint32_t)(TC_THREADS / 64) * 1024u : (uint32_t)(KSTEP / 4) * (ROWS / 32) * 512u;
#pragma unroll
        for (int i = 0; i < PASSES; ++i) {
            uint4 hi, lo;
            split4(v[i], hi, lo);
            sts128(s_hi + soff + STRIDE * i, hi);
            sts128(s_lo + soff + STRIDE * i, lo);
        }
    }
};

// ---- epilogue: TMEM -> registers -> per-warp smem transpose -> coalesced global -------------------------------------
// A thread owns one accumulator row (32 contiguous columns per tcgen05.ld): storing that directly makes every STG touch
// 32 different lines.  Each warp transposes its 32x32 block through a private, XOR-swizzled 4 KB scratch so that 8 lanes
// cover one 128 B row segment -> 4 full lines per STG.128.  The activation variant is a template parameter: with it as a
// run-time value the eight row groups of a chunk were separated by branches and ran strictly one after the other
// (LD -> tanh chain -> STG, ~200 cycles each; 1.7-2.2k cycles per chunk measured); branch-free they are issued together.
enum { EPI_PLAIN = 0, EPI_TANH, EPI_RELU, EPI_DTANH, EPI_DRELU, EPI_GENERIC };
struct EpiArgs {
    uint32_t scr, tmem_row;
    int lane, c_begin, c_end, m_base, n0;
    float* C;
    bool have_acc;
    long long* dbg;
};
template <int MODE>
__device__ __forceinline__ float epi_apply(float v, float h, int act, int act_in, bool has_h) {
    if (MODE == EPI_TANH) return tanh_fast(v);
    if (MODE == EPI_RELU) return fmaxf(v, 0.f);
    if (MODE == EPI_DTANH) return v * (1.0f - h * h);
    if (MODE == EPI_DRELU) return h > 0.f ? v : 0.f;
    if (MODE == EPI_GENERIC) {
        if (act == GYMRL_ACT_TANH) v = tanh_fast(v);
        else if (act == GYMRL_ACT_RELU) v = fmaxf(v, 0.f);
        if (has_h) {
            if (act_in == GYMRL_ACT_TANH) v *= (1.0f - h * h);
            else if (act_in == GYMRL_ACT_RELU) v = h > 0.f ? v : 0.f;
        }
    }
    return v;
}
template <int MODE>
__device__ __forceinline__ void epilogue_chunks(const TcGemmParams& p, const EpiArgs& ea) {
    constexpr bool USE_H = MODE == EPI_DTANH || MODE == EPI_DRELU || MODE == EPI_GENERIC;
    long long* dbg = ea.dbg;
    const int lane = ea.lane, cq = lane & 7, rsub = lane >> 3;
    const bool has_h = USE_H && p.H != nullptr;
#pragma unroll 1
    for (int c0 = ea.c_begin; c0 < ea.c_end; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = ea.tmem_row + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        // this lane's output columns after the transpose (same for all 8 row groups): bias loaded once per chunk
        const int n = ea.n0 + c0 + cq * 4;
        const bool n_ok = n < p.N;
        const float4 b4 = (p.bias && n_ok) ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 h4[8];
        if (has_h) {   // issued before the TMEM wait: the L2 latency of the h tile hides behind the transpose
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int m = min(ea.m_base + j * 4 + rsub, p.M - 1);
                h4[j] = n_ok ? __ldg(reinterpret_cast<const float4*>(p.H + (long long)m * p.ldh + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        TC_STAMP(300 + (c0 >> 5) * 4 + 0);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            sts128(ea.scr + lane * 128 + ((q ^ (lane & 7)) << 4),
                   ea.have_acc ? make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]) : make_uint4(0u, 0u, 0u, 0u));
        __syncwarp();
        TC_STAMP(300 + (c0 >> 5) * 4 + 1);
        float4 a4[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int row = j * 4 + rsub;
            a4[j] = lds128(ea.scr + row * 128 + ((cq ^ (row & 7)) << 4));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 hh = has_h ? h4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            a4[j].x = epi_apply<MODE>(a4[j].x + b4.x, hh.x, p.act, p.act_in, has_h);
            a4[j].y = epi_apply<MODE>(a4[j].y + b4.y, hh.y, p.act, p.act_in, has_h);
            a4[j].z = epi_apply<MODE>(a4[j].z + b4.z, hh.z, p.act, p.act_in, has_h);
            a4[j].w = epi_apply<MODE>(a4[j].w + b4.w, hh.w, p.act, p.act_in, has_h);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int m = ea.m_base + j * 4 + rsub;
            if (m < p.M && n_ok) *reinterpret_cast<float4*>(ea.C + (long long)m * p.ldc + n) = a4[j];
        }
        TC_STAMP(300 + (c0 >> 5) * 4 + 2);
        __syncwarp();   // the scratch is rewritten by the next chunk
    }
}

#define TC_LAUNCH_THREADS (TC_THREADS + 32)   // 8 producer / epilogue warps + 1 MMA-issue warp

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory"); }

template <int BN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(TC_LAUNCH_THREADS, 1) gemm_tf32x3_kernel(const TcGemmParams p) {
    constexpr int BM = 128;
    constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128;           // one image (hi or lo) of one 32-k slab
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[2];   // producers -> MMA warp: stage converted (TC_THREADS arrivals)
    __shared__ __align__(8) uint64_t bar_free[2];   // tensor core -> producers: the MMAs reading the stage retired (tcgen05.commit)
    __shared__ __align__(8) uint64_t bar_acc;       // tensor core -> epilogue: accumulator complete
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    long long* cdbg = (g_tc_cta_dbg && t == 0) ? g_tc_cta_dbg + 8ll * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
    TC_GSTAMP(0);
    if (cdbg) { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); cdbg[6] = smid; }
    const int kbeg = blockIdx.z * p.k_chunk;
    const int kend = min(p.K, kbeg + p.k_chunk);
    const int nslab = (kend - kbeg + 31) / 32;

    if (t == 0) {
        mbar_init(&bar_full[0], TC_THREADS);
        mbar_init(&bar_full[1], TC_THREADS);
        mbar_init(&bar_free[0], 1);
        mbar_init(&bar_free[1], 1);
        mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    // PDL: barrier init and the TMEM allocation above touch no global memory, so they overlap the predecessor's tail
    pdl_wait();
    pdl_launch_dependents();

    long long* dbg = (g_tc_dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (t == 0 || t == TC_THREADS)) ? g_tc_dbg + (t ? 512 : 0) : nullptr;
    TC_STAMP(0);
    TC_GSTAMP(1);

    if (warp == TC_THREADS / 32) {
        // ===== MMA-issue warp: one elected lane feeds the tensor core; issuing blocks for about the duration of the
        // MMAs (measured: 1.4k cycles per 12-MMA slab), which is why it must not share a thread with the converters =====
        if (lane == 0) {
            // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32, A=B=TF32, majors, N>>3, M>>4
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((A_KMAJOR ? 0u : 1u) << 15) | ((B_KMAJOR ? 0u : 1u) << 16) |
                                       ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < nslab; ++kb) {
                const int s = kb & 1, use = kb >> 1;
                TC_STAMP(8 + kb * 8 + 0);
                mbar_wait(&bar_full[s], (uint32_t)(use & 1));
                TC_STAMP(8 + kb * 8 + 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the producers' st.shared -> visible to the tensor core
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
                const uint32_t ah = st, al = st + A_BYTES, bh = st + 2 * A_BYTES, bl = st + 2 * A_BYTES + B_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j) {   // 4 k-steps of 8
                    // K-major: +32 B inside the 128 B swizzle span (LBO = 16 B unused, SBO = 1024 B between 8-row atoms)
                    // MN-major: k-atom pair j; LBO = 512 B between MN atoms, SBO = k-atom stride
                    const uint64_t dah = A_KMAJOR ? make_desc(ah + j * 32, 16, 1024) : make_desc_mn<BM>(ah, j);
                    const uint64_t dal = A_KMAJOR ? make_desc(al + j * 32, 16, 1024) : make_desc_mn<BM>(al, j);
                    const uint64_t dbh = B_KMAJOR ? make_desc(bh + j * 32, 16, 1024) : make_desc_mn<BN>(bh, j);
                    const uint64_t dbl = B_KMAJOR ? make_desc(bl + j * 32, 16, 1024) : make_desc_mn<BN>(bl, j);
                    umma_tf32(tmem_d, dal, dbh, IDESC, (kb | j) ? 1u : 0u);
                    umma_tf32(tmem_d, dah, dbl, IDESC, 1u);
                    umma_tf32(tmem_d, dah, dbh, IDESC, 1u);
                }
                umma_commit(&bar_free[s]);
                if (kb == nslab - 1) umma_commit(&bar_acc);
                TC_STAMP(8 + kb * 8 + 2);
            }
        }
    } else {
        // ===== producer warps: global -> registers -> hi/lo tf32 images in the stage ring; then the epilogue =====
        const uint32_t smem_base = smem_u32(smem);
        Stager<BM, A_KMAJOR> sa;
        Stager<BN, B_KMAJOR> sb;
        sa.init(p.A, p.lda, p.a_rows, m0, p.M, kbeg);
        sb.init(p.B, p.ldb, p.b_rows, n0, p.N, kbeg);
        float4 va0[Stager<BM, A_KMAJOR>::PASSES], vb0[Stager<BN, B_KMAJOR>::PASSES];
        float4 va1[Stager<BM, A_KMAJOR>::PASSES], vb1[Stager<BN, B_KMAJOR>::PASSES];
        float4 va2[Stager<BM, A_KMAJOR>::PASSES], vb2[Stager<BN, B_KMAJOR>::PASSES];
        const bool do_colsum = !A_KMAJOR && p.colsum != nullptr && blockIdx.x == 0;
        float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);

        auto consume = [&](int kb, const float4 (&va)[Stager<BM, A_KMAJOR>::PASSES], const float4 (&vb)[Stager<BN, B_KMAJOR>::PASSES]) {
            const int s = kb & 1, use = kb >> 1;
            TC_STAMP(8 + kb * 8 + 0);
            if (use >= 1) mbar_wait(&bar_free[s], (uint32_t)((use - 1) & 1));   // MMAs that read this stage have retired
            TC_STAMP(8 + kb * 8 + 1);
            const uint32_t st = smem_base + (uint32_t)s * STAGE_BYTES;
            sa.store(st, st + A_BYTES, va);
            sb.store(st + 2 * A_BYTES, st + 2 * A_BYTES + B_BYTES, vb);
            if (!A_KMAJOR && do_colsum) {   // db rides along with dW: the dY slab is already in registers
#pragma unroll
                for (int i = 0; i < Stager<BM, A_KMAJOR>::PASSES; ++i) {
                    csum.x += va[i].x; csum.y += va[i].y; csum.z += va[i].z; csum.w += va[i].w;
                }
            }
            // (the generic -> async proxy fence is executed by the MMA lane after it has acquired bar_full: here it would compile to
            //  MEMBAR.ALL.CTA and wait for this warp's prefetched global loads, i.e. cost a memory round trip per slab)
            mbar_arrive(&bar_full[s]);
            TC_STAMP(8 + kb * 8 + 2);
        };

        // register sets: the loads of the next slab(s) are in flight while slab k is converted.  The launch is 9 warps,
        // for which ptxas budgets 168 registers per thread: three sets fit for BN <= 128, two for BN = 256.
        if (BN <= 128) {
            if (nslab > 0) { sa.load(va0); sb.load(vb0); }
            if (nslab > 1) { sa.load(va1); sb.load(vb1); }
            for (int kb = 0; kb < nslab; kb += 3) {
                if (kb + 2 < nslab) { sa.load(va2); sb.load(vb2); }
                consume(kb, va0, vb0);
                if (kb + 1 < nslab) {
                    if (kb + 3 < nslab) { sa.load(va0); sb.load(vb0); }
                    consume(kb + 1, va1, vb1);
                }
                if (kb + 2 < nslab) {
                    if (kb + 4 < nslab) { sa.load(va1); sb.load(vb1); }
                    consume(kb + 2, va2, vb2);
                }
            }
        } else {
            if (nslab > 0) { sa.load(va0); sb.load(vb0); }
            for (int kb = 0; kb < nslab; kb += 2) {
                if (kb + 1 < nslab) { sa.load(va1); sb.load(vb1); }
                consume(kb, va0, vb0);
                if (kb + 1 < nslab) {
                    if (kb + 2 < nslab) { sa.load(va0); sb.load(vb0); }
                    consume(kb + 1, va1, vb1);
                }
            }
        }

        // ---- epilogue: TMEM -> registers -> smem transpose -> coalesced global ----
        TC_STAMP(1);
        TC_GSTAMP(2);
        if (nslab > 0) mbar_wait(&bar_acc, 0);
        TC_STAMP(2);
        TC_GSTAMP(3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // all MMAs have retired: the operand stages are free to reuse as scratch
        if (!A_KMAJOR && do_colsum) {
            // thread t summed columns 4*(t%32).. of the k-rows congruent to t/32 (mod 8): fold the 8 warps in fixed order
            float* red = reinterpret_cast<float*>(smem + 64 * 1024);
            *reinterpret_cast<float4*>(red + warp * BM + lane * 4) = csum;
            producers_sync();
            if (t < BM && m0 + t < p.M) {
                float sum = 0.f;
#pragma unroll
                for (int wv = 0; wv < TC_THREADS / 32; ++wv) sum += red[wv * BM + t];
                p.colsum[(long long)blockIdx.z * p.M + m0 + t] = sum;
            }
        }
        const int lane_q = warp & 3, col_half = warp >> 2;    // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        float* Cbase = p.C + (long long)blockIdx.z * p.c_split_stride;
        // A thread owns one accumulator row (32 contiguous columns per tcgen05.ld): storing that directly makes every
        // STG touch 32 different lines (measured: 12k cycles of epilogue per tile).  Each warp transposes its 32x32 block
        // through a private, XOR-swizzled 4 KB scratch so that 8 lanes cover one 128 B row segment -> 4 lines per STG.
        // (Prefetching the next chunk's tcgen05.ld while the current one is stored was tried and measured 8-20 % slower
        // on the same box, so the chunks stay strictly sequential.)
        const uint32_t scr = smem_base + (uint32_t)warp * 4096u;
        EpiArgs ea;
        ea.scr = scr; ea.tmem_row = tmem_d + ((uint32_t)(lane_q * 32) << 16); ea.lane = lane;
        ea.c_begin = col_half * (BN / 2); ea.c_end = (col_half + 1) * (BN / 2);
        ea.m_base = m0 + lane_q * 32; ea.n0 = n0; ea.C = Cbase; ea.have_acc = nslab > 0; ea.dbg = dbg;
        // the activation / derivative variant is uniform over the launch: pick a branch-free instantiation once
        if (p.H == nullptr) {
            if (p.act == GYMRL_ACT_TANH) epilogue_chunks<EPI_TANH>(p, ea);
            else if (p.act == GYMRL_ACT_RELU) epilogue_chunks<EPI_RELU>(p, ea);
            else epilogue_chunks<EPI_PLAIN>(p, ea);
        } else if (p.act == GYMRL_ACT_NONE && p.act_in == GYMRL_ACT_TANH) {
            epilogue_chunks<EPI_DTANH>(p, ea);
        } else if (p.act == GYMRL_ACT_NONE && p.act_in == GYMRL_ACT_RELU) {
            epilogue_chunks<EPI_DRELU>(p, ea);
        } else {
            epilogue_chunks<EPI_GENERIC>(p, ea);
        }
        TC_STAMP(3);
        TC_GSTAMP(4);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_GSTAMP(5);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN) : "memory");
    }
}


// =====================================================================================================================
// Variant with the A operand in tensor memory (tcgen05.mma "TS" form).
// The SS kernel above is bound by the 128 B/clk shared-memory port: per 32-k slab the tensor core reads 12 x (4 KB A + BN*32 B)
// and the converters write both operands' hi/lo images (240 KB at BN = 256 -> 1875 cycles against 1536 cycles of MMA).
// Here the converters write A's hi/lo images straight into TMEM with tcgen05.st (lane = output row, column = k), so A costs
// no shared-memory traffic at all (160 KB per slab -> 1250 cycles: the MMAs become the bound) and the stage ring holds B only,
// which makes room for a third stage.  TMEM: columns [0, BN) accumulator, then NST x 64 columns of A (hi 32 | lo 32).
// A thread owns one row of the tile and half of the slab's 32 k-values (two warps per TMEM lane quarter):
//   K-major A  (A[m][k]):  four LDG.128 from its own row (64 contiguous bytes);
//   MN-major A (A[k][m]):  sixteen LDG.32, coalesced across the warp (the transposition costs nothing: TMEM is written row-wise).
// =====================================================================================================================
template <int NREG>
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[NREG]) {
    static_assert(NREG == 16, "x16");
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// A-operand loader of one thread: 16 k-values of its row per slab.
template <bool KMAJOR>
struct ARowLoader {
    const float* ptr;
    long long ld;
    __device__ __forceinline__ void init(const float* g, int ld_, const int32_t* rows_, int row, int m_total, int kbeg, int khalf) {
        const int m = min(row, m_total - 1);   // out-of-range rows are clamped onto a valid one (their outputs are never written)
        ld = ld_;
        if (KMAJOR) ptr = g + (rows_ ? (long long)rows_[m] : (long long)m) * ld_ + kbeg + 16 * khalf;
        else ptr = g + (long long)(kbeg + 16 * khalf) * ld_ + m;
    }
    __device__ __forceinline__ void load(float (&v)[16]) {
        if (KMAJOR) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(ptr) + i);
                v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
            }
            ptr += 32;
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __ldg(ptr + (long long)i * ld);
            ptr += 32 * ld;
        }
    }
};

template <int BN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(TC_LAUNCH_THREADS, 1) gemm_tf32x3_ts_kernel(const TcGemmParams p) {
    constexpr int BM = 128;
    constexpr int NST = BN == 256 ? 3 : 4;                                // B-only stages: 64 KB each at BN = 256
    constexpr uint32_t B_BYTES = BN * 128;                                // one image (hi or lo) of one 32-k slab of B
    constexpr uint32_t STAGE_BYTES = 2 * B_BYTES;
    constexpr uint32_t TMEM_COLS = (BN + 64 * NST) <= 256 ? 256 : 512;
    static_assert(BN + 64 * NST <= 512, "TMEM columns");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[NST];   // producers -> MMA warp: stage converted (TC_THREADS arrivals)
    __shared__ __align__(8) uint64_t bar_free[NST];   // tensor core -> producers: the MMAs reading the stage retired
    __shared__ __align__(8) uint64_t bar_acc;         // tensor core -> epilogue: accumulator complete
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    long long* cdbg = (g_tc_cta_dbg && t == 0) ? g_tc_cta_dbg + 8ll * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
    TC_GSTAMP(0);
    if (cdbg) { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); cdbg[6] = smid; }
    const int kbeg = blockIdx.z * p.k_chunk;
    const int kend = min(p.K, kbeg + p.k_chunk);
    const int nslab = (kend - kbeg + 31) / 32;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) { mbar_init(&bar_full[s], TC_THREADS); mbar_init(&bar_free[s], 1); }
        mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    // PDL: barrier init and the TMEM allocation above touch no global memory, so they overlap the predecessor's tail
    pdl_wait();
    pdl_launch_dependents();
    long long* dbg = nullptr;
    TC_GSTAMP(1);

    if (warp == TC_THREADS / 32) {
        // ===== MMA-issue warp =====
        if (lane == 0) {
            // A from TMEM is always K-major in the descriptor
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((B_KMAJOR ? 0u : 1u) << 16) |
                                       ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            int s = 0, ph = 0;
            for (int kb = 0; kb < nslab; ++kb) {
                mbar_wait(&bar_full[s], (uint32_t)ph);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
                const uint32_t bh = st, bl = st + B_BYTES;
                const uint32_t a_hi = tmem_d + (uint32_t)(BN + 64 * s), a_lo = a_hi + 32u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {   // 4 k-steps of 8
                    const uint64_t dbh = B_KMAJOR ? make_desc(bh + j * 32, 16, 1024) : make_desc_mn<BN>(bh, j);
                    const uint64_t dbl = B_KMAJOR ? make_desc(bl + j * 32, 16, 1024) : make_desc_mn<BN>(bl, j);
                    umma_tf32_ts(tmem_d, a_lo + 8u * j, dbh, IDESC, (kb | j) ? 1u : 0u);
                    umma_tf32_ts(tmem_d, a_hi + 8u * j, dbl, IDESC, 1u);
                    umma_tf32_ts(tmem_d, a_hi + 8u * j, dbh, IDESC, 1u);
                }
                umma_commit(&bar_free[s]);
                if (kb == nslab - 1) umma_commit(&bar_acc);
                if (++s == NST) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ===== producer warps =====
        const uint32_t smem_base = smem_u32(smem);
        const int lane_q = warp & 3, khalf = warp >> 2;
        ARowLoader<A_KMAJOR> la;
        la.init(p.A, p.lda, A_KMAJOR ? p.a_rows : nullptr, m0 + lane_q * 32 + lane, p.M, kbeg, khalf);
        Stager<BN, B_KMAJOR> sb;
        sb.init(p.B, p.ldb, p.b_rows, n0, p.N, kbeg);
        float va0[16], va1[16];
        float4 vb0[Stager<BN, B_KMAJOR>::PASSES], vb1[Stager<BN, B_KMAJOR>::PASSES];
        const bool do_colsum = !A_KMAJOR && p.colsum != nullptr && blockIdx.x == 0;
        float csum = 0.f;
        const uint32_t a_lane = tmem_d + ((uint32_t)(lane_q * 32) << 16) + (uint32_t)(BN + 16 * khalf);
        int s = 0, ph = 0;

        auto consume = [&](int kb, const float (&va)[16], const float4 (&vb)[Stager<BN, B_KMAJOR>::PASSES]) {
            if (kb >= NST) mbar_wait(&bar_free[s], (uint32_t)(ph ^ 1));   // the MMAs that read this stage have retired
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                hi[i] = to_tf32(va[i]);
                lo[i] = to_tf32(va[i] - __uint_as_float(hi[i]));
            }
            const uint32_t ta = a_lane + (uint32_t)(64 * s);
            tmem_st16(ta, hi);
            tmem_st16(ta + 32u, lo);
            const uint32_t st = smem_base + (uint32_t)s * STAGE_BYTES;
            sb.store(st, st + B_BYTES, vb);
            if (!A_KMAJOR && do_colsum) {
#pragma unroll
                for (int i = 0; i < 16; ++i) csum += va[i];
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&bar_full[s]);     // proxy fence: MMA lane (see the SS kernel)
            if (++s == NST) { s = 0; ph ^= 1; }
        };

        if (nslab > 0) { la.load(va0); sb.load(vb0); }
        for (int kb = 0; kb < nslab; kb += 2) {
            if (kb + 1 < nslab) { la.load(va1); sb.load(vb1); }
            consume(kb, va0, vb0);
            if (kb + 1 < nslab) {
                if (kb + 2 < nslab) { la.load(va0); sb.load(vb0); }
                consume(kb + 1, va1, vb1);
            }
        }

        // ---- epilogue ----
        TC_GSTAMP(2);
        if (nslab > 0) mbar_wait(&bar_acc, 0);
        TC_GSTAMP(3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!A_KMAJOR && do_colsum) {
            // a thread summed its row's 16 k-values per slab: fold the two k-halves of a row in fixed order
            float* red = reinterpret_cast<float*>(smem + 32 * 1024);
            red[khalf * BM + lane_q * 32 + lane] = csum;
            producers_sync();
            if (t < BM && m0 + t < p.M) p.colsum[(long long)blockIdx.z * p.M + m0 + t] = red[t] + red[BM + t];
        }
        const int col_half = warp >> 2;
        float* Cbase = p.C + (long long)blockIdx.z * p.c_split_stride;
        EpiArgs ea;
        ea.scr = smem_base + (uint32_t)warp * 4096u; ea.tmem_row = tmem_d + ((uint32_t)(lane_q * 32) << 16); ea.lane = lane;
        ea.c_begin = col_half * (BN / 2); ea.c_end = (col_half + 1) * (BN / 2);
        ea.m_base = m0 + lane_q * 32; ea.n0 = n0; ea.C = Cbase; ea.have_acc = nslab > 0; ea.dbg = dbg;
        if (p.H == nullptr) {
            if (p.act == GYMRL_ACT_TANH) epilogue_chunks<EPI_TANH>(p, ea);
            else if (p.act == GYMRL_ACT_RELU) epilogue_chunks<EPI_RELU>(p, ea);
            else epilogue_chunks<EPI_PLAIN>(p, ea);
        } else if (p.act == GYMRL_ACT_NONE && p.act_in == GYMRL_ACT_TANH) {
            epilogue_chunks<EPI_DTANH>(p, ea);
        } else if (p.act == GYMRL_ACT_NONE && p.act_in == GYMRL_ACT_RELU) {
            epilogue_chunks<EPI_DRELU>(p, ea);
        } else {
            epilogue_chunks<EPI_GENERIC>(p, ea);
        }
        TC_GSTAMP(4);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_GSTAMP(5);
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TMEM_COLS) : "memory");
    }
}

template <int BN, bool AK, bool BKM>
static int launch_tc_ts(const TcGemmParams& p, int splits, cudaStream_t s) {
    constexpr int NST = BN == 256 ? 3 : 4;
    constexpr size_t STAGES = (size_t)NST * 2 * BN * 128;
    constexpr size_t SMEM = (STAGES > 40 * 1024 ? STAGES : 40 * 1024) + 1024;   // >= 32 KB epilogue scratch + colsum fold
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_ts_kernel<BN, AK, BKM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) GYMRL_FAIL(GYMRL_ECUDA, "cudaFuncSetAttribute(smem=%zu) failed: %s", SMEM, cudaGetErrorString(e));
        configured = true;
    }
    dim3 grid(p.N / BN, ceil_div(p.M, 128), splits);
    gymrl_launch_pdl(gemm_tf32x3_ts_kernel<BN, AK, BKM>, grid, dim3(TC_LAUNCH_THREADS), SMEM, s, p);
    gymrl_count_launch();
    return GYMRL_OK;
}


// =====================================================================================================================
// Warp-specialised persistent variant ("ws"): TMA-fed weights, register-split activations, 2-CTA MMA, overlapped epilogue.
//
// What bounded the kernels above (ncu + per-CTA timelines, profiles/r1_tc_gemm_ncu_full_v12_summary.md): (1) the 128 B/clk
// shared-memory port — per 32-k slab the converters write 96 KB of hi/lo images and the tensor core reads 144 KB (BN = 256);
// (2) 45 % of a CTA's life is prologue + epilogue with the tensor pipe idle.  This variant
//   * pairs two CTAs on one 256 x 256 tile (tcgen05.mma.cta_group::2): each CTA stages its own 128 A rows and HALF of B, so per
//     slab a CTA writes 64 KB and its tensor core reads 96 KB -> 1250 clk of port time against 1536 clk of MMA;
//   * takes B (a WEIGHT matrix) from pre-split hi/lo images (wimages.cu) with one 3-D TMA box per stage
//     (cp.async.bulk.tensor, SWIZZLE_128B = the UMMA K-major layout, [hi | lo] adjacent), so only A — 1/3 of the old
//     conversion work — passes through registers, on 4 producer warps instead of 8;
//   * is persistent (one CTA pair per SM pair, static round-robin over tiles) with TWO accumulators in TMEM (2 x 256 columns)
//     and 8 dedicated epilogue warps: the epilogue of tile i (tcgen05.ld -> smem transpose -> bias / act -> coalesced stores)
//     runs while the MMAs of tile i + 1 are issued.
// Roles (15 warps): 0-3 A producers | 4 TMA | 5 MMA issue (leader CTA only) | 6 relay (proxy fence) | 7-14 epilogue.
// mbarriers: full[s] (producers' elected lanes + TMA expect_tx -> MMA; lives in the leader CTA, the peer arrives remotely),
// free[s] (tcgen05.commit multicast to both CTAs -> producers / TMA), acc_full[b] (commit multicast -> epilogue),
// acc_empty[b] (epilogue warps of both CTAs -> MMA).  All waits are bounded (trap on timeout).
// =====================================================================================================================
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

namespace ws {
constexpr int BM = 128;     // rows per CTA; the tile width BN (256 or 128 output columns) is a template parameter
constexpr int PROD_WARPS = 4, EPI_WARPS = 8;
constexpr int TMA_WARP = 4, MMA_WARP = 5, RELAY_WARP = 6, EPI_WARP0 = 7;
constexpr int THREADS = (EPI_WARP0 + EPI_WARPS) * 32;   // 448
constexpr int PROD_THREADS = PROD_WARPS * 32;           // 128
constexpr int A_PASSES = BM * 8 / PROD_THREADS;         // float4 per producer thread per slab (8)

template <int NCTA, int BN>
struct Cfg {
    static constexpr int B_ROWS = BN / NCTA;                           // B rows (output columns) this CTA stages
    static constexpr uint32_t A_BYTES = BM * 128;                      // one image of one 32-k slab
    static constexpr uint32_t B_BYTES = B_ROWS * 128;
    static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // BN = 256: 64 KB (pair) / 96 KB (single CTA); BN = 128: 48 / 64 KB
    static constexpr int NST = (int)((227u * 1024u - 33u * 1024u) / STAGE_BYTES) > 4 ? 4 : (int)((227u * 1024u - 33u * 1024u) / STAGE_BYTES);
    static constexpr uint32_t EPI_SCRATCH = EPI_WARPS * 4096;
    static constexpr size_t SMEM = (size_t)NST * STAGE_BYTES + EPI_SCRATCH + 1024;
    static constexpr int TMEM_COLS = 2 * BN;                           // two accumulators
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// Arrive on a barrier addressed in the shared::cluster window (own CTA or the leader).  Deliberately WITHOUT .release.cluster:
// ptxas implements that qualifier as MEMBAR.ALL.GPU, which waits for every outstanding global load of the thread — it serialised
// the producers' prefetch (measured 1.26 us per slab instead of the tensor pipe's 0.78 us).  What the consumer needs is ordered
// by other means: the staged bytes are made visible to the async proxy by fence.proxy.async in every writing thread (completed
// before the __syncwarp that precedes the elected arrive), exactly as in CUTLASS' ClusterBarrier::arrive(cta_id).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
template <int NCTA>
__device__ __forceinline__ void mbar_wait_ws(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        uint32_t done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
template <int NCTA>
__device__ __forceinline__ void umma_tf32_ws(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (NCTA == 2)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                     ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// tcgen05.commit: the barrier at this shared-memory offset receives one arrival once every MMA issued so far has retired —
// in both CTAs of the pair for NCTA == 2 (multicast mask 0b11)
template <int NCTA>
__device__ __forceinline__ void umma_commit_ws(uint64_t* bar) {
    if (NCTA == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One 3-D box {32 k, B_ROWS rows, 2 images} of the weight images -> this CTA's stage; the bytes are accounted on the LEADER
// CTA's full barrier (mbar_cluster_addr), as the MMA that consumes both CTAs' stages is issued there.
template <int NCTA>
__device__ __forceinline__ void tma_load_b(uint32_t smem_dst, const CUtensorMap* map, int k0, int row0, uint32_t mbar_cluster_addr) {
    if (NCTA == 2)
        asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(k0), "r"(row0), "r"(0), "r"(mbar_cluster_addr) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(k0), "r"(row0), "r"(0), "r"(mbar_cluster_addr) : "memory");
}
}  // namespace ws

template <int NCTA, int BN>
__global__ void __launch_bounds__(ws::THREADS, 1) gemm3x_ws_kernel(const TcGemmParams p, const __grid_constant__ CUtensorMap tmap_b,
                                                                   int tiles_m, int tiles_n) {
    const int dbg_flags = p.bn_max;   // developer probe (GYMRL_TC_WS_DBG): 1 = producers skip their st.shared, 2 = only the hi*hi MMA, 4 = no TMA
    using C = ws::Cfg<NCTA, BN>;
    constexpr int NST = C::NST;
    constexpr int BM = ws::BM;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar_full[NST];
    __shared__ __align__(8) uint64_t bar_free[NST];
    __shared__ __align__(8) uint64_t bar_acc_full[2];
    __shared__ __align__(8) uint64_t bar_acc_empty[2];
    __shared__ __align__(8) uint64_t bar_staged[NST];   // this CTA's producers -> its relay lane, which fences and arrives on the leader's full[s]
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t rank = NCTA == 2 ? ws::cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x / NCTA, n_clusters = gridDim.x / NCTA;
    const int n_tiles = tiles_m * tiles_n;
    const int nslab = p.K / 32;
    const uint32_t smem_base = smem_u32(smem);
    // developer probe (tools/ws_cta_times.py): 32 %globaltimer slots per CTA
    long long* const wdbg = g_tc_cta_dbg ? g_tc_cta_dbg + 32ll * blockIdx.x : nullptr;
#define WS_STAMP(slot) do { if (wdbg) wdbg[(slot)] = globaltimer_ns(); } while (0)
    if (t == 0) { WS_STAMP(0); if (wdbg) { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); wdbg[3] = smid; } }

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) {
            // leader: the relay lane of every CTA of the pair + the expect_tx arrival of the TMA lane
            mbar_init(&bar_full[s], NCTA + 1);
            mbar_init(&bar_free[s], 1);
            mbar_init(&bar_staged[s], ws::PROD_WARPS);
        }
        mbar_init(&bar_acc_full[0], 1); mbar_init(&bar_acc_full[1], 1);
        mbar_init(&bar_acc_empty[0], NCTA * ws::EPI_WARPS); mbar_init(&bar_acc_empty[1], NCTA * ws::EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == ws::MMA_WARP) {   // the same warp of both CTAs allocates (and later frees) all 512 TMEM columns: two accumulators
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(C::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(C::TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    if (warp == ws::TMA_WARP && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (NCTA == 2) ws::cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    pdl_wait();
    pdl_launch_dependents();
    if (t == 0) WS_STAMP(1);

    // barrier addresses in the LEADER CTA, for the remote arrivals of the peer
    uint32_t full_leader[NST], acc_empty_leader[2];
#pragma unroll
    for (int s = 0; s < NST; ++s) full_leader[s] = NCTA == 2 ? ws::mapa_u32(smem_u32(&bar_full[s]), 0u) : smem_u32(&bar_full[s]);
#pragma unroll
    for (int b = 0; b < 2; ++b) acc_empty_leader[b] = NCTA == 2 ? ws::mapa_u32(smem_u32(&bar_acc_empty[b]), 0u) : smem_u32(&bar_acc_empty[b]);

    if (warp < ws::PROD_WARPS) {
        // ===== A producers: global fp32 -> registers -> hi / lo tf32 images (K-major SWIZZLE_128B atoms) =====
        const int r0 = t >> 3, c = t & 7;
        const uint32_t soff = (uint32_t)(r0 >> 3) * 1024u + (uint32_t)(r0 & 7) * 128u + (uint32_t)((c ^ (r0 & 7)) << 4);
        const float* ptr[ws::A_PASSES];
        int l_tile = cluster_id, l_kb = 0;           // loader position (runs ahead of the consumer position)
        auto set_tile = [&](int tile) {
            const int m0 = ((tile / tiles_n) * NCTA + (int)rank) * BM;
#pragma unroll
            for (int i = 0; i < ws::A_PASSES; ++i) {
                const int m = min(m0 + r0 + (ws::PROD_THREADS / 8) * i, p.M - 1);   // clamped: those rows are never stored
                ptr[i] = p.A + (long long)m * p.lda + c * 4;
            }
        };
        auto load = [&](float4 (&v)[ws::A_PASSES]) {
            if (l_kb == 0) set_tile(l_tile);
#pragma unroll
            for (int i = 0; i < ws::A_PASSES; ++i) {
                v[i] = __ldg(reinterpret_cast<const float4*>(ptr[i]));
                ptr[i] += 32;
            }
            if (++l_kb == nslab) { l_kb = 0; l_tile += n_clusters; }
        };
        uint32_t it = 0;
        auto consume = [&](const float4 (&v)[ws::A_PASSES]) {
            const uint32_t s = it % NST, ph = (it / NST) & 1u;
            ws::mbar_wait_ws<NCTA>(&bar_free[s], ph ^ 1u);       // the MMAs that read this stage have retired (passes on first use)
            const uint32_t st = smem_base + s * C::STAGE_BYTES;
            if (t == 0 && it == 0) { if (wdbg) wdbg[8] = (long long)(__float_as_uint(v[0].x) & 0) + globaltimer_ns(); }   // first operands arrived
#pragma unroll
            for (int i = 0; i < ws::A_PASSES; ++i) {
                uint4 hi, lo;
                split4(v[i], hi, lo);
                if (!(dbg_flags & 1) || hi.x == 0x12345u) {
                    sts128(st + soff + 2048u * i, hi);                 // 16 rows per pass = two 8-row atoms
                    sts128(st + C::A_BYTES + soff + 2048u * i, lo);
                }
            }
            // No fence.proxy.async here: it compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the membar waits for the warp's
            // outstanding global loads — the prefetched slabs — so every slab cost a full memory round trip (1.1-1.3 us measured,
            // also in the round-1 kernels).  The generic -> async proxy fence is executed by the relay lane, which has no loads in
            // flight, after it has acquired this barrier; it then arrives on the leader's full[s].
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_staged[s]);
            if (t == 0) { if (it == 0) WS_STAMP(9); if ((int)it == nslab - 1) WS_STAMP(10); if ((int)it == 2 * nslab - 1) WS_STAMP(11); }
            ++it;
        };
        const int my_tiles = cluster_id < n_tiles ? (n_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;
        const long long total = (long long)my_tiles * nslab;
        // two register sets: the loads of slab g + 1 are in flight while slab g is converted (a third set was measured: no gain —
        // once the proxy fence left these warps the producers sustain 0.7 us per slab, the tensor pipe 1.0-1.1 us)
        float4 v0[ws::A_PASSES], v1[ws::A_PASSES];
        if (total > 0) load(v0);
        for (long long g = 0; g < total; g += 2) {
            if (g + 1 < total) load(v1);
            consume(v0);
            if (g + 1 < total) {
                if (g + 2 < total) load(v0);
                consume(v1);
            }
        }
    } else if (warp == ws::TMA_WARP) {
        // ===== TMA: one 3-D box {32 k, B_ROWS, hi|lo} of the weight images per stage (lane 0); all 32 lanes also pull the A
        // slabs the producers will read PF slabs later into L2 (prefetch.global.L2, four 128 B row pieces per lane and slab): with
        // L2-cold activations the producers' one-slab-ahead register prefetch (~1 us) barely covers the HBM latency =====
        {
            constexpr int PF = 4;
            int pf_tile = cluster_id, pf_kb = 0;
            auto pf_issue = [&]() {
                if (pf_tile < n_tiles) {
                    const int m0 = ((pf_tile / tiles_n) * NCTA + (int)rank) * BM;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int m = min(m0 + lane * 4 + i, p.M - 1);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.A + (long long)m * p.lda + pf_kb * 32));
                    }
                    if (++pf_kb == nslab) { pf_kb = 0; pf_tile += n_clusters; }
                }
            };
            const bool do_pf = !(dbg_flags & 8);
            if (do_pf) {
                // the first slabs are fetched by the producers themselves at kernel start: begin past them
                for (int i = 0; i < 2 && pf_tile < n_tiles; ++i) { if (++pf_kb == nslab) { pf_kb = 0; pf_tile += n_clusters; } }
            }
            uint32_t it = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters) {
                const int row0 = (tile % tiles_n) * BN + (int)rank * C::B_ROWS;
                for (int kb = 0; kb < nslab; ++kb, ++it) {
                    if (lane == 0) {
                        const uint32_t s = it % NST, ph = (it / NST) & 1u;
                        ws::mbar_wait_ws<NCTA>(&bar_free[s], ph ^ 1u);
                        if (dbg_flags & 4) { if (leader) mbar_arrive(&bar_full[s]); }
                        else {
                            if (leader) ws::mbar_arrive_expect_tx(&bar_full[s], (uint32_t)NCTA * 2u * C::B_BYTES);
                            ws::tma_load_b<NCTA>(smem_base + s * C::STAGE_BYTES + 2 * C::A_BYTES, &tmap_b, kb * 32, row0, full_leader[s]);
                        }
                        if (it == 0) WS_STAMP(16);
                        if ((int)it == nslab - 1) WS_STAMP(17);
                    }
                    __syncwarp();
                    if (do_pf) {
                        if (it == 0) {   // the look-ahead is opened AFTER the first weight box is in flight (it delayed that box by 0.6 us)
#pragma unroll 1
                            for (int i = 0; i < PF; ++i) pf_issue();
                        }
                        pf_issue();
                    }
                }
            }
        }
        __syncwarp();    // reconverge before the (aligned) cluster barrier at the end
    } else if (warp == ws::MMA_WARP) {
        // ===== MMA issue: one lane of the leader CTA drives the tensor cores of both CTAs =====
        if (leader && lane == 0) {
            constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * NCTA) >> 4) << 24);
            uint32_t it = 0, acc_it = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++acc_it) {
                const uint32_t b = acc_it & 1u, aph = (acc_it >> 1) & 1u;
                ws::mbar_wait_ws<NCTA>(&bar_acc_empty[b], aph ^ 1u);      // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_d + b * (uint32_t)BN;
                for (int kb = 0; kb < nslab; ++kb, ++it) {
                    const uint32_t s = it % NST, ph = (it / NST) & 1u;
                    ws::mbar_wait_ws<NCTA>(&bar_full[s], ph);
                    if (it == 0) WS_STAMP(4);
                    if ((int)it == nslab) WS_STAMP(6);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_base + s * C::STAGE_BYTES;
                    const uint32_t ah = st, al = st + C::A_BYTES, bh = st + 2 * C::A_BYTES, bl = bh + C::B_BYTES;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint64_t dah = make_desc(ah + j * 32, 16, 1024), dal = make_desc(al + j * 32, 16, 1024);
                        const uint64_t dbh = make_desc(bh + j * 32, 16, 1024), dbl = make_desc(bl + j * 32, 16, 1024);
                        if (!(dbg_flags & 2)) {
                            ws::umma_tf32_ws<NCTA>(d, dal, dbh, IDESC, (kb | j) ? 1u : 0u);
                            ws::umma_tf32_ws<NCTA>(d, dah, dbl, IDESC, 1u);
                            ws::umma_tf32_ws<NCTA>(d, dah, dbh, IDESC, 1u);
                        } else {
                            ws::umma_tf32_ws<NCTA>(d, dah, dbh, IDESC, (kb | j) ? 1u : 0u);
                        }
                    }
                    ws::umma_commit_ws<NCTA>(&bar_free[s]);
                }
                ws::umma_commit_ws<NCTA>(&bar_acc_full[b]);
                if (acc_it == 0) WS_STAMP(5);
                if (acc_it == 1) WS_STAMP(7);
            }
        }
        __syncwarp();
    } else if (warp == ws::RELAY_WARP) {
        // ===== relay: wait for this CTA's producers, make their st.shared visible to this SM's async proxy (tensor core), then
        // arrive on the LEADER's full barrier (for the peer CTA: remotely; the leader's MMA makes this SM read this CTA's stage)
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = cluster_id; tile < n_tiles; tile += n_clusters)
                for (int kb = 0; kb < nslab; ++kb, ++it) {
                    const uint32_t s = it % NST, ph = (it / NST) & 1u;
                    ws::mbar_wait_ws<NCTA>(&bar_staged[s], ph);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    ws::mbar_arrive_cluster(full_leader[s]);
                }
        }
        __syncwarp();
    } else {
        // ===== epilogue warps: TMEM -> registers -> smem transpose -> bias / act / act' -> coalesced global =====
        const int ew = warp - ws::EPI_WARP0;
        const int lane_q = warp & 3, col_half = ew >> 2;     // a warp reaches TMEM lanes 32 * (warp id % 4) .. + 31 only
        uint32_t acc_it = 0;
        for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++acc_it) {
            const uint32_t b = acc_it & 1u, aph = (acc_it >> 1) & 1u;
            const int m0 = ((tile / tiles_n) * NCTA + (int)rank) * BM, n0 = (tile % tiles_n) * BN;
            ws::mbar_wait_ws<NCTA>(&bar_acc_full[b], aph);
            if (ew == 0 && lane == 0 && acc_it < 2) WS_STAMP(12 + 2 * acc_it);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            EpiArgs ea;
            ea.scr = smem_base + NST * C::STAGE_BYTES + (uint32_t)ew * 4096u;
            ea.tmem_row = tmem_d + b * (uint32_t)BN + ((uint32_t)(lane_q * 32) << 16);
            ea.lane = lane;
            ea.c_begin = col_half * (BN / 2); ea.c_end = (col_half + 1) * (BN / 2);
            ea.m_base = m0 + lane_q * 32; ea.n0 = n0; ea.C = p.C; ea.have_acc = true; ea.dbg = nullptr;
            if (p.H == nullptr) {
                if (p.act == GYMRL_ACT_TANH) epilogue_chunks<EPI_TANH>(p, ea);
                else if (p.act == GYMRL_ACT_RELU) epilogue_chunks<EPI_RELU>(p, ea);
                else epilogue_chunks<EPI_PLAIN>(p, ea);
            } else if (p.act == GYMRL_ACT_NONE && p.act_in == GYMRL_ACT_TANH) {
                epilogue_chunks<EPI_DTANH>(p, ea);
            } else if (p.act == GYMRL_ACT_NONE && p.act_in == GYMRL_ACT_RELU) {
                epilogue_chunks<EPI_DRELU>(p, ea);
            } else {
                epilogue_chunks<EPI_GENERIC>(p, ea);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) ws::mbar_arrive_cluster(acc_empty_leader[b]);
            if (ew == 0 && lane == 0 && acc_it < 2) WS_STAMP(13 + 2 * acc_it);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (NCTA == 2) ws::cluster_sync_all(); else __syncthreads();
    if (t == 0) WS_STAMP(2);
    if (warp == ws::MMA_WARP) {
        if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(C::TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(C::TMEM_COLS) : "memory");
    }
}

// ---- ws host side -----------------------------------------------------------------------------------------------
#include "wimages.cuh"
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return reinterpret_cast<PFN_encodeTiled>(f);
    }();
    return fn;
}

// 0 = off, 1 = single-CTA tiles (debug / A-B), 2 = CTA pairs (default)
static int ws_mode() {
    static const int m = [] { const char* e = getenv("GYMRL_TC_WS"); return e ? atoi(e) : 2; }();
    return m;
}

static unsigned long long g_ws_launches = 0;
extern "C" unsigned long long gymrl_debug_ws_launches(void) { return g_ws_launches; }   // developer probe (tests: the ws path really ran)

template <int NCTA, int BN>
static int launch_ws(const TcGemmParams& p, const float* b_hi, long long img_stride, int b_rows_total, cudaStream_t s) {
    using C = ws::Cfg<NCTA, BN>;
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) GYMRL_FAIL(GYMRL_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)p.K, (cuuint64_t)b_rows_total, 2};
    const cuuint64_t strides[2] = {(cuuint64_t)p.K * sizeof(float), (cuuint64_t)img_stride * sizeof(float)};
    const cuuint32_t box[3] = {32, (cuuint32_t)C::B_ROWS, 2};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(b_hi), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) GYMRL_FAIL(GYMRL_ECUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm3x_ws_kernel<NCTA, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
        if (e != cudaSuccess) GYMRL_FAIL(GYMRL_ECUDA, "cudaFuncSetAttribute(smem=%zu) failed: %s", C::SMEM, cudaGetErrorString(e));
        configured = true;
    }
    const int tiles_m = ceil_div(p.M, ws::BM * NCTA), tiles_n = p.N / BN;
    const int n_tiles = tiles_m * tiles_n;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(ws::THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (NCTA == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    // persistent grid = the number of clusters that can be CO-RESIDENT (not SMs / 2: a GPC with an odd number of usable TPC slots
    // leaves SMs that cannot host a pair, and a cluster that does not fit waits for a whole wave of persistent CTAs to finish)
    static int max_clusters = 0;
    if (max_clusters == 0) {
        cfg.gridDim = dim3(GYMRL_NUM_SMS / NCTA * NCTA);
        cfg.attrs = attr; cfg.numAttrs = na;
        int n = 0;
        if (NCTA == 2 && cudaOccupancyMaxActiveClusters(&n, gemm3x_ws_kernel<NCTA, BN>, &cfg) == cudaSuccess && n > 0) max_clusters = n;
        else max_clusters = GYMRL_NUM_SMS / NCTA;
        (void)cudaGetLastError();
        if (getenv("GYMRL_TC_VERBOSE")) fprintf(stderr, "[gymrl] gemm3x_ws_kernel<%d, %d>: %d co-resident clusters\n", NCTA, BN, max_clusters);
    }
    int clusters = max_clusters < n_tiles ? max_clusters : n_tiles;
    cfg.gridDim = dim3(clusters * NCTA);
    if (gymrl_pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    TcGemmParams pk = p;
    static const int dbg = [] { const char* e = getenv("GYMRL_TC_WS_DBG"); return e ? atoi(e) : 0; }();
    pk.bn_max = dbg;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm3x_ws_kernel<NCTA, BN>, pk, map, tiles_m, tiles_n);
    if (e != cudaSuccess) GYMRL_FAIL(GYMRL_ECUDA, "launch of gemm3x_ws_kernel<%d, %d> failed: %s", NCTA, BN, cudaGetErrorString(e));
    gymrl_count_launch();
    ++g_ws_launches;
    return GYMRL_OK;
}

// The warp-specialised kernel takes the GEMM when B is a registered weight matrix (pre-split images exist), A is K-major and
// ungathered, the output is at least 256 columns wide in whole tiles and there are enough tiles to occupy the chip.
static bool ws_try_launch(const TcGemmParams& p, bool a_kmajor, bool b_kmajor, int splits, cudaStream_t s, int* rc) {
    const int mode = ws_mode();
    if (mode == 0 || !a_kmajor || splits != 1 || p.a_rows || p.b_rows || p.N % 128 != 0 || p.K % 32 != 0 || p.K < 64) return false;
    if (p.ldb != (b_kmajor ? p.K : p.N)) return false;
    const int ncta = mode == 1 ? 1 : 2;
    // Measured (tools/ws_gemm_bench.py, M = 16384, in a CUDA graph): with >= 2 tiles per CTA pair the overlapped epilogue wins
    // (N = 512 forward: 32.3 us vs 35.9 us); with a single tile per pair nothing overlaps and the one-tile-per-CTA kernel's
    // shorter set-up wins (N = 256: 20.9 vs 19.6 us).  GYMRL_TC_WS_MIN_TILES overrides the threshold (A/B runs).
    // 128-wide pair tiles (BN = 128) serve the narrow layers of long batches (PPO-full: 131072 x 128 x 128 = 512 tiles, where the
    // one-tile-per-CTA kernel spends most of a CTA's life in set-up and epilogue): GYMRL_TC_WS_BN128_MIN_TILES, default 200.
    const char* mt_env = getenv("GYMRL_TC_WS_MIN_TILES");     // read per call: the test-suite flips it inside one process
    const long long min_tiles = mt_env ? atoll(mt_env) : 0ll;
    const long long need = min_tiles > 0 ? min_tiles : (ncta == 2 ? 100 : 200);
    const char* mt128_env = getenv("GYMRL_TC_WS_BN128_MIN_TILES");
    const long long need128 = mt128_env ? atoll(mt128_env) : 200ll;
    const long long rows_t = (long long)ceil_div(p.M, ws::BM * ncta);
    int bn = 0;
    if (p.N % 256 == 0 && rows_t * (p.N / 256) >= need) bn = 256;
    else if (ncta == 2 && need128 > 0 && rows_t * (p.N / 128) >= need128) bn = 128;
    if (bn == 0) return false;
    const float* hi = nullptr;
    long long stride = 0;
    // forward: B = W [N][K];  backward-input: B = W [K][N] read as W^T [N][K] from the transposed images
    if (!(b_kmajor ? wimg_lookup(p.B, p.N, p.K, false, &hi, &stride) : wimg_lookup(p.B, p.K, p.N, true, &hi, &stride))) return false;
    if (bn == 128) *rc = launch_ws<2, 128>(p, hi, stride, p.N, s);
    else *rc = ncta == 2 ? launch_ws<2, 256>(p, hi, stride, p.N, s) : launch_ws<1, 256>(p, hi, stride, p.N, s);
    return true;
}

// ---- host side ------------------------------------------------------------------------------------------
template <int BN, bool AK, bool BKM>
static int launch_tc(const TcGemmParams& p, int splits, cudaStream_t s) {
    constexpr size_t SMEM = 2 * (2 * 128 * 128 + 2 * BN * 128) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, AK, BKM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) GYMRL_FAIL(GYMRL_ECUDA, "cudaFuncSetAttribute(smem=%zu) failed: %s", SMEM, cudaGetErrorString(e));
        configured = true;
    }
    dim3 grid(p.N / BN, ceil_div(p.M, 128), splits);
    gymrl_launch_pdl(gemm_tf32x3_kernel<BN, AK, BKM>, grid, dim3(TC_LAUNCH_THREADS), SMEM, s, p);
    gymrl_count_launch();
    return GYMRL_OK;
}

static inline bool al16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

// Shape / alignment gate for the tensor-core path (anything else stays on the FFMA kernel in linear.cu).
bool tc_gemm_supported(const TcGemmParams& p, bool a_kmajor, bool b_kmajor) {
    if (p.N % 64 != 0 || p.K % 32 != 0 || p.k_chunk % 32 != 0) return false;
    if (!al16(p.A) || !al16(p.B) || !al16(p.C) || (p.lda & 3) || (p.ldb & 3) || (p.ldc & 3)) return false;
    if (!a_kmajor && (p.M % 4 != 0)) return false;
    if (p.H && (!al16(p.H) || (p.ldh & 3))) return false;
    if (p.bias && !al16(p.bias)) return false;
    return p.M >= 128;
}

int tc_gemm_launch(const TcGemmParams& p, bool a_kmajor, bool b_kmajor, int splits, cudaStream_t s) {
    {
        int rc = GYMRL_OK;
        if (ws_try_launch(p, a_kmajor, b_kmajor, splits, s, &rc)) return rc;
    }
    // widest tile that still gives (nearly) every SM a CTA: the rollout's M = 4096 forward would otherwise run 32 CTAs
    // on 148 SMs.  (tc_gemm_supported guarantees N % 64 == 0.)
    const long long mt = (long long)ceil_div(p.M, 128) * splits;
    const int bn_max = p.bn_max > 0 ? p.bn_max : 256;
    int bn = 64;
    if (bn_max >= 256 && p.N % 256 == 0 && mt * (p.N / 256) >= 120) bn = 256;
    else if (bn_max >= 128 && p.N % 128 == 0 && mt * (p.N / 128) >= 120) bn = 128;
    // A through tensor memory (TS kernel) for MN-major A, i.e. the dW GEMMs: there a thread's 16 k-values of its row are
    // coalesced loads and the variant measured 8 % faster (26.9 -> 24.7 us at 512 x 256 x 16384).  For K-major A the row-owned
    // 64 B loads touch 32 lines per instruction and it measured 10 % slower than the SS kernel (17.2 -> 18.9 us at
    // 16384 x 256 x 256), so those stay on the shared-memory path.  GYMRL_TC_ATMEM = 0 / 1 / 2: never / MN-major only / always.
    static const int atmem_env = [] { const char* e = getenv("GYMRL_TC_ATMEM"); return e ? atoi(e) : 1; }();
    const bool atmem = atmem_env == 2 ? (a_kmajor || p.a_rows == nullptr) : (atmem_env == 1 && !a_kmajor && p.a_rows == nullptr);
#define TC_DISPATCH(AK, BKM)                                                                                   \
    switch (bn) {                                                                                              \
        case 256: return atmem ? launch_tc_ts<256, AK, BKM>(p, splits, s) : launch_tc<256, AK, BKM>(p, splits, s); \
        case 128: return atmem ? launch_tc_ts<128, AK, BKM>(p, splits, s) : launch_tc<128, AK, BKM>(p, splits, s); \
        default: return atmem ? launch_tc_ts<64, AK, BKM>(p, splits, s) : launch_tc<64, AK, BKM>(p, splits, s);    \
    }
    if (a_kmajor && b_kmajor) { TC_DISPATCH(true, true) }
    if (a_kmajor && !b_kmajor) { TC_DISPATCH(true, false) }
    if (!a_kmajor && !b_kmajor) { TC_DISPATCH(false, false) }
#undef TC_DISPATCH
    GYMRL_FAIL(GYMRL_EINVAL, "unsupported operand major combination");
}
