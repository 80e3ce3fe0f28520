// common.cuh — shared device/host helpers for the gymrl_b200 CUDA library (sm_100a only).
//
// Nothing in here is part of the public ABI; see include/gymrl.h for that.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include <string.h>

#include "../../include/gymrl.h"

// ----------------------------------------------------------------------------------------------
// Error convention: every extern "C" entry point returns 0 or a negative GYMRL_E* code and never
// throws. The message is kept thread-local and read back through gymrl_last_error().
// ----------------------------------------------------------------------------------------------
void gymrl_set_error(const char* fmt, ...);

#define GYMRL_FAIL(code, ...)            \
    do {                                 \
        gymrl_set_error(__VA_ARGS__);    \
        return (code);                   \
    } while (0)

#define GYMRL_REQUIRE(cond, ...)                         \
    do {                                                 \
        if (!(cond)) GYMRL_FAIL(GYMRL_EINVAL, __VA_ARGS__); \
    } while (0)

#define GYMRL_CUDA(expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            GYMRL_FAIL(GYMRL_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                       __FILE__, __LINE__);                                                \
    } while (0)

// After a kernel launch: catches bad launch configs without synchronising (capture-safe).
#define GYMRL_LAUNCH_CHECK(name)                                                         \
    do {                                                                                 \
        cudaError_t _e = cudaGetLastError();                                             \
        if (_e != cudaSuccess)                                                           \
            GYMRL_FAIL(GYMRL_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

#define GYMRL_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids of streaming kernels are sized from this

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may be
// scheduled while its predecessor in the stream is still draining; it must execute pdl_wait() (griddepcontrol.wait: blocks
// until the predecessor grid has completed and its writes are visible) before it touches anything the predecessor wrote
// or still reads.  Every kernel launched through GYMRL_LAUNCH_PDL calls pdl_wait() as its first statement and
// pdl_launch_dependents() right after, so the only thing that overlaps is launch latency / CTA scheduling (~1-2 us per
// kernel boundary, x13 kernels x 320 minibatches per PPO update).  Without the attribute both instructions are no-ops.
// Opt-in with GYMRL_PDL=1: on B200 it measured neutral for the epoch graph (reduce.cu: gymrl_pdl_enabled), so launches are
// fully serialised by default.
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool gymrl_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t gymrl_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gymrl_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ----------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11). key = (seed_lo, seed_hi),
// counter = (entity, draw, stream, aux). The oracle carries an independent restatement
// (oracle/philox.py, oracle/lunar_lander.c) so device and CPU draw identical numbers.
// ----------------------------------------------------------------------------------------------
enum PhiloxStream : uint32_t {
    PHILOX_ENV_RESET = 0,    // per-episode reset draws            (entity = global env id, draw = episode*8+j)
    PHILOX_ENV_STEP = 1,     // per-step env noise (LunarLander)   (entity = global env id, draw = env step counter)
    PHILOX_ACTION = 2,       // action sampling noise              (entity = global env id, draw = call counter)
    PHILOX_PERMUTE = 3,      // minibatch permutation keys
    PHILOX_REPLAY = 4,       // replay / PER sampling
    PHILOX_NOISYNET = 5,     // NoisyLinear factorised noise
    PHILOX_UPDATE = 6,       // noise drawn inside update() (SAC rsample, TD3 smoothing)
};

struct u32x4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t philox_mulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = philox_mulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = philox_mulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    u32x4 out = {c0, c1, c2, c3};
    return out;
}

__host__ __device__ __forceinline__ u32x4 philox_draw(uint64_t seed, uint64_t entity, uint32_t draw, uint32_t stream) {
    return philox4x32_10((uint32_t)entity, draw, stream, (uint32_t)(entity >> 32), (uint32_t)seed,
                         (uint32_t)(seed >> 32));
}

// 53-bit uniform in [0,1): same construction as numpy's Generator.random().
__host__ __device__ __forceinline__ double u01_f64(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// 24-bit uniform in [0,1)
__host__ __device__ __forceinline__ float u01_f32(uint32_t a) { return (float)(a >> 8) * (1.0f / 16777216.0f); }
// 24-bit uniform in (0,1]  (safe for log)
__host__ __device__ __forceinline__ float u01_open0_f32(uint32_t a) {
    return ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// ----------------------------------------------------------------------------------------------
// Warp / block reductions
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

// Block-wide sum; result valid in thread 0. `scratch` must hold >= 32 elements of T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? scratch[threadIdx.x] : T(0);
    if (wid == 0) v = warp_sum(v);
    return v;
}
// Deterministic cross-block accumulation (replaces float / double atomicAdd of per-block sums, whose order — and so the last
// bits of the result — changed from run to run): every block publishes its K partial sums, the last block to arrive (ticket
// counter) folds all of them in block order (thread-strided, then the fixed block tree) and adds the totals to out[0..K).
// `v` must be valid in thread 0.  The scratch is one static buffer per translation unit: launches that use it are serialised
// by the stream they share (the library drives one stream per process).  Call from ALL threads of every block.
#define GYMRL_FOLD_MAX_BLOCKS 4096
static __device__ double g_fold_scratch[GYMRL_FOLD_MAX_BLOCKS * 8];
static __device__ unsigned int g_fold_ticket = 0;
template <int K, typename OutT>
__device__ __forceinline__ void ordered_block_accumulate(const double (&v)[K], OutT* out, double* scratch_smem /* >= 32 doubles */) {
    static_assert(K <= 8, "at most 8 values per block");
    __shared__ int s_last;
    const int nb = gridDim.x;
    if (nb == 1) {
        if (threadIdx.x == 0)
            for (int j = 0; j < K; ++j) out[j] = (OutT)((double)out[j] + v[j]);
        return;
    }
    if (threadIdx.x == 0) {
        for (int j = 0; j < K; ++j) g_fold_scratch[blockIdx.x * K + j] = v[j];
        __threadfence();
        const unsigned int t = atomicAdd(&g_fold_ticket, 1u);
        s_last = (t == (unsigned int)nb - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // all K values in one pass: thread-strided partial sums over the blocks (fixed order), one warp tree each, then the warps in
    // order — two block barriers in total (a block_sum per value cost ~1 us each in the last block of the fused-heads kernel)
    double a[K];
#pragma unroll
    for (int j = 0; j < K; ++j) a[j] = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x)
#pragma unroll
        for (int j = 0; j < K; ++j) a[j] += *((volatile double*)&g_fold_scratch[b * K + j]);
    __shared__ double s_part[32][K];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        a[j] = warp_sum(a[j]);
        if (lane == 0) s_part[wid][j] = a[j];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double t = 0.0;
        for (int w = 0; w < nw; ++w) t += s_part[w][threadIdx.x];
        out[threadIdx.x] = (OutT)((double)out[threadIdx.x] + t);
    }
    (void)scratch_smem;
    if (threadIdx.x == 0) g_fold_ticket = 0;
}

// tanh with fp32-grade accuracy (|rel err| < ~5e-7) in ~10 instructions: odd polynomial near 0, 1 - 2/(e^{2|x|}+1) elsewhere.
// (libdevice tanhf costs ~40 dependent instructions per element, which made the 4-warp epilogue the bottleneck.)
__device__ __forceinline__ float exp2f_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float tanh_fast(float x) {
    // branch-free: both forms are evaluated and selected (a per-element branch diverges inside almost every warp and
    // its BSSY/BSYNC overhead cost more than the ~7 extra instructions)
    const float ax = fabsf(x);
    const float x2 = x * x;
    float p = 62.0f / 2835.0f;
    p = fmaf(p, x2, -17.0f / 315.0f);
    p = fmaf(p, x2, 2.0f / 15.0f);
    p = fmaf(p, x2, -1.0f / 3.0f);
    const float small = fmaf(x * x2, p, x);
    const float e = exp2f_approx(ax * 2.885390081777927f);   // e^{2|x|}
    const float big = copysignf(fmaf(-2.0f, rcp_approx(e + 1.0f), 1.0f), x);
    return ax < 0.25f ? small : big;
}
#endif
