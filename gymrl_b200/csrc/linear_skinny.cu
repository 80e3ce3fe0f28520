// linear_skinny.cu — dense layers whose GEMM degenerates: tiny fan-out (policy/value/Q heads, N <= 16) or tiny fan-in
// (the observation layer, K <= 16).  A 64x64 tile is >90 % padding for these shapes (they cost ~30 % of the update in
// the round-1 launch list, profiles/r1_launches_ffma.md), while the work is a bandwidth-bound sweep over the [M, 256]
// activation: one pass, coalesced, weights in shared memory.
//   skinny-N forward   y[M][N]  = act(x[M][K] W[N][K]^T + b)        warp per row, lanes split K, shuffle reduce
//   skinny-N backward  dx[M][K] = (dy[M][N] W[N][K]) * act'(h)       thread per 4 outputs, W in smem
//   skinny-N dW/db     dW[N][K] = sum_m dy[m][n] x[m][k]             thread per k, row chunks -> partials -> reduce
//   small-K forward    y[M][N]  = act(x[M][K] W[N][K]^T + b)         thread per output, W in smem (row gather on x)
//   small-K dW/db      dW[N][K] = sum_m dy[m][n] x[idx[m]][k]        thread per n, row chunks -> partials -> reduce
// fp32 FMA throughout (these are < 1 % of the flops).
#include "common.cuh"
#include <cstdlib>

void gymrl_count_launch(int n = 1);

#define SK_MAXN 16
#define SK_MAXK 16

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == GYMRL_ACT_TANH) return tanh_fast(v);   // same epilogue function as the tensor-core path (|rel err| < 5e-7)
    if (act == GYMRL_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}

// ---- skinny-N forward: one warp per row ------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(256) skinny_fwd_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ rows,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         float* __restrict__ y, int ldy, int M, int K, int act) {
    extern __shared__ float sw[];   // [N][K]
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int m = blockIdx.x * wpb + warp; m < M; m += gridDim.x * wpb) {
        const float* xr = x + (size_t)(rows ? rows[m] : m) * ldx;
        float acc[N];
#pragma unroll
        for (int n = 0; n < N; ++n) acc[n] = 0.f;
        for (int k = lane * 4; k < K; k += 128) {   // K % 4 == 0 (checked by the launcher)
            const float4 v = *reinterpret_cast<const float4*>(xr + k);
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const float4 ww = *reinterpret_cast<const float4*>(&sw[n * K + k]);
                acc[n] = fmaf(v.x, ww.x, fmaf(v.y, ww.y, fmaf(v.z, ww.z, fmaf(v.w, ww.w, acc[n]))));
            }
        }
#pragma unroll
        for (int n = 0; n < N; ++n) acc[n] = warp_sum(acc[n]);
        if (lane == 0) {
#pragma unroll
            for (int n = 0; n < N; ++n) y[(size_t)m * ldy + n] = apply_act(acc[n] + (b ? b[n] : 0.f), act);
        }
    }
}

template <int N>
static void launch_skinny_fwd(const float* x, int ldx, const int32_t* rows, const float* w, const float* b, float* y, int ldy, int M,
                              int K, int act, cudaStream_t s) {
    // 8 warps per block; one row per warp while that still fits one resident wave (the launch is latency-bound: a warp that
    // walks 8 rows pays 8 dependent L2 round trips — 14.8 us for 8192 x 256 x 3 — GYMRL_SKINNY_ROWS restores that for A/B runs)
    static const int rows_per_warp = [] { const char* e = getenv("GYMRL_SKINNY_ROWS"); return e && atoi(e) > 0 ? atoi(e) : 1; }();
    int blocks = ceil_div(M, 8 * rows_per_warp);
    if (blocks > GYMRL_NUM_SMS * 8) blocks = GYMRL_NUM_SMS * 8;
    skinny_fwd_kernel<N><<<blocks, 256, (size_t)N * K * sizeof(float), s>>>(x, ldx, rows, w, b, y, ldy, M, K, act);
}

bool skinny_forward_supported(const float* x, int ldx, int N, int K) {
    return N >= 1 && N <= 8 && K % 4 == 0 && K <= 2048 && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
}
int skinny_forward(const float* x, int ldx, const int32_t* rows, const float* w, const float* b, float* y, int ldy, int M, int N, int K,
                   int act, cudaStream_t s) {
    switch (N) {
        case 1: launch_skinny_fwd<1>(x, ldx, rows, w, b, y, ldy, M, K, act, s); break;
        case 2: launch_skinny_fwd<2>(x, ldx, rows, w, b, y, ldy, M, K, act, s); break;
        case 3: launch_skinny_fwd<3>(x, ldx, rows, w, b, y, ldy, M, K, act, s); break;
        case 4: launch_skinny_fwd<4>(x, ldx, rows, w, b, y, ldy, M, K, act, s); break;
        case 5: launch_skinny_fwd<5>(x, ldx, rows, w, b, y, ldy, M, K, act, s); break;
        case 6: launch_skinny_fwd<6>(x, ldx, rows, w, b, y, ldy, M, K, act, s); break;
        case 7: launch_skinny_fwd<7>(x, ldx, rows, w, b, y, ldy, M, K, act, s); break;
        default: launch_skinny_fwd<8>(x, ldx, rows, w, b, y, ldy, M, K, act, s); break;
    }
    gymrl_count_launch();
    return GYMRL_OK;
}

// ---- skinny-N backward input: dx[m][k..k+3] = (sum_n dy[m][n] W[n][k..]) * act'(h) -------------------------------------
__global__ void __launch_bounds__(256) skinny_bwd_input_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ w,
                                                               const float* __restrict__ h, int ldh, float* __restrict__ dx, int lddx,
                                                               int M, int N, int K, int act_in, int accumulate) {
    extern __shared__ float sw[];   // [N][K]
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) sw[i] = w[i];
    __syncthreads();
    const int kq = K >> 2;
    const long long total = (long long)M * kq;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(t / kq), k = (int)(t % kq) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int n = 0; n < N; ++n) {
            const float g = dy[(size_t)m * lddy + n];
            const float4 ww = *reinterpret_cast<const float4*>(&sw[n * K + k]);
            acc.x = fmaf(g, ww.x, acc.x); acc.y = fmaf(g, ww.y, acc.y); acc.z = fmaf(g, ww.z, acc.z); acc.w = fmaf(g, ww.w, acc.w);
        }
        if (h) {
            const float4 hv = *reinterpret_cast<const float4*>(h + (size_t)m * ldh + k);
            if (act_in == GYMRL_ACT_TANH) {
                acc.x *= (1.0f - hv.x * hv.x); acc.y *= (1.0f - hv.y * hv.y); acc.z *= (1.0f - hv.z * hv.z); acc.w *= (1.0f - hv.w * hv.w);
            } else if (act_in == GYMRL_ACT_RELU) {
                acc.x = hv.x > 0.f ? acc.x : 0.f; acc.y = hv.y > 0.f ? acc.y : 0.f; acc.z = hv.z > 0.f ? acc.z : 0.f; acc.w = hv.w > 0.f ? acc.w : 0.f;
            }
        }
        float4* dst = reinterpret_cast<float4*>(dx + (size_t)m * lddx + k);
        if (accumulate) { const float4 o = *dst; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
        *dst = acc;
    }
}

bool skinny_backward_input_supported(int N, int K, const float* h, int ldh, const float* dx, int lddx) {
    const bool al = ((reinterpret_cast<uintptr_t>(dx) & 15) == 0) && (lddx % 4 == 0) && (!h || (((reinterpret_cast<uintptr_t>(h) & 15) == 0) && ldh % 4 == 0));
    return N >= 1 && N <= SK_MAXN && K % 4 == 0 && N * K <= 8192 && al;
}
int skinny_backward_input(const float* dy, int lddy, const float* w, const float* h, int ldh, float* dx, int lddx, int M, int N, int K,
                          int act_in, int accumulate, cudaStream_t s) {
    long long blocks = ceil_div_ll((long long)M * (K >> 2), 256 * 4);
    if (blocks > GYMRL_NUM_SMS * 8) blocks = GYMRL_NUM_SMS * 8;
    if (blocks < 1) blocks = 1;
    skinny_bwd_input_kernel<<<(int)blocks, 256, (size_t)N * K * sizeof(float), s>>>(dy, lddy, w, h, ldh, dx, lddx, M, N, K, act_in, accumulate);
    gymrl_count_launch();
    return GYMRL_OK;
}

// ---- skinny-N dW / db: thread per k, block per row chunk, partial [chunk][N][K] (+ [chunk][N]) --------------------------
template <int N>
__global__ void __launch_bounds__(256) skinny_dw_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                                        const int32_t* __restrict__ rows, float* __restrict__ part_w,
                                                        float* __restrict__ part_b, int M, int K, int rows_per_chunk) {
    const int m_beg = blockIdx.y * rows_per_chunk, m_end = min(M, m_beg + rows_per_chunk);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float sdy[32][N];
    float acc[N], accb[N];
#pragma unroll
    for (int n = 0; n < N; ++n) { acc[n] = 0.f; accb[n] = 0.f; }
    for (int m0 = m_beg; m0 < m_end; m0 += 32) {
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * N; i += blockDim.x) {
            const int mm = m0 + i / N;
            sdy[i / N][i % N] = mm < m_end ? dy[(size_t)mm * lddy + (i % N)] : 0.f;
        }
        __syncthreads();
        const int lim = min(32, m_end - m0);
        if (lim == 32 && k < K) {
            // full group: issue the 32 row loads back to back (the sweep over x is bandwidth-bound; one load in
            // flight per thread left it latency-bound at 64 blocks)
            float xv[32];
#pragma unroll
            for (int r = 0; r < 32; ++r) xv[r] = x[(size_t)(rows ? rows[m0 + r] : (m0 + r)) * ldx + k];
#pragma unroll
            for (int r = 0; r < 32; ++r) {
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    acc[n] = fmaf(sdy[r][n], xv[r], acc[n]);
                    accb[n] += sdy[r][n];
                }
            }
        } else {
            for (int r = 0; r < lim; ++r) {
                const float xv = k < K ? x[(size_t)(rows ? rows[m0 + r] : (m0 + r)) * ldx + k] : 0.f;
#pragma unroll
                for (int n = 0; n < N; ++n) {
                    acc[n] = fmaf(sdy[r][n], xv, acc[n]);
                    accb[n] += sdy[r][n];
                }
            }
        }
    }
    if (k < K) {
#pragma unroll
        for (int n = 0; n < N; ++n) part_w[((size_t)blockIdx.y * N + n) * K + k] = acc[n];
    }
    if (part_b && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) part_b[(size_t)blockIdx.y * N + n] = accb[n];
    }
}

bool skinny_dw_supported(int N) { return N >= 1 && N <= 8; }
// workspace floats: chunks * (N*K + N)
int skinny_dw(const float* dy, int lddy, const float* x, int ldx, const int32_t* rows, float* part_w, float* part_b, int M, int N, int K,
              int chunks, cudaStream_t s) {
    const int rpc = ceil_div(M, chunks);
    dim3 grid(ceil_div(K, 256), chunks);
#define SK_DW(NN) skinny_dw_kernel<NN><<<grid, 256, 0, s>>>(dy, lddy, x, ldx, rows, part_w, part_b, M, K, rpc)
    switch (N) {
        case 1: SK_DW(1); break; case 2: SK_DW(2); break; case 3: SK_DW(3); break; case 4: SK_DW(4); break;
        case 5: SK_DW(5); break; case 6: SK_DW(6); break; case 7: SK_DW(7); break; default: SK_DW(8); break;
    }
#undef SK_DW
    gymrl_count_launch();
    return GYMRL_OK;
}

// ---- small-K forward: thread per output column, rows looped; the block's x rows staged in smem once ---------------------
// One gather per block (rows_per_block x K floats, two dependent L2 round trips when the rows are index-gathered), then every
// thread walks the rows for its column: smem broadcast reads, K FMAs, activation, one coalesced 1 KB store per row and block.
// (The first version staged 16 rows at a time: four gather latencies and eight barriers per block, 14.5 us at M = 16384.)
constexpr int kSmallKRowsMax = 64;
template <int K>
__global__ void __launch_bounds__(256) smallk_fwd_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ rows,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         float* __restrict__ y, int ldy, int M, int N, int act, int rows_per_block) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m_beg = blockIdx.y * rows_per_block, m_end = min(M, m_beg + rows_per_block);
    const int nrows = m_end - m_beg;
    __shared__ __align__(16) float sx[kSmallKRowsMax][K];
    for (int i = threadIdx.x; i < nrows * K; i += blockDim.x) {
        const int mm = m_beg + i / K;
        sx[i / K][i % K] = x[(size_t)(rows ? rows[mm] : mm) * ldx + (i % K)];
    }
    float wr[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wr[k] = n < N ? w[(size_t)n * K + k] : 0.f;
    const float bias = (b && n < N) ? b[n] : 0.f;
    __syncthreads();
    if (n >= N) return;
    float* yp = y + (size_t)m_beg * ldy + n;
    int r = 0;
    for (; r + 4 <= nrows; r += 4) {
        float acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            acc[q] = bias;
#pragma unroll
            for (int k = 0; k < K; ++k) acc[q] = fmaf(sx[r + q][k], wr[k], acc[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) yp[(size_t)(r + q) * ldy] = apply_act(acc[q], act);
    }
    for (; r < nrows; ++r) {
        float acc = bias;
#pragma unroll
        for (int k = 0; k < K; ++k) acc = fmaf(sx[r][k], wr[k], acc);
        yp[(size_t)r * ldy] = apply_act(acc, act);
    }
}

// Vector variant (N % 4 == 0, 16 B-aligned y / ldy): a thread owns 4 adjacent output columns (their 4 x K weights live in
// registers) and walks every 4th row of the block, so one broadcast read of an x row (K floats) feeds 4K FMAs and the result
// leaves as one STG.128.  The scalar kernel above spends 8 shared-memory reads per output element and is bound by the
// shared-memory pipe (13.7 us for 16384 x 256 x 8); this one is bound by the activation math and the 16 MB it writes.
template <int K>
__global__ void __launch_bounds__(256) smallk_fwd4_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ rows,
                                                          const float* __restrict__ w, const float* __restrict__ b,
                                                          float* __restrict__ y, int ldy, int M, int N, int act, int rows_per_block) {
    const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;      // a warp shares rg: its x-row reads are broadcasts
    const int n = (blockIdx.x * 64 + cg) * 4;
    pdl_wait();
    pdl_launch_dependents();
    const int m_beg = blockIdx.y * rows_per_block, m_end = min(M, m_beg + rows_per_block);
    const int nrows = m_end - m_beg;
    __shared__ __align__(16) float sx[kSmallKRowsMax][K == 3 ? 4 : K];
    constexpr int KP = K == 3 ? 4 : K;
    for (int i = threadIdx.x; i < nrows * K; i += blockDim.x) {
        const int mm = m_beg + i / K;
        sx[i / K][i % K] = x[(size_t)(rows ? rows[mm] : mm) * ldx + (i % K)];
    }
    // the block's 256 x K weights: coalesced into shared memory (a thread's 4 x K values are 16 K bytes apart from its
    // neighbour's: read straight from global they cost 32 lines per load instruction), group pitch 4K+1 -> conflict-free reads
    __shared__ float sw[64][4 * K + 1];
    const int n_blk = blockIdx.x * 256;
    for (int i = threadIdx.x; i < 256 * K; i += blockDim.x) {
        const int col = i / K;
        sw[col >> 2][(col & 3) * K + (i % K)] = (n_blk + col) < N ? w[(size_t)n_blk * K + i] : 0.f;
    }
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N && b) bias = *reinterpret_cast<const float4*>(b + n);
    __syncthreads();
    if (n >= N) return;
    float wr[4][K];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < K; ++k) wr[c][k] = sw[cg][c * K + k];
    (void)KP;
#pragma unroll 2
    for (int r = rg; r < nrows; r += 4) {
        float xr[K];
#pragma unroll
        for (int k = 0; k < K; ++k) xr[k] = sx[r][k];
        float o[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int k = 0; k < K; ++k) o[c] = fmaf(xr[k], wr[c][k], o[c]);
        *reinterpret_cast<float4*>(y + (size_t)(m_beg + r) * ldy + n) =
            make_float4(apply_act(o[0], act), apply_act(o[1], act), apply_act(o[2], act), apply_act(o[3], act));
    }
}

// GYMRL_SMALLK_VEC=0 keeps the scalar small-K kernels (A/B comparison)
static bool smallk_vec_enabled() {
    static const bool on = [] { const char* e = getenv("GYMRL_SMALLK_VEC"); return !(e && e[0] == '0'); }();
    return on;
}
static int smallk_rows_target() {
    static const int v = [] { const char* e = getenv("GYMRL_SMALLK_BLOCKS_PER_SM"); return e ? atoi(e) : 2; }();
    return v < 1 ? 1 : v;
}
bool smallk_forward_supported(int K) { return K == 3 || K == 4 || K == 8; }
int smallk_forward(const float* x, int ldx, const int32_t* rows, const float* w, const float* b, float* y, int ldy, int M, int N, int K,
                   int act, cudaStream_t s) {
    // two blocks per SM when M allows it (measured on B200 at 16384 x 256 x 8: 7.4 us with 2, 8.1 with 4, 8.6 with 8 blocks per
    // SM - the per-block weight staging outweighs the extra latency hiding), at most kSmallKRowsMax rows each
    int rpb = ceil_div(M, smallk_rows_target() * GYMRL_NUM_SMS / ceil_div(N, 256));
    rpb = rpb < 8 ? 8 : (rpb > kSmallKRowsMax ? kSmallKRowsMax : (rpb + 7) / 8 * 8);
    dim3 grid(ceil_div(N, 256), ceil_div(M, rpb));
    const bool vec = smallk_vec_enabled() && (N % 4 == 0) && (ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) &&
                     (!b || (reinterpret_cast<uintptr_t>(b) & 15) == 0);
    if (vec) {
        if (K == 3) gymrl_launch_pdl(smallk_fwd4_kernel<3>, grid, dim3(256), 0, s, x, ldx, rows, w, b, y, ldy, M, N, act, rpb);
        else if (K == 4) gymrl_launch_pdl(smallk_fwd4_kernel<4>, grid, dim3(256), 0, s, x, ldx, rows, w, b, y, ldy, M, N, act, rpb);
        else gymrl_launch_pdl(smallk_fwd4_kernel<8>, grid, dim3(256), 0, s, x, ldx, rows, w, b, y, ldy, M, N, act, rpb);
    } else if (K == 3) smallk_fwd_kernel<3><<<grid, 256, 0, s>>>(x, ldx, rows, w, b, y, ldy, M, N, act, rpb);
    else if (K == 4) smallk_fwd_kernel<4><<<grid, 256, 0, s>>>(x, ldx, rows, w, b, y, ldy, M, N, act, rpb);
    else smallk_fwd_kernel<8><<<grid, 256, 0, s>>>(x, ldx, rows, w, b, y, ldy, M, N, act, rpb);
    gymrl_count_launch();
    return GYMRL_OK;
}

// ---- small-K dW / db: thread per output unit n, K (+1) accumulators, row chunks -> partials ------------------------------------
template <int K>
__global__ void __launch_bounds__(256) smallk_dw_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                                        const int32_t* __restrict__ rows, float* __restrict__ part_w,
                                                        float* __restrict__ part_b, int M, int N, int rows_per_chunk) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int m_beg = blockIdx.y * rows_per_chunk, m_end = min(M, m_beg + rows_per_chunk);
    __shared__ float sx[32][K];
    float acc[K], accb = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    for (int m0 = m_beg; m0 < m_end; m0 += 32) {
        const int lim = min(32, m_end - m0);
        // the dY column loads do not depend on the staged x rows: issue them before the barriers so that their L2 round
        // trip overlaps the (index -> row) gather instead of following it
        float g[32];
        if (n < N && lim == 32) {
#pragma unroll
            for (int r = 0; r < 32; ++r) g[r] = dy[(size_t)(m0 + r) * lddy + n];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) {
            const int mm = m0 + i / K;
            sx[i / K][i % K] = mm < m_end ? x[(size_t)(rows ? rows[mm] : mm) * ldx + (i % K)] : 0.f;
        }
        __syncthreads();
        if (n < N) {
            if (lim == 32) {
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    accb += g[r];
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] = fmaf(g[r], sx[r][k], acc[k]);
                }
            } else {
                for (int r = 0; r < lim; ++r) {
                    const float gv = dy[(size_t)(m0 + r) * lddy + n];
                    accb += gv;
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] = fmaf(gv, sx[r][k], acc[k]);
                }
            }
        }
    }
    if (n < N) {
#pragma unroll
        for (int k = 0; k < K; ++k) part_w[((size_t)blockIdx.y * N + n) * K + k] = acc[k];
        if (part_b) part_b[(size_t)blockIdx.y * N + n] = accb;
    }
}

// Vector variant (N % 4 == 0, 16 B-aligned dy / lddy): a thread owns 4 adjacent output units (4 x K + 4 accumulators), the
// block's four row groups walk every 4th row of the chunk (one LDG.128 of dY and one broadcast read of the x row feed 4K
// FMAs), and are folded through shared memory in a fixed order before the partial is written.
constexpr int kSmallKDwRows = 64;   // rows per chunk
template <int K>
__global__ void __launch_bounds__(256) smallk_dw4_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                                         const int32_t* __restrict__ rows, float* __restrict__ part_w,
                                                         float* __restrict__ part_b, int M, int N) {
    const int cg = threadIdx.x & 63, rg = threadIdx.x >> 6;
    const int n = (blockIdx.x * 64 + cg) * 4;
    pdl_wait();
    pdl_launch_dependents();
    const int m_beg = blockIdx.y * kSmallKDwRows, m_end = min(M, m_beg + kSmallKDwRows);
    const int nrows = m_end - m_beg;
    __shared__ __align__(16) float sx[kSmallKDwRows][K == 3 ? 4 : K];
    __shared__ float red[3][64][4 * K + 4 + 1];   // +1: odd pitch, conflict-free column-group writes
    // dY loads first: they do not depend on the staged x rows
    constexpr int RPT = kSmallKDwRows / 4;         // rows per thread
    float4 g[RPT];
    const bool ok = n < N;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rg + 4 * i;
        g[i] = (ok && r < nrows) ? *reinterpret_cast<const float4*>(dy + (size_t)(m_beg + r) * lddy + n) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = threadIdx.x; i < kSmallKDwRows * K; i += blockDim.x) {
        const int r = i / K, mm = m_beg + r;
        sx[r][i % K] = r < nrows ? x[(size_t)(rows ? rows[mm] : mm) * ldx + (i % K)] : 0.f;
    }
    __syncthreads();
    float acc[4][K], accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int k = 0; k < K; ++k) acc[c][k] = 0.f;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rg + 4 * i;
        float xr[K];
#pragma unroll
        for (int k = 0; k < K; ++k) xr[k] = sx[r][k];
        const float gv[4] = {g[i].x, g[i].y, g[i].z, g[i].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            accb[c] += gv[c];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[c][k] = fmaf(gv[c], xr[k], acc[c][k]);
        }
    }
    if (rg > 0) {
        float* q = red[rg - 1][cg];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < K; ++k) q[c * K + k] = acc[c][k];
            q[4 * K + c] = accb[c];
        }
    }
    __syncthreads();
    if (rg != 0 || !ok) return;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const float* q = red[s][cg];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[c][k] += q[c * K + k];
            accb[c] += q[4 * K + c];
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int k = 0; k < K; ++k) part_w[((size_t)blockIdx.y * N + n + c) * K + k] = acc[c][k];
        if (part_b) part_b[(size_t)blockIdx.y * N + n + c] = accb[c];
    }
}

bool smallk_dw_supported(int K) { return K == 3 || K == 4 || K == 8; }
// chunks the dW sweep will write for M rows (the caller sizes the partial buffers with it)
int smallk_dw_chunks(const float* dy, int lddy, int M, int N) {
    const bool vec = smallk_vec_enabled() && (N % 4 == 0) && (lddy % 4 == 0) && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
    if (vec && ceil_div(M, kSmallKDwRows) <= 512) return ceil_div(M, kSmallKDwRows);
    const int chunks = ceil_div(M, 32);
    return chunks > 512 ? 512 : chunks;
}
int smallk_dw(const float* dy, int lddy, const float* x, int ldx, const int32_t* rows, float* part_w, float* part_b, int M, int N, int K,
              int chunks, cudaStream_t s) {
    const bool vec = smallk_vec_enabled() && (N % 4 == 0) && (lddy % 4 == 0) && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
    if (vec && chunks == ceil_div(M, kSmallKDwRows)) {
        dim3 grid(ceil_div(N, 256), chunks);
        if (K == 3) gymrl_launch_pdl(smallk_dw4_kernel<3>, grid, dim3(256), 0, s, dy, lddy, x, ldx, rows, part_w, part_b, M, N);
        else if (K == 4) gymrl_launch_pdl(smallk_dw4_kernel<4>, grid, dim3(256), 0, s, dy, lddy, x, ldx, rows, part_w, part_b, M, N);
        else gymrl_launch_pdl(smallk_dw4_kernel<8>, grid, dim3(256), 0, s, dy, lddy, x, ldx, rows, part_w, part_b, M, N);
        gymrl_count_launch();
        return GYMRL_OK;
    }
    const int rpc = ceil_div(M, chunks);
    dim3 grid(ceil_div(N, 256), chunks);
    if (K == 3) smallk_dw_kernel<3><<<grid, 256, 0, s>>>(dy, lddy, x, ldx, rows, part_w, part_b, M, N, rpc);
    else if (K == 4) smallk_dw_kernel<4><<<grid, 256, 0, s>>>(dy, lddy, x, ldx, rows, part_w, part_b, M, N, rpc);
    else smallk_dw_kernel<8><<<grid, 256, 0, s>>>(dy, lddy, x, ldx, rows, part_w, part_b, M, N, rpc);
    gymrl_count_launch();
    return GYMRL_OK;
}

// ---- skinny-N fused backward: one sweep over x[M][K] yields dW, db partials AND dx = (dy W) * act'(x) ---------------------------
// The heads' input x is the previous layer's activation output, so it is both the dW operand and the h of act'(h).
// Warp per row (8 rows in flight per block), lane owns KC float4 column groups (K = 128 KC); W and the dW accumulators
// live in registers.  Block partials [chunk][N*K + N] are folded by reduce_pair_kernel (deterministic order).
template <int N, int KC>
__global__ void __launch_bounds__(256) skinny_bwd_fused_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx,
                                                               const float* __restrict__ w, float* __restrict__ part, float* __restrict__ dx,
                                                               int lddx, int M, int act_in, int accumulate_dx, int rows_per_chunk) {
    constexpr int K = 128 * KC;
    extern __shared__ float red[];   // [8 warps][N*K + N]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m_beg = blockIdx.x * rows_per_chunk, m_end = min(M, m_beg + rows_per_chunk);
    float4 wr[N][KC], acc[N][KC];
    float accb[N];
#pragma unroll
    for (int n = 0; n < N; ++n) {
        accb[n] = 0.f;
#pragma unroll
        for (int j = 0; j < KC; ++j) {
            wr[n][j] = *reinterpret_cast<const float4*>(w + (size_t)n * K + j * 128 + lane * 4);
            acc[n][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    for (int m = m_beg + warp; m < m_end; m += 8) {
        float g[N];
#pragma unroll
        for (int n = 0; n < N; ++n) g[n] = dy[(size_t)m * lddy + n];
        float4 xv[KC];
#pragma unroll
        for (int j = 0; j < KC; ++j) xv[j] = *reinterpret_cast<const float4*>(x + (size_t)m * ldx + j * 128 + lane * 4);
#pragma unroll
        for (int n = 0; n < N; ++n) accb[n] += g[n];
#pragma unroll
        for (int j = 0; j < KC; ++j) {
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int n = 0; n < N; ++n) {
                acc[n][j].x = fmaf(g[n], xv[j].x, acc[n][j].x); acc[n][j].y = fmaf(g[n], xv[j].y, acc[n][j].y);
                acc[n][j].z = fmaf(g[n], xv[j].z, acc[n][j].z); acc[n][j].w = fmaf(g[n], xv[j].w, acc[n][j].w);
                d.x = fmaf(g[n], wr[n][j].x, d.x); d.y = fmaf(g[n], wr[n][j].y, d.y);
                d.z = fmaf(g[n], wr[n][j].z, d.z); d.w = fmaf(g[n], wr[n][j].w, d.w);
            }
            if (dx) {
                if (act_in == GYMRL_ACT_TANH) {
                    d.x *= (1.0f - xv[j].x * xv[j].x); d.y *= (1.0f - xv[j].y * xv[j].y);
                    d.z *= (1.0f - xv[j].z * xv[j].z); d.w *= (1.0f - xv[j].w * xv[j].w);
                } else if (act_in == GYMRL_ACT_RELU) {
                    d.x = xv[j].x > 0.f ? d.x : 0.f; d.y = xv[j].y > 0.f ? d.y : 0.f;
                    d.z = xv[j].z > 0.f ? d.z : 0.f; d.w = xv[j].w > 0.f ? d.w : 0.f;
                }
                float4* dst = reinterpret_cast<float4*>(dx + (size_t)m * lddx + j * 128 + lane * 4);
                if (accumulate_dx) { const float4 o = *dst; d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w; }
                *dst = d;
            }
        }
    }
    // fold the 8 warps (fixed order) -> part[chunk][N*K + N]
    constexpr int STRIDE = N * K + N;
    constexpr int SSTRIDE = (STRIDE + 3) & ~3;   // keeps every warp's float4 stores 16 B-aligned
    float* mine = red + (size_t)warp * SSTRIDE;
#pragma unroll
    for (int n = 0; n < N; ++n) {
#pragma unroll
        for (int j = 0; j < KC; ++j) *reinterpret_cast<float4*>(mine + n * K + j * 128 + lane * 4) = acc[n][j];
        if (lane == 0) mine[N * K + n] = accb[n];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < STRIDE; i += 256) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) s += red[(size_t)wv * SSTRIDE + i];
        part[(size_t)blockIdx.x * STRIDE + i] = s;
    }
}

// out_w[i] (i < cnt_w) and out_b[i - cnt_w]: fold `splits` partial rows in a fixed order (deterministic).
// G = 8 split groups per output when there are many splits and few outputs (the skinny sweeps: 256-512 row chunks of a
// few hundred outputs, where one thread per output would walk the chunks serially: 20-40 us measured), else G = 1.
template <int G>
__global__ void __launch_bounds__(256) reduce_pair_kernel(const float* __restrict__ part_w, long long stride_w, const float* __restrict__ part_b,
                                                          long long stride_b, int splits, long long cnt_w, int cnt_b,
                                                          float* __restrict__ out_w, float* __restrict__ out_b, int accumulate) {
    constexpr int PER = 256 / G;   // outputs per block
    __shared__ float sm[G][PER];
    const int li = threadIdx.x % PER, g = threadIdx.x / PER;
    const long long i = (long long)blockIdx.x * PER + li;
    const bool is_w = i < cnt_w, is_b = !is_w && i < cnt_w + cnt_b && out_b != nullptr;
    float s = 0.f;
    if (is_w) {
        for (int k = g; k < splits; k += G) s += part_w[(long long)k * stride_w + i];
    } else if (is_b) {
        for (int k = g; k < splits; k += G) s += part_b[(long long)k * stride_b + (i - cnt_w)];
    }
    if (G > 1) {
        sm[g][li] = s;
        __syncthreads();
        if (g != 0) return;
        s = 0.f;
#pragma unroll
        for (int q = 0; q < G; ++q) s += sm[q][li];
    }
    if (is_w) out_w[i] = accumulate ? out_w[i] + s : s;
    else if (is_b) out_b[i - cnt_w] = accumulate ? out_b[i - cnt_w] + s : s;
}

bool gymrl_defer_reduce(const float* part, long long stride, int splits, long long count, float* out, int accumulate);

void reduce_pair(const float* part_w, long long stride_w, const float* part_b, long long stride_b, int splits, long long cnt_w, int cnt_b,
                 float* out_w, float* out_b, int accumulate, cudaStream_t s) {
    // inside a deferral scope (reduce.cu) the fold is recorded and runs with every other pending one in gymrl_reduce_flush
    if (gymrl_defer_reduce(part_w, stride_w, splits, cnt_w, out_w, accumulate)) {
        gymrl_defer_reduce(part_b, stride_b, splits, cnt_b, out_b, accumulate);
        return;
    }
    const long long total = cnt_w + (out_b ? cnt_b : 0);
    if (splits >= 32 && total <= 16384)
        reduce_pair_kernel<8><<<(unsigned)ceil_div_ll(total, 32), 256, 0, s>>>(part_w, stride_w, part_b, stride_b, splits, cnt_w, cnt_b, out_w,
                                                                               out_b, accumulate);
    else
        reduce_pair_kernel<1><<<(unsigned)ceil_div_ll(total, 256), 256, 0, s>>>(part_w, stride_w, part_b, stride_b, splits, cnt_w, cnt_b, out_w,
                                                                                out_b, accumulate);
    gymrl_count_launch();
}

bool skinny_bwd_fused_supported(const float* dy, const float* x, int ldx, const float* dx, int lddx, int N, int K) {
    const bool shape = N >= 1 && N <= 8 && (K == 128 || K == 256 || K == 512) && N * (K / 128) <= 16;
    const bool al = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldx % 4 == 0) &&
                    (!dx || (((reinterpret_cast<uintptr_t>(dx) & 15) == 0) && (lddx % 4 == 0)));
    return shape && al;
}
int skinny_bwd_fused_chunks(int M) { return ceil_div(M, 64) < 512 ? ceil_div(M, 64) : 512; }

template <int N, int KC>
static int launch_fused(const float* dy, int lddy, const float* x, int ldx, const float* w, float* part, float* dx, int lddx, int M,
                        int act_in, int acc_dx, int chunks, cudaStream_t s) {
    constexpr size_t SMEM = (size_t)8 * ((N * 128 * KC + N + 3) & ~3) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (SMEM > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(skinny_bwd_fused_kernel<N, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
            if (e != cudaSuccess) GYMRL_FAIL(GYMRL_ECUDA, "cudaFuncSetAttribute(smem=%zu) failed: %s", SMEM, cudaGetErrorString(e));
        }
        configured = true;
    }
    skinny_bwd_fused_kernel<N, KC><<<chunks, 256, SMEM, s>>>(dy, lddy, x, ldx, w, part, dx, lddx, M, act_in, acc_dx, ceil_div(M, chunks));
    gymrl_count_launch();
    return GYMRL_OK;
}

// part: chunks * (N*K + N) floats
int skinny_bwd_fused(const float* dy, int lddy, const float* x, int ldx, const float* w, float* part, float* dx, int lddx, int M, int N,
                     int K, int act_in, int acc_dx, int chunks, cudaStream_t s) {
    const int KC = K / 128;
#define SKF(NN, KK) if (N == NN && KC == KK) return launch_fused<NN, KK>(dy, lddy, x, ldx, w, part, dx, lddx, M, act_in, acc_dx, chunks, s)
    SKF(1, 1); SKF(2, 1); SKF(3, 1); SKF(4, 1); SKF(5, 1); SKF(6, 1); SKF(7, 1); SKF(8, 1);
    SKF(1, 2); SKF(2, 2); SKF(3, 2); SKF(4, 2); SKF(5, 2); SKF(6, 2); SKF(7, 2); SKF(8, 2);
    SKF(1, 4); SKF(2, 4); SKF(3, 4); SKF(4, 4);
#undef SKF
    GYMRL_FAIL(GYMRL_EINVAL, "skinny_bwd_fused: unsupported N=%d K=%d", N, K);
}

// ---- backward-input of a small fan-in layer (K <= 8 input columns, e.g. the critic's first layer over [state | action]:
// dQ/d(s, a) for the actor gradient, sac_pendulum.py:211-218 / td3_pendulum.py:215-217): dX[m][k] = sum_n dY[m][n] W[n][k].
// One warp per row: lanes stride over n (coalesced dY row), K accumulators per lane, W^T staged in shared memory, fixed-order
// butterfly reduction.  The shape used to fall through to the FFMA GEMM with 64-wide tiles (21 us for 4096 x 256 x 4).
#define SMALLK_DX_MAXK 8
__global__ void __launch_bounds__(256) smallk_dx_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ w,
                                                        const float* __restrict__ h, int ldh, float* __restrict__ dx, int lddx, int M, int N,
                                                        int K, int act_in, int accumulate) {
    extern __shared__ float wt[];   // [K][N]
    for (int x = threadIdx.x; x < N * K; x += blockDim.x) wt[(x % K) * N + (x / K)] = w[x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int m = blockIdx.x * wpb + warp; m < M; m += gridDim.x * wpb) {
        float acc[SMALLK_DX_MAXK];
#pragma unroll
        for (int k = 0; k < SMALLK_DX_MAXK; ++k) acc[k] = 0.f;
        const float* row = dy + (size_t)m * lddy;
        for (int n = lane; n < N; n += 32) {
            const float g = row[n];
#pragma unroll
            for (int k = 0; k < SMALLK_DX_MAXK; ++k)
                if (k < K) acc[k] = fmaf(g, wt[k * N + n], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < SMALLK_DX_MAXK; ++k) {
            if (k < K) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
        }
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < SMALLK_DX_MAXK; ++k) v = lane == k ? acc[k] : v;
        if (lane < K) {
            if (h) {
                const float hv = h[(size_t)m * ldh + lane];
                if (act_in == GYMRL_ACT_TANH) v *= (1.0f - hv * hv);
                else if (act_in == GYMRL_ACT_RELU) v = hv > 0.f ? v : 0.f;
            }
            float* dst = dx + (size_t)m * lddx + lane;
            *dst = accumulate ? *dst + v : v;
        }
    }
}
bool smallk_dx_supported(int N, int K) { return K >= 1 && K <= SMALLK_DX_MAXK && N >= 32 && (size_t)N * K * sizeof(float) <= 32768; }
int smallk_dx(const float* dy, int lddy, const float* w, const float* h, int ldh, float* dx, int lddx, int M, int N, int K, int act_in,
              int accumulate, cudaStream_t s) {
    int blocks = ceil_div(M, 8);
    if (blocks > GYMRL_NUM_SMS * 4) blocks = GYMRL_NUM_SMS * 4;
    smallk_dx_kernel<<<blocks, 256, (size_t)N * K * sizeof(float), s>>>(dy, lddy, w, h, ldh, dx, lddx, M, N, K, act_in, accumulate);
    gymrl_count_launch();
    return GYMRL_OK;
}
