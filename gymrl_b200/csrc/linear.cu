// linear.cu — dense layers of the actor/critic/Q networks, forward and backward.
//
// Replaces nn.Linear + activation inside ActorCritic.forward (algorithms/ppo_lunarlander.py:63-90),
// QNetwork (dqn_cartpole.py:53-65), DuelingNoisyNetwork (rainbow_dqn_cartpole.py:100-113), SAC/TD3
// Actor/Critic (sac_pendulum.py:49-125, td3_pendulum.py:49-93) and the autograd backward of the same.
//
// Round-1 kernel: fp32 FFMA register-tiled GEMM (the reference computes in fp32 and the parity bar is
// fp32-tight, so no bf16/tf32 shortcut is taken here; the 3xTF32 tcgen05 path is the next step, see
// DESIGN.md).  One template covers the three products
//   forward   Y  = X  W^T        A = X  (k-major, optional row gather)  B = W  (k-major)
//   backward  dX = dY W          A = dY (k-major)                        B = W  (reduction-major)
//   backward  dW = dY^T X        A = dY (reduction-major)                B = X  (reduction-major, optional row gather)
// with tiles staged in shared memory as [BK][BM+4]/[BK][BN+4] (double buffered, register prefetch of
// the next k-slab), 16x16 threads each owning a TMxTN micro-tile split in 4-wide groups so that all
// shared loads are conflict-free float4 and all global stores are 256 B contiguous per half-warp.
// Fused epilogues: +bias, tanh/relu (forward); * act'(h) from the stored activations (backward input);
// deterministic split-M partial sums + fused column sums for db (backward weight; reduced by
// reduce_partials_kernel so results do not depend on atomics ordering).
#include "common.cuh"
#include "linear_tc.cuh"

#include <cstdlib>
#include <cstring>

void gymrl_count_launch(int n = 1);

// linear_skinny.cu: degenerate shapes (heads with N <= 8/16 outputs, observation layer with K in {3,4,8})
bool skinny_forward_supported(const float* x, int ldx, int N, int K);
int skinny_forward(const float* x, int ldx, const int32_t* rows, const float* w, const float* b, float* y, int ldy, int M, int N, int K,
                   int act, cudaStream_t s);
bool skinny_backward_input_supported(int N, int K, const float* h, int ldh, const float* dx, int lddx);
int skinny_backward_input(const float* dy, int lddy, const float* w, const float* h, int ldh, float* dx, int lddx, int M, int N, int K,
                          int act_in, int accumulate, cudaStream_t s);
bool skinny_dw_supported(int N);
int skinny_dw(const float* dy, int lddy, const float* x, int ldx, const int32_t* rows, float* part_w, float* part_b, int M, int N, int K,
              int chunks, cudaStream_t s);
bool smallk_forward_supported(int K);
int smallk_forward(const float* x, int ldx, const int32_t* rows, const float* w, const float* b, float* y, int ldy, int M, int N, int K,
                   int act, cudaStream_t s);
bool smallk_dw_supported(int K);
bool smallk_dx_supported(int N, int K);
int smallk_dx(const float* dy, int lddy, const float* w, const float* h, int ldh, float* dx, int lddx, int M, int N, int K, int act_in,
              int accumulate, cudaStream_t s);
bool skinny_bwd_fused_supported(const float* dy, const float* x, int ldx, const float* dx, int lddx, int N, int K);
int skinny_bwd_fused_chunks(int M);
int skinny_bwd_fused(const float* dy, int lddy, const float* x, int ldx, const float* w, float* part, float* dx, int lddx, int M, int N,
                     int K, int act_in, int acc_dx, int chunks, cudaStream_t s);
void reduce_pair(const float* part_w, long long stride_w, const float* part_b, long long stride_b, int splits, long long cnt_w, int cnt_b,
                 float* out_w, float* out_b, int accumulate, cudaStream_t s);
int smallk_dw(const float* dy, int lddy, const float* x, int ldx, const int32_t* rows, float* part_w, float* part_b, int M, int N, int K,
              int chunks, cudaStream_t s);
int smallk_dw_chunks(const float* dy, int lddy, int M, int N);
static int g_skinny = 1;   // GYMRL_SKINNY=0 disables the degenerate-shape kernels (debug / A-B comparison)

// GEMM engine selection: 0 = fp32 FFMA tiles (this file), 1 = tcgen05 3xTF32 (linear_tc.cu) where the shape gate allows.
static int g_gemm_mode = -1;
static int gemm_mode() {
    if (g_gemm_mode < 0) {
        const char* e = getenv("GYMRL_GEMM");
        g_gemm_mode = (e && strcmp(e, "ffma") == 0) ? 0 : 1;
        const char* k = getenv("GYMRL_SKINNY");
        g_skinny = (k && strcmp(k, "0") == 0) ? 0 : 1;
    }
    return g_gemm_mode;
}
extern "C" int gymrl_set_gemm_mode(int mode) {
    GYMRL_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (ffma) or 1 (tcgen05 3xTF32)");
    g_gemm_mode = mode;
    return GYMRL_OK;
}
extern "C" int gymrl_get_gemm_mode(void) { return gemm_mode(); }

struct GemmParams {
    const float* A; int lda; const int32_t* a_rows;
    const float* B; int ldb; const int32_t* b_rows;
    float* C; int ldc;
    int M, N, K;
    const float* bias; int act;
    const float* H; int ldh; int act_in;
    int accumulate;
    int k_chunk;             // reduction elements handled per blockIdx.z
    float* rowsum;           // dW: partial db [splits][M]
    long long c_split_stride;  // elements between split slices of C
    int vecA, vecB;          // float4 global loads allowed
};

#define BK 16

template <int BM, int BN, int TM, int TN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256, 2) gemm_kernel(const GemmParams p) {
    constexpr int NT = 256;
    constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
    constexpr int A_F4 = BM * BK / 4 / NT, B_F4 = BN * BK / 4 / NT;  // float4 per thread per slab
    constexpr int GM = TM / 4, GN = TN / 4;
    static_assert(BM / TM == 16 && BN / TN == 16, "16x16 thread grid");
    __shared__ __align__(16) float As[2][BK][LDA_S];
    __shared__ __align__(16) float Bs[2][BK][LDB_S];

    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * p.k_chunk;
    const int kend = min(p.K, kbeg + p.k_chunk);
    const int nk = (kend - kbeg + BK - 1) / BK;

    float4 ra[A_F4], rb[B_F4];

    auto load_a = [&](int k0) {
#pragma unroll
        for (int q = 0; q < A_F4; ++q) {
            const int f = t + q * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A_KMAJOR) {
                const int row = f / (BK / 4), kq = f % (BK / 4);
                const int m = m0 + row, k = k0 + kq * 4;
                if (m < p.M && k < kend) {
                    const long long r = p.a_rows ? (long long)p.a_rows[m] : (long long)m;
                    const float* src = p.A + r * p.lda + k;
                    if (p.vecA && k + 3 < kend) v = *reinterpret_cast<const float4*>(src);
                    else {
                        v.x = src[0];
                        if (k + 1 < kend) v.y = src[1];
                        if (k + 2 < kend) v.z = src[2];
                        if (k + 3 < kend) v.w = src[3];
                    }
                }
            } else {
                const int kr = f / (BM / 4), cq = f % (BM / 4);
                const int k = k0 + kr, m = m0 + cq * 4;
                if (k < kend && m < p.M) {
                    const long long r = p.a_rows ? (long long)p.a_rows[k] : (long long)k;
                    const float* src = p.A + r * p.lda + m;
                    if (p.vecA && m + 3 < p.M) v = *reinterpret_cast<const float4*>(src);
                    else {
                        v.x = src[0];
                        if (m + 1 < p.M) v.y = src[1];
                        if (m + 2 < p.M) v.z = src[2];
                        if (m + 3 < p.M) v.w = src[3];
                    }
                }
            }
            ra[q] = v;
        }
    };
    auto load_b = [&](int k0) {
#pragma unroll
        for (int q = 0; q < B_F4; ++q) {
            const int f = t + q * NT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (B_KMAJOR) {
                const int row = f / (BK / 4), kq = f % (BK / 4);
                const int n = n0 + row, k = k0 + kq * 4;
                if (n < p.N && k < kend) {
                    const long long r = p.b_rows ? (long long)p.b_rows[n] : (long long)n;
                    const float* src = p.B + r * p.ldb + k;
                    if (p.vecB && k + 3 < kend) v = *reinterpret_cast<const float4*>(src);
                    else {
                        v.x = src[0];
                        if (k + 1 < kend) v.y = src[1];
                        if (k + 2 < kend) v.z = src[2];
                        if (k + 3 < kend) v.w = src[3];
                    }
                }
            } else {
                const int kr = f / (BN / 4), cq = f % (BN / 4);
                const int k = k0 + kr, n = n0 + cq * 4;
                if (k < kend && n < p.N) {
                    const long long r = p.b_rows ? (long long)p.b_rows[k] : (long long)k;
                    const float* src = p.B + r * p.ldb + n;
                    if (p.vecB && n + 3 < p.N) v = *reinterpret_cast<const float4*>(src);
                    else {
                        v.x = src[0];
                        if (n + 1 < p.N) v.y = src[1];
                        if (n + 2 < p.N) v.z = src[2];
                        if (n + 3 < p.N) v.w = src[3];
                    }
                }
            }
            rb[q] = v;
        }
    };
    auto store_ab = [&](int buf) {
#pragma unroll
        for (int q = 0; q < A_F4; ++q) {
            const int f = t + q * NT;
            if (A_KMAJOR) {
                const int row = f / (BK / 4), kq = f % (BK / 4);
                As[buf][kq * 4 + 0][row] = ra[q].x;
                As[buf][kq * 4 + 1][row] = ra[q].y;
                As[buf][kq * 4 + 2][row] = ra[q].z;
                As[buf][kq * 4 + 3][row] = ra[q].w;
            } else {
                const int kr = f / (BM / 4), cq = f % (BM / 4);
                *reinterpret_cast<float4*>(&As[buf][kr][cq * 4]) = ra[q];
            }
        }
#pragma unroll
        for (int q = 0; q < B_F4; ++q) {
            const int f = t + q * NT;
            if (B_KMAJOR) {
                const int row = f / (BK / 4), kq = f % (BK / 4);
                Bs[buf][kq * 4 + 0][row] = rb[q].x;
                Bs[buf][kq * 4 + 1][row] = rb[q].y;
                Bs[buf][kq * 4 + 2][row] = rb[q].z;
                Bs[buf][kq * 4 + 3][row] = rb[q].w;
            } else {
                const int kr = f / (BN / 4), cq = f % (BN / 4);
                *reinterpret_cast<float4*>(&Bs[buf][kr][cq * 4]) = rb[q];
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float rsum[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) rsum[i] = 0.f;
    const bool do_rowsum = (p.rowsum != nullptr) && blockIdx.x == 0;

    if (nk > 0) {
        load_a(kbeg);
        load_b(kbeg);
        store_ab(0);
    }
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
            load_a(kbeg + (kt + 1) * BK);
            load_b(kbeg + (kt + 1) * BK);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][g * (BM / GM) + ty * 4]);
                a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][g * (BN / GN) + tx * 4]);
                b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            if (do_rowsum) {
#pragma unroll
                for (int i = 0; i < TM; ++i) rsum[i] += a[i];
            }
        }
        if (kt + 1 < nk) store_ab(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue ----
    float* Cbase = p.C + (long long)blockIdx.z * p.c_split_stride;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * (BM / GM) + ty * 4 + (i % 4);
        if (m >= p.M) continue;
        if (do_rowsum && tx == 0) p.rowsum[(long long)blockIdx.z * p.M + m] = rsum[i];
#pragma unroll
        for (int g = 0; g < GN; ++g) {
            const int n = n0 + g * (BN / GN) + tx * 4;
            if (n >= p.N) continue;
            float o[4] = {acc[i][g * 4 + 0], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (n + c >= p.N) continue;
                float v = o[c];
                if (p.bias) v += p.bias[n + c];
                if (p.act == GYMRL_ACT_TANH) v = tanhf(v);
                else if (p.act == GYMRL_ACT_RELU) v = fmaxf(v, 0.f);
                if (p.H) {
                    const float h = p.H[(long long)m * p.ldh + n + c];
                    if (p.act_in == GYMRL_ACT_TANH) v *= (1.0f - h * h);
                    else if (p.act_in == GYMRL_ACT_RELU) v = h > 0.f ? v : 0.f;
                }
                o[c] = v;
            }
            float* dst = Cbase + (long long)m * p.ldc + n;
            const bool vec_ok = (n + 3 < p.N) && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cbase) & 15) == 0);
            if (vec_ok) {
                if (p.accumulate) {
                    const float4 old = *reinterpret_cast<const float4*>(dst);
                    o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w;
                }
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (n + c < p.N) dst[c] = p.accumulate ? dst[c] + o[c] : o[c];
            }
        }
    }
}

template <bool AK, bool BKM>
static void launch_gemm(const GemmParams& p, int splits, cudaStream_t s) {
    // big tiles once the grid still fills the chip, small tiles otherwise
    const long long big_ctas = (long long)ceil_div(p.M, 128) * ceil_div(p.N, 128) * splits;
    if (big_ctas >= 2 * GYMRL_NUM_SMS && p.N >= 128) {
        dim3 grid(ceil_div(p.N, 128), ceil_div(p.M, 128), splits);
        gemm_kernel<128, 128, 8, 8, AK, BKM><<<grid, 256, 0, s>>>(p);
    } else {
        dim3 grid(ceil_div(p.N, 64), ceil_div(p.M, 64), splits);
        gemm_kernel<64, 64, 4, 4, AK, BKM><<<grid, 256, 0, s>>>(p);
    }
    gymrl_count_launch();
}

static inline int aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

extern "C" int gymrl_linear_forward(const float* d_x, int ldx, const int32_t* d_row_index, const float* d_w,
                                    const float* d_b, float* d_y, int ldy, int M, int N, int K, int act, void* stream) {
    GYMRL_REQUIRE(d_x && d_w && d_y, "NULL pointer");
    GYMRL_REQUIRE(M > 0 && N > 0 && K > 0 && ldx >= K && ldy >= N, "bad shape M=%d N=%d K=%d ldx=%d ldy=%d", M, N, K, ldx, ldy);
    GYMRL_REQUIRE(act >= 0 && act <= GYMRL_ACT_RELU, "unknown activation %d", act);
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.A = d_x; p.lda = ldx; p.a_rows = d_row_index;
    p.B = d_w; p.ldb = K; p.b_rows = nullptr;
    p.C = d_y; p.ldc = ldy;
    p.M = M; p.N = N; p.K = K;
    p.bias = d_b; p.act = act;
    p.k_chunk = K; p.c_split_stride = 0;
    p.vecA = aligned16(d_x) && (ldx % 4 == 0);
    p.vecB = aligned16(d_w) && (K % 4 == 0);
    (void)gemm_mode();
    if (g_skinny && N <= 8 && skinny_forward_supported(d_x, ldx, N, K)) {
        skinny_forward(d_x, ldx, d_row_index, d_w, d_b, d_y, ldy, M, N, K, act, as_stream(stream));
        GYMRL_LAUNCH_CHECK("linear_forward(skinny)");
        return GYMRL_OK;
    }
    if (g_skinny && N >= 32 && smallk_forward_supported(K)) {
        smallk_forward(d_x, ldx, d_row_index, d_w, d_b, d_y, ldy, M, N, K, act, as_stream(stream));
        GYMRL_LAUNCH_CHECK("linear_forward(small-K)");
        return GYMRL_OK;
    }
    if (gemm_mode() == 1) {
        TcGemmParams t;
        memset(&t, 0, sizeof(t));
        t.A = d_x; t.lda = ldx; t.a_rows = d_row_index; t.B = d_w; t.ldb = K; t.C = d_y; t.ldc = ldy;
        t.M = M; t.N = N; t.K = K; t.bias = d_b; t.act = act; t.k_chunk = K;
        if (tc_gemm_supported(t, true, true)) {
            int rc = tc_gemm_launch(t, true, true, 1, as_stream(stream));
            if (rc != GYMRL_OK) return rc;
            GYMRL_LAUNCH_CHECK("linear_forward(tc)");
            return GYMRL_OK;
        }
    }
    launch_gemm<true, true>(p, 1, as_stream(stream));
    GYMRL_LAUNCH_CHECK("linear_forward");
    return GYMRL_OK;
}

extern "C" int gymrl_linear_backward_input(const float* d_dy, int lddy, const float* d_w, const float* d_h_in, int ldh,
                                           float* d_dx, int lddx, int M, int N, int K, int act_in, int accumulate,
                                           void* stream) {
    GYMRL_REQUIRE(d_dy && d_w && d_dx, "NULL pointer");
    GYMRL_REQUIRE(M > 0 && N > 0 && K > 0 && lddy >= N && lddx >= K, "bad shape");
    GYMRL_REQUIRE(!d_h_in || ldh >= K, "ldh < K");
    // dX[M][K] = sum_n dY[m][n] W[n][k]: GEMM (M x K) with reduction N
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.A = d_dy; p.lda = lddy; p.B = d_w; p.ldb = K;
    p.C = d_dx; p.ldc = lddx;
    p.M = M; p.N = K; p.K = N;
    p.H = d_h_in; p.ldh = ldh; p.act_in = act_in;
    p.accumulate = accumulate;
    p.k_chunk = N;
    p.vecA = aligned16(d_dy) && (lddy % 4 == 0);
    p.vecB = aligned16(d_w) && (K % 4 == 0);
    (void)gemm_mode();
    if (g_skinny && N <= 16 && skinny_backward_input_supported(N, K, d_h_in, ldh, d_dx, lddx)) {
        skinny_backward_input(d_dy, lddy, d_w, d_h_in, ldh, d_dx, lddx, M, N, K, act_in, accumulate, as_stream(stream));
        GYMRL_LAUNCH_CHECK("linear_backward_input(skinny)");
        return GYMRL_OK;
    }
    if (g_skinny && smallk_dx_supported(N, K)) {   // small fan-in (K <= 8): one warp per row
        smallk_dx(d_dy, lddy, d_w, d_h_in, ldh, d_dx, lddx, M, N, K, act_in, accumulate, as_stream(stream));
        GYMRL_LAUNCH_CHECK("linear_backward_input(smallk)");
        return GYMRL_OK;
    }
    if (gemm_mode() == 1 && !accumulate) {
        TcGemmParams t;
        memset(&t, 0, sizeof(t));
        t.A = d_dy; t.lda = lddy; t.B = d_w; t.ldb = K; t.C = d_dx; t.ldc = lddx;
        t.M = M; t.N = K; t.K = N; t.H = d_h_in; t.ldh = ldh; t.act_in = act_in; t.k_chunk = N;
        if (tc_gemm_supported(t, true, false) && (!d_h_in || (ldh % 4 == 0))) {
            int rc = tc_gemm_launch(t, true, false, 1, as_stream(stream));
            if (rc != GYMRL_OK) return rc;
            GYMRL_LAUNCH_CHECK("linear_backward_input(tc)");
            return GYMRL_OK;
        }
    }
    launch_gemm<true, false>(p, 1, as_stream(stream));
    GYMRL_LAUNCH_CHECK("linear_backward_input");
    return GYMRL_OK;
}

// ---- backward weight: split-M partials + deterministic reduction ---------------------------------
static int dw_splits(int M, int N, int K) {
    const int tiles = ceil_div(N, 64) * ceil_div(K, 64);
    int s = ceil_div(4 * GYMRL_NUM_SMS, tiles);
    const int max_s = ceil_div(M, 128);
    if (s > max_s) s = max_s;
    if (s > 64) s = 64;
    if (s < 1) s = 1;
    return s;
}

extern "C" size_t gymrl_linear_backward_weight_workspace(int M, int N, int K) {
    // FFMA / tensor-core split-M use at most 64 partial slices; the bandwidth-bound skinny sweeps (tiny fan-out or
    // fan-in) cut the rows into up to 512 chunks so that every SM streams
    const int s = (N <= 8 || K <= 8) ? 512 : 64;
    return (size_t)s * ((size_t)N * K + (size_t)N) * sizeof(float);
}

// partial[z][n] = sum over the z-th row chunk of dY[m][n]: column sums for db when the tensor-core dW path (which has no
// fused column sums) is taken; the chunks are then folded by reduce_partials_kernel (deterministic order).
__global__ void colsum_partial_kernel(const float* __restrict__ dy, int lddy, int M, int N, int rows_per_chunk,
                                      float* __restrict__ partial) {
    __shared__ float sm[8][33];
    const int n = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
    const int m_beg = blockIdx.y * rows_per_chunk, m_end = min(M, m_beg + rows_per_chunk);
    float s = 0.f;
    if (n < N)
        for (int m = m_beg + g; m < m_end; m += 8) s += dy[(size_t)m * lddy + n];
    sm[g][threadIdx.x & 31] = s;
    __syncthreads();
    if (g == 0 && n < N) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += sm[k][threadIdx.x];
        partial[(size_t)blockIdx.y * N + n] = tot;
    }
}

__global__ void reduce_partials_kernel(const float* __restrict__ ws, long long count, int splits, float* __restrict__ out,
                                       int accumulate) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += ws[(long long)k * count + i];
    out[i] = accumulate ? out[i] + s : s;
}

extern "C" int gymrl_linear_backward_weight(const float* d_dy, int lddy, const float* d_x, int ldx,
                                            const int32_t* d_row_index, float* d_dw, float* d_db, int M, int N, int K,
                                            int accumulate, void* d_workspace, size_t workspace_bytes, void* stream) {
    GYMRL_REQUIRE(d_dy && d_x && d_dw && d_workspace, "NULL pointer");
    GYMRL_REQUIRE(M > 0 && N > 0 && K > 0 && lddy >= N && ldx >= K, "bad shape");
    GYMRL_REQUIRE(workspace_bytes >= gymrl_linear_backward_weight_workspace(M, N, K), "workspace too small");
    cudaStream_t s = as_stream(stream);
    const int splits = dw_splits(M, N, K);
    float* ws = (float*)d_workspace;
    float* ws_db = ws + (size_t)splits * N * K;
    // dW[N][K] = sum_m dY[m][n] X[m][k]: GEMM (N x K) with reduction M (split across blockIdx.z)
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.A = d_dy; p.lda = lddy; p.B = d_x; p.ldb = ldx; p.b_rows = d_row_index;
    p.C = ws; p.ldc = K;
    p.M = N; p.N = K; p.K = M;
    int chunk = ceil_div(M, splits);
    chunk = ceil_div(chunk, BK) * BK;
    p.k_chunk = chunk;
    p.c_split_stride = (long long)N * K;
    p.rowsum = d_db ? ws_db : nullptr;
    p.vecA = aligned16(d_dy) && (lddy % 4 == 0);
    p.vecB = aligned16(d_x) && (ldx % 4 == 0);
    (void)gemm_mode();
    if (g_skinny && ((N <= 8 && skinny_dw_supported(N)) || (N >= 32 && smallk_dw_supported(K)))) {
        int chunks = ceil_div(M, 32);
        if (chunks > 512) chunks = 512;
        if (N > 8) chunks = smallk_dw_chunks(d_dy, lddy, M, N);
        float* part_w = ws;
        float* part_b = ws + (size_t)chunks * N * K;
        if (N <= 8) skinny_dw(d_dy, lddy, d_x, ldx, d_row_index, part_w, d_db ? part_b : nullptr, M, N, K, chunks, s);
        else smallk_dw(d_dy, lddy, d_x, ldx, d_row_index, part_w, d_db ? part_b : nullptr, M, N, K, chunks, s);
        GYMRL_LAUNCH_CHECK("linear_backward_weight(skinny)");
        reduce_pair(part_w, (long long)N * K, part_b, N, chunks, (long long)N * K, N, d_dw, d_db, accumulate, s);
        GYMRL_LAUNCH_CHECK("reduce_pair(skinny)");
        return GYMRL_OK;
    }
    if (gemm_mode() == 1) {
        TcGemmParams t;
        memset(&t, 0, sizeof(t));
        t.A = d_dy; t.lda = lddy; t.B = d_x; t.ldb = ldx; t.b_rows = d_row_index; t.C = ws; t.ldc = K;
        t.M = N; t.N = K; t.K = M;
        // column-tile width of the dW GEMM: narrower tiles need fewer split-K slices to fill the SMs, i.e. fewer partial tiles to
        // write and fold, at the price of re-reading dY once per column tile (GYMRL_DW_BN = 64 / 128 / 256 for A/B runs)
        static const int dw_bn = [] { const char* e = getenv("GYMRL_DW_BN"); const int v = e ? atoi(e) : 256; return (v == 64 || v == 128) ? v : 256; }();
        t.bn_max = dw_bn;
        const int tc_tiles = ceil_div(N, 128) * ceil_div(K, dw_bn);
        int tsplits = GYMRL_NUM_SMS / tc_tiles;
        if (tsplits > 64) tsplits = 64;
        if (tsplits > M / 256) tsplits = M / 256;
        if (tsplits < 1) tsplits = 1;
        int tchunk = ceil_div(ceil_div(M, tsplits), 32) * 32;
        t.k_chunk = tchunk; t.c_split_stride = (long long)N * K;
        float* ws_db2 = ws + (size_t)tsplits * N * K;   // behind the dW partial tiles
        t.colsum = d_db ? ws_db2 : nullptr;
        if ((M % 32 == 0) && tc_gemm_supported(t, false, false)) {
            int rc = tc_gemm_launch(t, false, false, tsplits, s);
            if (rc != GYMRL_OK) return rc;
            GYMRL_LAUNCH_CHECK("linear_backward_weight(tc)");
            reduce_pair(ws, (long long)N * K, ws_db2, N, tsplits, (long long)N * K, N, d_dw, d_db, accumulate, s);
            GYMRL_LAUNCH_CHECK("reduce_pair(tc)");
            return GYMRL_OK;
        }
    }
    {
        dim3 grid(ceil_div(K, 64), ceil_div(N, 64), splits);
        gemm_kernel<64, 64, 4, 4, false, false><<<grid, 256, 0, s>>>(p);
        gymrl_count_launch();
    }
    GYMRL_LAUNCH_CHECK("linear_backward_weight");
    const long long cnt = (long long)N * K;
    reduce_pair(ws, cnt, ws_db, N, splits, cnt, N, d_dw, d_db, accumulate, s);
    GYMRL_LAUNCH_CHECK("reduce_pair");
    return GYMRL_OK;
}

// Whole backward of one dense layer: dW, db and (optionally) dX = (dY W) * act'(x), x being the previous layer's
// activation output.  Skinny heads take the fused single-sweep kernel; every other shape is the two calls above.
extern "C" int gymrl_linear_backward(const float* d_dy, int lddy, const float* d_x, int ldx, const int32_t* d_row_index,
                                     const float* d_w, float* d_dw, float* d_db, float* d_dx, int lddx, int M, int N, int K,
                                     int act_in, int accumulate, void* d_workspace, size_t workspace_bytes, void* stream) {
    GYMRL_REQUIRE(d_dy && d_x && d_w && d_dw && d_workspace, "NULL pointer");
    GYMRL_REQUIRE(M > 0 && N > 0 && K > 0 && lddy >= N && ldx >= K, "bad shape");
    GYMRL_REQUIRE(!(d_dx && d_row_index), "dX of a row-gathered input is not defined (scatter): pass d_dx = NULL");
    GYMRL_REQUIRE(!d_dx || lddx >= K, "lddx < K");
    GYMRL_REQUIRE(workspace_bytes >= gymrl_linear_backward_weight_workspace(M, N, K), "workspace too small");
    (void)gemm_mode();
    if (g_skinny && !d_row_index && skinny_bwd_fused_supported(d_dy, d_x, ldx, d_dx, lddx, N, K)) {
        cudaStream_t s = as_stream(stream);
        const int chunks = skinny_bwd_fused_chunks(M);
        float* part = (float*)d_workspace;
        int rc = skinny_bwd_fused(d_dy, lddy, d_x, ldx, d_w, part, d_dx, lddx, M, N, K, act_in, 0, chunks, s);
        if (rc != GYMRL_OK) return rc;
        GYMRL_LAUNCH_CHECK("linear_backward(skinny fused)");
        const long long stride = (long long)N * K + N;
        reduce_pair(part, stride, part + (size_t)N * K, stride, chunks, (long long)N * K, N, d_dw, d_db, accumulate, s);
        GYMRL_LAUNCH_CHECK("reduce_pair(skinny fused)");
        return GYMRL_OK;
    }
    int rc = gymrl_linear_backward_weight(d_dy, lddy, d_x, ldx, d_row_index, d_dw, d_db, M, N, K, accumulate, d_workspace, workspace_bytes,
                                          stream);
    if (rc != GYMRL_OK || !d_dx) return rc;
    return gymrl_linear_backward_input(d_dy, lddy, d_w, d_x, ldx, d_dx, lddx, M, N, K, act_in, 0, stream);
}
