// mhc.cu — the row-wise ("glue") half of ppo_full's network: manifold hyper-connection stages, RMSNorm, SiLU.
//
// Reference: algorithms/ppo_full_lunarlander.py
//     sinkhorn_knopp_batched :76-103, ManifoldHyperConnectionFuse.mapping/process/depth_connection :144-194,
//     MHCBlock.forward :210-229, MHCBackbone.forward :251-267, RMSNorm :273-284, MLP :287-318.
// The dense layers in between (input_proj, linear1/2 of every block, the head layers) are the library's ordinary
// GEMMs (linear_tc.cu / linear_skinny.cu) with ACT_NONE: every non-linearity of this network sits in the kernels here,
// where a whole row is in registers anyway.
//
// One mHC stage, for a row h = (h_0, h_1) of n = 2 branches of width D (hv = the 2D-vector):
//     H_k   = sum_c g_c hv_c w[c][k],  k = 0..7                      g = mhc.norm.weight, w = mhc.w
//     r_    = 1 / (|hv| / sqrt(2D) + 1e-6)
//     t_k   = r_ * H_k * alpha[k<2 ? 0 : k<4 ? 1 : 2] + beta_k
//     pre_i = sigmoid(t_i), post_i = 2 sigmoid(t_{2+i}), E_ij = exp(t_{4+2i+j})
//     (u, v) = `sk_iters` Sinkhorn-Knopp iterations on E (no gradient), P_ij = u_i E_ij v_j
//     h_pre = sum_i pre_i h_i     -> z = linear(h_pre) (GEMM)  -> h_out = silu(z)
//     h'_i  = post_i h_out + sum_j P_ij h_j
// Layout: rows are contiguous [2][D] (or a [D] row read for both branches: the backbone input, ref :256-258).
// One warp per row, lane l owns elements {128 j + 4 l .. + 3} of each branch (float4 loads, fully coalesced); the
// 8 projections and the norm are warp-shuffle reductions; the 2x2 algebra runs redundantly on every lane.
// HBM-bound: a stage reads 2D + D floats and writes 2D + D + 8 per row in the forward pass.
// Parameter gradients (dw [2D][8], dg [2D], dalpha [3], dbeta [8]) are accumulated in registers over the rows a warp
// owns, reduced across the block in shared memory, written as per-block partials and summed by a second kernel in a
// fixed order (deterministic; no float atomics).
#include "common.cuh"
#include "rowwise.cuh"

void gymrl_count_launch(int n = 1);

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kMaxBlocks = 2 * GYMRL_NUM_SMS;

// 1 / (1 + e^-x) with ex2.approx / rcp.approx (a few ulp): the IEEE expf + division pair costs ~25 instructions per element and
// the SiLU / SiLU' of every activation element goes through it (parity: same golden tolerances).
__device__ __forceinline__ float sigmoidf_(float x) { return rcp_approx(1.0f + exp2f_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float silu_gradf_(float x) {
    const float s = sigmoidf_(x);
    return s * (1.0f + x * (1.0f - s));
}
// Per-lane copy of the stage parameters touching this lane's elements.
template <int NCH>
struct StageParams {
    float4 g[2][NCH];        // g at (branch, chunk)
    float4 w[2][NCH][8];     // w[c][k] for the 4 elements, per k
    float alpha[3], beta[8];
    __device__ void load(const float* __restrict__ gp, const float* __restrict__ wp, const float* __restrict__ ap,
                         const float* __restrict__ bp, int lane) {
        constexpr int D = 128 * NCH;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const int c = i * D + 128 * j + 4 * lane;
                g[i][j] = ld4(gp + c);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    w[i][j][k] = make_float4(wp[(size_t)(c + 0) * 8 + k], wp[(size_t)(c + 1) * 8 + k], wp[(size_t)(c + 2) * 8 + k],
                                             wp[(size_t)(c + 3) * 8 + k]);
            }
#pragma unroll
        for (int k = 0; k < 3; ++k) alpha[k] = ap[k];
#pragma unroll
        for (int k = 0; k < 8; ++k) beta[k] = bp[k];
    }
};

struct Coef {
    float pre[2], post[2], P[2][2], r_, H[8], s;
};

// mapping() of one row held in registers (ref :144-180)
template <int NCH>
__device__ __forceinline__ void stage_coefficients(const float4 (&h)[2][NCH], const StageParams<NCH>& sp, int sk_iters, Coef& c) {
    constexpr int D = 128 * NCH;
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const float4 gh = mul4(sp.g[i][j], h[i][j]);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += dot4(gh, sp.w[i][j][k]);
            acc[8] += dot4(h[i][j], h[i][j]);
        }
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = warp_sum(acc[k]);
    c.s = acc[8];
    const float r = sqrtf(acc[8]) / sqrtf((float)(2 * D));
    c.r_ = 1.0f / (r + 1e-6f);
    float t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        c.H[k] = acc[k];
        t[k] = c.r_ * acc[k] * sp.alpha[k < 2 ? 0 : (k < 4 ? 1 : 2)] + sp.beta[k];
    }
    c.pre[0] = sigmoidf_(t[0]); c.pre[1] = sigmoidf_(t[1]);
    c.post[0] = 2.0f * sigmoidf_(t[2]); c.post[1] = 2.0f * sigmoidf_(t[3]);
    const float e00 = expf(t[4]), e01 = expf(t[5]), e10 = expf(t[6]), e11 = expf(t[7]);
    float u0 = 1.f, u1 = 1.f, v0 = 1.f, v1 = 1.f;
    const float eps = 1e-8f;
    for (int it = 0; it < sk_iters; ++it) {
        // rcp.approx (1 ulp): an IEEE division is ~10 dependent instructions, and these 4 x sk_iters reciprocals are a serial
        // chain evaluated on every lane — they were two thirds of the stage kernels' instructions.  The iteration is a
        // contraction, so the 1-ulp errors do not accumulate (goldens: same tolerance as before).
        u0 = rcp_approx(e00 * v0 + e01 * v1 + eps);
        u1 = rcp_approx(e10 * v0 + e11 * v1 + eps);
        v0 = rcp_approx(e00 * u0 + e10 * u1 + eps);
        v1 = rcp_approx(e01 * u0 + e11 * u1 + eps);
    }
    c.P[0][0] = u0 * e00 * v0; c.P[0][1] = u0 * e01 * v1;
    c.P[1][0] = u1 * e10 * v0; c.P[1][1] = u1 * e11 * v1;
}

// ---------------------------------------------------------------------------------------------------------------
// forward: [post of the previous stage] -> [mapping + h_pre of this stage] -> [final branch sum + RMSNorm]
// ---------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
mhc_stage_fwd_kernel(const float* __restrict__ h_prev, int prev_row_stride, int prev_branch_stride, const float* __restrict__ z_prev,
                     const float* __restrict__ coef_prev, float* __restrict__ h_cur, const float* __restrict__ g,
                     const float* __restrict__ w, const float* __restrict__ alpha, const float* __restrict__ beta,
                     float* __restrict__ coef_cur, float* __restrict__ h_pre, const float* __restrict__ final_weight,
                     float* __restrict__ feat, int M, int sk_iters, float eps) {
    constexpr int D = 128 * NCH;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), nwarps = gridDim.x * kWarpsPerBlock;
    StageParams<NCH> sp;
    if (g) sp.load(g, w, alpha, beta, lane);
    float4 fw[NCH];
    if (final_weight) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) fw[j] = ld4(final_weight + 128 * j + 4 * lane);
    }
    for (int row = warp; row < M; row += nwarps) {
        float4 h[2][NCH];
        const float* hp = h_prev + (size_t)row * prev_row_stride;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NCH; ++j) h[i][j] = ld4(hp + (size_t)i * prev_branch_stride + 128 * j + 4 * lane);
        if (z_prev) {
            // depth_connection of the previous stage (ref :190-194): h'_i = post_i silu(z) + sum_j P_ij h_j
            const float4 c0 = ld4(coef_prev + (size_t)row * 24), c1 = ld4(coef_prev + (size_t)row * 24 + 4);
            float4 hn[2][NCH];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const float4 z = ld4(z_prev + (size_t)row * D + 128 * j + 4 * lane);
                const float4 ho = make_float4(siluf_(z.x), siluf_(z.y), siluf_(z.z), siluf_(z.w));
                hn[0][j] = axpy4(c0.z, ho, axpy4(c1.x, h[0][j], scale4(c1.y, h[1][j])));
                hn[1][j] = axpy4(c0.w, ho, axpy4(c1.z, h[0][j], scale4(c1.w, h[1][j])));
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    h[i][j] = hn[i][j];
                    st4(h_cur + (size_t)row * 2 * D + (size_t)i * D + 128 * j + 4 * lane, h[i][j]);
                }
        }
        if (g) {
            Coef c;
            stage_coefficients<NCH>(h, sp, sk_iters, c);
            if (lane == 0) {
                // the whole mapping of the row is kept (24 floats, the layout the backward kernels use): the backward pass reads it
                // back instead of re-running the 8 projections and the Sinkhorn chain on every row
                float* sc = coef_cur + (size_t)row * 24;
                st4(sc, make_float4(c.pre[0], c.pre[1], c.post[0], c.post[1]));
                st4(sc + 4, make_float4(c.P[0][0], c.P[0][1], c.P[1][0], c.P[1][1]));
                st4(sc + 8, make_float4(c.r_, c.s, 0.f, 0.f));
                st4(sc + 12, make_float4(0.f, 0.f, 0.f, 0.f));
                st4(sc + 16, make_float4(c.H[0], c.H[1], c.H[2], c.H[3]));
                st4(sc + 20, make_float4(c.H[4], c.H[5], c.H[6], c.H[7]));
            }
#pragma unroll
            for (int j = 0; j < NCH; ++j)
                st4(h_pre + (size_t)row * D + 128 * j + 4 * lane, axpy4(c.pre[0], h[0][j], scale4(c.pre[1], h[1][j])));
        }
        if (final_weight) {
            // h.sum(dim=2) -> RMSNorm (ref :263-267, :279-284)
            float4 x[NCH];
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                x[j] = add4(h[0][j], h[1][j]);
                ss += dot4(x[j], x[j]);
            }
            ss = warp_sum(ss);
            const float inv = rsqrtf(ss / (float)D + eps);
#pragma unroll
            for (int j = 0; j < NCH; ++j) st4(feat + (size_t)row * D + 128 * j + 4 * lane, mul4(scale4(inv, x[j]), fw[j]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward, part A (before the stage's GEMM backward): from dh' = dL/dh_{s+1}
//     dz = (sum_i post_i dh'_i) silu'(z);  dh_partial_j = sum_i P_ij dh'_i;  dpost_i = <dh'_i, silu(z)>;  dP_ij = <dh'_i, h_j>
// coef row [24] = pre0 pre1 post0 post1 | P00 P01 P10 P11 | r_ s dpost0 dpost1 | dP00 dP01 dP10 dP11 | H0..H7
// The forward pass left the row's mapping there (pre / post / P / r_ / s / H); this kernel adds the six inner products.  (Round 1
// re-ran the mapping here: 8 projections, 9 warp reductions and the 40-reciprocal Sinkhorn chain per row, with the stage
// parameters in 64+ registers per lane — 149 registers, one block per SM.  Now it is a streaming pass.)
// ---------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
mhc_stage_bwd_a_kernel(const float* __restrict__ h_in, int row_stride, int branch_stride, const float* __restrict__ z,
                       const float* __restrict__ dh_next, float* __restrict__ dz, float* __restrict__ dh_partial,
                       float* __restrict__ coef, int M) {
    constexpr int D = 128 * NCH;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), nwarps = gridDim.x * kWarpsPerBlock;
    for (int row = warp; row < M; row += nwarps) {
        float4 h[2][NCH], d[2][NCH];
        const float* hp = h_in + (size_t)row * row_stride;
        float* sc = coef + (size_t)row * 24;
        const float4 s0 = ld4(sc), s1 = ld4(sc + 4), s2 = ld4(sc + 8);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                h[i][j] = ld4(hp + (size_t)i * branch_stride + 128 * j + 4 * lane);
                d[i][j] = ld4(dh_next + (size_t)row * 2 * D + (size_t)i * D + 128 * j + 4 * lane);
            }
        const float post0 = s0.z, post1 = s0.w, P00 = s1.x, P01 = s1.y, P10 = s1.z, P11 = s1.w;
        float red[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) red[k] = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const float4 zz = ld4(z + (size_t)row * D + 128 * j + 4 * lane);
            const float4 ho = make_float4(siluf_(zz.x), siluf_(zz.y), siluf_(zz.z), siluf_(zz.w));
            const float4 sg = make_float4(silu_gradf_(zz.x), silu_gradf_(zz.y), silu_gradf_(zz.z), silu_gradf_(zz.w));
            const float4 dho = axpy4(post0, d[0][j], scale4(post1, d[1][j]));
            st4(dz + (size_t)row * D + 128 * j + 4 * lane, mul4(dho, sg));
            red[0] += dot4(d[0][j], ho); red[1] += dot4(d[1][j], ho);
            red[2] += dot4(d[0][j], h[0][j]); red[3] += dot4(d[0][j], h[1][j]);
            red[4] += dot4(d[1][j], h[0][j]); red[5] += dot4(d[1][j], h[1][j]);
            st4(dh_partial + (size_t)row * 2 * D + 128 * j + 4 * lane, axpy4(P00, d[0][j], scale4(P10, d[1][j])));
            st4(dh_partial + (size_t)row * 2 * D + D + 128 * j + 4 * lane, axpy4(P01, d[0][j], scale4(P11, d[1][j])));
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) red[k] = warp_sum(red[k]);
        if (lane == 0) {
            st4(sc + 8, make_float4(s2.x, s2.y, red[0], red[1]));
            st4(sc + 12, make_float4(red[2], red[3], red[4], red[5]));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward, part B (after the stage's GEMM backward produced dh_pre = dz @ W):
//     dh_i = dh_partial_i + pre_i dh_pre + g_c sum_k w[c][k] dH_k + hv_c dr / (sqrt(s) sqrt(2D))
//     parameter gradients dw, dg, dalpha, dbeta -> per-block partials
// partial row layout: [2D*8 dw | 2D dg | 3 dalpha | 8 dbeta]
// ---------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, NCH == 1 ? 2 : 1)
mhc_stage_bwd_b_kernel(const float* __restrict__ h_in, int row_stride, int branch_stride, const float* __restrict__ dh_pre,
                       const float* __restrict__ scratch, const float* __restrict__ dh_partial, float* __restrict__ dh_out,
                       float* __restrict__ dx0, const float* __restrict__ g, const float* __restrict__ w,
                       const float* __restrict__ alpha, float* __restrict__ partials, int M) {
    constexpr int D = 128 * NCH;
    constexpr int P = 2 * D * 8 + 2 * D + 11;
    constexpr int PS = (P + 3) & ~3;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warp = blockIdx.x * kWarpsPerBlock + wid, nwarps = gridDim.x * kWarpsPerBlock;
    extern __shared__ __align__(16) float s_acc[];   // [PS] block accumulator | [8][2D] the stage's w, k-major (lane-contiguous float4 reads)
    float* s_w = s_acc + PS;
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_acc[i] = 0.f;
    for (int i = threadIdx.x; i < 2 * D * 8; i += blockDim.x) s_w[(i & 7) * 2 * D + (i >> 3)] = w[i];
    __syncthreads();
    // The weights are read from shared memory inside the row loop (16 LDS.128 per row at D = 128) instead of living in 64 registers
    // per lane: with the 64 accumulator registers the kernel sat at 211 registers = one block per SM, i.e. 8 warps to hide a
    // row's HBM round trip; at <= 128 registers two blocks are resident.
    float4 gp[2][NCH];
    float al[3];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NCH; ++j) gp[i][j] = ld4(g + i * D + 128 * j + 4 * lane);
#pragma unroll
    for (int k = 0; k < 3; ++k) al[k] = alpha[k];
    float4 a_w[2][NCH][8], a_g[2][NCH];
    float a_ab[11];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            a_g[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 8; ++k) a_w[i][j][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
    for (int k = 0; k < 11; ++k) a_ab[k] = 0.f;

    for (int row = warp; row < M; row += nwarps) {
        float4 h[2][NCH], dp[NCH];
        const float* hp = h_in + (size_t)row * row_stride;
        float dpre0 = 0.f, dpre1 = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            h[0][j] = ld4(hp + 128 * j + 4 * lane);
            h[1][j] = ld4(hp + (size_t)branch_stride + 128 * j + 4 * lane);
            dp[j] = ld4(dh_pre + (size_t)row * D + 128 * j + 4 * lane);
            dpre0 += dot4(dp[j], h[0][j]);
            dpre1 += dot4(dp[j], h[1][j]);
        }
        dpre0 = warp_sum(dpre0);
        dpre1 = warp_sum(dpre1);
        const float* sc = scratch + (size_t)row * 24;
        const float4 s0 = ld4(sc), s1 = ld4(sc + 4), s2 = ld4(sc + 8), s3 = ld4(sc + 12), s4 = ld4(sc + 16), s5 = ld4(sc + 20);
        const float pre0 = s0.x, pre1 = s0.y, post0 = s0.z, post1 = s0.w;
        const float r_ = s2.x, ss = s2.y;
        const float H[8] = {s4.x, s4.y, s4.z, s4.w, s5.x, s5.y, s5.z, s5.w};
        float dt[8];
        dt[0] = dpre0 * pre0 * (1.0f - pre0);
        dt[1] = dpre1 * pre1 * (1.0f - pre1);
        dt[2] = s2.z * post0 * (1.0f - 0.5f * post0);
        dt[3] = s2.w * post1 * (1.0f - 0.5f * post1);
        dt[4] = s3.x * s1.x; dt[5] = s3.y * s1.y; dt[6] = s3.z * s1.z; dt[7] = s3.w * s1.w;   // dE E = dP P
        float dH[8], dr_ = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float a = al[k < 2 ? 0 : (k < 4 ? 1 : 2)];
            dH[k] = dt[k] * r_ * a;
            dr_ += dt[k] * H[k] * a;
            a_ab[3 + k] += dt[k];
            a_ab[k < 2 ? 0 : (k < 4 ? 1 : 2)] += dt[k] * r_ * H[k];
        }
        const float rt_s = sqrtf(ss), rt_nd = sqrtf((float)(2 * D));
        const float dr = -r_ * r_ * dr_;
        const float norm_coef = ss > 0.f ? dr / (rt_s * rt_nd) : 0.f;
        const float prei[2] = {pre0, pre1};
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float4 wd = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 gh = mul4(gp[i][j], h[i][j]);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    wd = axpy4(dH[k], *reinterpret_cast<const float4*>(s_w + k * 2 * D + i * D + 128 * j + 4 * lane), wd);
                    a_w[i][j][k] = axpy4(dH[k], gh, a_w[i][j][k]);
                }
                a_g[i][j] = add4(a_g[i][j], mul4(h[i][j], wd));
                float4 dh = ld4(dh_partial + (size_t)row * 2 * D + (size_t)i * D + 128 * j + 4 * lane);
                dh = axpy4(prei[i], dp[j], dh);
                dh = add4(dh, mul4(gp[i][j], wd));
                dh = axpy4(norm_coef, h[i][j], dh);
                h[i][j] = dh;   // reuse as the output register
            }
        if (dx0) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) st4(dx0 + (size_t)row * D + 128 * j + 4 * lane, add4(h[0][j], h[1][j]));
        } else {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NCH; ++j) st4(dh_out + (size_t)row * 2 * D + (size_t)i * D + 128 * j + 4 * lane, h[i][j]);
        }
    }
    // block reduction of the parameter-gradient accumulators (shared-memory atomics: 8 warps, fixed data, order-
    // independent only up to float rounding -> use a fixed warp order instead)
    for (int wsel = 0; wsel < kWarpsPerBlock; ++wsel) {
        if (wid == wsel) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int c = i * D + 128 * j + 4 * lane;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        s_acc[(c + 0) * 8 + k] += a_w[i][j][k].x; s_acc[(c + 1) * 8 + k] += a_w[i][j][k].y;
                        s_acc[(c + 2) * 8 + k] += a_w[i][j][k].z; s_acc[(c + 3) * 8 + k] += a_w[i][j][k].w;
                    }
                    s_acc[2 * D * 8 + c + 0] += a_g[i][j].x; s_acc[2 * D * 8 + c + 1] += a_g[i][j].y;
                    s_acc[2 * D * 8 + c + 2] += a_g[i][j].z; s_acc[2 * D * 8 + c + 3] += a_g[i][j].w;
                }
            if (lane == 0) {
                // alpha / beta sums are identical on every lane of the warp
#pragma unroll
                for (int k = 0; k < 11; ++k) s_acc[2 * D * 8 + 2 * D + k] += a_ab[k];
            }
        }
        __syncthreads();
    }
    float* out = partials + (size_t)blockIdx.x * P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) out[i] = s_acc[i];
}

// ---------------------------------------------------------------------------------------------------------------
// RMSNorm (+ optional SiLU in front, + optional "sum of two branches" in front), row-wise, `groups` groups of width W
//   y = a * rsqrt(mean(a^2) + eps) * weight,  a = silu(x) | x | x_0 + x_1
// ---------------------------------------------------------------------------------------------------------------
template <int NCH, bool SILU, bool SUM2>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rmsnorm_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ weight, float* __restrict__ y, int ldy, int M,
                   int groups, float eps) {
    constexpr int W = 128 * NCH;
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * kWarpsPerBlock;
    for (long long item = warp; item < (long long)M * groups; item += nwarps) {
        const int row = (int)(item / groups), gi = (int)(item % groups);
        const float* xp = x + (size_t)row * ldx + (SUM2 ? 0 : (size_t)gi * W);
        float4 a[NCH];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float4 v = ld4(xp + 128 * j + 4 * lane);
            if (SUM2) v = add4(v, ld4(xp + W + 128 * j + 4 * lane));
            if (SILU) v = make_float4(siluf_(v.x), siluf_(v.y), siluf_(v.z), siluf_(v.w));
            a[j] = v;
            ss += dot4(v, v);
        }
        ss = warp_sum(ss);
        const float inv = rsqrtf(ss / (float)W + eps);
#pragma unroll
        for (int j = 0; j < NCH; ++j)
            st4(y + (size_t)row * ldy + (size_t)gi * W + 128 * j + 4 * lane,
                mul4(scale4(inv, a[j]), ld4(weight + (size_t)gi * W + 128 * j + 4 * lane)));
    }
}

// dx = inv (dy w - nh mean(dy w nh)) [* silu'(x)],  dweight += dy nh;  SUM2 writes dx to both branches
template <int NCH, bool SILU, bool SUM2>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rmsnorm_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ weight, const float* __restrict__ dy, int lddy,
                   float* __restrict__ dx, int lddx, float* __restrict__ partials, int M, int groups, float eps) {
    constexpr int W = 128 * NCH;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long warp = (long long)blockIdx.x * kWarpsPerBlock + wid, nwarps = (long long)gridDim.x * kWarpsPerBlock;
    extern __shared__ float s_acc[];   // [groups * W]
    const int P = groups * W;
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
    // a warp visits items warp, warp + nwarps, ...: with nwarps a multiple of `groups` it always sees the same group
    float4 acc[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int my_group = (int)(warp % groups);
    for (long long item = warp; item < (long long)M * groups; item += nwarps) {
        const int row = (int)(item / groups), gi = (int)(item % groups);
        const float* xp = x + (size_t)row * ldx + (SUM2 ? 0 : (size_t)gi * W);
        float4 xv[NCH], a[NCH], dyw[NCH], dyv[NCH];
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float4 v = ld4(xp + 128 * j + 4 * lane);
            if (SUM2) v = add4(v, ld4(xp + W + 128 * j + 4 * lane));
            xv[j] = v;
            if (SILU) v = make_float4(siluf_(v.x), siluf_(v.y), siluf_(v.z), siluf_(v.w));
            a[j] = v;
            ss += dot4(v, v);
        }
        ss = warp_sum(ss);
        const float inv = rsqrtf(ss / (float)W + eps);
        float m = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            dyv[j] = ld4(dy + (size_t)row * lddy + (size_t)gi * W + 128 * j + 4 * lane);
            dyw[j] = mul4(dyv[j], ld4(weight + (size_t)gi * W + 128 * j + 4 * lane));
            a[j] = scale4(inv, a[j]);   // n-hat
            m += dot4(dyw[j], a[j]);
            acc[j] = add4(acc[j], mul4(dyv[j], a[j]));
        }
        m = warp_sum(m) / (float)W;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float4 d = scale4(inv, axpy4(-m, a[j], dyw[j]));
            if (SILU) d = mul4(d, make_float4(silu_gradf_(xv[j].x), silu_gradf_(xv[j].y), silu_gradf_(xv[j].z), silu_gradf_(xv[j].w)));
            if (SUM2) {
                st4(dx + (size_t)row * lddx + 128 * j + 4 * lane, d);
                st4(dx + (size_t)row * lddx + W + 128 * j + 4 * lane, d);
            } else {
                st4(dx + (size_t)row * lddx + (size_t)gi * W + 128 * j + 4 * lane, d);
            }
        }
    }
    for (int wsel = 0; wsel < kWarpsPerBlock; ++wsel) {
        if (wid == wsel) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float* p = s_acc + (size_t)my_group * W + 128 * j + 4 * lane;
                p[0] += acc[j].x; p[1] += acc[j].y; p[2] += acc[j].z; p[3] += acc[j].w;
            }
        }
        __syncthreads();
    }
    float* out = partials + (size_t)blockIdx.x * P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) out[i] = s_acc[i];
}

int grid_for_rows(long long rows) {
    long long b = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (b > kMaxBlocks) b = kMaxBlocks;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

extern "C" size_t gymrl_mhc_workspace_bytes(int D, int head_width, int head_groups) {
    size_t a = (size_t)kMaxBlocks * (size_t)(2 * D * 8 + 2 * D + 11) * sizeof(float);
    size_t b = (size_t)kMaxBlocks * (size_t)(head_width * head_groups) * sizeof(float);
    size_t c = (size_t)kMaxBlocks * (size_t)D * sizeof(float);
    size_t m = a > b ? a : b;
    return m > c ? m : c;
}

extern "C" int gymrl_mhc_stage_forward(const float* h_prev, int prev_row_stride, int prev_branch_stride, const float* z_prev,
                                       const float* coef_prev, float* h_cur, const float* g, const float* w, const float* alpha,
                                       const float* beta, float* coef_cur, float* h_pre, const float* final_weight, float* feat,
                                       int M, int D, int sk_iters, float eps, void* stream) {
    GYMRL_REQUIRE(h_prev != nullptr && M >= 0, "NULL input");
    GYMRL_REQUIRE(D == 128 || D == 256, "mHC kernels are built for mhc_dim 128 or 256 (got %d)", D);
    GYMRL_REQUIRE(!z_prev || (coef_prev && h_cur), "previous-stage inputs incomplete");
    GYMRL_REQUIRE(!g || (w && alpha && beta && coef_cur && h_pre), "next-stage parameters incomplete");
    GYMRL_REQUIRE(!final_weight || feat, "feat is NULL");
    GYMRL_REQUIRE(prev_row_stride % 4 == 0 && prev_branch_stride % 4 == 0, "strides must be multiples of 4 floats");
    if (M == 0) return GYMRL_OK;
    cudaStream_t s = as_stream(stream);
    const int grid = grid_for_rows(M);
    if (D == 128)
        mhc_stage_fwd_kernel<1><<<grid, kWarpsPerBlock * 32, 0, s>>>(h_prev, prev_row_stride, prev_branch_stride, z_prev, coef_prev, h_cur, g,
                                                                     w, alpha, beta, coef_cur, h_pre, final_weight, feat, M, sk_iters, eps);
    else
        mhc_stage_fwd_kernel<2><<<grid, kWarpsPerBlock * 32, 0, s>>>(h_prev, prev_row_stride, prev_branch_stride, z_prev, coef_prev, h_cur, g,
                                                                     w, alpha, beta, coef_cur, h_pre, final_weight, feat, M, sk_iters, eps);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("mhc_stage_forward");
    return GYMRL_OK;
}

extern "C" int gymrl_mhc_stage_backward_a(const float* h, int row_stride, int branch_stride, const float* z, const float* dh_next,
                                          float* dz, float* dh_partial, float* coef, int M, int D, void* stream) {
    GYMRL_REQUIRE(h && z && dh_next && dz && dh_partial && coef, "NULL argument");
    GYMRL_REQUIRE(D == 128 || D == 256, "mHC kernels are built for mhc_dim 128 or 256 (got %d)", D);
    if (M == 0) return GYMRL_OK;
    cudaStream_t s = as_stream(stream);
    const int grid = grid_for_rows(M);
    if (D == 128) mhc_stage_bwd_a_kernel<1><<<grid, kWarpsPerBlock * 32, 0, s>>>(h, row_stride, branch_stride, z, dh_next, dz, dh_partial, coef, M);
    else mhc_stage_bwd_a_kernel<2><<<grid, kWarpsPerBlock * 32, 0, s>>>(h, row_stride, branch_stride, z, dh_next, dz, dh_partial, coef, M);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("mhc_stage_backward_a");
    return GYMRL_OK;
}

extern "C" int gymrl_mhc_stage_backward_b(const float* h, int row_stride, int branch_stride, const float* dh_pre, const float* scratch,
                                          const float* dh_partial, float* dh, float* dx0, const float* g, const float* w,
                                          const float* alpha, float* dg, float* dw, float* dalpha, float* dbeta, void* workspace,
                                          size_t workspace_bytes, int accumulate, int M, int D, void* stream) {
    GYMRL_REQUIRE(h && dh_pre && scratch && dh_partial && (dh || dx0) && g && w && alpha && dg && dw && dalpha && dbeta && workspace,
                  "NULL argument");
    GYMRL_REQUIRE(D == 128 || D == 256, "mHC kernels are built for mhc_dim 128 or 256 (got %d)", D);
    const int P = 2 * D * 8 + 2 * D + 11;
    const int grid = grid_for_rows(M > 0 ? M : 1);
    GYMRL_REQUIRE(workspace_bytes >= (size_t)grid * P * sizeof(float), "workspace too small: need %zu bytes",
                  (size_t)grid * P * sizeof(float));
    cudaStream_t s = as_stream(stream);
    const size_t smem = ((size_t)((P + 3) & ~3) + (size_t)2 * D * 8) * sizeof(float);   // accumulator + the stage's w (k-major)
    if (D == 128) {
        mhc_stage_bwd_b_kernel<1><<<grid, kWarpsPerBlock * 32, smem, s>>>(h, row_stride, branch_stride, dh_pre, scratch, dh_partial, dh, dx0,
                                                                          g, w, alpha, (float*)workspace, M);
    } else {
        static bool attr = false;
        if (!attr) {
            GYMRL_CUDA(cudaFuncSetAttribute(mhc_stage_bwd_b_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr = true;
        }
        mhc_stage_bwd_b_kernel<2><<<grid, kWarpsPerBlock * 32, smem, s>>>(h, row_stride, branch_stride, dh_pre, scratch, dh_partial, dh, dx0,
                                                                          g, w, alpha, (float*)workspace, M);
    }
    launch_reduce_blocks((const float*)workspace, grid, P, dw, 2 * D * 8, dg, 2 * D, dalpha, 3, dbeta, 8,
                                                         accumulate, s);
    gymrl_count_launch(2);
    GYMRL_LAUNCH_CHECK("mhc_stage_backward_b");
    return GYMRL_OK;
}

template <bool SILU, bool SUM2>
static int launch_rms_fwd(const float* x, int ldx, const float* weight, float* y, int ldy, int M, int W, int groups, float eps, cudaStream_t s) {
    const int grid = grid_for_rows((long long)M * groups);
    if (W == 128) rmsnorm_fwd_kernel<1, SILU, SUM2><<<grid, kWarpsPerBlock * 32, 0, s>>>(x, ldx, weight, y, ldy, M, groups, eps);
    else rmsnorm_fwd_kernel<2, SILU, SUM2><<<grid, kWarpsPerBlock * 32, 0, s>>>(x, ldx, weight, y, ldy, M, groups, eps);
    return 0;
}

extern "C" int gymrl_rmsnorm_forward(const float* x, int ldx, int sum2, int silu, const float* weight, float* y, int ldy, int M, int W,
                                     int groups, float eps, void* stream) {
    GYMRL_REQUIRE(x && weight && y, "NULL argument");
    GYMRL_REQUIRE(W == 128 || W == 256, "RMSNorm kernels are built for widths 128 and 256 (got %d)", W);
    GYMRL_REQUIRE(groups >= 1 && !(sum2 && (silu || groups != 1)), "unsupported RMSNorm variant");
    GYMRL_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0, "leading dimensions must be multiples of 4 floats");
    if (M == 0) return GYMRL_OK;
    cudaStream_t s = as_stream(stream);
    if (sum2) launch_rms_fwd<false, true>(x, ldx, weight, y, ldy, M, W, groups, eps, s);
    else if (silu) launch_rms_fwd<true, false>(x, ldx, weight, y, ldy, M, W, groups, eps, s);
    else launch_rms_fwd<false, false>(x, ldx, weight, y, ldy, M, W, groups, eps, s);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("rmsnorm_forward");
    return GYMRL_OK;
}

template <bool SILU, bool SUM2>
static void launch_rms_bwd(const float* x, int ldx, const float* weight, const float* dy, int lddy, float* dx, int lddx, float* partials,
                           int grid, int M, int W, int groups, float eps, cudaStream_t s) {
    const size_t smem = (size_t)groups * W * sizeof(float);
    if (W == 128) rmsnorm_bwd_kernel<1, SILU, SUM2><<<grid, kWarpsPerBlock * 32, smem, s>>>(x, ldx, weight, dy, lddy, dx, lddx, partials, M, groups, eps);
    else rmsnorm_bwd_kernel<2, SILU, SUM2><<<grid, kWarpsPerBlock * 32, smem, s>>>(x, ldx, weight, dy, lddy, dx, lddx, partials, M, groups, eps);
}

extern "C" int gymrl_rmsnorm_backward(const float* x, int ldx, int sum2, int silu, const float* weight, const float* dy, int lddy,
                                      float* dx, int lddx, float* dweight, void* workspace, size_t workspace_bytes, int accumulate,
                                      int M, int W, int groups, float eps, void* stream) {
    GYMRL_REQUIRE(x && weight && dy && dx && dweight && workspace, "NULL argument");
    GYMRL_REQUIRE(W == 128 || W == 256, "RMSNorm kernels are built for widths 128 and 256 (got %d)", W);
    GYMRL_REQUIRE(groups >= 1 && groups <= 8 && (kWarpsPerBlock % groups) == 0 && !(sum2 && (silu || groups != 1)),
                  "unsupported RMSNorm variant");
    const int P = groups * W;
    const int grid = grid_for_rows((long long)(M > 0 ? M : 1) * groups);
    GYMRL_REQUIRE(workspace_bytes >= (size_t)grid * P * sizeof(float), "workspace too small: need %zu bytes",
                  (size_t)grid * P * sizeof(float));
    cudaStream_t s = as_stream(stream);
    float* partials = (float*)workspace;
    if (sum2) launch_rms_bwd<false, true>(x, ldx, weight, dy, lddy, dx, lddx, partials, grid, M, W, groups, eps, s);
    else if (silu) launch_rms_bwd<true, false>(x, ldx, weight, dy, lddy, dx, lddx, partials, grid, M, W, groups, eps, s);
    else launch_rms_bwd<false, false>(x, ldx, weight, dy, lddy, dx, lddx, partials, grid, M, W, groups, eps, s);
    launch_reduce_blocks(partials, grid, P, dweight, P, nullptr, 0, nullptr, 0, nullptr, 0, accumulate, s);
    gymrl_count_launch(2);
    GYMRL_LAUNCH_CHECK("rmsnorm_backward");
    return GYMRL_OK;
}
