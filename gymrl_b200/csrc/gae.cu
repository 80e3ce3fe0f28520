// gae.cu — GAE / returns and advantage normalisation statistics.
//
// dialect 0: PPOTrainer.compute_gae (algorithms/ppo_lunarlander.py:179-196) and compute_advantages
//   (ppo_full_lunarlander.py:507-535).  The reference runs the recurrence in float64;
//   here fp32 in/out with a float64 recurrence, evaluated as a chunked scan over time:
//     A_t = delta_t + c_t A_{t+1}  is affine in A_{t+1}, so a chunk [t0,t1) composes to
//     A_{t0} = D + C A_{t1}.  A block owns 32 consecutive envs (lanes = envs: every [T][N] row access
//     is a coalesced 128 B line) and its W warps own W time chunks: pass 1 folds each chunk to (C, D),
//     the carries are resolved through shared memory, pass 2 re-walks the chunk (L1/L2 hits) writing
//     adv/ret.  4096 envs x 128 steps -> 128 blocks x 8 warps instead of a 128-step serial walk.
// dialect 1: ReplayBuffer_on_policy.compute_advantage (utils/buffer.py:21-35).  That loop runs in
//   float32 (NumPy weak-scalar promotion), in the order ((f32(gamma*lam) * gae) * (1-d)) + delta with
//   delta = (r + (f32(gamma) * v') * (1-dw)) - v; reproduced bit-for-bit by a serial per-env walk
//   (this file is built with -fmad=false).
#include "common.cuh"

void gymrl_count_launch(int n = 1);

#define GAE_MAX_WARPS 8

__global__ void __launch_bounds__(32 * GAE_MAX_WARPS)
gae_scan_kernel(const float* __restrict__ reward, const float* __restrict__ value, const float* __restrict__ v_last,
                const uint8_t* __restrict__ done, float* __restrict__ adv, float* __restrict__ ret, int T, int N,
                double gamma, double lam_a, double lam_c, int two_streams, int coef_f32, int boot_f32) {
    __shared__ double sC[2][GAE_MAX_WARPS][32], sD[2][GAE_MAX_WARPS][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int n = blockIdx.x * 32 + lane;
    const bool valid = n < N;
    const int chunk = (T + W - 1) / W;
    const int t0 = w * chunk, t1 = min(T, t0 + chunk);
    // NumPy >= 2 (NEP 50) rounding of the reference loops, see the dialect notes in include/gymrl.h:
    //  coef_f32: gamma*lam*(1-d) is a float32 scalar (ppo_lunarlander.py:181 makes `dones` float32)
    //  boot_f32: gamma*V_{t+1} is a float32 product (ppo_full stores values as 0-dim float32 tensors)
    const double gla = coef_f32 ? (double)(float)(gamma * lam_a) : gamma * lam_a;
    const double glc = coef_f32 ? (double)(float)(gamma * lam_c) : gamma * lam_c;
    const float gamma_f = (float)gamma;
    // pass 1: fold the chunk backwards into (C, D) for both lambda streams
    double Ca = 1.0, Da = 0.0, Cc = 1.0, Dc = 0.0;
    if (valid) {
        for (int t = t1 - 1; t >= t0; --t) {
            const size_t k = (size_t)t * N + n;
            const double nd = 1.0 - (double)done[k];
            const float vnf = (t == T - 1) ? v_last[n] : value[k + N];
            const double boot = boot_f32 ? (double)(gamma_f * vnf) : gamma * (double)vnf;
            const double delta = (double)reward[k] + boot * nd - (double)value[k];
            const double ca = gla * nd;
            Da = delta + ca * Da;
            Ca = ca * Ca;
            if (two_streams) {
                const double cc = glc * nd;
                Dc = delta + cc * Dc;
                Cc = cc * Cc;
            }
        }
    }
    sC[0][w][lane] = Ca; sD[0][w][lane] = Da;
    sC[1][w][lane] = Cc; sD[1][w][lane] = Dc;
    __syncthreads();
    // incoming A at t1 for this chunk = composition of all later chunks applied to A_T = 0
    double Ain_a = 0.0, Ain_c = 0.0;
    for (int ww = W - 1; ww > w; --ww) {
        Ain_a = sD[0][ww][lane] + sC[0][ww][lane] * Ain_a;
        if (two_streams) Ain_c = sD[1][ww][lane] + sC[1][ww][lane] * Ain_c;
    }
    // pass 2: re-walk the chunk with the resolved carry
    if (valid) {
        double Aa = Ain_a, Ac = Ain_c;
        for (int t = t1 - 1; t >= t0; --t) {
            const size_t k = (size_t)t * N + n;
            const double nd = 1.0 - (double)done[k];
            const float vnf = (t == T - 1) ? v_last[n] : value[k + N];
            const double boot = boot_f32 ? (double)(gamma_f * vnf) : gamma * (double)vnf;
            const double v = (double)value[k];
            const double delta = (double)reward[k] + boot * nd - v;
            Aa = delta + gla * nd * Aa;
            if (two_streams) Ac = delta + glc * nd * Ac;
            else Ac = Aa;
            adv[k] = (float)Aa;
            ret[k] = (float)(Ac + v);
        }
    }
}

__global__ void gae_utils_kernel(const float* __restrict__ reward, const float* __restrict__ value,
                                 const float* __restrict__ v_next, const uint8_t* __restrict__ done,
                                 const uint8_t* __restrict__ dw, float* __restrict__ adv, float* __restrict__ ret, int T,
                                 int N, float gamma, float gl) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float gae = 0.0f;
    bool first = true;
    for (int t = T - 1; t >= 0; --t) {
        const size_t k = (size_t)t * N + n;
        const float d = (float)done[k], w = (float)(dw ? dw[k] : done[k]);
        const float v = value[k];
        const float delta = (reward[k] + (gamma * v_next[k]) * (1.0f - w)) - v;
        // first iteration: gamma*lamda*0.0 is a Python float 0.0, times (1-d) -> 0
        gae = first ? (0.0f * (1.0f - d) + delta) : ((gl * gae) * (1.0f - d) + delta);
        first = false;
        adv[k] = gae;
        ret[k] = gae + v;
    }
}

extern "C" int gymrl_gae(const float* d_reward, const float* d_value, const float* d_v_last_or_next, const uint8_t* d_done,
                         const uint8_t* d_dw, float* d_adv, float* d_ret, int T, int N, double gamma, double lam_actor,
                         double lam_critic, int dialect, void* stream) {
    GYMRL_REQUIRE(d_reward && d_value && d_v_last_or_next && d_done && d_adv && d_ret, "NULL pointer");
    GYMRL_REQUIRE(T > 0 && N > 0, "bad shape T=%d N=%d", T, N);
    cudaStream_t s = as_stream(stream);
    if (dialect == 0 || dialect == 2) {
        int W = GAE_MAX_WARPS;
        while (W > 1 && T < 8 * W) W >>= 1;
        const int two = lam_actor != lam_critic;
        gae_scan_kernel<<<ceil_div(N, 32), 32 * W, 0, s>>>(d_reward, d_value, d_v_last_or_next, d_done, d_adv, d_ret, T, N,
                                                           gamma, lam_actor, lam_critic, two, dialect == 0, dialect == 2);
    } else if (dialect == 1) {
        // float32(gamma * lamda): the Python-float product is rounded once when it meets the f32 array
        const float gl = (float)(gamma * lam_actor);
        gae_utils_kernel<<<ceil_div(N, 128), 128, 0, s>>>(d_reward, d_value, d_v_last_or_next, d_done, d_dw, d_adv, d_ret, T,
                                                         N, (float)gamma, gl);
    } else {
        GYMRL_FAIL(GYMRL_EINVAL, "unknown GAE dialect %d", dialect);
    }
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("gae");
    return GYMRL_OK;
}

// ---- sum / sum of squares (float64 accumulation) and in-place normalisation ---------------------
__global__ void sum_sumsq_kernel(const float* __restrict__ x, long long n, double* __restrict__ sums) {
    __shared__ double scratch[32];
    double s = 0.0, q = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double v = (double)x[i];
        s += v;
        q += v * v;
    }
    s = block_sum(s, scratch);
    q = block_sum(q, scratch);
    const double v[2] = {s, q};
    ordered_block_accumulate<2>(v, sums, scratch);     // fixed block order: bitwise reproducible moments
}

extern "C" int gymrl_sum_sumsq(const float* d_x, long long n, double* d_sums, void* stream) {
    GYMRL_REQUIRE(d_x && d_sums && n > 0, "bad arguments");
    const int threads = 256;
    const int blocks = (int)(ceil_div_ll(n, threads) < GYMRL_NUM_SMS * 4 ? ceil_div_ll(n, threads) : GYMRL_NUM_SMS * 4);
    sum_sumsq_kernel<<<blocks, threads, 0, as_stream(stream)>>>(d_x, n, d_sums);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("sum_sumsq");
    return GYMRL_OK;
}

__global__ void normalize_kernel(float* __restrict__ x, long long n, const double* __restrict__ sums, double count, int ddof,
                                 float eps) {
    const double mean = sums[0] / count;
    double var = (sums[1] - count * mean * mean) / (count - (double)ddof);
    if (var < 0.0) var = 0.0;
    const double stdv = sqrt(var);
    const double denom = stdv + (double)eps;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = (float)(((double)x[i] - mean) / denom);
}

extern "C" int gymrl_normalize_inplace(float* d_x, long long n, const double* d_sums, double count, int ddof, float eps,
                                       void* stream) {
    GYMRL_REQUIRE(d_x && d_sums && n > 0 && count > (double)ddof, "bad arguments");
    const int threads = 256;
    const int blocks = (int)(ceil_div_ll(n, threads) < GYMRL_NUM_SMS * 8 ? ceil_div_ll(n, threads) : GYMRL_NUM_SMS * 8);
    normalize_kernel<<<blocks, threads, 0, as_stream(stream)>>>(d_x, n, d_sums, count, ddof, eps);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("normalize");
    return GYMRL_OK;
}
