// normalize.cu — running observation normalisation and reward scaling of the utils path.
//
// Reference: utils/normalization.py  RunningMeanStd :4-22, Normalization :25-35, RewardScaling :38-52, applied once
// per env step by utils/runner.py:112,125-126.  The reference keeps ONE statistic per agent and feeds it one
// observation at a time.  Here a step delivers N observations (one per env copy):
//   * N <= 32: they are fed one after the other in env order with the reference's exact update rule and mixed
//     precision (mean float32, S float64, std float64; first-sample quirk mean = std = x, q16), so N = 1 reproduces
//     the reference bit for bit;
//   * N  > 32: the batch mean / M2 are reduced in float64 and merged (Chan et al.), n += N.
// Normalisation then uses the statistic after the whole batch.
// state layout (float64): [0] n, [1 .. D] mean, [1+D .. 2D] S, [1+2D .. 3D] std.
#include "common.cuh"

void gymrl_count_launch(int n = 1);

namespace {

__global__ void running_update_kernel(const float* __restrict__ x, int N, int D, double* __restrict__ st) {
    __shared__ double s_a[32], s_b[32];
    const int d = blockIdx.x;
    double* mean = st + 1 + d;
    double* S = st + 1 + D + d;
    double* sd = st + 1 + 2 * D + d;
    const double n0 = st[0];
    if (N <= 32) {
        if (threadIdx.x == 0) {
            double n = n0, Sv = *S, sdv = *sd;
            float m = (float)*mean;
            for (int i = 0; i < N; ++i) {
                const float xi = x[(size_t)i * D + d];
                n += 1.0;
                if (n == 1.0) {
                    m = xi;
                    sdv = (double)xi;
                } else {
                    const float old = m;
                    m = old + (xi - old) / (float)n;
                    Sv = Sv + (double)((xi - old) * (xi - m));
                    sdv = sqrt(Sv / n);
                }
            }
            *mean = (double)m; *S = Sv; *sd = sdv;
        }
        return;
    }
    double sum = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) sum += (double)x[(size_t)i * D + d];
    sum = block_sum(sum, s_a);
    __shared__ double s_mean;
    if (threadIdx.x == 0) s_mean = sum / N;
    __syncthreads();
    const double bm = s_mean;
    double m2 = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double dv = (double)x[(size_t)i * D + d] - bm;
        m2 += dv * dv;
    }
    m2 = block_sum(m2, s_b);
    if (threadIdx.x == 0) {
        const double n1 = n0 + N, delta = bm - *mean;
        const double mnew = n0 == 0.0 ? bm : *mean + delta * N / n1;
        const double Snew = n0 == 0.0 ? m2 : *S + m2 + delta * delta * n0 * N / n1;
        *mean = mnew; *S = Snew; *sd = sqrt(Snew / n1);
    }
}

__global__ void running_count_kernel(double* st, int N) { st[0] += (double)N; }

__global__ void running_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, long long total, int D,
                                         const double* __restrict__ st, int center) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int d = (int)(i % D);
    const float m = (float)st[1 + d];
    const double sd = st[1 + 2 * D + d];
    const float c = center ? x[i] - m : x[i];
    y[i] = (float)((double)c / (sd + 1e-8));
}

// R = gamma R + r per env; statistic over R (shape 1) fed in env order; out = r / (std + 1e-8)
__global__ void reward_scaling_kernel(const float* __restrict__ r, float* __restrict__ out, double* __restrict__ R,
                                      const uint8_t* __restrict__ reset, double gamma, double* __restrict__ st, int N) {
    __shared__ double s_a[32], s_b[32];
    __shared__ double s_mean, s_std;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double prev = (reset && reset[i]) ? 0.0 : R[i];
        R[i] = gamma * prev + (double)r[i];
    }
    __syncthreads();
    if (N <= 32) {
        if (threadIdx.x == 0) {
            double n = st[0], Sv = st[2], sdv = st[3];
            float m = (float)st[1];
            for (int i = 0; i < N; ++i) {
                const float xi = (float)R[i];    // RunningMeanStd.update casts to float32 (:13)
                n += 1.0;
                if (n == 1.0) { m = xi; sdv = (double)xi; }
                else {
                    const float old = m;
                    m = old + (xi - old) / (float)n;
                    Sv = Sv + (double)((xi - old) * (xi - m));
                    sdv = sqrt(Sv / n);
                }
            }
            st[0] = n; st[1] = (double)m; st[2] = Sv; st[3] = sdv;
            s_std = sdv;
        }
    } else {
        double sum = 0.0;
        for (int i = threadIdx.x; i < N; i += blockDim.x) sum += (double)(float)R[i];
        sum = block_sum(sum, s_a);
        if (threadIdx.x == 0) s_mean = sum / N;
        __syncthreads();
        double m2 = 0.0;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const double dv = (double)(float)R[i] - s_mean;
            m2 += dv * dv;
        }
        m2 = block_sum(m2, s_b);
        if (threadIdx.x == 0) {
            const double n0 = st[0], n1 = n0 + N, delta = s_mean - st[1];
            st[1] = n0 == 0.0 ? s_mean : st[1] + delta * N / n1;
            st[2] = n0 == 0.0 ? m2 : st[2] + m2 + delta * delta * n0 * N / n1;
            st[3] = sqrt(st[2] / n1);
            st[0] = n1;
            s_std = st[3];
        }
    }
    __syncthreads();
    const double sd = s_std;
    for (int i = threadIdx.x; i < N; i += blockDim.x) out[i] = (float)((double)r[i] / (sd + 1e-8));
}

}  // namespace

extern "C" int gymrl_running_stats_update(const float* d_x, int N, int D, double* d_state, void* stream) {
    GYMRL_REQUIRE(d_x && d_state && N >= 1 && D >= 1, "bad arguments");
    cudaStream_t s = as_stream(stream);
    running_update_kernel<<<D, N <= 32 ? 32 : 256, 0, s>>>(d_x, N, D, d_state);
    running_count_kernel<<<1, 1, 0, s>>>(d_state, N);
    gymrl_count_launch(2);
    GYMRL_LAUNCH_CHECK("running_stats_update");
    return GYMRL_OK;
}

extern "C" int gymrl_running_normalize(const float* d_x, float* d_y, int N, int D, const double* d_state, int center, void* stream) {
    GYMRL_REQUIRE(d_x && d_y && d_state && N >= 1 && D >= 1, "bad arguments");
    const long long total = (long long)N * D;
    running_normalize_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, as_stream(stream)>>>(d_x, d_y, total, D, d_state, center);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("running_normalize");
    return GYMRL_OK;
}

extern "C" int gymrl_reward_scaling(const float* d_r, float* d_out, double* d_R, const uint8_t* d_reset, double gamma,
                                    double* d_state, int N, void* stream) {
    GYMRL_REQUIRE(d_r && d_out && d_R && d_state && N >= 1, "bad arguments");
    reward_scaling_kernel<<<1, N <= 32 ? 32 : 256, 0, as_stream(stream)>>>(d_r, d_out, d_R, d_reset, gamma, d_state, N);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("reward_scaling");
    return GYMRL_OK;
}
