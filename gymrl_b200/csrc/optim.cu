// optim.cu — gradient norm, fused clip + Adam, Polyak/hard target sync, minibatch permutation
// (SURVEY §8 a9, a16, a17).
//
//  - gymrl_adam_step: torch.optim.Adam's single-tensor update (lerp first moment, addcmul second
//    moment, bias-corrected step, eps added after the sqrt) on one flat fp32 buffer, with
//    nn.utils.clip_grad_norm_ (algorithms/ppo_lunarlander.py:304-306, rainbow_dqn_cartpole.py:344)
//    or the per-element clamp of algorithms/dqn_cartpole.py:163-165 folded into the gradient read.
//    Streaming kernel: 16 B of reads + 12 B of writes per parameter... the whole state of these
//    networks (0.27-0.8 MB x 4 arrays) is L2 resident, so this is launch/latency bound; it exists to
//    replace ~40 eager multi-tensor launches per optimizer step with one.
//  - gymrl_polyak: target <- tau*source + (1-tau)*target.
//  - gymrl_random_permutation: bijective mixing network + cycle walking (no sort, O(1) per index).
#include "common.cuh"
#include "wimages.cuh"

void gymrl_count_launch(int n = 1);

__global__ void grad_sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
    __shared__ double scratch[32];
    double q = 0.0;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
    if (vec) {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
            const float4 v = g4[i];
            q += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
        }
        for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
            q += (double)g[i] * g[i];
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
            q += (double)g[i] * g[i];
    }
    q = block_sum(q, scratch);
    const double v[1] = {q};
    ordered_block_accumulate<1>(v, out, scratch);      // fixed block order: bitwise reproducible norm
}

extern "C" int gymrl_grad_sumsq(const float* d_grad, long long n, double* d_sumsq, void* stream) {
    GYMRL_REQUIRE(d_grad && d_sumsq && n > 0, "bad arguments");
    const int threads = 256;
    long long blocks = ceil_div_ll(n, (long long)threads * 4);
    if (blocks > GYMRL_NUM_SMS * 2) blocks = GYMRL_NUM_SMS * 2;
    if (blocks < 1) blocks = 1;
    grad_sumsq_kernel<<<(int)blocks, threads, 0, as_stream(stream)>>>(d_grad, n, d_sumsq);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("grad_sumsq");
    return GYMRL_OK;
}

__global__ void adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, long long n, const double* __restrict__ lr, float beta1, float beta2,
                            float eps, const int32_t* __restrict__ step, const double* __restrict__ sumsq, float max_norm,
                            float clampv, float grad_scale) {
    const int t = *step + 1;
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    const float step_size = (float)(-(*lr / bc1));
    const float bc2_sqrt = (float)sqrt(bc2);
    const float w1 = (float)(1.0 - (double)beta1), w2 = (float)(1.0 - (double)beta2);
    float coef = grad_scale;
    if (sumsq) {
        const float total_norm = (float)sqrt(*sumsq * (double)grad_scale * (double)grad_scale);
        const float c = max_norm / (total_norm + 1e-6f);
        coef *= fminf(c, 1.0f);
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float g = grad[i] * coef;
        if (clampv > 0.f) g = fminf(fmaxf(g, -clampv), clampv);
        float mi = m[i], vi = v[i];
        mi = mi + w1 * (g - mi);
        vi = vi * beta2 + (w2 * g) * g;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        param[i] = param[i] + (step_size * mi) / denom;
        m[i] = mi;
        v[i] = vi;
    }
}
__global__ void adam_post_kernel(int32_t* step, double* sumsq) {
    *step += 1;
    if (sumsq) *sumsq = 0.0;
}

extern "C" int gymrl_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, long long n,
                               const double* d_lr, float beta1, float beta2, float eps, int32_t* d_step,
                               const double* d_sumsq, float max_norm, float clamp, float grad_scale, void* stream) {
    GYMRL_REQUIRE(d_param && d_grad && d_exp_avg && d_exp_avg_sq && d_lr && d_step && n > 0, "bad arguments");
    cudaStream_t s = as_stream(stream);
    const int threads = 256;
    long long blocks = ceil_div_ll(n, threads);
    if (blocks > GYMRL_NUM_SMS * 4) blocks = GYMRL_NUM_SMS * 4;
    adam_kernel<<<(int)blocks, threads, 0, s>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, d_lr, beta1, beta2, eps, d_step,
                                               d_sumsq, max_norm, clamp, grad_scale);
    adam_post_kernel<<<1, 1, 0, s>>>(d_step, const_cast<double*>(d_sumsq));
    gymrl_count_launch(2);
    GYMRL_LAUNCH_CHECK("adam_step");
    wimg_refresh_if_registered(d_param, s);     // pre-split tf32 images of the weights follow every parameter write
    return GYMRL_OK;
}

// Clip + Adam with the global norm taken from per-block sums of squares (gymrl_reduce_flush) and the step counter advanced
// by the last block to finish: one launch instead of grad_sumsq + adam + adam_post.  Every block folds the partials in the
// same fixed order (thread-strided, then the block tree), so all blocks see the same norm bit for bit.
__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, const double* __restrict__ lr, float beta1,
                                                        float beta2, float eps, int32_t* __restrict__ step,
                                                        const double* __restrict__ sumsq_partials, int n_partials, float max_norm,
                                                        float grad_scale, uint32_t* __restrict__ done_counter, const WImgDev wimg) {
    __shared__ double scratch[32];
    __shared__ float s_coef;
    pdl_wait();
    pdl_launch_dependents();
    const int t = *step + 1;
    double q = 0.0;
    for (int i = threadIdx.x; i < n_partials; i += blockDim.x) q += sumsq_partials[i];
    q = block_sum(q, scratch);
    if (threadIdx.x == 0) {
        const float total_norm = (float)sqrt(q * (double)grad_scale * (double)grad_scale);
        s_coef = grad_scale * fminf(max_norm / (total_norm + 1e-6f), 1.0f);
    }
    __syncthreads();
    const float coef = s_coef;
    const double bc1 = 1.0 - pow((double)beta1, (double)t);
    const double bc2 = 1.0 - pow((double)beta2, (double)t);
    const float step_size = (float)(-(*lr / bc1));
    const float bc2_sqrt = (float)sqrt(bc2);
    const float w1 = (float)(1.0 - (double)beta1), w2 = (float)(1.0 - (double)beta2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float g = grad[i] * coef;
        float mi = m[i], vi = v[i];
        mi = mi + w1 * (g - mi);
        vi = vi * beta2 + (w2 * g) * g;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        const float pn = param[i] + (step_size * mi) / denom;
        param[i] = pn;
        m[i] = mi;
        v[i] = vi;
        if (wimg.images && i < wimg.n) wimg_store(wimg, i, pn);   // pre-split tf32 images of the weights follow the write, same pass
    }
    // every block read *step before it got here, so the last one to arrive may advance it
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(done_counter, 1u) == gridDim.x - 1) {
            *step = t;
            *done_counter = 0u;
        }
    }
}

extern "C" int gymrl_clip_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, long long n,
                                    const double* d_lr, float beta1, float beta2, float eps, int32_t* d_step,
                                    const double* d_sumsq_partials, int n_partials, float max_norm, float grad_scale,
                                    uint32_t* d_done_counter, void* stream) {
    GYMRL_REQUIRE(d_param && d_grad && d_exp_avg && d_exp_avg_sq && d_lr && d_step && d_sumsq_partials && d_done_counter && n > 0,
                  "bad arguments");
    GYMRL_REQUIRE(n_partials > 0 && max_norm > 0.f, "clip_adam_step needs the sum-of-squares partials and a positive max_norm");
    const int threads = 256;
    long long blocks = ceil_div_ll(n, threads);
    if (blocks > GYMRL_NUM_SMS * 4) blocks = GYMRL_NUM_SMS * 4;
    const WImgDev wimg = wimg_device_view(d_param);     // registered buffer: the images are re-split inside the Adam pass
    gymrl_launch_pdl(clip_adam_kernel, dim3((int)blocks), dim3(threads), 0, as_stream(stream), d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, d_lr,
                     beta1, beta2, eps, d_step, d_sumsq_partials, n_partials, max_norm, grad_scale, d_done_counter, wimg);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("clip_adam_step");
    return GYMRL_OK;
}

__global__ void polyak_kernel(float* __restrict__ target, const float* __restrict__ source, long long n, float tau) {
    const float omt = (float)(1.0 - (double)tau);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        target[i] = tau * source[i] + omt * target[i];
}

extern "C" int gymrl_polyak(float* d_target, const float* d_source, long long n, float tau, void* stream) {
    GYMRL_REQUIRE(d_target && d_source && n > 0, "bad arguments");
    const int threads = 256;
    long long blocks = ceil_div_ll(n, threads);
    if (blocks > GYMRL_NUM_SMS * 4) blocks = GYMRL_NUM_SMS * 4;
    polyak_kernel<<<(int)blocks, threads, 0, as_stream(stream)>>>(d_target, d_source, n, tau);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("polyak");
    wimg_refresh_if_registered(d_target, as_stream(stream));
    return GYMRL_OK;
}

// ---- random permutation ---------------------------------------------------------------------------
__global__ void permutation_kernel(int32_t* __restrict__ perm, int n, int bits, uint64_t seed, uint32_t draw,
                                   const uint32_t* __restrict__ draw_base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (draw_base) draw += *draw_base;
    // keys are recomputed per thread (2 Philox calls) so that the draw counter can live on the device
    // and a captured CUDA graph produces a fresh permutation on every replay
    const u32x4 a = philox_draw(seed, 0, draw, PHILOX_PERMUTE);
    const u32x4 b = philox_draw(seed, 1, draw, PHILOX_PERMUTE);
    const uint32_t mulk[4] = {a.x | 1u, a.y | 1u, a.z | 1u, a.w | 1u};
    const uint32_t addk[4] = {b.x, b.y, b.z, b.w};
    const uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    const int sh = bits > 1 ? bits / 2 : 1;
    uint32_t x = (uint32_t)i;
    do {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            x = (x * mulk[r]) & mask;
            x ^= x >> (sh + (r & 1));
            x = (x + addk[r]) & mask;
        }
    } while (x >= (uint32_t)n);
    perm[i] = (int32_t)x;
}

extern "C" int gymrl_random_permutation(int32_t* d_perm, int n, uint64_t seed, uint32_t draw, const uint32_t* d_draw_base,
                                        void* stream) {
    GYMRL_REQUIRE(d_perm && n > 0, "bad arguments");
    int bits = 1;
    while ((1ll << bits) < (long long)n) ++bits;
    permutation_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(d_perm, n, bits, seed, draw, d_draw_base);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("random_permutation");
    return GYMRL_OK;
}

__global__ void counter_add_kernel(uint32_t* c, uint32_t inc) { *c += inc; }
extern "C" int gymrl_counter_add(uint32_t* d_counter, uint32_t inc, void* stream) {
    GYMRL_REQUIRE(d_counter, "NULL counter");
    counter_add_kernel<<<1, 1, 0, as_stream(stream)>>>(d_counter, inc);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("counter_add");
    return GYMRL_OK;
}

// dst[i] = src[block * n + i], block read from device memory: lets one captured minibatch graph walk
// through the epoch's permutation (replaces indices[start:end], algorithms/ppo_lunarlander.py:264-266).
__global__ void slice_i32_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ src, int n,
                                 const uint32_t* __restrict__ block) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[(size_t)(*block) * n + i];
}
extern "C" int gymrl_slice_i32(int32_t* d_dst, const int32_t* d_src, int n, const uint32_t* d_block_index, void* stream) {
    GYMRL_REQUIRE(d_dst && d_src && d_block_index && n > 0, "bad arguments");
    slice_i32_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(d_dst, d_src, n, d_block_index);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("slice_i32");
    return GYMRL_OK;
}
