// sample.cu — action selection kernels.
//
//  - gymrl_sample_categorical: Categorical(logits).sample()/log_prob/entropy as torch computes them
//    (ActorCritic.get_action, algorithms/ppo_lunarlander.py:92-104; ppo_full_lunarlander.py:395-407).
//  - gymrl_select_eps_greedy: DQNTrainer.select_action (algorithms/dqn_cartpole.py:124-133).
//  - gymrl_sample_tanh_gaussian: Actor.sample/get_action (algorithms/sac_pendulum.py:76-98).
//  - gymrl_add_gaussian_noise_clip: TD3 exploration / target smoothing (algorithms/td3_pendulum.py:157-170,194-204).
//
// One thread per env row; n_actions/act_dim are tiny (<= 16), rows are read as contiguous floats so a
// warp covers 32 consecutive rows (512 B for A = 4).  Noise comes either from the caller (parity
// tests feed the reference's own draws) or from Philox keyed by (seed, global env id, draw).
#include "common.cuh"

void gymrl_count_launch(int n = 1);

#define MAX_A 16

__device__ __forceinline__ void normalized_logits(const float* z, int A, float* ln, float* p) {
    float m = z[0];
    for (int j = 1; j < A; ++j) m = fmaxf(m, z[j]);
    float s = 0.f;
    for (int j = 0; j < A; ++j) s += expf(z[j] - m);
    const float lse = logf(s) + m;
    float m2 = -INFINITY;
    for (int j = 0; j < A; ++j) { ln[j] = z[j] - lse; m2 = fmaxf(m2, ln[j]); }
    float s2 = 0.f;
    for (int j = 0; j < A; ++j) { p[j] = expf(ln[j] - m2); s2 += p[j]; }
    for (int j = 0; j < A; ++j) p[j] = p[j] / s2;
}

__global__ void sample_categorical_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ noise,
                                          int32_t* __restrict__ action, float* __restrict__ logp,
                                          float* __restrict__ entropy, const float* __restrict__ value_in, int ldv,
                                          float* __restrict__ value_out, int n, int A, uint64_t seed, uint64_t first_id,
                                          uint32_t draw, const uint32_t* __restrict__ draw_base, int deterministic) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (draw_base) draw += *draw_base;
    float z[MAX_A], ln[MAX_A], p[MAX_A];
    for (int j = 0; j < A; ++j) z[j] = logits[(size_t)i * ld + j];
    normalized_logits(z, A, ln, p);
    int best = 0;
    if (deterministic) {
        float bv = z[0];
        for (int j = 1; j < A; ++j) if (z[j] > bv) { bv = z[j]; best = j; }
    } else {
        float bv = -INFINITY;
        for (int j = 0; j < A; ++j) {
            float q;
            if (noise) {
                q = noise[(size_t)i * A + j];
            } else {
                const u32x4 r = philox_draw(seed, first_id + i, draw, PHILOX_ACTION | ((uint32_t)(j >> 2) << 8));
                const uint32_t w = (j & 3) == 0 ? r.x : ((j & 3) == 1 ? r.y : ((j & 3) == 2 ? r.z : r.w));
                q = -logf(u01_open0_f32(w));
            }
            const float val = p[j] / q;
            if (val > bv) { bv = val; best = j; }
        }
    }
    action[i] = best;
    if (logp) logp[i] = ln[best];
    if (value_out) value_out[i] = value_in[(size_t)i * ldv];
    if (entropy) {
        float h = 0.f;
        for (int j = 0; j < A; ++j) h += fmaxf(ln[j], -3.402823466e+38f) * p[j];
        entropy[i] = -h;
    }
}

extern "C" int gymrl_sample_categorical(const float* d_logits, int ld_logits, const float* d_noise, int32_t* d_action,
                                        float* d_logp, float* d_entropy, const float* d_value_in, int ld_value_in,
                                        float* d_value_out, int n, int n_actions, uint64_t seed,
                                        uint64_t first_id, uint32_t draw, const uint32_t* d_draw_base, int deterministic, void* stream) {
    GYMRL_REQUIRE(d_logits && d_action, "NULL logits/action");
    GYMRL_REQUIRE(n > 0 && n_actions > 0 && n_actions <= MAX_A, "bad n=%d or n_actions=%d (max %d)", n, n_actions, MAX_A);
    GYMRL_REQUIRE(ld_logits >= n_actions, "ld_logits < n_actions");
    GYMRL_REQUIRE(!d_value_out || d_value_in, "value_out without value_in");
    sample_categorical_kernel<<<ceil_div(n, 128), 128, 0, as_stream(stream)>>>(
        d_logits, ld_logits, d_noise, d_action, d_logp, d_entropy, d_value_in, ld_value_in, d_value_out, n, n_actions, seed,
        first_id, draw, d_draw_base, deterministic);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("sample_categorical");
    return GYMRL_OK;
}

// ---- rollout tail in one launch: actor head + critic head + Categorical sample ----------------------------------------------
// h [n][2H] = (actor | critic) head-trunk activations; logits = Wa h_a + ba (Wa [A][H]), V = Wc h_c + bc; then exactly what
// sample_categorical_kernel does on those logits.  One warp per row; the dot products repeat skinny_fwd_kernel's arithmetic
// (lane-strided float4 FMAs in the same order, the same butterfly), so logits, actions and log-probs are bit-identical to the
// three separate launches (gymrl_linear_forward x 2 + gymrl_sample_categorical) this replaces in the PPO rollout graph.
#define PHS_MAX_A 8
__global__ void __launch_bounds__(256) policy_heads_sample_kernel(const float* __restrict__ h, int ldh, const float* __restrict__ Wa,
                                                                  const float* __restrict__ ba, const float* __restrict__ Wc,
                                                                  const float* __restrict__ bc, int H, int A, int32_t* __restrict__ action,
                                                                  float* __restrict__ logp, float* __restrict__ entropy,
                                                                  float* __restrict__ value, float* __restrict__ lv_out, int n,
                                                                  uint64_t seed, uint64_t first_id, uint32_t draw,
                                                                  const uint32_t* __restrict__ draw_base, int deterministic) {
    extern __shared__ __align__(16) float sw[];   // [A + 1][H]: Wa rows, then Wc
    for (int i = threadIdx.x; i < A * H; i += blockDim.x) sw[i] = Wa[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) sw[A * H + i] = Wc[i];
    __syncthreads();
    if (draw_base) draw += *draw_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    for (int m = blockIdx.x * wpb + warp; m < n; m += gridDim.x * wpb) {
        const float* hp = h + (size_t)m * ldh;
        float acc[PHS_MAX_A + 1];
#pragma unroll
        for (int a = 0; a <= PHS_MAX_A; ++a) acc[a] = 0.f;
        for (int k = lane * 4; k < H; k += 128) {
            const float4 va = *reinterpret_cast<const float4*>(hp + k);
            const float4 vc = *reinterpret_cast<const float4*>(hp + H + k);
#pragma unroll
            for (int a = 0; a < PHS_MAX_A; ++a) {
                if (a < A) {
                    const float4 ww = *reinterpret_cast<const float4*>(&sw[a * H + k]);
                    acc[a] = fmaf(va.x, ww.x, fmaf(va.y, ww.y, fmaf(va.z, ww.z, fmaf(va.w, ww.w, acc[a]))));
                }
            }
            const float4 wc = *reinterpret_cast<const float4*>(&sw[A * H + k]);
            acc[PHS_MAX_A] = fmaf(vc.x, wc.x, fmaf(vc.y, wc.y, fmaf(vc.z, wc.z, fmaf(vc.w, wc.w, acc[PHS_MAX_A]))));
        }
        float z[PHS_MAX_A], ln[PHS_MAX_A], p[PHS_MAX_A];
#pragma unroll
        for (int a = 0; a < PHS_MAX_A; ++a) {
            if (a < A) z[a] = warp_sum(acc[a]) + ba[a];
        }
        const float V = warp_sum(acc[PHS_MAX_A]) + bc[0];
        normalized_logits(z, A, ln, p);
        int best = 0;
        if (deterministic) {
            float bv = z[0];
            for (int j = 1; j < A; ++j) if (z[j] > bv) { bv = z[j]; best = j; }
        } else {
            float bv = -INFINITY;
            for (int j = 0; j < A; ++j) {
                const u32x4 r = philox_draw(seed, first_id + m, draw, PHILOX_ACTION | ((uint32_t)(j >> 2) << 8));
                const uint32_t w = (j & 3) == 0 ? r.x : ((j & 3) == 1 ? r.y : ((j & 3) == 2 ? r.z : r.w));
                const float q = -logf(u01_open0_f32(w));
                const float val = p[j] / q;
                if (val > bv) { bv = val; best = j; }
            }
        }
        if (lane == 0) {
            action[m] = best;
            if (logp) logp[m] = ln[best];
            if (value) value[m] = V;
            if (entropy) {
                float e = 0.f;
                for (int j = 0; j < A; ++j) e += fmaxf(ln[j], -3.402823466e+38f) * p[j];
                entropy[m] = -e;
            }
            if (lv_out) {
                for (int j = 0; j < A; ++j) lv_out[(size_t)m * 8 + j] = z[j];
                lv_out[(size_t)m * 8 + A] = V;
            }
        }
    }
}

extern "C" int gymrl_policy_heads_sample(const float* d_h, int ldh, const float* d_Wa, const float* d_ba, const float* d_Wc,
                                         const float* d_bc, int H, int n_actions, int32_t* d_action, float* d_logp, float* d_entropy,
                                         float* d_value, float* d_lv_out, int n, uint64_t seed, uint64_t first_id, uint32_t draw,
                                         const uint32_t* d_draw_base, int deterministic, void* stream) {
    GYMRL_REQUIRE(d_h && d_Wa && d_ba && d_Wc && d_bc && d_action, "NULL pointer");
    GYMRL_REQUIRE(n > 0 && n_actions > 0 && n_actions <= 7, "bad n=%d or n_actions=%d (max 7)", n, n_actions);
    GYMRL_REQUIRE(H > 0 && H % 4 == 0 && H <= 2048 && ldh >= 2 * H && ldh % 4 == 0 && ((reinterpret_cast<uintptr_t>(d_h) & 15) == 0),
                  "h must be a 16-byte aligned [n][>= 2H] matrix with H a multiple of 4");
    int blocks = ceil_div(n, 8);
    if (blocks > GYMRL_NUM_SMS * 8) blocks = GYMRL_NUM_SMS * 8;
    policy_heads_sample_kernel<<<blocks, 256, (size_t)(n_actions + 1) * H * sizeof(float), as_stream(stream)>>>(
        d_h, ldh, d_Wa, d_ba, d_Wc, d_bc, H, n_actions, d_action, d_logp, d_entropy, d_value, d_lv_out, n, seed, first_id, draw,
        d_draw_base, deterministic);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("policy_heads_sample");
    return GYMRL_OK;
}

__global__ void eps_greedy_kernel(const float* __restrict__ q, int ld, int32_t* __restrict__ action, int n, int A,
                                  float eps, uint64_t seed, uint64_t first_id, uint32_t draw, const uint32_t* __restrict__ draw_base,
                                  const float* __restrict__ eps_dev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (draw_base) draw += *draw_base;
    if (eps_dev) eps = *eps_dev;     // host-scheduled epsilon in a device slot: a captured lockstep graph follows the decay
    const u32x4 r = philox_draw(seed, first_id + i, draw, PHILOX_ACTION);
    int a;
    if (u01_f32(r.x) < eps) {
        a = (int)(u01_f32(r.y) * (float)A);
        if (a >= A) a = A - 1;
    } else {
        a = 0;
        float bv = q[(size_t)i * ld];
        for (int j = 1; j < A; ++j) {
            const float v = q[(size_t)i * ld + j];
            if (v > bv) { bv = v; a = j; }
        }
    }
    action[i] = a;
}

extern "C" int gymrl_select_eps_greedy(const float* d_q, int ld_q, int32_t* d_action, int n, int n_actions, float eps,
                                       uint64_t seed, uint64_t first_id, uint32_t draw, const uint32_t* d_draw_base, void* stream) {
    GYMRL_REQUIRE(d_q && d_action, "NULL q/action");
    GYMRL_REQUIRE(n > 0 && n_actions > 0 && ld_q >= n_actions, "bad shape");
    eps_greedy_kernel<<<ceil_div(n, 128), 128, 0, as_stream(stream)>>>(d_q, ld_q, d_action, n, n_actions, eps, seed,
                                                                      first_id, draw, d_draw_base, nullptr);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("eps_greedy");
    return GYMRL_OK;
}

extern "C" int gymrl_select_eps_greedy_dev(const float* d_q, int ld_q, int32_t* d_action, int n, int n_actions, const float* d_eps,
                                           uint64_t seed, uint64_t first_id, uint32_t draw, const uint32_t* d_draw_base, void* stream) {
    GYMRL_REQUIRE(d_q && d_action && d_eps, "NULL q/action/eps");
    GYMRL_REQUIRE(n > 0 && n_actions > 0 && ld_q >= n_actions, "bad shape");
    eps_greedy_kernel<<<ceil_div(n, 128), 128, 0, as_stream(stream)>>>(d_q, ld_q, d_action, n, n_actions, 0.0f, seed,
                                                                      first_id, draw, d_draw_base, d_eps);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("eps_greedy_dev");
    return GYMRL_OK;
}

// N(0,1) pair by Box-Muller from two 24-bit uniforms
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float u1 = u01_open0_f32(a), u2 = u01_f32(b);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincosf(6.283185307179586f * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t id, uint32_t draw, uint32_t stream, int j) {
    const u32x4 r = philox_draw(seed, id, draw, stream | ((uint32_t)(j >> 2) << 8));
    float z0, z1;
    if ((j & 2) == 0) box_muller(r.x, r.y, z0, z1);
    else box_muller(r.z, r.w, z0, z1);
    return (j & 1) == 0 ? z0 : z1;
}

__global__ void tanh_gaussian_kernel(const float* __restrict__ mean, const float* __restrict__ log_std, int ld,
                                     const float* __restrict__ noise, float* __restrict__ action,
                                     float* __restrict__ logp, float* __restrict__ pre_tanh, int n, int A, float bound,
                                     float ls_min, float ls_max, uint64_t seed, uint64_t first_id, uint32_t draw,
                                     const uint32_t* __restrict__ draw_base, int deterministic) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (draw_base) draw += *draw_base;
    float lp = 0.f;
    for (int j = 0; j < A; ++j) {
        const float mu = mean[(size_t)i * ld + j];
        if (deterministic) {
            action[(size_t)i * A + j] = tanhf(mu) * bound;
            if (pre_tanh) pre_tanh[(size_t)i * A + j] = mu;
            continue;
        }
        const float ls = fminf(fmaxf(log_std[(size_t)i * ld + j], ls_min), ls_max);
        const float std = expf(ls);
        const float xi = noise ? noise[(size_t)i * A + j] : philox_normal(seed, first_id + i, draw, PHILOX_ACTION, j);
        const float x = mu + std * xi;  // Normal.rsample(): loc + eps * scale
        const float t = tanhf(x);
        action[(size_t)i * A + j] = t * bound;
        if (pre_tanh) pre_tanh[(size_t)i * A + j] = x;
        // Normal.log_prob: -((x-mu)^2)/(2 var) - log(scale) - log(sqrt(2 pi))
        const float var = std * std;
        float l = -((x - mu) * (x - mu)) / (2.0f * var) - logf(std) - 0.9189385332046727f;
        l -= logf(bound * (1.0f - t * t) + 1e-6f);
        lp += l;
    }
    if (logp && !deterministic) logp[i] = lp;
}

extern "C" int gymrl_sample_tanh_gaussian(const float* d_mean, const float* d_log_std, int ld, const float* d_noise,
                                          float* d_action, float* d_logp, float* d_pre_tanh, int n, int act_dim,
                                          float bound, float log_std_min, float log_std_max, uint64_t seed,
                                          uint64_t first_id, uint32_t draw, const uint32_t* d_draw_base, int deterministic, void* stream) {
    GYMRL_REQUIRE(d_mean && d_action, "NULL mean/action");
    GYMRL_REQUIRE(deterministic || d_log_std, "log_std required for stochastic sampling");
    GYMRL_REQUIRE(n > 0 && act_dim > 0 && ld >= act_dim, "bad shape");
    tanh_gaussian_kernel<<<ceil_div(n, 128), 128, 0, as_stream(stream)>>>(
        d_mean, d_log_std, ld, d_noise, d_action, d_logp, d_pre_tanh, n, act_dim, bound, log_std_min, log_std_max, seed,
        first_id, draw, d_draw_base, deterministic);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("tanh_gaussian");
    return GYMRL_OK;
}

__global__ void gaussian_noise_clip_kernel(const float* __restrict__ mu, const float* __restrict__ noise,
                                           float* __restrict__ action, int n, int A, float sigma, float noise_clip,
                                           float bound, uint64_t seed, uint64_t first_id, uint32_t draw,
                                           const uint32_t* __restrict__ draw_base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (draw_base) draw += *draw_base;
    for (int j = 0; j < A; ++j) {
        const float xi = noise ? noise[(size_t)i * A + j] : philox_normal(seed, first_id + i, draw, PHILOX_UPDATE, j);
        float nz = xi * sigma;
        if (noise_clip > 0.f) nz = fminf(fmaxf(nz, -noise_clip), noise_clip);
        const float a = mu[(size_t)i * A + j] + nz;
        action[(size_t)i * A + j] = fminf(fmaxf(a, -bound), bound);
    }
}

extern "C" int gymrl_add_gaussian_noise_clip(const float* d_mu, const float* d_noise, float* d_action, int n,
                                             int act_dim, float sigma, float noise_clip, float bound, uint64_t seed,
                                             uint64_t first_id, uint32_t draw, const uint32_t* d_draw_base, void* stream) {
    GYMRL_REQUIRE(d_mu && d_action, "NULL mu/action");
    GYMRL_REQUIRE(n > 0 && act_dim > 0, "bad shape");
    gaussian_noise_clip_kernel<<<ceil_div(n, 128), 128, 0, as_stream(stream)>>>(d_mu, d_noise, d_action, n, act_dim, sigma,
                                                                                noise_clip, bound, seed, first_id, draw, d_draw_base);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("gaussian_noise_clip");
    return GYMRL_OK;
}

// out[i] ~ N(0,1) from Philox (entity = entity0 + i/4): the noise buffers of Actor.sample's rsample (sac :80) and of
// torch.randn_like (td3 :195), kept so the backward pass can reuse the exact draw.
__global__ void fill_normal_kernel(float* __restrict__ out, int n, uint64_t seed, uint64_t entity0, uint32_t draw,
                                   const uint32_t* __restrict__ draw_base, uint32_t stream_id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (draw_base) draw += *draw_base;
    out[i] = philox_normal(seed, entity0 + (uint64_t)(i >> 2), draw, stream_id, i & 3);
}
extern "C" int gymrl_fill_normal(float* d_out, int n, uint64_t seed, uint64_t entity0, uint32_t draw, const uint32_t* d_draw_base,
                                 void* stream) {
    GYMRL_REQUIRE(d_out && n > 0, "bad arguments");
    fill_normal_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(d_out, n, seed, entity0, draw, d_draw_base, PHILOX_UPDATE);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("fill_normal");
    return GYMRL_OK;
}
