// env.cu — env handle, CartPole-v1 and Pendulum-v1 lockstep kernels, episode bookkeeping.
//
// Replaces gym.make(...).reset()/.step() at the reference call sites listed in include/gymrl.h.
// The arithmetic restates gymnasium's classic_control/cartpole.py and pendulum.py (third-party,
// absent from the reference tree; see oracle/envs_np.py for the CPU restatement this is checked
// against).  State is float64 like gymnasium's; observations are the float32 casts.
//
// Mapping: one thread per env instance (state is 2-4 doubles, so a warp per env would idle 31
// lanes); a warp therefore owns 32 consecutive envs, reads/writes the SoA state planes and the
// [N][D] observation rows fully coalesced, and uses one ballot per warp to aggregate the
// finished-episode bookkeeping into a single atomic.  Compiled with -fmad=false so the float64
// update rounds exactly like the NumPy restatement.
#include "env.cuh"

#include <atomic>
#include <string>
#include <vector>

// ---- ABI basics --------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

void gymrl_set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}
void gymrl_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

extern "C" int gymrl_version(void) { return GYMRL_ABI_VERSION; }
extern "C" const char* gymrl_last_error(void) { return g_last_error.c_str(); }
extern "C" uint64_t gymrl_launch_count(void) { return g_launches.load(); }

// ---- constants ---------------------------------------------------------------------------------
#define CP_MAX_STEPS 500
#define PD_MAX_STEPS 200
#define LL_MAX_STEPS 1000
#define LL_STATE_DOUBLES 128

extern "C" int gymrl_env_info(int kind, int* obs_dim, int* act_dim, int* n_actions, int* max_episode_steps,
                              float* action_bound, int* state_doubles) {
    int od, ad, na, ms, sd;
    float ab;
    switch (kind) {
        case GYMRL_ENV_CARTPOLE: od = 4; ad = 0; na = 2; ms = CP_MAX_STEPS; ab = 0.f; sd = 8; break;
        case GYMRL_ENV_PENDULUM: od = 3; ad = 1; na = 0; ms = PD_MAX_STEPS; ab = 2.f; sd = 6; break;
        case GYMRL_ENV_LUNARLANDER: od = 8; ad = 0; na = 4; ms = LL_MAX_STEPS; ab = 0.f; sd = LL_STATE_DOUBLES; break;
        default: GYMRL_FAIL(GYMRL_EINVAL, "unknown env kind %d", kind);
    }
    if (obs_dim) *obs_dim = od;
    if (act_dim) *act_dim = ad;
    if (n_actions) *n_actions = na;
    if (max_episode_steps) *max_episode_steps = ms;
    if (action_bound) *action_bound = ab;
    if (state_doubles) *state_doubles = sd;
    return GYMRL_OK;
}

// ---- CartPole-v1 -------------------------------------------------------------------------------
__device__ __forceinline__ void cartpole_reset_state(uint64_t seed, uint64_t id, uint32_t episode, double s[4]) {
    const u32x4 r0 = philox_draw(seed, id, episode * 8u + 0u, PHILOX_ENV_RESET);
    const u32x4 r1 = philox_draw(seed, id, episode * 8u + 1u, PHILOX_ENV_RESET);
    s[0] = -0.05 + (0.05 - -0.05) * u01_f64(r0.x, r0.y);
    s[1] = -0.05 + (0.05 - -0.05) * u01_f64(r0.z, r0.w);
    s[2] = -0.05 + (0.05 - -0.05) * u01_f64(r1.x, r1.y);
    s[3] = -0.05 + (0.05 - -0.05) * u01_f64(r1.z, r1.w);
}

__global__ void cartpole_reset_kernel(gymrl_env e, const uint8_t* __restrict__ mask, float* __restrict__ obs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n) return;
    if (mask && !mask[i]) return;
    double s[4];
    const uint32_t ep = e.episode[i];
    cartpole_reset_state(e.seed, e.first_id + i, ep, s);
    e.episode[i] = ep + 1;
    e.elapsed[i] = 0;
    e.ep_return[i] = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) e.state[(size_t)k * e.n + i] = s[k];
    if (obs) reinterpret_cast<float4*>(obs)[i] = make_float4((float)s[0], (float)s[1], (float)s[2], (float)s[3]);
}

__global__ void cartpole_step_kernel(gymrl_env e, const int32_t* __restrict__ action, float* __restrict__ obs,
                                     float* __restrict__ next_obs, float* __restrict__ reward,
                                     uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated,
                                     uint8_t* __restrict__ done_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < e.n;
    bool done = false;
    float fin_ret = 0.f;
    int fin_len = 0;
    if (valid) {
        const double gravity = 9.8, masscart = 1.0, masspole = 0.1, total_mass = masspole + masscart, length = 0.5,
                     polemass_length = masspole * length, force_mag = 10.0, tau = 0.02;
        const double theta_thr = 12 * 2 * 3.141592653589793 / 360, x_thr = 2.4;
        double x = e.state[i], x_dot = e.state[(size_t)e.n + i], theta = e.state[(size_t)2 * e.n + i],
               theta_dot = e.state[(size_t)3 * e.n + i];
        const double force = action[i] == 1 ? force_mag : -force_mag;
        const double costheta = cos(theta), sintheta = sin(theta);
        const double temp = (force + polemass_length * (theta_dot * theta_dot) * sintheta) / total_mass;
        const double thetaacc = (gravity * sintheta - costheta * temp) /
                                (length * (4.0 / 3.0 - masspole * (costheta * costheta) / total_mass));
        const double xacc = temp - polemass_length * thetaacc * costheta / total_mass;
        x = x + tau * x_dot;
        x_dot = x_dot + tau * xacc;
        theta = theta + tau * theta_dot;
        theta_dot = theta_dot + tau * thetaacc;
        const bool term = x < -x_thr || x > x_thr || theta < -theta_thr || theta > theta_thr;
        const int el = e.elapsed[i] + 1;
        const bool trunc = el >= CP_MAX_STEPS;
        const double ret = e.ep_return[i] + 1.0;
        const float4 o = make_float4((float)x, (float)x_dot, (float)theta, (float)theta_dot);
        if (next_obs) reinterpret_cast<float4*>(next_obs)[i] = o;
        reward[i] = 1.0f;
        terminated[i] = term;
        truncated[i] = trunc;
        if (done_out) done_out[i] = term || trunc;
        e.stepctr[i] += 1;
        done = term || trunc;
        if (done) {
            fin_ret = (float)ret;
            fin_len = el;
            double s[4];
            const uint32_t ep = e.episode[i];
            cartpole_reset_state(e.seed, e.first_id + i, ep, s);
            e.episode[i] = ep + 1;
            e.elapsed[i] = 0;
            e.ep_return[i] = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) e.state[(size_t)k * e.n + i] = s[k];
            reinterpret_cast<float4*>(obs)[i] = make_float4((float)s[0], (float)s[1], (float)s[2], (float)s[3]);
        } else {
            e.elapsed[i] = el;
            e.ep_return[i] = ret;
            e.state[i] = x;
            e.state[(size_t)e.n + i] = x_dot;
            e.state[(size_t)2 * e.n + i] = theta;
            e.state[(size_t)3 * e.n + i] = theta_dot;
            reinterpret_cast<float4*>(obs)[i] = o;
        }
    }
    episode_ring_push(done, fin_ret, fin_len, e.ring_ret, e.ring_len, e.ring_count);
}

// ---- Pendulum-v1 -------------------------------------------------------------------------------
__device__ __forceinline__ void pendulum_reset_state(uint64_t seed, uint64_t id, uint32_t episode, double s[2]) {
    const u32x4 r0 = philox_draw(seed, id, episode * 8u + 0u, PHILOX_ENV_RESET);
    const double pi = 3.141592653589793;
    s[0] = -pi + (pi - -pi) * u01_f64(r0.x, r0.y);
    s[1] = -1.0 + (1.0 - -1.0) * u01_f64(r0.z, r0.w);
}
__device__ __forceinline__ void pendulum_obs(const double s[2], float* o) {
    o[0] = (float)cos(s[0]);
    o[1] = (float)sin(s[0]);
    o[2] = (float)s[1];
}

__global__ void pendulum_reset_kernel(gymrl_env e, const uint8_t* __restrict__ mask, float* __restrict__ obs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n) return;
    if (mask && !mask[i]) return;
    double s[2];
    const uint32_t ep = e.episode[i];
    pendulum_reset_state(e.seed, e.first_id + i, ep, s);
    e.episode[i] = ep + 1;
    e.elapsed[i] = 0;
    e.ep_return[i] = 0.0;
    e.state[i] = s[0];
    e.state[(size_t)e.n + i] = s[1];
    if (obs) pendulum_obs(s, obs + (size_t)3 * i);
}

__global__ void pendulum_step_kernel(gymrl_env e, const float* __restrict__ action, float* __restrict__ obs,
                                     float* __restrict__ next_obs, float* __restrict__ reward,
                                     uint8_t* __restrict__ terminated, uint8_t* __restrict__ truncated,
                                     uint8_t* __restrict__ done_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < e.n;
    bool done = false;
    float fin_ret = 0.f;
    int fin_len = 0;
    if (valid) {
        const double max_speed = 8.0, dt = 0.05, g = 10.0, m = 1.0, l = 1.0, pi = 3.141592653589793;
        const double th = e.state[i], thdot = e.state[(size_t)e.n + i];
        const float uf = fminf(fmaxf(action[i], -2.0f), 2.0f);
        const double u = (double)uf;
        // angle_normalize(x) = ((x + pi) % (2 pi)) - pi with Python's sign-of-divisor modulo
        double md = fmod(th + pi, 2 * pi);
        if (md != 0.0 && md < 0.0) md += 2 * pi;
        const double an = md - pi;
        const double costs = an * an + 0.1 * (thdot * thdot) + 0.001 * (u * u);
        double newthdot = thdot + (3 * g / (2 * l) * sin(th) + 3.0 / (m * (l * l)) * u) * dt;
        newthdot = fmin(fmax(newthdot, -max_speed), max_speed);
        const double newth = th + newthdot * dt;
        double s[2] = {newth, newthdot};
        float o[3];
        pendulum_obs(s, o);
        if (next_obs) { next_obs[(size_t)3 * i] = o[0]; next_obs[(size_t)3 * i + 1] = o[1]; next_obs[(size_t)3 * i + 2] = o[2]; }
        const int el = e.elapsed[i] + 1;
        const bool trunc = el >= PD_MAX_STEPS;
        const double ret = e.ep_return[i] + (-costs);
        reward[i] = (float)(-costs);
        terminated[i] = 0;
        truncated[i] = trunc;
        if (done_out) done_out[i] = trunc;
        e.stepctr[i] += 1;
        done = trunc;
        if (done) {
            fin_ret = (float)ret;
            fin_len = el;
            const uint32_t ep = e.episode[i];
            pendulum_reset_state(e.seed, e.first_id + i, ep, s);
            e.episode[i] = ep + 1;
            e.elapsed[i] = 0;
            e.ep_return[i] = 0.0;
            pendulum_obs(s, o);
        } else {
            e.elapsed[i] = el;
            e.ep_return[i] = ret;
        }
        e.state[i] = s[0];
        e.state[(size_t)e.n + i] = s[1];
        obs[(size_t)3 * i] = o[0]; obs[(size_t)3 * i + 1] = o[1]; obs[(size_t)3 * i + 2] = o[2];
    }
    episode_ring_push(done, fin_ret, fin_len, e.ring_ret, e.ring_len, e.ring_count);
}

// ---- classic get/set state: [N][S + 4] = physics state, elapsed, episode, stepctr, ep_return --------
__global__ void classic_get_state_kernel(gymrl_env e, int S, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n) return;
    double* o = out + (size_t)i * (S + 4);
    for (int k = 0; k < S; ++k) o[k] = e.state[(size_t)k * e.n + i];
    o[S] = e.elapsed[i]; o[S + 1] = e.episode[i]; o[S + 2] = e.stepctr[i]; o[S + 3] = e.ep_return[i];
}
__global__ void classic_set_state_kernel(gymrl_env e, int S, const double* __restrict__ in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e.n) return;
    const double* o = in + (size_t)i * (S + 4);
    for (int k = 0; k < S; ++k) e.state[(size_t)k * e.n + i] = o[k];
    e.elapsed[i] = (int32_t)o[S]; e.episode[i] = (uint32_t)o[S + 1]; e.stepctr[i] = (uint32_t)o[S + 2]; e.ep_return[i] = o[S + 3];
}

// ---- handle lifecycle ---------------------------------------------------------------------------
static int classic_state_dim(int kind) { return kind == GYMRL_ENV_CARTPOLE ? 4 : 2; }

extern "C" int gymrl_env_create(gymrl_env** out, int kind, int n_envs, uint64_t seed, uint64_t first_env_id) {
    GYMRL_REQUIRE(out != nullptr, "out is NULL");
    GYMRL_REQUIRE(kind >= 0 && kind <= 2, "unknown env kind %d", kind);
    GYMRL_REQUIRE(n_envs > 0, "n_envs must be positive (got %d)", n_envs);
    int dev = -1;
    GYMRL_CUDA(cudaGetDevice(&dev));
    gymrl_env* e = new gymrl_env();
    memset(e, 0, sizeof(*e));
    e->kind = kind; e->n = n_envs; e->seed = seed; e->first_id = first_env_id; e->device = dev;
    const size_t n = (size_t)n_envs;
#define ENV_ALLOC(ptr, bytes)                                                    \
    do {                                                                         \
        cudaError_t _e = cudaMalloc((void**)&(ptr), (bytes));                    \
        if (_e != cudaSuccess) {                                                 \
            gymrl_env_destroy(e);                                                \
            GYMRL_FAIL(GYMRL_ENOMEM, "cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(_e)); \
        }                                                                        \
        cudaMemset((ptr), 0, (bytes));                                           \
    } while (0)
    ENV_ALLOC(e->elapsed, n * sizeof(int32_t));
    ENV_ALLOC(e->episode, n * sizeof(uint32_t));
    ENV_ALLOC(e->stepctr, n * sizeof(uint32_t));
    ENV_ALLOC(e->ep_return, n * sizeof(double));
    ENV_ALLOC(e->ring_ret, GYMRL_EP_RING * sizeof(float));
    ENV_ALLOC(e->ring_len, GYMRL_EP_RING * sizeof(int32_t));
    ENV_ALLOC(e->ring_count, 2 * sizeof(unsigned long long));   // [0] finished episodes, [1] LunarLander dropped-manifold events
    if (kind == GYMRL_ENV_LUNARLANDER) {
        int rc = lunar_alloc(e);
        if (rc != GYMRL_OK) { gymrl_env_destroy(e); return rc; }
    } else {
        ENV_ALLOC(e->state, n * classic_state_dim(kind) * sizeof(double));
    }
#undef ENV_ALLOC
    *out = e;
    return GYMRL_OK;
}

extern "C" int gymrl_env_destroy(gymrl_env* e) {
    if (!e) return GYMRL_OK;
    cudaFree(e->state); cudaFree(e->elapsed); cudaFree(e->episode); cudaFree(e->stepctr); cudaFree(e->ep_return);
    cudaFree(e->ring_ret); cudaFree(e->ring_len); cudaFree(e->ring_count);
    lunar_free(e);
    delete e;
    return GYMRL_OK;
}

extern "C" int gymrl_env_reset(gymrl_env* e, const uint8_t* d_mask, float* d_obs, void* stream) {
    GYMRL_REQUIRE(e != nullptr, "env is NULL");
    cudaStream_t s = as_stream(stream);
    const int threads = 128, blocks = ceil_div(e->n, threads);
    switch (e->kind) {
        case GYMRL_ENV_CARTPOLE: cartpole_reset_kernel<<<blocks, threads, 0, s>>>(*e, d_mask, d_obs); break;
        case GYMRL_ENV_PENDULUM: pendulum_reset_kernel<<<blocks, threads, 0, s>>>(*e, d_mask, d_obs); break;
        default: return lunar_reset(e, d_mask, d_obs, s);
    }
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("env_reset");
    return GYMRL_OK;
}

extern "C" int gymrl_env_step(gymrl_env* e, const void* d_actions, float* d_obs, float* d_next_obs, float* d_reward,
                              uint8_t* d_terminated, uint8_t* d_truncated, uint8_t* d_done, void* stream) {
    GYMRL_REQUIRE(e != nullptr, "env is NULL");
    GYMRL_REQUIRE(d_actions && d_obs && d_reward && d_terminated && d_truncated, "NULL output/input pointer");
    cudaStream_t s = as_stream(stream);
    const int threads = 128, blocks = ceil_div(e->n, threads);
    switch (e->kind) {
        case GYMRL_ENV_CARTPOLE:
            cartpole_step_kernel<<<blocks, threads, 0, s>>>(*e, (const int32_t*)d_actions, d_obs, d_next_obs, d_reward,
                                                           d_terminated, d_truncated, d_done);
            break;
        case GYMRL_ENV_PENDULUM:
            pendulum_step_kernel<<<blocks, threads, 0, s>>>(*e, (const float*)d_actions, d_obs, d_next_obs, d_reward,
                                                           d_terminated, d_truncated, d_done);
            break;
        default:
            return lunar_step(e, (const int32_t*)d_actions, d_obs, d_next_obs, d_reward, d_terminated, d_truncated, d_done, s);
    }
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("env_step");
    return GYMRL_OK;
}

extern "C" int gymrl_env_get_state(gymrl_env* e, double* d_state, void* stream) {
    GYMRL_REQUIRE(e && d_state, "NULL argument");
    cudaStream_t s = as_stream(stream);
    if (e->kind == GYMRL_ENV_LUNARLANDER) return lunar_get_state(e, d_state, s);
    classic_get_state_kernel<<<ceil_div(e->n, 128), 128, 0, s>>>(*e, classic_state_dim(e->kind), d_state);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("env_get_state");
    return GYMRL_OK;
}
extern "C" int gymrl_env_set_state(gymrl_env* e, const double* d_state, void* stream) {
    GYMRL_REQUIRE(e && d_state, "NULL argument");
    cudaStream_t s = as_stream(stream);
    if (e->kind == GYMRL_ENV_LUNARLANDER) return lunar_set_state(e, d_state, s);
    classic_set_state_kernel<<<ceil_div(e->n, 128), 128, 0, s>>>(*e, classic_state_dim(e->kind), d_state);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("env_set_state");
    return GYMRL_OK;
}

extern "C" int gymrl_env_set_profile(gymrl_env* e, long long* d_prof) {
    GYMRL_REQUIRE(e != nullptr, "env is NULL");
    GYMRL_REQUIRE(e->kind == GYMRL_ENV_LUNARLANDER || d_prof == nullptr, "profile counters exist for LunarLander only");
    e->prof = d_prof;
    return GYMRL_OK;
}

extern "C" int gymrl_env_set_solver(gymrl_env* e, int variant) {
    GYMRL_REQUIRE(e != nullptr, "env is NULL");
    GYMRL_REQUIRE(variant == 0 || variant == 2 || variant == 3, "unknown solver variant %d (0, 2 or 3)", variant);
    GYMRL_REQUIRE(e->kind == GYMRL_ENV_LUNARLANDER || variant == 0, "solver variants exist for LunarLander only");
    e->solver = variant;
    return GYMRL_OK;
}
extern "C" int gymrl_env_get_solver(gymrl_env* e, int* variant) {
    GYMRL_REQUIRE(e != nullptr && variant != nullptr, "NULL argument");
    *variant = e->solver;
    return GYMRL_OK;
}

extern "C" int gymrl_env_overflow_count(gymrl_env* e, uint64_t* count, void* stream) {
    GYMRL_REQUIRE(e != nullptr && count != nullptr, "NULL argument");
    unsigned long long c = 0;
    cudaStream_t s = as_stream(stream);
    GYMRL_CUDA(cudaMemcpyAsync(&c, e->ring_count + 1, sizeof(c), cudaMemcpyDeviceToHost, s));
    GYMRL_CUDA(cudaStreamSynchronize(s));
    *count = c;
    return GYMRL_OK;
}

extern "C" int gymrl_env_episode_stats(gymrl_env* e, int last_k, double* mean_return, double* mean_length,
                                       uint64_t* total_episodes, void* stream) {
    GYMRL_REQUIRE(e != nullptr, "env is NULL");
    GYMRL_REQUIRE(last_k > 0 && last_k <= GYMRL_EP_RING, "last_k must be in [1, %d]", GYMRL_EP_RING);
    cudaStream_t s = as_stream(stream);
    unsigned long long count = 0;
    std::vector<float> ret(GYMRL_EP_RING);
    std::vector<int32_t> len(GYMRL_EP_RING);
    GYMRL_CUDA(cudaMemcpyAsync(&count, e->ring_count, sizeof(count), cudaMemcpyDeviceToHost, s));
    GYMRL_CUDA(cudaMemcpyAsync(ret.data(), e->ring_ret, GYMRL_EP_RING * sizeof(float), cudaMemcpyDeviceToHost, s));
    GYMRL_CUDA(cudaMemcpyAsync(len.data(), e->ring_len, GYMRL_EP_RING * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    GYMRL_CUDA(cudaStreamSynchronize(s));
    const unsigned long long k = count < (unsigned long long)last_k ? count : (unsigned long long)last_k;
    double sr = 0.0, sl = 0.0;
    for (unsigned long long j = 0; j < k; ++j) {
        const unsigned long long slot = (count - 1 - j) % GYMRL_EP_RING;
        sr += ret[slot];
        sl += len[slot];
    }
    if (mean_return) *mean_return = k ? sr / (double)k : 0.0;
    if (mean_length) *mean_length = k ? sl / (double)k : 0.0;
    if (total_episodes) *total_episodes = count;
    return GYMRL_OK;
}
