// replay.cu — device-resident replay memory: uniform ring, n-step fold, prioritized
// sum-tree with TD-error write-back.
//
//  a10  ReplayBuffer.push/sample (algorithms/dqn_cartpole.py:68-88, same class in sac/td3/ddpg):
//       deque(maxlen=capacity) + random.sample WITHOUT replacement  ->  SoA ring in HBM, device cursor,
//       indices = first B images of a keyed random bijection on [0, size) (cycle-walking), row gathers.
//  a12  PrioritizedNStepBuffer.store_transition/_get_n_step_transition (rainbow_dqn_cartpole.py:179-218):
//       per-env window of the last n transitions (never cleared at episode ends), folded back to
//       front; emitted straight into the ring.
//  a11  SumTree (rainbow_dqn_cartpole.py:116-152) / PrioritizedNStepBuffer.sample/update_priorities (:220-261):
//       the SAME binary-heap layout as the reference (2*cap-1 float64 nodes, leaf i at cap-1+i, "go left iff
//       v <= tree[left]"), so non-power-of-two capacities reproduce the reference's rotated prefix order
//       (SURVEY q4).  One thread per sample walks root->leaf (the tree is L2 resident: 2^21 leaves = 32 MB);
//       updates walk leaf->root with float64 atomics; duplicates inside a batch resolve last-writer-wins in
//       batch order like the reference's Python loop (q6); priority_max is a true max over the leaves (q5).
// These are latency-bound pointer walks over an L2-resident tree, not bandwidth-bound streams; the
// algorithmic bytes per sample are in DESIGN.md.
#include "common.cuh"

void gymrl_count_launch(int n = 1);

// ---- keyed bijection on [0, n) (same construction as gymrl_random_permutation) ----------------------
__device__ __forceinline__ uint32_t keyed_bijection(uint32_t i, uint32_t n, uint64_t seed, uint32_t draw, uint32_t stream) {
    int bits = 1;
    while ((1ull << bits) < (unsigned long long)n) ++bits;
    const u32x4 a = philox_draw(seed, 0, draw, stream);
    const u32x4 b = philox_draw(seed, 1, draw, stream);
    const uint32_t mulk[4] = {a.x | 1u, a.y | 1u, a.z | 1u, a.w | 1u};
    const uint32_t addk[4] = {b.x, b.y, b.z, b.w};
    const uint32_t mask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
    const int sh = bits > 1 ? bits / 2 : 1;
    uint32_t x = i;
    do {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            x = (x * mulk[r]) & mask;
            x ^= x >> (sh + (r & 1));
            x = (x + addk[r]) & mask;
        }
    } while (x >= n);
    return x;
}

// ring state: int32[2] = {cursor, size}
__global__ void replay_sample_kernel(int32_t* __restrict__ idx, int B, const int32_t* __restrict__ ring_state, uint64_t seed,
                                     uint32_t draw, const uint32_t* __restrict__ draw_base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (draw_base) draw += *draw_base;
    const uint32_t size = (uint32_t)ring_state[1];
    // random.sample(buffer, min(B, len)): without replacement; rows beyond `size` repeat (caller guards len >= B)
    idx[i] = size == 0 ? 0 : (int32_t)keyed_bijection((uint32_t)i % size, size, seed, draw, PHILOX_REPLAY);
}

extern "C" int gymrl_replay_sample_indices(int32_t* d_idx, int batch, const int32_t* d_ring_state, uint64_t seed, uint32_t draw,
                                           const uint32_t* d_draw_base, void* stream) {
    GYMRL_REQUIRE(d_idx && d_ring_state && batch > 0, "bad arguments");
    replay_sample_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(d_idx, batch, d_ring_state, seed, draw, d_draw_base);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("replay_sample_indices");
    return GYMRL_OK;
}

// Store n rows of width w (floats or int32, both 4 B) at ring positions (cursor + i) % capacity.
__global__ void replay_store_kernel(float* __restrict__ dst, const float* __restrict__ src, int n, int w, int capacity,
                                    const int32_t* __restrict__ ring_state) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * w) return;
    const int i = (int)(t / w), c = (int)(t % w);
    const int pos = (int)(((long long)ring_state[0] + i) % capacity);
    dst[(size_t)pos * w + c] = src[t];
}
__global__ void replay_store_u8_kernel(float* __restrict__ dst, const uint8_t* __restrict__ src, int n, int capacity,
                                       const int32_t* __restrict__ ring_state) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pos = (int)(((long long)ring_state[0] + i) % capacity);
    dst[pos] = (float)src[i];
}
__global__ void replay_advance_kernel(int32_t* ring_state, int n, int capacity) {
    ring_state[0] = (int32_t)(((long long)ring_state[0] + n) % capacity);
    ring_state[1] = min(capacity, ring_state[1] + n);
}

extern "C" int gymrl_replay_store(void* d_dst, const void* d_src, int n, int width, int src_is_u8, int capacity,
                                  const int32_t* d_ring_state, void* stream) {
    GYMRL_REQUIRE(d_dst && d_src && d_ring_state && n > 0 && width > 0 && capacity >= n, "bad arguments");
    cudaStream_t s = as_stream(stream);
    if (src_is_u8) {
        GYMRL_REQUIRE(width == 1, "u8 source must have width 1");
        replay_store_u8_kernel<<<ceil_div(n, 256), 256, 0, s>>>((float*)d_dst, (const uint8_t*)d_src, n, capacity, d_ring_state);
    } else {
        const long long tot = (long long)n * width;
        replay_store_kernel<<<(unsigned)ceil_div_ll(tot, 256), 256, 0, s>>>((float*)d_dst, (const float*)d_src, n, width, capacity,
                                                                           d_ring_state);
    }
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("replay_store");
    return GYMRL_OK;
}

extern "C" int gymrl_replay_advance(int32_t* d_ring_state, int n, int capacity, void* stream) {
    GYMRL_REQUIRE(d_ring_state && n > 0 && capacity > 0, "bad arguments");
    replay_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(d_ring_state, n, capacity);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("replay_advance");
    return GYMRL_OK;
}

// dst[i] = [ a[ia(i)][0:wa] , b[ib(i)][0:wb] ]  (4-byte elements) — row gather and torch.cat([state, action], 1)
// (Critic.forward, algorithms/sac_pendulum.py:112) in one pass.  b nullable.
__global__ void gather_concat_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ a, int wa, int lda,
                                     const int32_t* __restrict__ ia, const float* __restrict__ b, int wb, int ldb,
                                     const int32_t* __restrict__ ib, int n) {
    const int w = wa + wb;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * w) return;
    const int i = (int)(t / w), c = (int)(t % w);
    float v;
    if (c < wa) v = a[(size_t)(ia ? ia[i] : i) * lda + c];
    else v = b[(size_t)(ib ? ib[i] : i) * ldb + (c - wa)];
    dst[(size_t)i * ldd + c] = v;
}

extern "C" int gymrl_gather_concat(void* d_dst, int ld_dst, const void* d_a, int width_a, int ld_a, const int32_t* d_idx_a,
                                   const void* d_b, int width_b, int ld_b, const int32_t* d_idx_b, int n, void* stream) {
    GYMRL_REQUIRE(d_dst && d_a && n > 0 && width_a > 0 && width_b >= 0, "bad arguments");
    GYMRL_REQUIRE(width_b == 0 || d_b, "b is NULL");
    const long long tot = (long long)n * (width_a + width_b);
    gather_concat_kernel<<<(unsigned)ceil_div_ll(tot, 256), 256, 0, as_stream(stream)>>>(
        (float*)d_dst, ld_dst, (const float*)d_a, width_a, ld_a, d_idx_a, (const float*)d_b, width_b, ld_b, d_idx_b, n);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("gather_concat");
    return GYMRL_OK;
}

// ---- n-step window -----------------------------------------------------------------------------------
// Window planes are [n][N]-major (slot-major) so a warp's 32 envs touch contiguous rows. `pushed` counts
// lockstep pushes so far; slot of the k-th oldest entry = (pushed + k) % n once the window is full.
__global__ void nstep_push_kernel(float* __restrict__ w_obs, int32_t* __restrict__ w_act, float* __restrict__ w_rew,
                                  float* __restrict__ w_nobs, uint8_t* __restrict__ w_term, uint8_t* __restrict__ w_done,
                                  const float* __restrict__ obs, const int32_t* __restrict__ act, const float* __restrict__ rew,
                                  const float* __restrict__ nobs, const uint8_t* __restrict__ term, const uint8_t* __restrict__ done,
                                  int N, int D, int n_steps, double gamma, const int32_t* __restrict__ pushed_ptr,
                                  float* __restrict__ r_obs, int32_t* __restrict__ r_act, float* __restrict__ r_rew,
                                  float* __restrict__ r_nobs, float* __restrict__ r_term, int capacity,
                                  const int32_t* __restrict__ ring_state) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const int pushed = *pushed_ptr;
    const int slot = pushed % n_steps;
    // overwrite the oldest slot with the new transition (deque(maxlen=n).append)
    for (int d = 0; d < D; ++d) {
        w_obs[((size_t)slot * N + e) * D + d] = obs[(size_t)e * D + d];
        w_nobs[((size_t)slot * N + e) * D + d] = nobs[(size_t)e * D + d];
    }
    w_act[(size_t)slot * N + e] = act[e];
    w_rew[(size_t)slot * N + e] = rew[e];
    w_term[(size_t)slot * N + e] = term[e];
    w_done[(size_t)slot * N + e] = done[e];
    if (pushed + 1 < n_steps) return;  // window not yet full: nothing emitted
    // fold back to front: R = r + gamma (1-d) R; (s', terminal) from the EARLIEST done in the window
    const int oldest = (slot + 1) % n_steps, newest = slot;
    int ns_slot = newest;
    float terminal = (float)w_term[(size_t)newest * N + e];
    double R = 0.0;
    for (int k = n_steps - 1; k >= 0; --k) {
        const int s = (oldest + k) % n_steps;
        const double r = (double)w_rew[(size_t)s * N + e];
        const int d = w_done[(size_t)s * N + e];
        R = r + gamma * (double)(1 - d) * R;
        if (d) { ns_slot = s; terminal = (float)w_term[(size_t)s * N + e]; }
    }
    const int pos = (int)(((long long)ring_state[0] + e) % capacity);
    for (int d = 0; d < D; ++d) {
        r_obs[(size_t)pos * D + d] = w_obs[((size_t)oldest * N + e) * D + d];
        r_nobs[(size_t)pos * D + d] = w_nobs[((size_t)ns_slot * N + e) * D + d];
    }
    r_act[pos] = w_act[(size_t)oldest * N + e];
    r_rew[pos] = (float)R;
    r_term[pos] = terminal;
}
__global__ void counter_inc_kernel(int32_t* c) { *c += 1; }

extern "C" int gymrl_nstep_push(float* w_obs, int32_t* w_act, float* w_rew, float* w_nobs, uint8_t* w_term, uint8_t* w_done,
                                const float* d_obs, const int32_t* d_act, const float* d_rew, const float* d_nobs,
                                const uint8_t* d_term, const uint8_t* d_done, int n_envs, int obs_dim, int n_steps, double gamma,
                                int32_t* d_pushed, float* r_obs, int32_t* r_act, float* r_rew, float* r_nobs, float* r_term,
                                int capacity, const int32_t* d_ring_state, void* stream) {
    GYMRL_REQUIRE(w_obs && w_act && w_rew && w_nobs && w_term && w_done && d_obs && d_act && d_rew && d_nobs && d_term && d_done,
                  "NULL window/input pointer");
    GYMRL_REQUIRE(d_pushed && r_obs && r_act && r_rew && r_nobs && r_term && d_ring_state, "NULL ring pointer");
    GYMRL_REQUIRE(n_envs > 0 && obs_dim > 0 && n_steps > 0 && capacity >= n_envs, "bad shape");
    cudaStream_t s = as_stream(stream);
    nstep_push_kernel<<<ceil_div(n_envs, 128), 128, 0, s>>>(w_obs, w_act, w_rew, w_nobs, w_term, w_done, d_obs, d_act, d_rew, d_nobs,
                                                          d_term, d_done, n_envs, obs_dim, n_steps, gamma, d_pushed, r_obs, r_act,
                                                          r_rew, r_nobs, r_term, capacity, d_ring_state);
    counter_inc_kernel<<<1, 1, 0, s>>>(d_pushed);
    gymrl_count_launch(2);
    GYMRL_LAUNCH_CHECK("nstep_push");
    return GYMRL_OK;
}

// ---- sum tree ----------------------------------------------------------------------------------------
__device__ __forceinline__ void sumtree_set_leaf(double* tree, int capacity, int data_index, double priority) {
    int node = data_index + capacity - 1;
    const double change = priority - tree[node];
    tree[node] = priority;
    while (node != 0) {
        node = (node - 1) / 2;
        atomicAdd(&tree[node], change);
    }
}

// pass 1: winner[leaf] = last batch position that writes it; pass 2: only winners apply (last-writer-wins)
__global__ void sumtree_mark_kernel(int32_t* __restrict__ winner, const int32_t* __restrict__ idx, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicMax(&winner[idx[i]], i);
}
__global__ void sumtree_apply_kernel(double* __restrict__ tree, int capacity, int32_t* __restrict__ winner,
                                     const int32_t* __restrict__ idx, const double* __restrict__ prio64,
                                     const float* __restrict__ td, int n, float eps, float alpha, float clip_max) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int leaf = idx[i];
    if (winner[leaf] != i) return;
    double p;
    if (prio64) p = prio64[i];
    else {
        // rainbow :259  (np.abs(td) + 0.01) ** alpha  — a float32 expression under NumPy >= 2;
        // ddqn_per :142-147  min(|td| + 1e-4, 1) ** 0.6
        float a = fabsf(td[i]) + eps;
        if (clip_max > 0.f) a = fminf(a, clip_max);
        p = (double)powf(a, alpha);
    }
    sumtree_set_leaf(tree, capacity, leaf, p);
    winner[leaf] = -1;
}

extern "C" int gymrl_sumtree_update(double* d_tree, int capacity, const int32_t* d_idx, const double* d_priority,
                                    const float* d_td_error, int n, float eps, float alpha, float clip_max, int32_t* d_winner_scratch,
                                    void* stream) {
    GYMRL_REQUIRE(d_tree && d_idx && d_winner_scratch && n > 0 && capacity > 0, "bad arguments");
    GYMRL_REQUIRE((d_priority != nullptr) != (d_td_error != nullptr), "pass exactly one of priority / td_error");
    cudaStream_t s = as_stream(stream);
    sumtree_mark_kernel<<<ceil_div(n, 256), 256, 0, s>>>(d_winner_scratch, d_idx, n);
    sumtree_apply_kernel<<<ceil_div(n, 256), 256, 0, s>>>(d_tree, capacity, d_winner_scratch, d_idx, d_priority, d_td_error, n, eps,
                                                         alpha, clip_max);
    gymrl_count_launch(2);
    GYMRL_LAUNCH_CHECK("sumtree_update");
    return GYMRL_OK;
}

// Store path: leaves [cursor, cursor+n) get priority max(leaves) (1.0 for the very first item), rainbow :201-202.
__global__ void sumtree_max_kernel(const double* __restrict__ tree, int capacity, double* __restrict__ out) {
    __shared__ double scratch[32];
    double m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += gridDim.x * blockDim.x) m = fmax(m, tree[capacity - 1 + i]);
    m = warp_max(m);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) scratch[wid] = m;
    __syncthreads();
    if (wid == 0) {
        m = lane < (blockDim.x >> 5) ? scratch[lane] : 0.0;
        m = warp_max(m);
        if (lane == 0) {  // priorities are >= 0: the float64 bit pattern orders like an unsigned integer
            atomicMax(reinterpret_cast<unsigned long long*>(out), (unsigned long long)__double_as_longlong(m));
        }
    }
}
__global__ void sumtree_store_kernel(double* __restrict__ tree, int capacity, int n, const int32_t* __restrict__ ring_state,
                                     const double* __restrict__ maxp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int pos = (int)(((long long)ring_state[0] + i) % capacity);
    const double p = (ring_state[1] == 0 && *maxp == 0.0) ? 1.0 : *maxp;
    sumtree_set_leaf(tree, capacity, pos, p);
}
__global__ void zero_double_kernel(double* p) { *p = 0.0; }

extern "C" int gymrl_sumtree_store_new(double* d_tree, int capacity, int n, const int32_t* d_ring_state, double* d_max_scratch,
                                       void* stream) {
    GYMRL_REQUIRE(d_tree && d_ring_state && d_max_scratch && n > 0 && capacity >= n, "bad arguments");
    cudaStream_t s = as_stream(stream);
    zero_double_kernel<<<1, 1, 0, s>>>(d_max_scratch);
    int blocks = ceil_div(capacity, 1024);
    if (blocks > GYMRL_NUM_SMS * 2) blocks = GYMRL_NUM_SMS * 2;
    sumtree_max_kernel<<<blocks, 256, 0, s>>>(d_tree, capacity, d_max_scratch);
    sumtree_store_kernel<<<ceil_div(n, 256), 256, 0, s>>>(d_tree, capacity, n, d_ring_state, d_max_scratch);
    gymrl_count_launch(3);
    GYMRL_LAUNCH_CHECK("sumtree_store_new");
    return GYMRL_OK;
}

// Stratified sampling (rainbow :220-256): v_i = U(seg*i, seg*(i+1)); descend; w_i = (size * p_i / total)^(-beta), then / max w.
__global__ void sumtree_sample_kernel(const double* __restrict__ tree, int capacity, int B, const double* __restrict__ uniforms,
                                      const int32_t* __restrict__ ring_state, const double* __restrict__ beta_ptr, int32_t* __restrict__ out_idx,
                                      float* __restrict__ out_w, double* __restrict__ out_prio, unsigned int* __restrict__ wmax_bits,
                                      int return_tree_index, uint64_t seed, uint32_t draw, const uint32_t* __restrict__ draw_base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (draw_base) draw += *draw_base;
    const double total = tree[0];
    const double segment = total / (double)B;
    double u;
    if (uniforms) u = uniforms[i];
    else {
        const u32x4 r = philox_draw(seed, (uint64_t)i, draw, PHILOX_REPLAY);
        u = u01_f64(r.x, r.y);
    }
    const double a = segment * (double)i, b = segment * (double)(i + 1);
    double v = a + (b - a) * u;  // np.random.uniform(a, b)
    if (return_tree_index & 2) v = u;  // flag bit 1: `uniforms` holds raw prefix values (SumTree.get_index(v))
    const int tree_capacity = 2 * capacity - 1;
    int parent = 0;
    while (true) {
        const int left = 2 * parent + 1;
        if (left >= tree_capacity) break;
        const double lv = tree[left];
        if (v <= lv) parent = left;
        else { v -= lv; parent = left + 1; }
    }
    const double priority = tree[parent];
    const int data_index = parent - capacity + 1;
    out_idx[i] = (return_tree_index & 1) ? parent : data_index;
    if (out_prio) out_prio[i] = priority;
    const double prob = priority / total;
    const float w = (float)pow((double)ring_state[1] * prob, -(*beta_ptr));
    out_w[i] = w;
    atomicMax(wmax_bits, __float_as_uint(w));  // w > 0
}
__global__ void sumtree_normalize_kernel(float* __restrict__ w, int B, unsigned int* __restrict__ wmax_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) w[i] = w[i] / __uint_as_float(*wmax_bits);
}
__global__ void zero_u32_kernel(unsigned int* p) { *p = 0u; }

extern "C" int gymrl_sumtree_sample(const double* d_tree, int capacity, int batch, const double* d_uniforms,
                                    const int32_t* d_ring_state, const double* d_beta, int32_t* d_out_idx, float* d_out_is_weight,
                                    double* d_out_priority, uint32_t* d_scratch_u32, int return_tree_index, uint64_t seed,
                                    uint32_t draw, const uint32_t* d_draw_base, void* stream) {
    GYMRL_REQUIRE(d_tree && d_ring_state && d_beta && d_out_idx && d_out_is_weight && d_scratch_u32 && batch > 0 && capacity > 0, "bad arguments");
    cudaStream_t s = as_stream(stream);
    zero_u32_kernel<<<1, 1, 0, s>>>(d_scratch_u32);
    sumtree_sample_kernel<<<ceil_div(batch, 128), 128, 0, s>>>(d_tree, capacity, batch, d_uniforms, d_ring_state, d_beta, d_out_idx,
                                                              d_out_is_weight, d_out_priority, d_scratch_u32, return_tree_index,
                                                              seed, draw, d_draw_base);
    sumtree_normalize_kernel<<<ceil_div(batch, 256), 256, 0, s>>>(d_out_is_weight, batch, d_scratch_u32);
    gymrl_count_launch(3);
    GYMRL_LAUNCH_CHECK("sumtree_sample");
    return GYMRL_OK;
}
