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
//       updates write the leaves and recompute the touched ancestors level by level (one block, no atomics on the
//       tree, bitwise reproducible); duplicates inside a batch resolve last-writer-wins in batch order like the
//       reference's Python loop (q6); priority_max is a true max over the leaves (q5).
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

// One launch for a whole lockstep's transitions: obs, next_obs [n][obs_dim], action [n][act_width] (4-byte elements), reward [n],
// done (uint8 -> float) written at ring rows (cursor + i) % capacity, and the {cursor, size} state advanced by the block that
// finishes last (d_done_ctr: one zero-initialised word, left zero).  Replaces 5 x gymrl_replay_store + gymrl_replay_advance.
__global__ void replay_store_all_kernel(float* __restrict__ r_obs, float* __restrict__ r_nobs, float* __restrict__ r_act,
                                        float* __restrict__ r_rew, float* __restrict__ r_done, const float* __restrict__ obs,
                                        const float* __restrict__ nobs, const float* __restrict__ act, const float* __restrict__ rew,
                                        const uint8_t* __restrict__ done, int n, int D, int AW, int capacity,
                                        int32_t* __restrict__ ring_state, unsigned int* __restrict__ done_ctr) {
    const int cursor = ring_state[0];
    const int W = 2 * D + AW + 2;                 // words per transition
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < (long long)n * W) {
        const int i = (int)(t / W), c = (int)(t % W);
        const size_t pos = (size_t)(((long long)cursor + i) % capacity);
        if (c < D) r_obs[pos * D + c] = obs[(size_t)i * D + c];
        else if (c < 2 * D) r_nobs[pos * D + (c - D)] = nobs[(size_t)i * D + (c - D)];
        else if (c < 2 * D + AW) r_act[pos * AW + (c - 2 * D)] = act[(size_t)i * AW + (c - 2 * D)];
        else if (c == 2 * D + AW) r_rew[pos] = rew[i];
        else r_done[pos] = (float)done[i];
    }
    __shared__ bool s_last;
    __syncthreads();                              // every thread of the block has read the cursor
    if (threadIdx.x == 0) { __threadfence(); s_last = atomicAdd(done_ctr, 1u) == gridDim.x - 1; }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        ring_state[0] = (int32_t)(((long long)cursor + n) % capacity);
        ring_state[1] = min(capacity, ring_state[1] + n);
        *done_ctr = 0u;
    }
}
extern "C" int gymrl_replay_store_all(float* r_obs, float* r_next_obs, void* r_action, float* r_reward, float* r_done,
                                      const float* d_obs, const float* d_next_obs, const void* d_action, const float* d_reward,
                                      const uint8_t* d_done, int n, int obs_dim, int act_width, int capacity, int32_t* d_ring_state,
                                      uint32_t* d_done_ctr, void* stream) {
    GYMRL_REQUIRE(r_obs && r_next_obs && r_action && r_reward && r_done && d_obs && d_next_obs && d_action && d_reward && d_done,
                  "NULL ring / input pointer");
    GYMRL_REQUIRE(d_ring_state && d_done_ctr && n > 0 && obs_dim > 0 && act_width > 0 && capacity >= n, "bad arguments");
    const long long tot = (long long)n * (2 * obs_dim + act_width + 2);
    replay_store_all_kernel<<<(unsigned)ceil_div_ll(tot, 256), 256, 0, as_stream(stream)>>>(
        r_obs, r_next_obs, (float*)r_action, r_reward, r_done, d_obs, d_next_obs, (const float*)d_action, d_reward, d_done, n, obs_dim,
        act_width, capacity, d_ring_state, d_done_ctr);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("replay_store_all");
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
                                  const float* __restrict__ nobs, const uint8_t* __restrict__ term, const uint8_t* __restrict__ trunc,
                                  const uint8_t* __restrict__ done, int N, int D, int n_steps, double gamma, int32_t* __restrict__ pushed_ptr,
                                  float* __restrict__ r_obs, int32_t* __restrict__ r_act, float* __restrict__ r_rew,
                                  float* __restrict__ r_nobs, float* __restrict__ r_term, int capacity,
                                  const int32_t* __restrict__ ring_state) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int pushed = *pushed_ptr;
    // the push counter is advanced by the block that finishes last (pushed_ptr[1]: blocks done, zero between calls)
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(pushed_ptr + 1, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) { pushed_ptr[0] = pushed + 1; pushed_ptr[1] = 0; }
    if (e >= N) return;
    const int slot = pushed % n_steps;
    // overwrite the oldest slot with the new transition (deque(maxlen=n).append)
    for (int d = 0; d < D; ++d) {
        w_obs[((size_t)slot * N + e) * D + d] = obs[(size_t)e * D + d];
        w_nobs[((size_t)slot * N + e) * D + d] = nobs[(size_t)e * D + d];
    }
    w_act[(size_t)slot * N + e] = act[e];
    w_rew[(size_t)slot * N + e] = rew[e];
    // `terminal` of the reference's train loop (rainbow :376): terminated and not cut by the time limit
    w_term[(size_t)slot * N + e] = trunc ? (uint8_t)(term[e] & (trunc[e] ? 0 : 1)) : term[e];
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

extern "C" int gymrl_nstep_push(float* w_obs, int32_t* w_act, float* w_rew, float* w_nobs, uint8_t* w_term, uint8_t* w_done,
                                const float* d_obs, const int32_t* d_act, const float* d_rew, const float* d_nobs,
                                const uint8_t* d_term, const uint8_t* d_trunc, const uint8_t* d_done, int n_envs, int obs_dim, int n_steps,
                                double gamma,
                                int32_t* d_pushed, float* r_obs, int32_t* r_act, float* r_rew, float* r_nobs, float* r_term,
                                int capacity, const int32_t* d_ring_state, void* stream) {
    GYMRL_REQUIRE(w_obs && w_act && w_rew && w_nobs && w_term && w_done && d_obs && d_act && d_rew && d_nobs && d_term && d_done,
                  "NULL window/input pointer");
    GYMRL_REQUIRE(d_pushed && r_obs && r_act && r_rew && r_nobs && r_term && d_ring_state, "NULL ring pointer");
    GYMRL_REQUIRE(n_envs > 0 && obs_dim > 0 && n_steps > 0 && capacity >= n_envs, "bad shape");
    cudaStream_t s = as_stream(stream);
    nstep_push_kernel<<<ceil_div(n_envs, 128), 128, 0, s>>>(w_obs, w_act, w_rew, w_nobs, w_term, w_done, d_obs, d_act, d_rew, d_nobs,
                                                          d_term, d_trunc, d_done, n_envs, obs_dim, n_steps, gamma, d_pushed, r_obs,
                                                          r_act, r_rew, r_nobs, r_term, capacity, d_ring_state);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("nstep_push");
    return GYMRL_OK;
}

// ---- sum tree ----------------------------------------------------------------------------------------
// Batched leaf writes.  Round 1 walked leaf -> root with a float64 atomicAdd(change) per level: 8192 leaves x 21 levels of atomics
// that all meet at the root (measured 53-55 us per batch, and the sums depended on the arrival order).  Now every internal node is
// always exactly fl(left + right) — a pure function of the leaves below it — which makes a batch embarrassingly parallel:
//   mark    one thread per item: winner[leaf] = last batch position that writes it (duplicates: last writer wins, like the
//           reference's Python loop) and expected[s] = number of distinct touched leaves under sub-root s (the node's ancestor
//           at depth T = min(13, deepest complete internal level));
//   set     one warp per item: the winner writes its leaf and arrives at its sub-root; the warp whose arrival completes a
//           sub-root recomputes that whole subtree (<= 256 leaves) bottom-up, a warp-synchronous level at a time; the block that
//           finishes last rebuilds the <= 8191 nodes above level T in shared memory from row T.
// No atomics on the tree, no grid-wide barrier, bitwise reproducible.  Against the reference's incremental
// `tree[parent] += change` only the summation order differs (tests: leaves exact, sums rtol 1e-12).
// Scratch (d_winner_scratch): int32[capacity + GYMRL_SUMTREE_SCRATCH_EXTRA] = winner[capacity] (-1) | expected[8192] | arrive[8192]
// | blocks_done (all 0); every call leaves it in that state.
#define ST_TOP_MAX 13
#define ST_ROWS (1 << ST_TOP_MAX)
#define ST_EXTRA (2 * ST_ROWS + 1)
#define ST_SROWS 1024
__host__ __device__ __forceinline__ int heap_depth(int node) {
    int d = 0;
    for (unsigned v = (unsigned)node + 1u; v > 1u; v >>= 1) ++d;
    return d;
}
// deepest level d <= ST_TOP_MAX that is complete and internal (2^(d+1) <= capacity); -1 if the tree is a single leaf
__host__ __device__ __forceinline__ int st_top_level(int capacity) {
    int d = -1;
    while (d + 1 <= ST_TOP_MAX && (2ll << (d + 1)) <= (long long)capacity) ++d;
    return d;
}
__device__ __forceinline__ int st_sub_root(int node, int T) {
    int d = heap_depth(node);
    while (d > T) { node = (node - 1) >> 1; --d; }
    return node;
}
__device__ __forceinline__ int st_leaf_of(const int32_t* idx, const int32_t* ring_state, int i, int capacity) {
    return idx ? idx[i] : (int)(((long long)ring_state[0] + i) % capacity);
}

__global__ void sumtree_mark_kernel(int32_t* __restrict__ scratch, const int32_t* __restrict__ idx, const int32_t* __restrict__ ring_state,
                                    int n, int capacity) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int leaf = st_leaf_of(idx, ring_state, i, capacity);
    const int old = atomicMax(&scratch[leaf], i);
    const int T = st_top_level(capacity);
    if (old == -1 && T >= 0) atomicAdd(&scratch[capacity + st_sub_root(leaf + capacity - 1, T) - ((1 << T) - 1)], 1);
}

// mode 0: leaf <- priority (prio64[i], or from the TD error);  mode 1: leaf <- the store rule (max of the leaves, 1.0 when empty)
__global__ void __launch_bounds__(256) sumtree_set_kernel(double* __restrict__ tree, int capacity, int n, int mode, int32_t* __restrict__ scratch,
                                                          const int32_t* __restrict__ idx, const double* __restrict__ prio64,
                                                          const float* __restrict__ td, float eps, float alpha, float clip_max,
                                                          const int32_t* __restrict__ ring_state, double* __restrict__ maxp) {
    __shared__ double s_row[ST_SROWS];    // (a fold of) row T of the tree — last block only
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int T = st_top_level(capacity);
    const int maxd = heap_depth(2 * capacity - 2);
    int32_t* winner = scratch;
    int32_t* expected = scratch + capacity;
    int32_t* arrive = expected + ST_ROWS;
    int32_t* blocks_done = arrive + ST_ROWS;
    if (i < n) {
        int sub = -1;   // sub-root this warp recomputes
        if (lane == 0) {
            const int leaf = st_leaf_of(idx, ring_state, i, capacity);
            if (winner[leaf] == i) {
                double p;
                if (mode == 1) {
                    const double m = *maxp;
                    p = (ring_state[1] == 0 && m == 0.0) ? 1.0 : m;    // rainbow :201: 1.0 for the very first item, else max(leaves)
                } else if (prio64) p = prio64[i];
                else {
                    // rainbow :259  (np.abs(td) + 0.01) ** alpha  — a float32 expression under NumPy >= 2;
                    // ddqn_per :142-147  min(|td| + 1e-4, 1) ** 0.6
                    float a = fabsf(td[i]) + eps;
                    if (clip_max > 0.f) a = fminf(a, clip_max);
                    p = (double)powf(a, alpha);
                }
                const int node = leaf + capacity - 1;
                tree[node] = p;
                winner[leaf] = -1;
                if (T >= 0) {
                    const int root = st_sub_root(node, T);
                    const int r = root - ((1 << T) - 1);
                    __threadfence();
                    if (atomicAdd(&arrive[r], 1) + 1 == expected[r]) {   // every touched leaf under this sub-root is written
                        arrive[r] = 0; expected[r] = 0;
                        sub = root;
                    }
                }
            }
        }
        sub = __shfl_sync(0xffffffffu, sub, 0);
        if (sub >= 0) {
            __threadfence();
            const int H = maxd - T;                       // relative depth of the deepest leaves under a sub-root
            if (H >= 1 && H <= 8) {
                // fast path (<= 256 leaves): the row above the deepest leaves in registers (1, 2 or 4 nodes per lane, one L2 round
                // trip), then pairwise in-lane and by shuffles — left + right at every node, like the generic loop below
                const int R = 1 << (H - 1);
                const int P = R >= 32 ? R / 32 : 1;
                const long long root1 = (long long)sub + 1;
                const long long first = (root1 << (H - 1)) - 1;
                double v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = 0.0;
                    const int x = lane * P + j;
                    if (j < P && x < R) {
                        const long long nd = first + x;
                        if (nd < capacity - 1) { v[j] = __ldcg(&tree[2 * nd + 1]) + __ldcg(&tree[2 * nd + 2]); tree[nd] = v[j]; }
                        else v[j] = __ldcg(&tree[nd]);    // a leaf one level up (capacities that are not a power of two)
                    }
                }
                double cur = v[0];
                int k = H - 1;
                if (P == 4) {
                    const long long f2 = (root1 << (H - 2)) - 1, f3 = (root1 << (H - 3)) - 1;
                    const double a0 = v[0] + v[1], a1 = v[2] + v[3];
                    tree[f2 + 2 * lane] = a0; tree[f2 + 2 * lane + 1] = a1;
                    cur = a0 + a1;
                    tree[f3 + lane] = cur;
                    k = H - 3;
                } else if (P == 2) {
                    cur = v[0] + v[1];
                    tree[(root1 << (H - 2)) - 1 + lane] = cur;
                    k = H - 2;
                }
                const int width = 1 << k;                 // lanes [0, width) hold the nodes of relative depth k
                for (int st = 1; st < width; st <<= 1) {
                    const double right = __shfl_down_sync(0xffffffffu, cur, st);
                    --k;
                    if (lane < width && (lane % (2 * st)) == 0) {
                        cur = cur + right;
                        tree[(root1 << k) - 1 + lane / (2 * st)] = cur;
                    }
                }
            } else {
                for (int k = H - 1; k >= 0; --k) {
                    const long long first = (((long long)sub + 1) << k) - 1;
                    for (int x = lane; x < (1 << k); x += 32) {
                        const long long nd = first + x;
                        if (nd < capacity - 1) tree[nd] = __ldcg(&tree[2 * nd + 1]) + __ldcg(&tree[2 * nd + 2]);   // internal nodes only
                    }
                    __syncwarp();
                }
            }
        }
    }
    // ---- the block that finishes last rebuilds the levels above T from row T ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); s_last = atomicAdd(blocks_done, 1) == (int)gridDim.x - 1; }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        *blocks_done = 0;
        if (mode == 1) *maxp = 0.0;   // consumed: sumtree_max_kernel accumulates into it with atomicMax
    }
    if (T < 1) return;
    // Rows wider than the 1024-entry shared buffer are first folded three levels in registers: a thread owns eight adjacent
    // nodes of row T and writes their parents / grandparents / great-grandparent (T is 11..13 there).
    int top = T;
    if ((1 << T) > ST_SROWS) {
        const int rows = 1 << T, base = rows - 1;
        // all of a thread's row-T nodes are requested before the first parent is written (the stores would otherwise fence the
        // next group's loads behind them: four dependent L2 round trips instead of one)
        constexpr int G = ST_ROWS / 8 / 256;                // groups of 8 nodes per thread (4 at T = 13)
        double v[G][8];
#pragma unroll
        for (int gI = 0; gI < G; ++gI) {
            const int x = threadIdx.x + gI * 256;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[gI][j] = x < rows / 8 ? __ldcg(&tree[base + 8 * x + j]) : 0.0;
        }
#pragma unroll
        for (int gI = 0; gI < G; ++gI) {
            const int x = threadIdx.x + gI * 256;
            if (x < rows / 8) {
                double a[4], b[2];
#pragma unroll
                for (int j = 0; j < 4; ++j) { a[j] = v[gI][2 * j] + v[gI][2 * j + 1]; tree[(rows / 2 - 1) + 4 * x + j] = a[j]; }
#pragma unroll
                for (int j = 0; j < 2; ++j) { b[j] = a[2 * j] + a[2 * j + 1]; tree[(rows / 4 - 1) + 2 * x + j] = b[j]; }
                const double c = b[0] + b[1];
                tree[(rows / 8 - 1) + x] = c;
                s_row[x] = c;
            }
        }
        top = T - 3;
    } else {
        const int rows = 1 << T, base = rows - 1;
        for (int x = threadIdx.x; x < rows; x += blockDim.x) s_row[x] = __ldcg(&tree[base + x]);
    }
    __syncthreads();
    for (int d = top - 1; d >= 0; --d) {
        const int cnt = 1 << d;
        double v[ST_SROWS / 2 / 256];
#pragma unroll
        for (int j = 0; j < ST_SROWS / 2 / 256; ++j) {
            const int x = threadIdx.x + j * 256;
            if (x < cnt) v[j] = s_row[2 * x] + s_row[2 * x + 1];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ST_SROWS / 2 / 256; ++j) {
            const int x = threadIdx.x + j * 256;
            if (x < cnt) { s_row[x] = v[j]; tree[cnt - 1 + x] = v[j]; }
        }
        __syncthreads();
    }
}

static int sumtree_launch_set(double* d_tree, int capacity, int n, int mode, int32_t* d_scratch, const int32_t* d_idx,
                              const double* d_priority, const float* d_td_error, float eps, float alpha, float clip_max,
                              const int32_t* d_ring_state, double* d_maxp, cudaStream_t s) {
    sumtree_mark_kernel<<<ceil_div(n, 256), 256, 0, s>>>(d_scratch, d_idx, d_ring_state, n, capacity);
    sumtree_set_kernel<<<ceil_div(n, 8), 256, 0, s>>>(d_tree, capacity, n, mode, d_scratch, d_idx, d_priority, d_td_error, eps, alpha,
                                                        clip_max, d_ring_state, d_maxp);
    gymrl_count_launch(2);
    return GYMRL_OK;
}

extern "C" int gymrl_sumtree_scratch_ints(int capacity) { return capacity + ST_EXTRA; }

extern "C" int gymrl_sumtree_update(double* d_tree, int capacity, const int32_t* d_idx, const double* d_priority,
                                    const float* d_td_error, int n, float eps, float alpha, float clip_max, int32_t* d_winner_scratch,
                                    void* stream) {
    GYMRL_REQUIRE(d_tree && d_idx && d_winner_scratch && n > 0 && capacity > 0, "bad arguments");
    GYMRL_REQUIRE((d_priority != nullptr) != (d_td_error != nullptr), "pass exactly one of priority / td_error");
    int rc = sumtree_launch_set(d_tree, capacity, n, 0, d_winner_scratch, d_idx, d_priority, d_td_error, eps, alpha, clip_max, nullptr, nullptr,
                                as_stream(stream));
    if (rc != GYMRL_OK) return rc;
    GYMRL_LAUNCH_CHECK("sumtree_update");
    return GYMRL_OK;
}

// Store path: leaves [cursor, cursor+n) get priority max(leaves) (1.0 for the very first item), rainbow :201-202.
// d_max_scratch must be zero on the first call; the set kernel leaves it zero again.
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

extern "C" int gymrl_sumtree_store_new(double* d_tree, int capacity, int n, const int32_t* d_ring_state, double* d_max_scratch,
                                       int32_t* d_winner_scratch, void* stream) {
    GYMRL_REQUIRE(d_tree && d_ring_state && d_max_scratch && d_winner_scratch && n > 0 && capacity >= n, "bad arguments");
    cudaStream_t s = as_stream(stream);
    int blocks = ceil_div(capacity, 1024);
    if (blocks > GYMRL_NUM_SMS * 2) blocks = GYMRL_NUM_SMS * 2;
    sumtree_max_kernel<<<blocks, 256, 0, s>>>(d_tree, capacity, d_max_scratch);
    gymrl_count_launch();
    int rc = sumtree_launch_set(d_tree, capacity, n, 1, d_winner_scratch, nullptr, nullptr, nullptr, 0.f, 0.f, 0.f, d_ring_state, d_max_scratch, s);
    if (rc != GYMRL_OK) return rc;
    GYMRL_LAUNCH_CHECK("sumtree_store_new");
    return GYMRL_OK;
}

// Stratified sampling (rainbow :220-256): v_i = U(seg*i, seg*(i+1)); descend; w_i = (size * p_i / total)^(-beta), then / max w.
__global__ void sumtree_sample_kernel(const double* __restrict__ tree, int capacity, int B, const double* __restrict__ uniforms,
                                      const int32_t* __restrict__ ring_state, const double* __restrict__ beta_ptr, int32_t* __restrict__ out_idx,
                                      float* __restrict__ out_w, double* __restrict__ out_prio, unsigned int* __restrict__ wmax_bits,
                                      int return_tree_index, uint64_t seed, uint32_t draw, const uint32_t* __restrict__ draw_base) {
    // levels 0 .. TL of the tree (<= 2047 nodes) are staged in shared memory: the descent's first 11 of ~21 dependent loads
    __shared__ double s_top[2047];
    const int TL = min(10, st_top_level(capacity));
    const int staged = TL >= 0 ? (2 << TL) - 1 : 0;
    for (int x = threadIdx.x; x < staged; x += blockDim.x) s_top[x] = tree[x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        if (draw_base) draw += *draw_base;
        const double total = staged > 0 ? s_top[0] : tree[0];
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
            const double lv = left < staged ? s_top[left] : tree[left];
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
    // is_weight /= is_weight.max() (rainbow :243) by the block that finishes last: scratch[0] = max bits, scratch[1] = blocks done;
    // both are left zero for the next call (no separate zeroing / normalising launches)
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(wmax_bits + 1, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        __threadfence();
        const float wmax = __uint_as_float(atomicOr(wmax_bits, 0u));
        for (int k0 = 0; k0 < B; k0 += 16 * blockDim.x) {     // 16 independent loads in flight per thread
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { const int k = k0 + j * blockDim.x + threadIdx.x; v[j] = k < B ? __ldcg(out_w + k) : 0.f; }
#pragma unroll
            for (int j = 0; j < 16; ++j) { const int k = k0 + j * blockDim.x + threadIdx.x; if (k < B) out_w[k] = v[j] / wmax; }
        }
        __syncthreads();
        if (threadIdx.x == 0) { wmax_bits[0] = 0u; wmax_bits[1] = 0u; }
    }
}
extern "C" int gymrl_sumtree_sample(const double* d_tree, int capacity, int batch, const double* d_uniforms,
                                    const int32_t* d_ring_state, const double* d_beta, int32_t* d_out_idx, float* d_out_is_weight,
                                    double* d_out_priority, uint32_t* d_scratch_u32, int return_tree_index, uint64_t seed,
                                    uint32_t draw, const uint32_t* d_draw_base, void* stream) {
    GYMRL_REQUIRE(d_tree && d_ring_state && d_beta && d_out_idx && d_out_is_weight && d_scratch_u32 && batch > 0 && capacity > 0, "bad arguments");
    cudaStream_t s = as_stream(stream);
    // d_scratch_u32: 2 words, zero on the first call (the kernel leaves them zero)
    sumtree_sample_kernel<<<ceil_div(batch, 256), 256, 0, s>>>(d_tree, capacity, batch, d_uniforms, d_ring_state, d_beta, d_out_idx,
                                                              d_out_is_weight, d_out_priority, d_scratch_u32, return_tree_index,
                                                              seed, draw, d_draw_base);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("sumtree_sample");
    return GYMRL_OK;
}
