// comm.cu — the path's one collective: the sum of the flat fp32 gradient over the ranks of ONE node, as a
// one-shot reduction over NVLink peer memory fused with the sum of squares the global-norm clip needs.
//
// The reference has no collectives (single process); SURVEY §8b proposes `gymrl_comm_init / gymrl_allreduce_grads`
// for the env-sharded engine.  Round 1 used ncclAllReduce: 320 blocking 0.8 MB all-reduces per PPO update, ~46 us each at 8 GPUs
// (latency bound: NCCL's ring/tree steps for < 1 MB), followed by a separate norm pass.  Here every rank
//   1. publishes its gradient in a peer-visible staging buffer (cudaMalloc'ed by this library, exported with cudaIpc),
//   2. raises a per-block flag in every peer's flag table (st.release.sys over NVLink),
//   3. waits for the same block's flag from every peer, then reads the block's slice from all W staging buffers
//      (ld.global.cv through the peer mappings), adds them in rank order 0..W-1 — the same order on every rank, so all ranks
//      hold bit-identical sums — writes the reduced slice to a local buffer and the slice's sum of squares to partials[block];
// gymrl_clip_adam_step then consumes (reduced, partials).  No grid-wide synchronisation: block b of a rank only waits for
// block b of its peers.  Two staging buffers alternate by launch parity, so a rank that runs ahead never overwrites a slice a
// slower peer still reads (it cannot pass barrier i+1 before every peer has finished reading in launch i).  The launch
// counter lives in device memory and is advanced by the kernel, so the kernel replays unchanged inside a CUDA graph.
// All spin waits are bounded and trap on timeout: a lost peer fails loudly instead of hanging the GPU.
#include "common.cuh"

#include <new>
#include <vector>

void gymrl_count_launch(int n = 1);

#define COMM_MAX_WORLD 16
#define COMM_THREADS 256

struct gymrl_comm {
    int rank = 0, world = 1, n_blocks = 0, device = 0;
    long long n = 0, n_pad = 0;         // floats in the gradient, padded to a multiple of 4 * n_blocks
    size_t bytes = 0;                   // size of the local allocation
    uint8_t* local = nullptr;           // [staging0 | staging1 | flags]
    uint8_t* peer[COMM_MAX_WORLD] = {}; // peer[r] = mapping of rank r's allocation (peer[rank] == local)
    bool opened[COMM_MAX_WORLD] = {};
    uint32_t* epoch = nullptr;          // device: number of completed launches
    uint32_t* done = nullptr;           // device: blocks finished in the current launch
    // device-side table of peer base pointers
    uint8_t** d_peer = nullptr;
    long long* d_dbg = nullptr;         // developer probe: %globaltimer stamps of block 0 (gymrl_debug_comm_stamps)
};

struct CommLayout {
    size_t staging_bytes;   // one staging buffer
    size_t flags_off;       // byte offset of the flag table: [world][n_blocks] uint32
};
static CommLayout comm_layout(long long n_pad, int world, int n_blocks) {
    CommLayout L;
    L.staging_bytes = ((size_t)n_pad * 4 + 255) & ~(size_t)255;
    L.flags_off = 2 * L.staging_bytes;
    (void)world; (void)n_blocks;
    return L;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_cv4(const float4* p) {
    float4 v;
    asm volatile("ld.global.cv.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// grid = n_blocks, block = COMM_THREADS.  slice = n_pad / n_blocks floats (a multiple of 4).
__global__ void __launch_bounds__(COMM_THREADS) peer_reduce_sumsq_kernel(const float* __restrict__ grad, long long n, long long n_pad,
                                                                        float* __restrict__ reduced, double* __restrict__ partials,
                                                                        uint8_t* const* __restrict__ peers, int rank, int world,
                                                                        size_t staging_bytes, size_t flags_off,
                                                                        uint32_t* __restrict__ epoch_ctr, uint32_t* __restrict__ done_ctr,
                                                                        long long* __restrict__ dbg) {
    __shared__ double scratch[32];
#define COMM_STAMP(k) do { if (dbg && blockIdx.x == 0 && threadIdx.x == 0) { long long _v; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_v)); dbg[k] = _v; } } while (0)
    COMM_STAMP(0);
    const int b = blockIdx.x, nb = gridDim.x, t = threadIdx.x;
    const uint32_t e = *epoch_ctr + 1u;             // every block reads it before the last block of this launch advances it
    const int parity = (int)(e & 1u);
    const long long slice = n_pad / nb, lo = (long long)b * slice;
    const int n4 = (int)(slice >> 2);
    float4* my_stage = reinterpret_cast<float4*>(peers[rank] + (size_t)parity * staging_bytes) + (lo >> 2);
    // 1. publish this rank's slice (zero-padded past n)
    for (int i = t; i < n4; i += COMM_THREADS) {
        const long long g0 = lo + 4ll * i;
        float4 v;
        if (g0 + 3 < n) v = *reinterpret_cast<const float4*>(grad + g0);
        else {
            v.x = g0 < n ? grad[g0] : 0.f; v.y = g0 + 1 < n ? grad[g0 + 1] : 0.f;
            v.z = g0 + 2 < n ? grad[g0 + 2] : 0.f; v.w = 0.f;
        }
        my_stage[i] = v;
    }
    // bar.sync orders the block's staging writes before thread t's st.release.sys (release is cumulative over what the
    // releasing thread has observed), so no per-thread system fence is needed
    __syncthreads();
    COMM_STAMP(1);
    // 2. raise flag[rank][b] in every rank's table, 3. wait for flag[r][b] of every rank r in the local table
    if (t < world) {
        uint32_t* remote = reinterpret_cast<uint32_t*>(peers[t] + flags_off) + (size_t)rank * nb + b;
        st_release_sys(remote, e);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(peers[rank] + flags_off) + (size_t)t * nb + b;
        uint32_t it = 0;
        // flags only grow; a peer that already ran ahead by one launch shows e + 1 (it cannot be further ahead: it needs our flag)
        while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
            if (++it > (1u << 26)) __trap();     // ~ seconds: a rank died or the launch sequences diverged
            __nanosleep(20);
        }
    }
    __syncthreads();
    COMM_STAMP(2);
    // 4. fixed-order sum over ranks + sum of squares of the reduced slice
    double q = 0.0;
    for (int i = t; i < n4; i += COMM_THREADS) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int r0 = 0; r0 < world; r0 += 8) {      // up to eight peer loads in flight: one NVLink round trip for a whole node
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (r0 + j < world)
                    v[j] = ld_cv4(reinterpret_cast<const float4*>(peers[r0 + j] + (size_t)parity * staging_bytes) + (lo >> 2) + i);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (r0 + j < world) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
        }
        const long long g0 = lo + 4ll * i;
        if (g0 + 3 < n) *reinterpret_cast<float4*>(reduced + g0) = acc;
        else {
            if (g0 < n) reduced[g0] = acc.x;
            if (g0 + 1 < n) reduced[g0 + 1] = acc.y;
            if (g0 + 2 < n) reduced[g0 + 2] = acc.z;
        }
        q += (double)acc.x * acc.x + (double)acc.y * acc.y + (double)acc.z * acc.z + (double)acc.w * acc.w;
    }
    q = block_sum(q, scratch);
    COMM_STAMP(3);
    if (t == 0) {
        partials[b] = q;
        __threadfence();
        if (atomicAdd(done_ctr, 1u) == (uint32_t)nb - 1u) {
            *epoch_ctr = e;
            *done_ctr = 0u;
        }
    }
}

extern "C" int gymrl_comm_create(gymrl_comm** out, int rank, int world, long long n_floats, int n_blocks) {
    GYMRL_REQUIRE(out && world >= 1 && world <= COMM_MAX_WORLD && rank >= 0 && rank < world && n_floats > 0, "bad arguments");
    if (n_blocks <= 0) n_blocks = (int)((n_floats / 4 + COMM_THREADS - 1) / COMM_THREADS);   // one float4 per thread: every peer load of a slice is in flight at once
    if (n_blocks < 1) n_blocks = 1;
    if (n_blocks > 1024) n_blocks = 1024;
    GYMRL_REQUIRE(n_blocks <= 1024, "n_blocks too large");
    gymrl_comm* c = new (std::nothrow) gymrl_comm();
    GYMRL_REQUIRE(c, "out of host memory");
    c->rank = rank; c->world = world; c->n_blocks = n_blocks; c->n = n_floats;
    const long long q = 4ll * n_blocks;
    c->n_pad = (n_floats + q - 1) / q * q;
    GYMRL_CUDA(cudaGetDevice(&c->device));
    const CommLayout L = comm_layout(c->n_pad, world, n_blocks);
    c->bytes = L.flags_off + (size_t)world * n_blocks * sizeof(uint32_t);
    GYMRL_CUDA(cudaMalloc(&c->local, c->bytes));
    GYMRL_CUDA(cudaMemset(c->local, 0, c->bytes));
    GYMRL_CUDA(cudaMalloc(&c->epoch, 2 * sizeof(uint32_t)));
    GYMRL_CUDA(cudaMemset(c->epoch, 0, 2 * sizeof(uint32_t)));
    c->done = c->epoch + 1;
    GYMRL_CUDA(cudaMalloc(&c->d_peer, COMM_MAX_WORLD * sizeof(uint8_t*)));
    c->peer[rank] = c->local;
    GYMRL_CUDA(cudaDeviceSynchronize());
    *out = c;
    return GYMRL_OK;
}

extern "C" int gymrl_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int gymrl_comm_get_handle(gymrl_comm* c, void* handle_out) {
    GYMRL_REQUIRE(c && handle_out, "bad arguments");
    cudaIpcMemHandle_t h;
    GYMRL_CUDA(cudaIpcGetMemHandle(&h, c->local));
    memcpy(handle_out, &h, sizeof(h));
    return GYMRL_OK;
}

// handles: world x gymrl_comm_handle_bytes() bytes, in rank order (every rank's gymrl_comm_get_handle output, exchanged by the
// host program — torch.distributed.all_gather in gymrl_b200/dist.py).
extern "C" int gymrl_comm_open(gymrl_comm* c, const void* handles) {
    GYMRL_REQUIRE(c && handles, "bad arguments");
    const uint8_t* hb = static_cast<const uint8_t*>(handles);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hb + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        GYMRL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[r] = static_cast<uint8_t*>(p);
        c->opened[r] = true;
    }
    GYMRL_CUDA(cudaMemcpy(c->d_peer, c->peer, COMM_MAX_WORLD * sizeof(uint8_t*), cudaMemcpyHostToDevice));
    return GYMRL_OK;
}

extern "C" int gymrl_comm_n_partials(const gymrl_comm* c) { return c ? c->n_blocks : 0; }

// reduced[i] = sum_r grad_r[i] (rank order, identical bits on every rank); partials[b] = sum of squares of block b's slice of
// `reduced`.  Every rank of the communicator must enqueue the same sequence of these calls.  Capture-safe.
extern "C" int gymrl_comm_allreduce_sumsq(gymrl_comm* c, const float* d_grad, float* d_reduced, double* d_sumsq_partials, void* stream) {
    GYMRL_REQUIRE(c && d_grad && d_reduced && d_sumsq_partials, "bad arguments");
    GYMRL_REQUIRE((reinterpret_cast<uintptr_t>(d_grad) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_reduced) & 15) == 0,
                  "gradient buffers must be 16-byte aligned");
    for (int r = 0; r < c->world; ++r) GYMRL_REQUIRE(c->peer[r] != nullptr, "gymrl_comm_open has not been called (rank %d unmapped)", r);
    const CommLayout L = comm_layout(c->n_pad, c->world, c->n_blocks);
    peer_reduce_sumsq_kernel<<<c->n_blocks, COMM_THREADS, 0, as_stream(stream)>>>(d_grad, c->n, c->n_pad, d_reduced, d_sumsq_partials,
                                                                                 c->d_peer, c->rank, c->world, L.staging_bytes,
                                                                                 L.flags_off, c->epoch, c->done, c->d_dbg);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("peer_reduce_sumsq");
    return GYMRL_OK;
}

// developer probe (not part of the ABI): d_buf[0..3] = %globaltimer (ns) of block 0 at entry / published / peers arrived / reduced
extern "C" int gymrl_debug_comm_stamps(gymrl_comm* c, long long* d_buf) {
    if (!c) return GYMRL_EINVAL;
    c->d_dbg = d_buf;
    return GYMRL_OK;
}

extern "C" int gymrl_comm_destroy(gymrl_comm* c) {
    if (!c) return GYMRL_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r] && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->local) cudaFree(c->local);
    if (c->epoch) cudaFree(c->epoch);
    if (c->d_peer) cudaFree(c->d_peer);
    delete c;
    return GYMRL_OK;
}
