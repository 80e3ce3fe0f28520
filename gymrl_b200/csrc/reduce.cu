// reduce.cu — deferred, merged reduction of parameter-gradient partials.
//
// Every backward kernel of the dense layers leaves its parameter gradients as per-CTA (split-K / row-chunk) partials
// that a small kernel folds in a fixed order.  One launch per layer costs more in launch latency than in work
// (3-10 us each for a few hundred KB), so a caller may open a *deferral scope*: inside it gymrl_linear_backward*,
// gymrl_ppo_heads_fused ... record their pending sums instead of launching the fold, and gymrl_reduce_flush folds all
// of them in ONE launch — optionally leaving the per-block sums of squares of the final gradient, which
// gymrl_clip_adam_step turns into the global norm of nn.utils.clip_grad_norm_ (algorithms/ppo_lunarlander.py:304-306)
// without another pass over the gradient.
// Contract of the scope: each recorded backward must have its own workspace, alive until the flush has run.
// The fold order is fixed (split groups of a block, then group order): results are bit-reproducible run to run.
#include "common.cuh"

void gymrl_count_launch(int n = 1);

#include <cstdlib>
bool gymrl_pdl_enabled() {
    // off unless GYMRL_PDL=1: measured neutral on B200 (whole PPO minibatch 205.7 us serialised vs 207.1 us with PDL edges in the
    // epoch graph — kernel nodes of one graph already launch back to back), so the plain stream order stays the default
    static const bool on = [] { const char* e = getenv("GYMRL_PDL"); return e && e[0] == '1'; }();
    return on;
}

namespace {

constexpr int kMaxSegs = 24;
constexpr int kGroups = 8;

struct Seg {
    const float* part;   // partial k of output i at part[k * stride + i]
    long long stride;
    float* out;
    int splits, count;
    int accumulate;
    int per_block;       // outputs per block: 256 (one thread per output) or 32 (8 split groups per output)
    int block_begin;     // first block of this segment
};
struct SegTable {
    Seg seg[kMaxSegs];
    int nseg;
};

thread_local bool t_defer = false;
thread_local SegTable t_table;

__global__ void __launch_bounds__(256) reduce_segments_kernel(const SegTable tab, double* __restrict__ sumsq_partials) {
    __shared__ float sm[kGroups][32];
    __shared__ double scratch[32];
    pdl_wait();
    pdl_launch_dependents();
    int si = 0;
#pragma unroll 1
    for (int q = 1; q < tab.nseg; ++q)
        if ((int)blockIdx.x >= tab.seg[q].block_begin) si = q;
    const Seg& sg = tab.seg[si];
    const int blk = blockIdx.x - sg.block_begin;
    float v = 0.f;
    bool owner = false;
    if (sg.per_block == 256) {
        const int i = blk * 256 + threadIdx.x;
        if (i < sg.count) {
            const float* q = sg.part + i;
            float s = 0.f;
            int k = 0;
            for (; k + 4 <= sg.splits; k += 4, q += 4 * sg.stride) {
                const float v0 = q[0], v1 = q[sg.stride], v2 = q[2 * sg.stride], v3 = q[3 * sg.stride];
                s += v0; s += v1; s += v2; s += v3;
            }
            for (; k < sg.splits; ++k, q += sg.stride) s += *q;
            v = sg.accumulate ? sg.out[i] + s : s;
            sg.out[i] = v;
            owner = true;
        }
    } else {
        const int li = threadIdx.x & 31, g = threadIdx.x >> 5;
        const int i = blk * 32 + li;
        const int per = (sg.splits + kGroups - 1) / kGroups;
        const int k0 = g * per, k1 = min(sg.splits, k0 + per);
        float s = 0.f;
        if (i < sg.count) {
            const float* q = sg.part + (long long)k0 * sg.stride + i;
            int k = k0;
            for (; k + 4 <= k1; k += 4, q += 4 * sg.stride) {
                const float v0 = q[0], v1 = q[sg.stride], v2 = q[2 * sg.stride], v3 = q[3 * sg.stride];
                s += v0; s += v1; s += v2; s += v3;
            }
            for (; k < k1; ++k, q += sg.stride) s += *q;
        }
        sm[g][li] = s;
        __syncthreads();
        if (g == 0 && i < sg.count) {
            float t = sm[0][li];
#pragma unroll
            for (int q = 1; q < kGroups; ++q) t += sm[q][li];
            v = sg.accumulate ? sg.out[i] + t : t;
            sg.out[i] = v;
            owner = true;
        }
    }
    if (sumsq_partials) {   // uniform over the launch
        double q = owner ? (double)v * (double)v : 0.0;
        q = block_sum(q, scratch);
        if (threadIdx.x == 0) sumsq_partials[blockIdx.x] = q;
    }
}

int launch_table(const SegTable& tab_in, double* d_sumsq_partials, int capacity, int* n_blocks, cudaStream_t s) {
    SegTable tab = tab_in;
    int blocks = 0;
    for (int i = 0; i < tab.nseg; ++i) {
        Seg& g = tab.seg[i];
        g.per_block = (g.splits >= 32 && g.count <= 16384) ? 32 : 256;
        g.block_begin = blocks;
        blocks += ceil_div(g.count, g.per_block);
    }
    if (n_blocks) *n_blocks = blocks;
    if (blocks == 0) return GYMRL_OK;
    GYMRL_REQUIRE(!d_sumsq_partials || capacity >= blocks, "sum-of-squares partial buffer too small: need %d entries", blocks);
    gymrl_launch_pdl(reduce_segments_kernel, dim3(blocks), dim3(256), 0, s, tab, d_sumsq_partials);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("reduce_segments");
    return GYMRL_OK;
}

}  // namespace

// Internal hook for the backward kernels' host code: inside a deferral scope record `out[i] (+)= sum_k part[k*stride + i]`
// and return true; outside return false (the caller launches its own fold).
bool gymrl_defer_reduce(const float* part, long long stride, int splits, long long count, float* out, int accumulate) {
    if (!t_defer) return false;
    if (out == nullptr || count <= 0) return true;   // nothing to fold
    if (t_table.nseg >= kMaxSegs || count > 0x7fffffffll) {
        t_defer = false;   // cannot hold it: poison the scope so the flush reports it
        t_table.nseg = -1;
        return true;
    }
    Seg& g = t_table.seg[t_table.nseg++];
    g.part = part; g.stride = stride; g.out = out; g.splits = splits; g.count = (int)count; g.accumulate = accumulate;
    g.per_block = 0; g.block_begin = 0;
    return true;
}

extern "C" int gymrl_reduce_defer_begin(void) {
    GYMRL_REQUIRE(!t_defer, "a deferral scope is already open on this thread");
    t_defer = true;
    t_table.nseg = 0;
    return GYMRL_OK;
}

extern "C" int gymrl_reduce_flush(double* d_sumsq_partials, int capacity, int* n_partials, long long* n_outputs, void* stream) {
    const bool poisoned = t_table.nseg < 0;
    GYMRL_REQUIRE(t_defer || poisoned, "no deferral scope is open on this thread");
    t_defer = false;
    if (poisoned) {
        t_table.nseg = 0;
        GYMRL_FAIL(GYMRL_EINVAL, "more than %d pending reductions in one deferral scope", kMaxSegs);
    }
    int blocks = 0;
    long long outputs = 0;
    for (int i = 0; i < t_table.nseg; ++i) outputs += t_table.seg[i].count;
    if (n_outputs) *n_outputs = outputs;
    const int rc = launch_table(t_table, d_sumsq_partials, capacity, &blocks, as_stream(stream));
    t_table.nseg = 0;
    if (n_partials) *n_partials = blocks;
    return rc;
}
