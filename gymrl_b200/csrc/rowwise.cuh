// rowwise.cuh — helpers shared by the row-wise (one warp per row) kernels: float4 arithmetic and the fixed-order
// reduction of per-block parameter-gradient partials.
#pragma once
#include "common.cuh"

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 axpy4(float a, float4 x, float4 y) {
    return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}
__device__ __forceinline__ float4 scale4(float a, float4 x) { return make_float4(a * x.x, a * x.y, a * x.z, a * x.w); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }


// out[p] (+)= sum_b partials[b][p] in a fixed order: 8 groups of a block each sum a contiguous slice of the blocks for 32
// consecutive outputs (coalesced 128 B lines), the 8 slice sums are then added in group order (deterministic, no atomics).
constexpr int kReduceGroups = 8;
__global__ void __launch_bounds__(32 * kReduceGroups)
reduce_blocks_kernel(const float* __restrict__ partials, int nblk, int P, float* __restrict__ o0, int n0, float* __restrict__ o1, int n1,
                     float* __restrict__ o2, int n2, float* __restrict__ o3, int n3, int accumulate) {
    __shared__ float s_part[kReduceGroups][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const int per = (nblk + kReduceGroups - 1) / kReduceGroups;
    const int b0 = grp * per, b1 = min(nblk, b0 + per);
    float s = 0.f;
    if (p < P) {
        const float* q = partials + (size_t)b0 * P + p;
        int b = b0;
        for (; b + 4 <= b1; b += 4, q += 4 * (size_t)P) {
            const float v0 = q[0], v1 = q[P], v2 = q[2 * (size_t)P], v3 = q[3 * (size_t)P];
            s += v0; s += v1; s += v2; s += v3;
        }
        for (; b < b1; ++b, q += P) s += *q;
    }
    s_part[grp][lane] = s;
    __syncthreads();
    if (grp != 0 || p >= P) return;
    float t = s_part[0][lane];
#pragma unroll
    for (int g = 1; g < kReduceGroups; ++g) t += s_part[g][lane];
    float* dst;
    int q = p;
    if (q < n0) dst = o0 + q;
    else if ((q -= n0) < n1) dst = o1 + q;
    else if ((q -= n1) < n2) dst = o2 + q;
    else { q -= n2; dst = o3 + q; }
    (void)n3;
    *dst = accumulate ? *dst + t : t;
}

inline void launch_reduce_blocks(const float* partials, int nblk, int P, float* o0, int n0, float* o1, int n1, float* o2, int n2,
                                 float* o3, int n3, int accumulate, cudaStream_t s) {
    reduce_blocks_kernel<<<(P + 31) / 32, 32 * kReduceGroups, 0, s>>>(partials, nblk, P, o0, n0, o1, n1, o2, n2, o3, n3, accumulate);
}

}  // namespace
