// l2_write_probe.cu — how fast can one SM / all SMs store to an L2-resident buffer?  (developer micro-benchmark that sizes
// the GEMM epilogue: a 128x256 fp32 tile is 128 KB per CTA.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gymrl_b200/lib/l2_write_probe tools/probes/l2_write_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

// every CTA writes its own `bytes_per_cta` region `reps` times with coalesced STG.128 (a warp covers 512 contiguous bytes)
__global__ void write_kernel(float4* out, size_t f4_per_cta, int reps, long long* cycles) {
    float4* base = out + (size_t)blockIdx.x * f4_per_cta;
    const float4 v = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
        for (size_t i = threadIdx.x; i < f4_per_cta; i += blockDim.x) base[i] = v;
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
// same but each thread stores a row-owned 16 B piece: lane l writes row l (row pitch `pitch_f4`), i.e. 32 lines per STG
__global__ void write_rows_kernel(float4* out, size_t f4_per_cta, int reps, long long* cycles) {
    float4* base = out + (size_t)blockIdx.x * f4_per_cta;
    const float4 v = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    const long long t0 = clock64();
    // region viewed as rows of 64 float4 (1 KB); a warp owns 32 rows at a time, lane = row, loop over the 64 pieces
    const size_t rows = f4_per_cta / 64;
    for (int r = 0; r < reps; ++r)
        for (size_t r0 = (size_t)warp * 32; r0 < rows; r0 += (size_t)nw * 32)
            for (int c = 0; c < 64; ++c) base[(r0 + lane) * 64 + c] = v;
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    const size_t bytes_per_cta = 128 * 1024;
    float4* buf;
    long long* cyc;
    cudaMalloc(&buf, 1024 * bytes_per_cta);
    cudaMalloc(&cyc, 1024 * sizeof(long long));
    long long h[1024];
    for (int rows = 0; rows < 2; ++rows)
        for (int ctas : {1, 8, 37, 74, 148, 296})
            for (int threads : {256, 1024}) {
                const int reps = 8;
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0); cudaEventCreate(&e1);
                for (int w = 0; w < 2; ++w) {
                    cudaEventRecord(e0);
                    if (rows) write_rows_kernel<<<ctas, threads>>>(buf, bytes_per_cta / 16, reps, cyc);
                    else write_kernel<<<ctas, threads>>>(buf, bytes_per_cta / 16, reps, cyc);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                cudaMemcpy(h, cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
                long long mx = 0; double avg = 0;
                for (int i = 0; i < ctas; ++i) { mx = h[i] > mx ? h[i] : mx; avg += h[i]; }
                avg /= ctas;
                const double bytes = (double)bytes_per_cta * reps;
                printf("%s ctas %3d threads %4d: %.1f B/clk/CTA (avg), %.1f (slowest); kernel %.1f us -> %.2f TB/s aggregate\n",
                       rows ? "row-owned " : "coalesced ", ctas, threads, bytes / avg, bytes / mx, ms * 1e3, bytes * ctas / (ms * 1e-3) / 1e12);
            }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
