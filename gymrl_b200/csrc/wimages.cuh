// wimages.cuh — internal interface of the weight-image registry (wimages.cu).
#pragma once
#include "common.cuh"

// If `w` is the start of a registered matrix of exactly (rows x cols): *hi_image = its hi image (transposed = the W^T images),
// the lo image lives *image_stride floats behind it.  Host-side, takes a mutex; O(#registered buffers).
bool wimg_lookup(const float* w, int rows, int cols, bool transposed, const float** hi_image, long long* image_stride);
// Re-split the images of the parameter buffer starting at d_param_base, if one is registered (no-op otherwise).
void wimg_refresh_if_registered(const float* d_param_base, cudaStream_t st);

// Device-side view of a registered buffer, for kernels that write parameters and want to refresh the images in the same pass
// (clip_adam_kernel): images == nullptr when the buffer is not registered.
struct WImgDev {
    float* images;
    const int* mats;     // device [n_mats][3] = offset, rows, cols
    int n_mats;
    long long n;
};
WImgDev wimg_device_view(const float* d_param_base);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t wimg_to_tf32_rn(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// write the hi / lo (and transposed) images of parameter element i whose new value is x
__device__ __forceinline__ void wimg_store(const WImgDev& w, long long i, float x) {
    const uint32_t hi = wimg_to_tf32_rn(x);
    const uint32_t lo = wimg_to_tf32_rn(x - __uint_as_float(hi));
    uint32_t* im = reinterpret_cast<uint32_t*>(w.images);
    im[i] = hi;
    im[w.n + i] = lo;
    for (int m = 0; m < w.n_mats; ++m) {
        const long long off = w.mats[3 * m];
        const int rows = w.mats[3 * m + 1], cols = w.mats[3 * m + 2];
        const long long j = i - off;
        if (j >= 0 && j < (long long)rows * cols) {
            const int r = (int)(j / cols), c = (int)(j - (long long)r * cols);
            const long long t = off + (long long)c * rows + r;
            im[2 * w.n + t] = hi;
            im[3 * w.n + t] = lo;
            break;
        }
    }
}
#endif
