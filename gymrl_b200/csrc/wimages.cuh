// wimages.cuh — internal interface of the weight-image registry (wimages.cu).
#pragma once
#include "common.cuh"

// If `w` is the start of a registered matrix of exactly (rows x cols): *hi_image = its hi image (transposed = the W^T images),
// the lo image lives *image_stride floats behind it.  Host-side, takes a mutex; O(#registered buffers).
bool wimg_lookup(const float* w, int rows, int cols, bool transposed, const float** hi_image, long long* image_stride);
// Re-split the images of the parameter buffer starting at d_param_base, if one is registered (no-op otherwise).
void wimg_refresh_if_registered(const float* d_param_base, cudaStream_t st);
