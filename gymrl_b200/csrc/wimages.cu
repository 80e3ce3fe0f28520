// wimages.cu — pre-split tf32 images of the weight matrices, the TMA-side operand of the warp-specialised GEMM (linear_tc.cu).
//
// The 3xTF32 GEMM needs every operand as hi = tf32(x) and lo = tf32(x - hi).  Activations change every launch and are split in
// registers on their way into shared memory; WEIGHTS change once per optimizer step, so their images are produced once per step
// here and then fetched by TMA (cp.async.bulk.tensor, SWIZZLE_128B) straight into the UMMA operand layout — no register pass,
// no converter warps for B.  Per registered flat parameter buffer of n floats the image buffer holds 4 n floats:
//     [ hi | lo | hi^T | lo^T ]     same element offsets as the flat buffer; for a registered matrix (offset, rows, cols) the
//                                   transposed images hold W^T ([cols][rows], row-major) at the same offset
// hi^T / lo^T feed the backward-input GEMM dX = dY W as a K-major operand (B[n = in-feature][k = out-feature]).
// The optimiser / target-sync entry points (gymrl_adam_step, gymrl_clip_adam_step, gymrl_polyak) refresh the images of a
// registered buffer themselves (wimg_refresh_if_registered), so library-side parameter writes can never leave them stale; a host
// program that writes parameters behind the library's back (load_state_dict, broadcast) calls gymrl_weight_images_refresh.
#include "common.cuh"
#include "wimages.cuh"

#include <mutex>
#include <vector>

void gymrl_count_launch(int n = 1);

namespace {

struct WSet {
    const float* flat;
    long long n;
    float* images;
    int n_mats;
    int* d_mats;                 // device copy of [n_mats][3] = offset, rows, cols
    std::vector<int> mats;
};
std::vector<WSet> g_sets;
std::mutex g_mu;

__global__ void wimg_refresh_kernel(const float* __restrict__ flat, WImgDev w) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n) return;
    wimg_store(w, i, flat[i]);
}

int refresh_set(const WSet& s, cudaStream_t st) {
    WImgDev w;
    w.images = s.images; w.mats = s.d_mats; w.n_mats = s.n_mats; w.n = s.n;
    wimg_refresh_kernel<<<(int)ceil_div_ll(s.n, 256), 256, 0, st>>>(s.flat, w);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("weight_images_refresh");
    return GYMRL_OK;
}

}  // namespace

bool wimg_lookup(const float* w, int rows, int cols, bool transposed, const float** hi_image, long long* image_stride) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (const WSet& s : g_sets) {
        if (w < s.flat || w >= s.flat + s.n) continue;
        const long long off = w - s.flat;
        for (int m = 0; m < s.n_mats; ++m) {
            if (s.mats[3 * m] == off && s.mats[3 * m + 1] == rows && s.mats[3 * m + 2] == cols) {
                *hi_image = s.images + (transposed ? 2 * s.n : 0) + off;
                *image_stride = s.n;
                return true;
            }
        }
        return false;
    }
    return false;
}

WImgDev wimg_device_view(const float* d_param_base) {
    WImgDev w;
    w.images = nullptr; w.mats = nullptr; w.n_mats = 0; w.n = 0;
    std::lock_guard<std::mutex> lk(g_mu);
    for (const WSet& s : g_sets)
        if (s.flat == d_param_base) { w.images = s.images; w.mats = s.d_mats; w.n_mats = s.n_mats; w.n = s.n; break; }
    return w;
}

void wimg_refresh_if_registered(const float* d_param_base, cudaStream_t st) {
    WSet hit;
    bool found = false;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (const WSet& s : g_sets)
            if (s.flat == d_param_base) { hit = s; found = true; break; }
    }
    if (found) (void)refresh_set(hit, st);
}

extern "C" int gymrl_weight_images_register(const float* d_flat, long long n_floats, float* d_images, const int* mats, int n_mats) {
    GYMRL_REQUIRE(d_flat && d_images && n_floats > 0 && n_floats % 4 == 0 && n_mats >= 0 && (n_mats == 0 || mats), "bad arguments");
    GYMRL_REQUIRE((reinterpret_cast<uintptr_t>(d_images) & 15) == 0, "image buffer must be 16-byte aligned");
    for (int m = 0; m < n_mats; ++m)
        GYMRL_REQUIRE(mats[3 * m] >= 0 && mats[3 * m] % 4 == 0 && mats[3 * m + 1] > 0 && mats[3 * m + 2] > 0 &&
                      (long long)mats[3 * m] + (long long)mats[3 * m + 1] * mats[3 * m + 2] <= n_floats, "matrix %d out of range", m);
    WSet s;
    s.flat = d_flat; s.n = n_floats; s.images = d_images; s.n_mats = n_mats; s.d_mats = nullptr;
    s.mats.assign(mats, mats + 3 * n_mats);
    if (n_mats > 0) {
        GYMRL_CUDA(cudaMalloc(&s.d_mats, sizeof(int) * 3 * n_mats));
        GYMRL_CUDA(cudaMemcpy(s.d_mats, mats, sizeof(int) * 3 * n_mats, cudaMemcpyHostToDevice));
    }
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t k = 0; k < g_sets.size(); ++k)
        if (g_sets[k].flat == d_flat) {
            if (g_sets[k].d_mats) cudaFree(g_sets[k].d_mats);
            g_sets.erase(g_sets.begin() + k);
            break;
        }
    g_sets.push_back(s);
    return GYMRL_OK;
}

extern "C" int gymrl_weight_images_refresh(const float* d_flat, void* stream) {
    GYMRL_REQUIRE(d_flat, "NULL pointer");
    WSet hit;
    bool found = false;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (const WSet& s : g_sets)
            if (s.flat == d_flat) { hit = s; found = true; break; }
    }
    GYMRL_REQUIRE(found, "no weight images registered for this parameter buffer");
    return refresh_set(hit, as_stream(stream));
}

extern "C" int gymrl_weight_images_unregister(const float* d_flat) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t k = 0; k < g_sets.size(); ++k)
        if (g_sets[k].flat == d_flat) {
            if (g_sets[k].d_mats) cudaFree(g_sets[k].d_mats);
            g_sets.erase(g_sets.begin() + k);
            return GYMRL_OK;
        }
    return GYMRL_OK;
}
