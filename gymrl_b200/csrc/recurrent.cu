// recurrent.cu — building blocks of the recurrent PPO / PPG family: the sequence-minibatch gather of
// algorithms/ppo_lstm_lunarlander.py:682-707 and the GRU cell of torch.nn.GRU as the reference uses it
// (ppo_rnn_lunarlander.py:124-139 MLPRNN, ppo_lstm_lunarlander.py:449-492 URNN, ppg_rnn_lunarlander.py:330-395).
//
//  gymrl_seq_gather:        out[b][t][:] = src[seq_index[b] * L + t][:]     (states.view(S, L, -1)[perm[start:end]])
//  gymrl_gru_cell_forward:  r = sigmoid(gi_r + gh_r), z = sigmoid(gi_z + gh_z), n = tanh(gi_n + r * gh_n), h' = (1 - z) n + z h
//                           with gi = x W_ih^T + b_ih and gh = h W_hh^T + b_hh computed by gymrl_linear_forward ([B][3H], gate
//                           order r | z | n as in torch); the gates are saved for the backward pass
//  gymrl_gru_cell_backward: dgi, dgh ([B][3H]) and the direct part of dh (z * dh'); the caller adds dgh W_hh and runs the
//                           weight gradients through gymrl_linear_backward_weight (BPTT = this pair per time step)
// Element-wise, one thread per (row, hidden unit): bandwidth bound, 7 reads + 4 writes of 4 B per element forward.
#include "common.cuh"

void gymrl_count_launch(int n = 1);

__global__ void seq_gather_kernel(const float* __restrict__ src, int lds, const int32_t* __restrict__ seq_index, int L, int D,
                                  float* __restrict__ out, int ldo, long long total) {
    // one thread per output float4 (D % 4 == 0) or float
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int per_row = D;
    const long long row = i / per_row;
    const int c = (int)(i - row * per_row);
    const int b = (int)(row / L), t = (int)(row - (long long)b * L);
    const long long srow = (long long)seq_index[b] * L + t;
    out[row * ldo + c] = src[srow * lds + c];
}

extern "C" int gymrl_seq_gather(const float* d_src, int ld_src, const int32_t* d_seq_index, int n_seq, int seq_len, int width,
                                float* d_out, int ld_out, void* stream) {
    GYMRL_REQUIRE(d_src && d_seq_index && d_out && n_seq > 0 && seq_len > 0 && width > 0 && ld_src >= width && ld_out >= width,
                  "bad arguments");
    const long long total = (long long)n_seq * seq_len * width;
    seq_gather_kernel<<<(int)ceil_div_ll(total, 256), 256, 0, as_stream(stream)>>>(d_src, ld_src, d_seq_index, seq_len, width, d_out,
                                                                                  ld_out, total);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("seq_gather");
    return GYMRL_OK;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void gru_cell_forward_kernel(const float* __restrict__ gi, int ldgi, const float* __restrict__ gh, int ldgh,
                                        const float* __restrict__ h, int ldh, float* __restrict__ h_out, int ldo,
                                        float* __restrict__ gates, int B, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, j = i - b * H;
    const float* gib = gi + (size_t)b * ldgi;
    const float* ghb = gh + (size_t)b * ldgh;
    const float r = sigmoidf_(gib[j] + ghb[j]);
    const float z = sigmoidf_(gib[H + j] + ghb[H + j]);
    const float n = tanhf(gib[2 * H + j] + r * ghb[2 * H + j]);
    const float hp = h[(size_t)b * ldh + j];
    h_out[(size_t)b * ldo + j] = (1.0f - z) * n + z * hp;
    if (gates) {
        float* g = gates + (size_t)b * 3 * H;
        g[j] = r; g[H + j] = z; g[2 * H + j] = n;
    }
}

extern "C" int gymrl_gru_cell_forward(const float* d_gi, int ld_gi, const float* d_gh, int ld_gh, const float* d_h, int ld_h,
                                      float* d_h_out, int ld_out, float* d_gates, int batch, int hidden, void* stream) {
    GYMRL_REQUIRE(d_gi && d_gh && d_h && d_h_out && batch > 0 && hidden > 0, "bad arguments");
    GYMRL_REQUIRE(ld_gi >= 3 * hidden && ld_gh >= 3 * hidden && ld_h >= hidden && ld_out >= hidden, "leading dimensions too small");
    gru_cell_forward_kernel<<<ceil_div(batch * hidden, 256), 256, 0, as_stream(stream)>>>(d_gi, ld_gi, d_gh, ld_gh, d_h, ld_h, d_h_out,
                                                                                         ld_out, d_gates, batch, hidden);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("gru_cell_forward");
    return GYMRL_OK;
}

__global__ void gru_cell_backward_kernel(const float* __restrict__ dh_out, int lddo, const float* __restrict__ gates,
                                         const float* __restrict__ gh, int ldgh, const float* __restrict__ h, int ldh,
                                         float* __restrict__ dgi, int lddgi, float* __restrict__ dgh, int lddgh,
                                         float* __restrict__ dh, int lddh, int accumulate_dh, int B, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, j = i - b * H;
    const float* g = gates + (size_t)b * 3 * H;
    const float r = g[j], z = g[H + j], n = g[2 * H + j];
    const float hp = h[(size_t)b * ldh + j];
    const float ghn = gh[(size_t)b * ldgh + 2 * H + j];
    const float d = dh_out[(size_t)b * lddo + j];
    // h' = (1 - z) n + z h
    const float dn = d * (1.0f - z);
    const float dz = d * (hp - n);
    const float dan = dn * (1.0f - n * n);         // pre-activation of n:  gi_n + r * gh_n
    const float dr = dan * ghn;
    const float daz = dz * z * (1.0f - z);
    const float dar = dr * r * (1.0f - r);
    float* dgib = dgi + (size_t)b * lddgi;
    float* dghb = dgh + (size_t)b * lddgh;
    dgib[j] = dar;           dghb[j] = dar;
    dgib[H + j] = daz;       dghb[H + j] = daz;
    dgib[2 * H + j] = dan;   dghb[2 * H + j] = dan * r;
    if (dh) {
        float* p = dh + (size_t)b * lddh + j;
        *p = (accumulate_dh ? *p : 0.0f) + d * z;
    }
}

extern "C" int gymrl_gru_cell_backward(const float* d_dh_out, int ld_dh_out, const float* d_gates, const float* d_gh, int ld_gh,
                                       const float* d_h, int ld_h, float* d_dgi, int ld_dgi, float* d_dgh, int ld_dgh, float* d_dh,
                                       int ld_dh, int accumulate_dh, int batch, int hidden, void* stream) {
    GYMRL_REQUIRE(d_dh_out && d_gates && d_gh && d_h && d_dgi && d_dgh && batch > 0 && hidden > 0, "bad arguments");
    GYMRL_REQUIRE(ld_gh >= 3 * hidden && ld_dgi >= 3 * hidden && ld_dgh >= 3 * hidden && ld_h >= hidden && ld_dh_out >= hidden &&
                  (!d_dh || ld_dh >= hidden), "leading dimensions too small");
    gru_cell_backward_kernel<<<ceil_div(batch * hidden, 256), 256, 0, as_stream(stream)>>>(d_dh_out, ld_dh_out, d_gates, d_gh, ld_gh, d_h,
                                                                                          ld_h, d_dgi, ld_dgi, d_dgh, ld_dgh, d_dh, ld_dh,
                                                                                          accumulate_dh, batch, hidden);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("gru_cell_backward");
    return GYMRL_OK;
}
