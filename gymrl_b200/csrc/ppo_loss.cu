// ppo_loss.cu — PPO clipped-surrogate / dual-clip / ERC-masked / value-clip losses with their analytic
// gradients wrt (logits, V) in one pass.
//
//  mode GYMRL_PPO_DUALCLIP : PPOTrainer.update, algorithms/ppo_lunarlander.py:278-300
//        L = -mean(where(A<0, max(min(rA, clip(r)A), c A), min(rA, clip(r)A))) + vc mean((V-R)^2) - ec mean(H)
//  mode GYMRL_PPO_FULL     : update_model, algorithms/ppo_full_lunarlander.py:586-633
//        mask = 1[1-b_lo < H/(H_old+1e-8) < 1+b_hi];  L = mean(-min(clamp(r,0,c)A, clip(r)A) mask)
//        + mean(vc mask (V-R)^2) - ec mean(H mask)         (plain .mean(): masked rows count as zeros)
//  flag GYMRL_PPO_VALUE_CLIP: V_c = V_old + clamp(V - V_old, -e_lo, +e_hi); value term max((V-R)^2, (V_c-R)^2)
//        (algorithms/ppo_lstm_lunarlander.py:763-771)
// Autograd conventions reproduced: torch.min/max split the gradient evenly on exact ties; clamp passes
// the gradient on the closed interval.  One thread per sample; logits rows are contiguous (A <= 16);
// the minibatch permutation is applied through row_index so action/logp_old/adv/ret are read in place
// (one 4 B gather each - the rows they share a sector with are other samples of the same minibatch
// only by chance, so this kernel is gather-latency bound, not bandwidth bound).  The five metric
// sums are block-reduced and added with one atomic per block.
#include "common.cuh"
#include "rowwise.cuh"

void gymrl_count_launch(int n = 1);
bool gymrl_defer_reduce(const float* part, long long stride, int splits, long long count, float* out, int accumulate);

#define MAX_A 16

struct PpoSample {
    float dlogits[MAX_A];
    float dvalue;
    float m_pol, m_val, m_ent, m_clip, m_kl, m_erc;
};

// One sample of the loss and its gradient wrt (logits, V).  `A` may be a compile-time constant at the call site.
__device__ __forceinline__ void ppo_sample(const float* z, float V, int a, float lpo, float Adv, float R, float ent_old_r,
                                           float val_old_r, const gymrl_ppo_cfg& cfg, float ent_coef, float invB, int A,
                                           PpoSample& o, float invN = -1.0f) {
    // invN: normaliser of the masked terms (policy / value / entropy / clip_frac) — 1/B for the plain .mean() of ppo_full
    // (masked rows count as zeros), 1/mask.sum() for ppo_lstm's masked_mean (GYMRL_PPO_MASKED_MEAN)
    if (invN < 0.0f) invN = invB;
    // loops run to MAX_A with a guard so that every array index is a compile-time constant (arrays stay in registers)
    float ln[MAX_A], p[MAX_A];
    float mx = z[0];
#pragma unroll
    for (int j = 1; j < MAX_A; ++j)
        if (j < A) mx = fmaxf(mx, z[j]);
    // softmax numerators through ex2.approx (relative error ~|z - max| * 6e-8, far inside the 1e-5 parity bar) and the
    // probabilities as e_j / s instead of a second exponential per action: the loss is evaluated per sample on a serial
    // chain (on every lane in the fused heads kernel), where the eight libdevice expf calls were ~80 instructions
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
        if (j < A) {
            p[j] = exp2f_approx((z[j] - mx) * 1.4426950408889634f);
            s += p[j];
        }
    const float lse = logf(s) + mx;
    const float inv_s = 1.0f / s;
    float H = 0.f;
    float lp = 0.f;
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
        if (j < A) {
            ln[j] = z[j] - lse;
            p[j] *= inv_s;
            H -= p[j] * ln[j];
            lp = j == a ? ln[j] : lp;
        }
    const float ratio = expf(lp - lpo);
    const float lo = 1.0f - cfg.clip_eps_min, hi = 1.0f + cfg.clip_eps_max;
    const bool in_range = ratio >= lo && ratio <= hi;
    const float surr2 = fminf(fmaxf(ratio, lo), hi) * Adv;
    const int mode = cfg.mode & 3;
    float mask = 1.0f;
    float obj, g;  // objective (to maximise) and d obj / d ratio
    if (mode == GYMRL_PPO_FULL) {
        const float er = H / (ent_old_r + 1e-8f);
        mask = (er > 1.0f - cfg.erc_low && er < 1.0f + cfg.erc_high) ? 1.0f : 0.0f;
        const float cr = fminf(fmaxf(ratio, 0.0f), cfg.dual_clip);
        const float surr1 = cr * Adv;
        const float g1 = (ratio >= 0.0f && ratio <= cfg.dual_clip) ? Adv : 0.0f;
        const float g2 = in_range ? Adv : 0.0f;
        obj = fminf(surr1, surr2);
        g = surr1 < surr2 ? g1 : (surr1 == surr2 ? 0.5f * (g1 + g2) : g2);
    } else {
        const float surr1 = ratio * Adv;
        const float g2 = in_range ? Adv : 0.0f;
        const float min_surr = fminf(surr1, surr2);
        float gm = surr1 < surr2 ? Adv : (surr1 == surr2 ? 0.5f * (Adv + g2) : g2);
        obj = min_surr;
        g = gm;
        if (Adv < 0.0f) {
            const float dc = cfg.dual_clip * Adv;
            obj = fmaxf(min_surr, dc);
            g = min_surr > dc ? gm : (min_surr == dc ? 0.5f * gm : 0.0f);
        }
    }
    // value term
    const float e1 = V - R;
    float vterm = e1 * e1, dv = 2.0f * e1;
    if (cfg.mode & GYMRL_PPO_VALUE_CLIP) {
        const float vo = val_old_r;
        const float dlt = V - vo;
        const float vcl = vo + fminf(fmaxf(dlt, -cfg.vclip_eps_min), cfg.vclip_eps_max);
        const float e2 = vcl - R;
        const float l2 = e2 * e2;
        const float pass = (dlt >= -cfg.vclip_eps_min && dlt <= cfg.vclip_eps_max) ? 1.0f : 0.0f;
        const float d2 = 2.0f * e2 * pass;
        if (l2 > vterm) { vterm = l2; dv = d2; }
        else if (l2 == vterm) { dv = 0.5f * (dv + d2); }
    }
    // gradients: L = -obj*mask/N + vc*mask*vterm/N - ec*mask*H/N   (N = B, or mask.sum() under masked_mean)
    const float dL_dlp = -(g * ratio) * mask * invN;
    const float dL_dH = -ent_coef * mask * invN;
#pragma unroll
    for (int j = 0; j < MAX_A; ++j)
        if (j < A) {
            const float dlp = (j == a ? 1.0f : 0.0f) - p[j];
            const float dH = -p[j] * (ln[j] + H);
            o.dlogits[j] = dL_dlp * dlp + dL_dH * dH;
        }
    o.dvalue = cfg.value_coef * mask * dv * invN;
    o.m_pol = -obj * mask * invN;
    o.m_val = cfg.value_coef * mask * vterm * invN;
    o.m_ent = H * mask * invN;
    o.m_clip = ((ratio < lo || ratio > hi) ? 1.0f : 0.0f) * mask * invN;
    o.m_kl = (lpo - lp) * invB;
    o.m_erc = (1.0f - mask) * invB;
}

// mask.sum() of the ERC mask over the minibatch (the denominator of masked_mean): same entropy arithmetic as ppo_sample.
__global__ void ppo_mask_count_kernel(const float* __restrict__ logits, int ldl, const int32_t* __restrict__ row_index,
                                      const float* __restrict__ ent_old, float* __restrict__ count, int B, int A, gymrl_ppo_cfg cfg) {
    __shared__ double dscratch[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float mask = 0.f;
    if (i < B) {
        const int r = row_index ? row_index[i] : i;
        float mx = logits[(size_t)i * ldl];
        for (int j = 1; j < A; ++j) mx = fmaxf(mx, logits[(size_t)i * ldl + j]);
        float p[MAX_A], s = 0.f;
#pragma unroll
        for (int j = 0; j < MAX_A; ++j)
            if (j < A) { p[j] = exp2f_approx((logits[(size_t)i * ldl + j] - mx) * 1.4426950408889634f); s += p[j]; }
        const float lse = logf(s) + mx, inv_s = 1.0f / s;
        float H = 0.f;
#pragma unroll
        for (int j = 0; j < MAX_A; ++j)
            if (j < A) { const float pj = p[j] * inv_s; H -= pj * (logits[(size_t)i * ldl + j] - lse); }
        const float er = H / (ent_old[r] + 1e-8f);
        mask = (er > 1.0f - cfg.erc_low && er < 1.0f + cfg.erc_high) ? 1.0f : 0.0f;
    }
    double m = block_sum((double)mask, dscratch);
    const double v[1] = {m};
    ordered_block_accumulate<1>(v, count, dscratch);
}

__global__ void ppo_loss_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ value, int ldv,
                                const int32_t* __restrict__ row_index, const int32_t* __restrict__ action,
                                const float* __restrict__ logp_old, const float* __restrict__ adv,
                                const float* __restrict__ ret, const float* __restrict__ ent_old,
                                const float* __restrict__ val_old, float* __restrict__ dlogits, int lddl,
                                float* __restrict__ dvalue, int lddv, float* __restrict__ metrics, int B, int A,
                                gymrl_ppo_cfg cfg) {
    __shared__ float scratch[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float invB = 1.0f / (float)B;
    float invN = invB;
    if (cfg.mode & GYMRL_PPO_MASKED_MEAN) {     // masked_mean (ppo_lstm :646-655): sum / mask.sum(), 0 when nothing is unmasked
        const float cnt = *cfg.d_mask_count;
        invN = cnt > 0.0f ? 1.0f / cnt : 0.0f;
    }
    const float ent_coef = cfg.d_entropy_coef ? *cfg.d_entropy_coef : cfg.entropy_coef;
    float m_pol = 0.f, m_val = 0.f, m_ent = 0.f, m_clip = 0.f, m_kl = 0.f, m_erc = 0.f;
    if (i < B) {
        const int r = row_index ? row_index[i] : i;
        float z[MAX_A];
#pragma unroll
        for (int j = 0; j < MAX_A; ++j)
            if (j < A) z[j] = logits[(size_t)i * ldl + j];
        PpoSample o;
        ppo_sample(z, value[(size_t)i * ldv], action[r], logp_old[r], adv[r], ret[r], ent_old ? ent_old[r] : 0.f,
                   val_old ? val_old[r] : 0.f, cfg, ent_coef, invB, A, o, invN);
#pragma unroll
        for (int j = 0; j < MAX_A; ++j)
            if (j < A) dlogits[(size_t)i * lddl + j] = o.dlogits[j];
        dvalue[(size_t)i * lddv] = o.dvalue;
        m_pol = o.m_pol; m_val = o.m_val; m_ent = o.m_ent; m_clip = o.m_clip; m_kl = o.m_kl; m_erc = o.m_erc;
    }
    m_pol = block_sum(m_pol, scratch);
    m_val = block_sum(m_val, scratch);
    m_ent = block_sum(m_ent, scratch);
    m_clip = block_sum(m_clip, scratch);
    m_kl = block_sum(m_kl, scratch);
    m_erc = block_sum(m_erc, scratch);
    if (metrics) {
        __shared__ double dscratch[32];
        const double v[8] = {m_pol, m_val, m_ent, m_clip, m_kl, m_erc, m_pol + m_val - ent_coef * m_ent, blockIdx.x == 0 ? 1.0 : 0.0};
        ordered_block_accumulate<8>(v, metrics, dscratch);   // fixed block order (reporting only, but reproducible)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused output heads + loss + heads backward (one sweep over the [M][2H] head-trunk activations).
//   h = (h_a | h_c): logits = Wa h_a + ba (A x H), V = Wc h_c + bc;  loss as above;
//   dh_a = (Wa^T dlogits) act'(h_a), dh_c = dV Wc act'(h_c);  dWa += dlogits (x) h_a, dba += dlogits, dWc += dV h_c, dbc += dV.
// Replaces skinny_fwd<A> + skinny_fwd<1> + ppo_loss + 2 x skinny_bwd_fused (+ their reductions): those five launches each
// re-read (half of) the [M][2H] activation and are latency-bound at ~10 us each; this reads it once and writes dh once.
// One warp per row, lane l owns elements {128 j + 4 l ..+3} of each half; the A + 1 dot products are warp reductions, the
// per-sample loss runs redundantly on every lane (registers only), weight-gradient sums stay in registers over the rows
// a warp owns and leave through per-block partials + a fixed-order reduction (deterministic).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kHeadsWarps = 8;
constexpr int kHeadsMaxBlocks = 4 * GYMRL_NUM_SMS;

template <int NCH, int A>
__global__ void __launch_bounds__(kHeadsWarps * 32, 2)
ppo_heads_fused_kernel(const float* __restrict__ h, int ldh, const float* __restrict__ Wa, const float* __restrict__ ba,
                       const float* __restrict__ Wc, const float* __restrict__ bc, const int32_t* __restrict__ row_index,
                       const int32_t* __restrict__ action, const float* __restrict__ logp_old, const float* __restrict__ adv,
                       const float* __restrict__ ret, const float* __restrict__ ent_old, const float* __restrict__ val_old,
                       float* __restrict__ dh, int lddh, int act_tanh, float* __restrict__ lv_out, float* __restrict__ partials,
                       float* __restrict__ metrics, int M, gymrl_ppo_cfg cfg) {
    constexpr int H = 128 * NCH;
    constexpr int P = A * H + A + H + 1;     // [dWa | dba | dWc | dbc]
    constexpr int PS = (P + 3) & ~3;
    static_assert(A % 4 == 0, "slice offsets must stay 16-byte aligned");
    extern __shared__ __align__(16) float s_acc[];
    __shared__ float s_met[kHeadsWarps][6];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warp = blockIdx.x * kHeadsWarps + wid, nwarps = gridDim.x * kHeadsWarps;
    pdl_wait();
    pdl_launch_dependents();
    float4 wa[A][NCH], wc[NCH], gwa[A][NCH], gwc[NCH];
    float bav[A];
    // every lane evaluates the same loss, so the scalar sums (A + 1 bias gradients, 6 metrics) are spread over lanes 0..A+6:
    // one accumulator register per lane instead of A + 7 on all of them (the kernel is at its 128-register budget)
    float lane_acc = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
#pragma unroll
        for (int a = 0; a < A; ++a) {
            wa[a][j] = ld4(Wa + (size_t)a * H + 128 * j + 4 * lane);
            gwa[a][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        wc[j] = ld4(Wc + 128 * j + 4 * lane);
        gwc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int a = 0; a < A; ++a) bav[a] = ba[a];
    const float bcv = bc[0];
    const float invB = 1.0f / (float)M;
    const float ent_coef = cfg.d_entropy_coef ? *cfg.d_entropy_coef : cfg.entropy_coef;
    for (int row = warp; row < M; row += nwarps) {
        const float* hp = h + (size_t)row * ldh;
        float4 ha[NCH], hc[NCH];
        float z[A + 1];
#pragma unroll
        for (int a = 0; a <= A; ++a) z[a] = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            ha[j] = ld4(hp + 128 * j + 4 * lane);
            hc[j] = ld4(hp + H + 128 * j + 4 * lane);
#pragma unroll
            for (int a = 0; a < A; ++a) z[a] += dot4(ha[j], wa[a][j]);
            z[A] += dot4(hc[j], wc[j]);
        }
        const int r = row_index ? row_index[row] : row;
        const int act = action[r];
        const float lpo = logp_old[r], Adv = adv[r], R = ret[r];
        const float eo = ent_old ? ent_old[r] : 0.f, vo = val_old ? val_old[r] : 0.f;
#pragma unroll
        for (int a = 0; a <= A; ++a) z[a] = warp_sum(z[a]);
#pragma unroll
        for (int a = 0; a < A; ++a) z[a] += bav[a];
        const float V = z[A] + bcv;
        PpoSample o;
        ppo_sample(z, V, act, lpo, Adv, R, eo, vo, cfg, ent_coef, invB, A, o);
        if (lv_out && lane == 0) {
#pragma unroll
            for (int a = 0; a < A; ++a) lv_out[(size_t)row * 8 + a] = z[a];
            lv_out[(size_t)row * 8 + A] = V;
        }
        {
            float v = o.m_erc;                       // lane A + 6
#pragma unroll
            for (int a = 0; a < A; ++a) v = lane == a ? o.dlogits[a] : v;
            v = lane == A ? o.dvalue : v;
            v = lane == A + 1 ? o.m_pol : v;
            v = lane == A + 2 ? o.m_val : v;
            v = lane == A + 3 ? o.m_ent : v;
            v = lane == A + 4 ? o.m_clip : v;
            v = lane == A + 5 ? o.m_kl : v;
            lane_acc += v;
        }
        float* dp = dh + (size_t)row * lddh;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float4 da = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int a = 0; a < A; ++a) {
                da = axpy4(o.dlogits[a], wa[a][j], da);
                gwa[a][j] = axpy4(o.dlogits[a], ha[j], gwa[a][j]);
            }
            float4 dc = scale4(o.dvalue, wc[j]);
            gwc[j] = axpy4(o.dvalue, hc[j], gwc[j]);
            if (act_tanh) {
                da = mul4(da, make_float4(1.f - ha[j].x * ha[j].x, 1.f - ha[j].y * ha[j].y, 1.f - ha[j].z * ha[j].z, 1.f - ha[j].w * ha[j].w));
                dc = mul4(dc, make_float4(1.f - hc[j].x * hc[j].x, 1.f - hc[j].y * hc[j].y, 1.f - hc[j].z * hc[j].z, 1.f - hc[j].w * hc[j].w));
            }
            st4(dp + 128 * j + 4 * lane, da);
            st4(dp + H + 128 * j + 4 * lane, dc);
        }
    }
    // every warp writes its register sums into its own [P] slice; then all threads fold the slices in warp order (the same
    // order of additions as the old one-warp-at-a-time accumulation, which cost eight barrier rounds at the end of the kernel)
    {
        float* mine = s_acc + (size_t)wid * PS;                   // slice stride PS: P rounded up to 16 bytes
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
#pragma unroll
            for (int a = 0; a < A; ++a) st4(mine + (size_t)a * H + 128 * j + 4 * lane, gwa[a][j]);
            st4(mine + A * H + A + 128 * j + 4 * lane, gwc[j]);   // A * H + A is a multiple of 4 floats
        }
        if (lane < A) mine[A * H + lane] = lane_acc;
        else if (lane == A) mine[A * H + A + H] = lane_acc;
        else if (lane < A + 7) s_met[wid][lane - A - 1] = lane_acc;
    }
    __syncthreads();
    float* out = partials + (size_t)blockIdx.x * P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < kHeadsWarps; ++w) sum += s_acc[(size_t)w * PS + i];
        out[i] = sum;
    }
    if (metrics) {
        __shared__ double dscratch[32];
        float m[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (threadIdx.x == 0)
            for (int w = 0; w < kHeadsWarps; ++w)
                for (int k = 0; k < 6; ++k) m[k] += s_met[w][k];
        const double v[8] = {m[0], m[1], m[2], m[3], m[4], m[5], m[0] + m[1] - ent_coef * m[2], blockIdx.x == 0 ? 1.0 : 0.0};
        ordered_block_accumulate<8>(v, metrics, dscratch);   // fixed block order (reporting only, but reproducible)
    }
}

extern "C" size_t gymrl_ppo_heads_workspace_bytes(int H, int A) {
    return (size_t)kHeadsMaxBlocks * (size_t)(A * H + A + H + 1) * sizeof(float);
}

extern "C" int gymrl_ppo_heads_fused(const float* d_h, int ldh, const float* d_Wa, const float* d_ba, const float* d_Wc,
                                     const float* d_bc, const int32_t* d_row_index, const int32_t* d_action,
                                     const float* d_logp_old, const float* d_adv, const float* d_ret, const float* d_entropy_old,
                                     const float* d_value_old, float* d_dh, int lddh, int act_in, float* d_dWa, float* d_dba,
                                     float* d_dWc, float* d_dbc, float* d_lv_out, float* d_metrics, void* d_workspace,
                                     size_t workspace_bytes, int accumulate, int batch, int H, int n_actions,
                                     const gymrl_ppo_cfg* cfg, void* stream) {
    GYMRL_REQUIRE(cfg != nullptr, "cfg is NULL");
    GYMRL_REQUIRE(d_h && d_Wa && d_ba && d_Wc && d_bc && d_action && d_logp_old && d_adv && d_ret && d_dh && d_dWa && d_dba && d_dWc &&
                  d_dbc && d_workspace, "NULL pointer");
    GYMRL_REQUIRE(batch > 0 && (H == 128 || H == 256) && n_actions == 4, "fused heads are built for H in {128, 256} and 4 actions "
                  "(got H=%d, A=%d): use gymrl_linear_* + gymrl_ppo_loss otherwise", H, n_actions);
    GYMRL_REQUIRE(act_in == 0 || act_in == 1, "act_in must be GYMRL_ACT_NONE or GYMRL_ACT_TANH");
    GYMRL_REQUIRE(ldh % 4 == 0 && lddh % 4 == 0, "leading dimensions must be multiples of 4 floats");
    GYMRL_REQUIRE((cfg->mode & 3) <= GYMRL_PPO_FULL, "unknown PPO mode %d", cfg->mode);
    GYMRL_REQUIRE((cfg->mode & 3) != GYMRL_PPO_FULL || d_entropy_old, "GYMRL_PPO_FULL needs d_entropy_old");
    GYMRL_REQUIRE(!(cfg->mode & GYMRL_PPO_VALUE_CLIP) || d_value_old, "VALUE_CLIP needs d_value_old");
    GYMRL_REQUIRE(!(cfg->mode & GYMRL_PPO_MASKED_MEAN), "MASKED_MEAN needs a pass over the whole minibatch first: use gymrl_ppo_loss");
    const int A = 4;
    const int P = A * H + A + H + 1;
    int grid = ceil_div(batch, kHeadsWarps * 8);           // >= 8 rows per warp so the per-warp set-up amortises
    if (grid > kHeadsMaxBlocks) grid = kHeadsMaxBlocks;
    if (grid < 1) grid = 1;
    GYMRL_REQUIRE(workspace_bytes >= (size_t)grid * P * sizeof(float), "workspace too small: need %zu bytes", (size_t)grid * P * sizeof(float));
    cudaStream_t s = as_stream(stream);
    const size_t smem = (size_t)kHeadsWarps * ((P + 3) & ~3) * sizeof(float);   // one [P] slice per warp (41 KB at H = 256)
    float* partials = (float*)d_workspace;
    if (H == 256)
        gymrl_launch_pdl(ppo_heads_fused_kernel<2, 4>, dim3(grid), dim3(kHeadsWarps * 32), smem, s, d_h, ldh, d_Wa, d_ba, d_Wc, d_bc, d_row_index, d_action, d_logp_old, d_adv,
                                                                           d_ret, d_entropy_old, d_value_old, d_dh, lddh, act_in, d_lv_out, partials,
                                                                           d_metrics, batch, *cfg);
    else
        gymrl_launch_pdl(ppo_heads_fused_kernel<1, 4>, dim3(grid), dim3(kHeadsWarps * 32), smem, s, d_h, ldh, d_Wa, d_ba, d_Wc, d_bc, d_row_index, d_action, d_logp_old, d_adv,
                                                                           d_ret, d_entropy_old, d_value_old, d_dh, lddh, act_in, d_lv_out, partials,
                                                                           d_metrics, batch, *cfg);
    if (gymrl_defer_reduce(partials, P, grid, A * H, d_dWa, accumulate)) {
        gymrl_defer_reduce(partials + A * H, P, grid, A, d_dba, accumulate);
        gymrl_defer_reduce(partials + A * H + A, P, grid, H, d_dWc, accumulate);
        gymrl_defer_reduce(partials + A * H + A + H, P, grid, 1, d_dbc, accumulate);
        gymrl_count_launch(1);
    } else {
        launch_reduce_blocks(partials, grid, P, d_dWa, A * H, d_dba, A, d_dWc, H, d_dbc, 1, accumulate, s);
        gymrl_count_launch(2);
    }
    GYMRL_LAUNCH_CHECK("ppo_heads_fused");
    return GYMRL_OK;
}

extern "C" int gymrl_ppo_loss(const float* d_logits, int ld_logits, const float* d_value, int ld_value,
                              const int32_t* d_row_index, const int32_t* d_action, const float* d_logp_old,
                              const float* d_adv, const float* d_ret, const float* d_entropy_old,
                              const float* d_value_old, float* d_dlogits, int ld_dlogits, float* d_dvalue, int ld_dvalue,
                              float* d_metrics, int batch, int n_actions, const gymrl_ppo_cfg* cfg, void* stream) {
    GYMRL_REQUIRE(cfg != nullptr, "cfg is NULL");
    GYMRL_REQUIRE(d_logits && d_value && d_action && d_logp_old && d_adv && d_ret && d_dlogits && d_dvalue, "NULL pointer");
    GYMRL_REQUIRE(batch > 0 && n_actions > 0 && n_actions <= MAX_A, "bad batch=%d / n_actions=%d", batch, n_actions);
    GYMRL_REQUIRE((cfg->mode & 3) <= GYMRL_PPO_FULL, "unknown PPO mode %d", cfg->mode);
    GYMRL_REQUIRE((cfg->mode & 3) != GYMRL_PPO_FULL || d_entropy_old, "GYMRL_PPO_FULL needs d_entropy_old");
    GYMRL_REQUIRE(!(cfg->mode & GYMRL_PPO_VALUE_CLIP) || d_value_old, "VALUE_CLIP needs d_value_old");
    if (cfg->mode & GYMRL_PPO_MASKED_MEAN) {
        GYMRL_REQUIRE((cfg->mode & 3) == GYMRL_PPO_FULL && cfg->d_mask_count, "MASKED_MEAN needs GYMRL_PPO_FULL and cfg.d_mask_count");
        GYMRL_CUDA(cudaMemsetAsync(cfg->d_mask_count, 0, sizeof(float), as_stream(stream)));
        ppo_mask_count_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(d_logits, ld_logits, d_row_index, d_entropy_old,
                                                                                  cfg->d_mask_count, batch, n_actions, *cfg);
        gymrl_count_launch();
    }
    ppo_loss_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(
        d_logits, ld_logits, d_value, ld_value, d_row_index, d_action, d_logp_old, d_adv, d_ret, d_entropy_old, d_value_old,
        d_dlogits, ld_dlogits, d_dvalue, ld_dvalue, d_metrics, batch, n_actions, *cfg);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("ppo_loss");
    return GYMRL_OK;
}
