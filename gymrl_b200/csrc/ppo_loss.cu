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

void gymrl_count_launch(int n = 1);

#define MAX_A 16

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
    const float ent_coef = cfg.d_entropy_coef ? *cfg.d_entropy_coef : cfg.entropy_coef;
    float m_pol = 0.f, m_val = 0.f, m_ent = 0.f, m_clip = 0.f, m_kl = 0.f, m_erc = 0.f;
    if (i < B) {
        const int r = row_index ? row_index[i] : i;
        float z[MAX_A], ln[MAX_A], p[MAX_A];
        for (int j = 0; j < A; ++j) z[j] = logits[(size_t)i * ldl + j];
        float mx = z[0];
        for (int j = 1; j < A; ++j) mx = fmaxf(mx, z[j]);
        float s = 0.f;
        for (int j = 0; j < A; ++j) s += expf(z[j] - mx);
        const float lse = logf(s) + mx;
        float H = 0.f;
        for (int j = 0; j < A; ++j) {
            ln[j] = z[j] - lse;
            p[j] = expf(ln[j]);
            H -= p[j] * ln[j];
        }
        const int a = action[r];
        const float lp = ln[a], lpo = logp_old[r], Adv = adv[r], R = ret[r], V = value[(size_t)i * ldv];
        const float ratio = expf(lp - lpo);
        const float lo = 1.0f - cfg.clip_eps_min, hi = 1.0f + cfg.clip_eps_max;
        const bool in_range = ratio >= lo && ratio <= hi;
        const float surr2 = fminf(fmaxf(ratio, lo), hi) * Adv;
        const int mode = cfg.mode & 3;
        float mask = 1.0f;
        float obj, g;  // objective (to maximise) and d obj / d ratio
        if (mode == GYMRL_PPO_FULL) {
            const float er = H / (ent_old[r] + 1e-8f);
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
            const float vo = val_old[r];
            const float dlt = V - vo;
            const float vcl = vo + fminf(fmaxf(dlt, -cfg.vclip_eps_min), cfg.vclip_eps_max);
            const float e2 = vcl - R;
            const float l2 = e2 * e2;
            const float pass = (dlt >= -cfg.vclip_eps_min && dlt <= cfg.vclip_eps_max) ? 1.0f : 0.0f;
            const float d2 = 2.0f * e2 * pass;
            if (l2 > vterm) { vterm = l2; dv = d2; }
            else if (l2 == vterm) { dv = 0.5f * (dv + d2); }
        }
        // gradients: L = -obj*mask/B + vc*mask*vterm/B - ec*mask*H/B
        const float dL_dlp = -(g * ratio) * mask * invB;
        const float dL_dH = -ent_coef * mask * invB;
        for (int j = 0; j < A; ++j) {
            const float dlp = (j == a ? 1.0f : 0.0f) - p[j];
            const float dH = -p[j] * (ln[j] + H);
            dlogits[(size_t)i * lddl + j] = dL_dlp * dlp + dL_dH * dH;
        }
        dvalue[(size_t)i * lddv] = cfg.value_coef * mask * dv * invB;
        m_pol = -obj * mask * invB;
        m_val = cfg.value_coef * mask * vterm * invB;
        m_ent = H * mask * invB;
        m_clip = ((ratio < lo || ratio > hi) ? 1.0f : 0.0f) * mask * invB;
        m_kl = (lpo - lp) * invB;
        m_erc = (1.0f - mask) * invB;
    }
    m_pol = block_sum(m_pol, scratch);
    m_val = block_sum(m_val, scratch);
    m_ent = block_sum(m_ent, scratch);
    m_clip = block_sum(m_clip, scratch);
    m_kl = block_sum(m_kl, scratch);
    m_erc = block_sum(m_erc, scratch);
    if (threadIdx.x == 0 && metrics) {
        atomicAdd(&metrics[0], m_pol);
        atomicAdd(&metrics[1], m_val);
        atomicAdd(&metrics[2], m_ent);
        atomicAdd(&metrics[3], m_clip);
        atomicAdd(&metrics[4], m_kl);
        atomicAdd(&metrics[5], m_erc);
        atomicAdd(&metrics[6], m_pol + m_val - ent_coef * m_ent);
        if (blockIdx.x == 0) atomicAdd(&metrics[7], 1.0f);
    }
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
    ppo_loss_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(
        d_logits, ld_logits, d_value, ld_value, d_row_index, d_action, d_logp_old, d_adv, d_ret, d_entropy_old, d_value_old,
        d_dlogits, ld_dlogits, d_dvalue, ld_dvalue, d_metrics, batch, n_actions, *cfg);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("ppo_loss");
    return GYMRL_OK;
}
