// qlearn.cu — TD targets, TD losses and their output gradients for the value-based / actor-critic
// off-policy trainers, plus NoisyLinear helpers (a18).
//
//  gymrl_dqn_loss        DQNTrainer.update (algorithms/dqn_cartpole.py:135-168): y = r + g max_a' Qt(s') (1-d), MSE
//                        RainbowDQNTrainer.update (rainbow_dqn_cartpole.py:311-361): double-Q with online argmax,
//                        y = R + g^n (1-terminal) Qt(s')[a*], loss = mean(w td^2), td exported for the PER write-back;
//                        optional dueling aggregation Q = V + A - mean(A) (rainbow :100-113) fused fwd + bwd.
//  gymrl_twin_q_target   SAC (sac_pendulum.py:233-237) / TD3 (td3_pendulum.py:194-204) target
//  gymrl_twin_q_loss     mse(q1,y) + mse(q2,y) and d/dq1, d/dq2 (sac :239-242, td3 :206-209)
//  gymrl_sac_actor_grad  d/d(mean, log_std) of mean(alpha*logpi - min(Q1,Q2)) through the reparameterised
//                        tanh-Gaussian (sac :76-87, :248-251), given dQmin/da from the critic's input gradient
//  gymrl_td3_actor_grad  d/d(pre-tanh) of -mean Q1(s, tanh(.)*bound) (td3 :215-217)
//  gymrl_sac_alpha_step  loss -mean(log_alpha*(logpi+Hbar)) + Adam on the float64 scalar (sac :257-263)
//  gymrl_noisy_*         factorised Gaussian noise f(x)=sign(x)sqrt|x| and W = mu + sigma*outer(eps_out, eps_in)
//                        (rainbow :51-97)
// All are one-thread-per-sample elementwise kernels over B <= 16K rows of <= 16 floats: launch/latency
// bound by construction; they exist to keep the whole update on the device inside one CUDA graph.
#include "common.cuh"

void gymrl_count_launch(int n = 1);

#define MAX_A 16

// ------------------------------------------------------------------------------------------------ DQN family
struct DqnArgs {
    const float* q; int ldq;             // online Q(s) [B][A]  (or advantage stream if dueling)
    const float* v; int ldv;             // dueling value stream [B][1] (nullable)
    const float* qn_t; int ldqt;         // target net on s' [B][A] (advantage stream if dueling)
    const float* vn_t; int ldvt;         // dueling value stream of the target net on s'
    const float* qn_o; int ldqo;         // online net on s' (double-Q action selection; nullable -> max over target)
    const float* vn_o; int ldvo;
    const int32_t* row_index;            // gathers action / reward / done
    const int32_t* action; const float* reward; const float* done; const float* is_weight;
    float* dq; int lddq; float* dv; int lddv;
    float* td_error; float* loss_acc;    // loss_acc[0] += loss, [1] += 1
    int B, A; float gamma_n;
};

__device__ __forceinline__ void dueling_q(const float* adv, const float* v, int A, float* q) {
    float m = 0.f;
    for (int j = 0; j < A; ++j) m += adv[j];
    m = m / (float)A;
    for (int j = 0; j < A; ++j) q[j] = v[0] + (adv[j] - m);
}

__global__ void dqn_loss_kernel(DqnArgs p) {
    __shared__ float scratch[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float li = 0.f;
    if (i < p.B) {
        const int r = p.row_index ? p.row_index[i] : i;
        const int A = p.A;
        float q[MAX_A], qt[MAX_A], qo[MAX_A];
        if (p.v) {
            dueling_q(p.q + (size_t)i * p.ldq, p.v + (size_t)i * p.ldv, A, q);
            dueling_q(p.qn_t + (size_t)i * p.ldqt, p.vn_t + (size_t)i * p.ldvt, A, qt);
            if (p.qn_o) dueling_q(p.qn_o + (size_t)i * p.ldqo, p.vn_o + (size_t)i * p.ldvo, A, qo);
        } else {
            for (int j = 0; j < A; ++j) { q[j] = p.q[(size_t)i * p.ldq + j]; qt[j] = p.qn_t[(size_t)i * p.ldqt + j]; }
            if (p.qn_o) for (int j = 0; j < A; ++j) qo[j] = p.qn_o[(size_t)i * p.ldqo + j];
        }
        float next_q;
        if (p.qn_o) {  // double DQN: a* = argmax online(s') (first max), evaluated by the target net
            int best = 0;
            for (int j = 1; j < A; ++j) if (qo[j] > qo[best]) best = j;
            next_q = qt[best];
        } else {
            next_q = qt[0];
            for (int j = 1; j < A; ++j) next_q = fmaxf(next_q, qt[j]);
        }
        const int a = p.action[r];
        // dqn:  rewards + gamma * next_q * (1 - dones);  rainbow: reward + gamma^n * (1 - terminal) * next_q
        const float y = p.qn_o ? (p.reward[r] + (p.gamma_n * (1.0f - p.done[r])) * next_q)
                               : (p.reward[r] + (p.gamma_n * next_q) * (1.0f - p.done[r]));
        const float td = q[a] - y;
        const float w = p.is_weight ? p.is_weight[i] : 1.0f;
        li = (td * td * w) / (float)p.B;
        const float g = 2.0f * td * w / (float)p.B;
        if (p.td_error) p.td_error[i] = td;
        if (p.v) {  // Q = V + A - mean(A):  dA_j = dQ_j - mean_k dQ_k = g (1[j==a] - 1/A);  dV = sum_j dQ_j = g
            for (int j = 0; j < A; ++j) p.dq[(size_t)i * p.lddq + j] = g * ((j == a ? 1.0f : 0.0f) - 1.0f / (float)A);
            p.dv[(size_t)i * p.lddv] = g;
        } else {
            for (int j = 0; j < A; ++j) p.dq[(size_t)i * p.lddq + j] = (j == a) ? g : 0.0f;
        }
    }
    li = block_sum(li, scratch);
    if (threadIdx.x == 0 && p.loss_acc) {
        atomicAdd(&p.loss_acc[0], li);
        if (blockIdx.x == 0) atomicAdd(&p.loss_acc[1], 1.0f);
    }
}

extern "C" int gymrl_dqn_loss(const float* d_q, int ld_q, const float* d_v, int ld_v, const float* d_qnext_target, int ld_qt,
                              const float* d_vnext_target, int ld_vt, const float* d_qnext_online, int ld_qo,
                              const float* d_vnext_online, int ld_vo, const int32_t* d_row_index, const int32_t* d_action,
                              const float* d_reward, const float* d_done, const float* d_is_weight, float* d_dq, int ld_dq,
                              float* d_dv, int ld_dv, float* d_td_error, float* d_loss_acc, int batch, int n_actions,
                              float gamma_n, void* stream) {
    GYMRL_REQUIRE(d_q && d_qnext_target && d_action && d_reward && d_done && d_dq, "NULL pointer");
    GYMRL_REQUIRE(batch > 0 && n_actions > 0 && n_actions <= MAX_A, "bad shape");
    GYMRL_REQUIRE(!d_v || (d_vnext_target && d_dv && (!d_qnext_online || d_vnext_online)), "dueling needs all value streams");
    DqnArgs p;
    p.q = d_q; p.ldq = ld_q; p.v = d_v; p.ldv = ld_v; p.qn_t = d_qnext_target; p.ldqt = ld_qt; p.vn_t = d_vnext_target; p.ldvt = ld_vt;
    p.qn_o = d_qnext_online; p.ldqo = ld_qo; p.vn_o = d_vnext_online; p.ldvo = ld_vo; p.row_index = d_row_index;
    p.action = d_action; p.reward = d_reward; p.done = d_done; p.is_weight = d_is_weight; p.dq = d_dq; p.lddq = ld_dq;
    p.dv = d_dv; p.lddv = ld_dv; p.td_error = d_td_error; p.loss_acc = d_loss_acc; p.B = batch; p.A = n_actions; p.gamma_n = gamma_n;
    dqn_loss_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(p);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("dqn_loss");
    return GYMRL_OK;
}

// ------------------------------------------------------------------------------------------------ twin-Q (SAC / TD3)
__global__ void twin_q_target_kernel(const float* __restrict__ reward, const float* __restrict__ done,
                                     const int32_t* __restrict__ row_index, const float* __restrict__ q1t, int ld1,
                                     const float* __restrict__ q2t, int ld2, const float* __restrict__ logp_next,
                                     const double* __restrict__ log_alpha, float gamma, float* __restrict__ y, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int r = row_index ? row_index[i] : i;
    float tq = fminf(q1t[(size_t)i * ld1], q2t[(size_t)i * ld2]);
    if (logp_next) {
        // alpha is a float64 0-dim tensor; alpha * next_log_probs promotes to float32
        const float alpha = (float)exp(*log_alpha);
        tq = tq - alpha * logp_next[i];
    }
    y[i] = reward[r] + (gamma * (1.0f - done[r])) * tq;
}

extern "C" int gymrl_twin_q_target(const float* d_reward, const float* d_done, const int32_t* d_row_index, const float* d_q1t,
                                   int ld_q1t, const float* d_q2t, int ld_q2t, const float* d_logp_next, const double* d_log_alpha,
                                   float gamma, float* d_y, int batch, void* stream) {
    GYMRL_REQUIRE(d_reward && d_done && d_q1t && d_q2t && d_y && batch > 0, "bad arguments");
    GYMRL_REQUIRE(!d_logp_next || d_log_alpha, "entropy term needs log_alpha");
    twin_q_target_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(d_reward, d_done, d_row_index, d_q1t, ld_q1t, d_q2t,
                                                                              ld_q2t, d_logp_next, d_log_alpha, gamma, d_y, batch);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("twin_q_target");
    return GYMRL_OK;
}

__global__ void twin_q_loss_kernel(const float* __restrict__ q1, int ld1, const float* __restrict__ q2, int ld2,
                                   const float* __restrict__ y, float* __restrict__ dq1, int ldd1, float* __restrict__ dq2,
                                   int ldd2, float* __restrict__ loss_acc, int B) {
    __shared__ float scratch[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float li = 0.f;
    if (i < B) {
        const float e1 = q1[(size_t)i * ld1] - y[i], e2 = q2[(size_t)i * ld2] - y[i];
        li = (e1 * e1 + e2 * e2) / (float)B;
        dq1[(size_t)i * ldd1] = 2.0f * e1 / (float)B;
        dq2[(size_t)i * ldd2] = 2.0f * e2 / (float)B;
    }
    li = block_sum(li, scratch);
    if (threadIdx.x == 0 && loss_acc) {
        atomicAdd(&loss_acc[0], li);
        if (blockIdx.x == 0) atomicAdd(&loss_acc[1], 1.0f);
    }
}

extern "C" int gymrl_twin_q_loss(const float* d_q1, int ld_q1, const float* d_q2, int ld_q2, const float* d_y, float* d_dq1,
                                 int ld_dq1, float* d_dq2, int ld_dq2, float* d_loss_acc, int batch, void* stream) {
    GYMRL_REQUIRE(d_q1 && d_q2 && d_y && d_dq1 && d_dq2 && batch > 0, "bad arguments");
    twin_q_loss_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(d_q1, ld_q1, d_q2, ld_q2, d_y, d_dq1, ld_dq1, d_dq2, ld_dq2,
                                                                            d_loss_acc, batch);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("twin_q_loss");
    return GYMRL_OK;
}

// dL/dq1, dL/dq2 of  L = mean(alpha*logpi - min(q1, q2))  (the -min part; ties split evenly like torch.min)
__global__ void min_q_grad_kernel(const float* __restrict__ q1, int ld1, const float* __restrict__ q2, int ld2,
                                  float* __restrict__ dq1, int ldd1, float* __restrict__ dq2, int ldd2, int B, int q1_only,
                                  float* __restrict__ acc) {
    __shared__ float scratch[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float li = 0.f;
    if (i < B) {
        const float g = -1.0f / (float)B;
        if (q1_only) {
            dq1[(size_t)i * ldd1] = g;
            li = q1 ? q1[(size_t)i * ld1] * g : 0.f;
        } else {
            const float a = q1[(size_t)i * ld1], b = q2[(size_t)i * ld2];
            dq1[(size_t)i * ldd1] = a < b ? g : (a == b ? 0.5f * g : 0.0f);
            dq2[(size_t)i * ldd2] = b < a ? g : (a == b ? 0.5f * g : 0.0f);
            li = fminf(a, b) * g;
        }
    }
    li = block_sum(li, scratch);
    if (threadIdx.x == 0 && acc) atomicAdd(acc, li);  // -mean(min Q) (or -mean Q1)
}

extern "C" int gymrl_min_q_grad(const float* d_q1, int ld_q1, const float* d_q2, int ld_q2, float* d_dq1, int ld_dq1, float* d_dq2,
                                int ld_dq2, int batch, int q1_only, float* d_acc, void* stream) {
    GYMRL_REQUIRE(d_dq1 && batch > 0 && (q1_only || (d_q1 && d_q2 && d_dq2)), "bad arguments");
    min_q_grad_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(d_q1, ld_q1, d_q2, ld_q2, d_dq1, ld_dq1, d_dq2, ld_dq2, batch,
                                                                           q1_only, d_acc);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("min_q_grad");
    return GYMRL_OK;
}

// SAC actor: given pre-tanh x, noise xi, log_std (pre-clamp), dQpart/da (the critic's input gradient wrt the action
// columns, already scaled by -1/B), write d/d mean and d/d log_std of mean(alpha*logpi - minQ); also sum logpi for alpha.
__global__ void sac_actor_grad_kernel(const float* __restrict__ x, const float* __restrict__ xi, const float* __restrict__ log_std,
                                      int ld, const float* __restrict__ dq_da, int ldg, const double* __restrict__ log_alpha,
                                      float bound, float ls_min, float ls_max, float* __restrict__ dmean, float* __restrict__ dls,
                                      int ldo, const float* __restrict__ logp, float* __restrict__ acc, int B, int A) {
    __shared__ float scratch[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float lsum = 0.f, lp_i = 0.f;
    if (i < B) {
        const float alpha = (float)exp(*log_alpha);
        const float invB = 1.0f / (float)B;
        for (int j = 0; j < A; ++j) {
            const float xv = x[(size_t)i * A + j];
            const float t = tanhf(xv);
            const float omt2 = 1.0f - t * t;
            const float ls_raw = log_std[(size_t)i * ld + j];
            const float pass = (ls_raw >= ls_min && ls_raw <= ls_max) ? 1.0f : 0.0f;
            const float sigma = expf(fminf(fmaxf(ls_raw, ls_min), ls_max));
            // d logpi / dx = d/dx [ -log(bound (1 - tanh^2 x) + 1e-6) ] = 2 bound t (1-t^2) / (bound (1-t^2) + 1e-6)
            const float dlp_dx = 2.0f * bound * t * omt2 / (bound * omt2 + 1e-6f);
            const float dL_dx = alpha * invB * dlp_dx + dq_da[(size_t)i * ldg + j] * bound * omt2;
            dmean[(size_t)i * ldo + j] = dL_dx;
            // log_std: -log(sigma) term gives -1; x = mu + sigma*xi gives sigma*xi; both only inside the clamp range
            dls[(size_t)i * ldo + j] = pass * (alpha * invB * (-1.0f) + dL_dx * sigma * xi[(size_t)i * A + j]);
        }
        lp_i = logp[i];
        lsum = alpha * lp_i * invB;
    }
    lsum = block_sum(lsum, scratch);
    float lps = block_sum(lp_i, scratch);
    if (acc) {
        __shared__ double dscratch[32];
        // acc[0]: alpha * mean(logpi) part of the actor loss; acc[1]: sum(logpi), which the alpha step consumes -> fixed block order
        const double v[2] = {lsum, lps};
        ordered_block_accumulate<2>(v, acc, dscratch);
    }
}

extern "C" int gymrl_sac_actor_grad(const float* d_pre_tanh, const float* d_noise, const float* d_log_std, int ld_log_std,
                                    const float* d_dq_daction, int ld_dq, const double* d_log_alpha, float bound, float log_std_min,
                                    float log_std_max, float* d_dmean, float* d_dlog_std, int ld_out, const float* d_logp,
                                    float* d_acc, int batch, int act_dim, void* stream) {
    GYMRL_REQUIRE(d_pre_tanh && d_noise && d_log_std && d_dq_daction && d_log_alpha && d_dmean && d_dlog_std && d_logp, "NULL pointer");
    GYMRL_REQUIRE(batch > 0 && act_dim > 0, "bad shape");
    sac_actor_grad_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(d_pre_tanh, d_noise, d_log_std, ld_log_std, d_dq_daction,
                                                                               ld_dq, d_log_alpha, bound, log_std_min, log_std_max,
                                                                               d_dmean, d_dlog_std, ld_out, d_logp, d_acc, batch, act_dim);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("sac_actor_grad");
    return GYMRL_OK;
}

// TD3 / DDPG actor: a = tanh(z) * bound; dL/dz = dL/da * bound * (1 - tanh^2 z), with tanh(z) = a / bound
__global__ void tanh_bound_grad_kernel(const float* __restrict__ action, const float* __restrict__ dq_da, int ldg,
                                       float* __restrict__ dz, int ldo, float bound, int B, int A) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * A) return;
    const int i = t / A, j = t % A;
    const float th = action[(size_t)i * A + j] / bound;
    dz[(size_t)i * ldo + j] = dq_da[(size_t)i * ldg + j] * bound * (1.0f - th * th);
}

extern "C" int gymrl_tanh_bound_grad(const float* d_action, const float* d_dq_daction, int ld_dq, float* d_dz, int ld_dz, float bound,
                                     int batch, int act_dim, void* stream) {
    GYMRL_REQUIRE(d_action && d_dq_daction && d_dz && batch > 0 && act_dim > 0, "bad arguments");
    tanh_bound_grad_kernel<<<ceil_div(batch * act_dim, 256), 256, 0, as_stream(stream)>>>(d_action, d_dq_daction, ld_dq, d_dz, ld_dz, bound,
                                                                                         batch, act_dim);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("tanh_bound_grad");
    return GYMRL_OK;
}

// y[i][j] = tanh(z[i][j]) * bound   (TD3 Actor.forward output, td3_pendulum.py:59-62)
__global__ void tanh_bound_kernel(const float* __restrict__ z, int ldz, float* __restrict__ a, float bound, int B, int A) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * A) return;
    a[t] = tanhf(z[(size_t)(t / A) * ldz + (t % A)]) * bound;
}
extern "C" int gymrl_tanh_bound(const float* d_z, int ld_z, float* d_action, float bound, int batch, int act_dim, void* stream) {
    GYMRL_REQUIRE(d_z && d_action && batch > 0 && act_dim > 0, "bad arguments");
    tanh_bound_kernel<<<ceil_div(batch * act_dim, 256), 256, 0, as_stream(stream)>>>(d_z, ld_z, d_action, bound, batch, act_dim);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("tanh_bound");
    return GYMRL_OK;
}

// alpha_loss = -mean(log_alpha * (logpi + Hbar)); d/dlog_alpha = -(mean(logpi) + Hbar); Adam on a float64 scalar.
// state = {exp_avg, exp_avg_sq, step}; acc[1] = sum(logpi) from sac_actor_grad.
__global__ void sac_alpha_step_kernel(double* log_alpha, double* state, const float* acc, int B, double target_entropy, double lr,
                                      float* loss_out) {
    const double mean_lp = (double)acc[1] / (double)B;
    const double g = -(mean_lp + target_entropy);
    if (loss_out) *loss_out = (float)(-(*log_alpha) * (mean_lp + target_entropy));
    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
    double m = state[0], v = state[1];
    const double t = state[2] + 1.0;
    m = m + (1.0 - b1) * (g - m);
    v = v * b2 + (1.0 - b2) * g * g;
    const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
    const double denom = sqrt(v) / sqrt(bc2) + eps;
    *log_alpha = *log_alpha + (-(lr / bc1) * m) / denom;
    state[0] = m; state[1] = v; state[2] = t;
}

extern "C" int gymrl_sac_alpha_step(double* d_log_alpha, double* d_adam_state, const float* d_acc, int batch, double target_entropy,
                                    double lr, float* d_loss_out, void* stream) {
    GYMRL_REQUIRE(d_log_alpha && d_adam_state && d_acc && batch > 0, "bad arguments");
    sac_alpha_step_kernel<<<1, 1, 0, as_stream(stream)>>>(d_log_alpha, d_adam_state, d_acc, batch, target_entropy, lr, d_loss_out);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("sac_alpha_step");
    return GYMRL_OK;
}

// ------------------------------------------------------------------------------------------------ discrete SAC (sac_cartpole)
// softmax policy over A actions: p = softmax(z), lp = log(p + 1e-8), H = -sum p lp   (sac_cartpole.py:96-99, :166-168, :191-193)
__device__ __forceinline__ void softmax_entropy(const float* z, int A, float* p, float* lp, float& H) {
    float mx = z[0];
    for (int j = 1; j < A; ++j) mx = fmaxf(mx, z[j]);
    float s = 0.f;
    for (int j = 0; j < A; ++j) { p[j] = expf(z[j] - mx); s += p[j]; }
    H = 0.f;
    for (int j = 0; j < A; ++j) {
        p[j] = p[j] / s;
        lp[j] = logf(p[j] + 1e-8f);
        H -= p[j] * lp[j];
    }
}

// y = r + gamma (1 - d) (sum_a p_a min(Q1t, Q2t)_a + alpha H(p)),  p = softmax(logits of the actor on s')   (ref :164-176)
__global__ void sacd_target_kernel(const float* __restrict__ logits, int ldz, const float* __restrict__ q1t, int ld1,
                                   const float* __restrict__ q2t, int ld2, const float* __restrict__ reward,
                                   const float* __restrict__ done, const int32_t* __restrict__ row_index,
                                   const float* __restrict__ log_alpha, float gamma, float* __restrict__ y, int B, int A) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const int r = row_index ? row_index[i] : i;
    float p[MAX_A], lp[MAX_A], H;
    softmax_entropy(logits + (size_t)i * ldz, A, p, lp, H);
    float mq = 0.f;
    for (int j = 0; j < A; ++j) mq += p[j] * fminf(q1t[(size_t)i * ld1 + j], q2t[(size_t)i * ld2 + j]);
    const float alpha = expf(*log_alpha);
    const float next_value = mq + alpha * H;
    y[i] = reward[r] + (gamma * (1.0f - done[r])) * next_value;
}

// critic losses mse(Q1(s)[a], y), mse(Q2(s)[a], y) and their gradients (zero for the actions not taken)   (ref :178-189)
__global__ void sacd_critic_loss_kernel(const float* __restrict__ q1, int ld1, const float* __restrict__ q2, int ld2,
                                        const int32_t* __restrict__ action, const int32_t* __restrict__ row_index,
                                        const float* __restrict__ y, float* __restrict__ dq1, int ldd1, float* __restrict__ dq2,
                                        int ldd2, float* __restrict__ loss_acc, int B, int A) {
    __shared__ float scratch[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float l1 = 0.f, l2 = 0.f;
    if (i < B) {
        const int r = row_index ? row_index[i] : i;
        const int a = action[r];
        const float e1 = q1[(size_t)i * ld1 + a] - y[i], e2 = q2[(size_t)i * ld2 + a] - y[i];
        l1 = e1 * e1 / (float)B;
        l2 = e2 * e2 / (float)B;
        for (int j = 0; j < A; ++j) {
            dq1[(size_t)i * ldd1 + j] = j == a ? 2.0f * e1 / (float)B : 0.0f;
            dq2[(size_t)i * ldd2 + j] = j == a ? 2.0f * e2 / (float)B : 0.0f;
        }
    }
    l1 = block_sum(l1, scratch);
    l2 = block_sum(l2, scratch);
    if (threadIdx.x == 0 && loss_acc) { atomicAdd(&loss_acc[0], l1); atomicAdd(&loss_acc[1], l2); }
}

// actor loss mean(-alpha H - sum_a p_a min(Q1, Q2)_a) and its gradient wrt the logits (through softmax and log(p + 1e-8));
// acc[0] += actor loss, acc[1] += sum_i H_i (consumed by the alpha step)   (ref :191-200)
__global__ void sacd_actor_grad_kernel(const float* __restrict__ logits, int ldz, const float* __restrict__ q1, int ld1,
                                       const float* __restrict__ q2, int ld2, const float* __restrict__ log_alpha,
                                       float* __restrict__ dlogits, int lddz, float* __restrict__ acc, int B, int A) {
    __shared__ float scratch[32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float li = 0.f, hi = 0.f;
    if (i < B) {
        float p[MAX_A], lp[MAX_A], H;
        softmax_entropy(logits + (size_t)i * ldz, A, p, lp, H);
        const float alpha = expf(*log_alpha);
        float g[MAX_A], mq = 0.f, pg = 0.f;
        for (int j = 0; j < A; ++j) {
            const float m = fminf(q1[(size_t)i * ld1 + j], q2[(size_t)i * ld2 + j]);
            mq += p[j] * m;
            g[j] = alpha * (lp[j] + p[j] / (p[j] + 1e-8f)) - m;     // dL_i / dp_j
            pg += p[j] * g[j];
        }
        for (int j = 0; j < A; ++j) dlogits[(size_t)i * lddz + j] = p[j] * (g[j] - pg) / (float)B;
        li = (-alpha * H - mq) / (float)B;
        hi = H;
    }
    li = block_sum(li, scratch);
    hi = block_sum(hi, scratch);
    if (threadIdx.x == 0 && acc) { atomicAdd(&acc[0], li); atomicAdd(&acc[1], hi); }
}

// alpha_loss = mean(exp(log_alpha) (H - target).detach()); one float32 Adam step on log_alpha (state = {exp_avg, exp_avg_sq, step})
__global__ void sacd_alpha_step_kernel(float* log_alpha, float* state, const float* acc, int B, float target_entropy, float lr,
                                       float* loss_out) {
    const float alpha = expf(*log_alpha);
    const float mean_dh = acc[1] / (float)B - target_entropy;
    const float g = alpha * mean_dh;
    if (loss_out) *loss_out = alpha * mean_dh;
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    float m = state[0], v = state[1];
    const float t = state[2] + 1.0f;
    m = m + (1.0f - b1) * (g - m);
    v = v * b2 + (1.0f - b2) * g * g;
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    const float step_size = (float)(-((double)lr / bc1));
    const float denom = sqrtf(v) / (float)sqrt(bc2) + eps;
    *log_alpha = *log_alpha + (step_size * m) / denom;
    state[0] = m; state[1] = v; state[2] = t;
}

extern "C" int gymrl_sac_discrete_target(const float* d_logits_next, int ld_logits, const float* d_q1t, int ld_q1t, const float* d_q2t,
                                         int ld_q2t, const float* d_reward, const float* d_done, const int32_t* d_row_index,
                                         const float* d_log_alpha, float gamma, float* d_y, int batch, int n_actions, void* stream) {
    GYMRL_REQUIRE(d_logits_next && d_q1t && d_q2t && d_reward && d_done && d_log_alpha && d_y && batch > 0, "bad arguments");
    GYMRL_REQUIRE(n_actions >= 1 && n_actions <= MAX_A, "n_actions out of range");
    sacd_target_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(d_logits_next, ld_logits, d_q1t, ld_q1t, d_q2t, ld_q2t, d_reward,
                                                                            d_done, d_row_index, d_log_alpha, gamma, d_y, batch, n_actions);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("sac_discrete_target");
    return GYMRL_OK;
}

extern "C" int gymrl_sac_discrete_critic_loss(const float* d_q1, int ld_q1, const float* d_q2, int ld_q2, const int32_t* d_action,
                                              const int32_t* d_row_index, const float* d_y, float* d_dq1, int ld_dq1, float* d_dq2,
                                              int ld_dq2, float* d_loss_acc, int batch, int n_actions, void* stream) {
    GYMRL_REQUIRE(d_q1 && d_q2 && d_action && d_y && d_dq1 && d_dq2 && batch > 0, "bad arguments");
    GYMRL_REQUIRE(n_actions >= 1 && n_actions <= MAX_A, "n_actions out of range");
    sacd_critic_loss_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(d_q1, ld_q1, d_q2, ld_q2, d_action, d_row_index, d_y, d_dq1,
                                                                                 ld_dq1, d_dq2, ld_dq2, d_loss_acc, batch, n_actions);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("sac_discrete_critic_loss");
    return GYMRL_OK;
}

extern "C" int gymrl_sac_discrete_actor_grad(const float* d_logits, int ld_logits, const float* d_q1, int ld_q1, const float* d_q2,
                                             int ld_q2, const float* d_log_alpha, float* d_dlogits, int ld_dlogits, float* d_acc,
                                             int batch, int n_actions, void* stream) {
    GYMRL_REQUIRE(d_logits && d_q1 && d_q2 && d_log_alpha && d_dlogits && batch > 0, "bad arguments");
    GYMRL_REQUIRE(n_actions >= 1 && n_actions <= MAX_A, "n_actions out of range");
    sacd_actor_grad_kernel<<<ceil_div(batch, 256), 256, 0, as_stream(stream)>>>(d_logits, ld_logits, d_q1, ld_q1, d_q2, ld_q2, d_log_alpha,
                                                                                d_dlogits, ld_dlogits, d_acc, batch, n_actions);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("sac_discrete_actor_grad");
    return GYMRL_OK;
}

extern "C" int gymrl_sac_discrete_alpha_step(float* d_log_alpha, float* d_adam_state, const float* d_acc, int batch, float target_entropy,
                                             float lr, float* d_loss_out, void* stream) {
    GYMRL_REQUIRE(d_log_alpha && d_adam_state && d_acc && batch > 0, "bad arguments");
    sacd_alpha_step_kernel<<<1, 1, 0, as_stream(stream)>>>(d_log_alpha, d_adam_state, d_acc, batch, target_entropy, lr, d_loss_out);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("sac_discrete_alpha_step");
    return GYMRL_OK;
}

// ------------------------------------------------------------------------------------------------ NoisyLinear
// eps = f(xi), f(x) = sign(x) sqrt|x|, xi ~ N(0,1)  (scale_noise, rainbow :76-79); d_xi optional pre-drawn normals
__global__ void noisy_sample_kernel(float* __restrict__ eps, const float* __restrict__ xi_in, int n, uint64_t seed, uint64_t entity,
                                    uint32_t draw, const uint32_t* __restrict__ draw_base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (draw_base) draw += *draw_base;
    float xi;
    if (xi_in) xi = xi_in[i];
    else {
        const u32x4 r = philox_draw(seed, entity + (uint64_t)(i >> 1), draw, PHILOX_NOISYNET);
        const float u1 = u01_open0_f32(r.x), u2 = u01_f32(r.y);
        const float rad = sqrtf(-2.0f * logf(u1));
        float s, c;
        sincosf(6.283185307179586f * u2, &s, &c);
        xi = (i & 1) ? rad * s : rad * c;
    }
    const float sg = xi > 0.f ? 1.0f : (xi < 0.f ? -1.0f : 0.0f);
    eps[i] = sg * sqrtf(fabsf(xi));
}
extern "C" int gymrl_noisy_sample(float* d_eps, const float* d_xi, int n, uint64_t seed, uint64_t entity, uint32_t draw,
                                  const uint32_t* d_draw_base, void* stream) {
    GYMRL_REQUIRE(d_eps && n > 0, "bad arguments");
    noisy_sample_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(d_eps, d_xi, n, seed, entity, draw, d_draw_base);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("noisy_sample");
    return GYMRL_OK;
}

// W = mu + sigma * outer(eps_out, eps_in); b = b_mu + b_sigma * eps_out   (NoisyLinear.forward, rainbow :89-97)
__global__ void noisy_compose_kernel(const float* __restrict__ w_mu, const float* __restrict__ w_sigma,
                                     const float* __restrict__ eps_in, const float* __restrict__ eps_out,
                                     const float* __restrict__ b_mu, const float* __restrict__ b_sigma, float* __restrict__ w,
                                     float* __restrict__ b, int N, int K) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N * K) {
        const int n = t / K, k = t % K;
        w[t] = w_mu[t] + w_sigma[t] * (eps_out[n] * eps_in[k]);
    }
    if (t < N) b[t] = b_mu[t] + b_sigma[t] * eps_out[t];
}
extern "C" int gymrl_noisy_compose(const float* d_w_mu, const float* d_w_sigma, const float* d_eps_in, const float* d_eps_out,
                                   const float* d_b_mu, const float* d_b_sigma, float* d_w, float* d_b, int N, int K, void* stream) {
    GYMRL_REQUIRE(d_w_mu && d_w_sigma && d_eps_in && d_eps_out && d_b_mu && d_b_sigma && d_w && d_b && N > 0 && K > 0, "bad arguments");
    noisy_compose_kernel<<<ceil_div(N * K, 256), 256, 0, as_stream(stream)>>>(d_w_mu, d_w_sigma, d_eps_in, d_eps_out, d_b_mu, d_b_sigma,
                                                                              d_w, d_b, N, K);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("noisy_compose");
    return GYMRL_OK;
}

// backward of the composition: dmu (+)= dW, dsigma (+)= dW * outer(eps_out, eps_in); same for the bias
__global__ void noisy_backward_kernel(const float* __restrict__ dw, const float* __restrict__ db, const float* __restrict__ eps_in,
                                      const float* __restrict__ eps_out, float* __restrict__ dw_mu, float* __restrict__ dw_sigma,
                                      float* __restrict__ db_mu, float* __restrict__ db_sigma, int N, int K, int accumulate) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N * K) {
        const int n = t / K, k = t % K;
        const float g = dw[t], gs = g * (eps_out[n] * eps_in[k]);
        dw_mu[t] = accumulate ? dw_mu[t] + g : g;
        dw_sigma[t] = accumulate ? dw_sigma[t] + gs : gs;
    }
    if (t < N) {
        const float g = db[t], gs = g * eps_out[t];
        db_mu[t] = accumulate ? db_mu[t] + g : g;
        db_sigma[t] = accumulate ? db_sigma[t] + gs : gs;
    }
}
extern "C" int gymrl_noisy_backward(const float* d_dw, const float* d_db, const float* d_eps_in, const float* d_eps_out,
                                    float* d_dw_mu, float* d_dw_sigma, float* d_db_mu, float* d_db_sigma, int N, int K,
                                    int accumulate, void* stream) {
    GYMRL_REQUIRE(d_dw && d_db && d_eps_in && d_eps_out && d_dw_mu && d_dw_sigma && d_db_mu && d_db_sigma && N > 0 && K > 0, "bad arguments");
    noisy_backward_kernel<<<ceil_div(N * K, 256), 256, 0, as_stream(stream)>>>(d_dw, d_db, d_eps_in, d_eps_out, d_dw_mu, d_dw_sigma, d_db_mu,
                                                                               d_db_sigma, N, K, accumulate);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("noisy_backward");
    return GYMRL_OK;
}

// ---- fused NoisyNet refresh: everything reset_noise() + the weight composition do for up to two NoisyLinear layers that share
// their input width (the dueling head: advantage [A][K] and value [1][K]) in ONE single-block launch — the four factorised
// noise vectors (same Philox keys and draw order as gymrl_noisy_sample: entity_base + {0,1,2,3} * entity_stride for in0, out0,
// in1, out1), W = mu + sigma * outer(eps_out, eps_in), b = b_mu + b_sigma * eps_out, and the advance of the device draw counter.
// Replaces 4 x noisy_sample + counter_add + 2 x noisy_compose (7 launches of <= 768 threads per network forward).
struct NoisyLayerArgs {
    const float *w_mu, *w_sigma, *b_mu, *b_sigma;
    float *eps_in, *eps_out, *w, *b;
    int N;
};
__device__ __forceinline__ float noisy_eps_draw(uint64_t seed, uint64_t entity, int i, uint32_t draw) {
    const u32x4 r = philox_draw(seed, entity + (uint64_t)(i >> 1), draw, PHILOX_NOISYNET);
    const float u1 = u01_open0_f32(r.x), u2 = u01_f32(r.y);
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    const float xi = (i & 1) ? rad * sn : rad * cs;
    const float sg = xi > 0.f ? 1.0f : (xi < 0.f ? -1.0f : 0.0f);
    return sg * sqrtf(fabsf(xi));
}
__global__ void __launch_bounds__(256) noisy_refresh_kernel(const NoisyLayerArgs l0, const NoisyLayerArgs l1, int n_layers, int K, int noisy,
                                                            uint64_t seed, uint64_t entity_base, uint64_t entity_stride, uint32_t draw,
                                                            uint32_t* __restrict__ draw_base, int counter_inc) {
    const int t = threadIdx.x;
    if (draw_base) draw += *draw_base;
    for (int l = 0; l < n_layers; ++l) {
        const NoisyLayerArgs& L = l ? l1 : l0;
        const uint64_t e_in = entity_base + (uint64_t)(2 * l) * entity_stride, e_out = entity_base + (uint64_t)(2 * l + 1) * entity_stride;
        for (int i = t; i < K; i += blockDim.x) L.eps_in[i] = noisy ? noisy_eps_draw(seed, e_in, i, draw) : 0.0f;
        for (int i = t; i < L.N; i += blockDim.x) L.eps_out[i] = noisy ? noisy_eps_draw(seed, e_out, i, draw) : 0.0f;
    }
    __syncthreads();
    for (int l = 0; l < n_layers; ++l) {
        const NoisyLayerArgs& L = l ? l1 : l0;
        for (int x = t; x < L.N * K; x += blockDim.x) {
            const int n = x / K, k = x % K;
            L.w[x] = L.w_mu[x] + L.w_sigma[x] * (L.eps_out[n] * L.eps_in[k]);
        }
        for (int n = t; n < L.N; n += blockDim.x) L.b[n] = L.b_mu[n] + L.b_sigma[n] * L.eps_out[n];
    }
    if (t == 0 && draw_base && counter_inc) *draw_base += (uint32_t)counter_inc;   // every thread read it before the barrier above
}
extern "C" int gymrl_noisy_refresh(const float* d_w_mu0, const float* d_w_sigma0, const float* d_b_mu0, const float* d_b_sigma0,
                                   float* d_eps_in0, float* d_eps_out0, float* d_w0, float* d_b0, int N0, const float* d_w_mu1,
                                   const float* d_w_sigma1, const float* d_b_mu1, const float* d_b_sigma1, float* d_eps_in1,
                                   float* d_eps_out1, float* d_w1, float* d_b1, int N1, int K, int noisy, uint64_t seed,
                                   uint64_t entity_base, uint64_t entity_stride, uint32_t draw, uint32_t* d_draw_base, int counter_inc,
                                   void* stream) {
    GYMRL_REQUIRE(d_w_mu0 && d_w_sigma0 && d_b_mu0 && d_b_sigma0 && d_eps_in0 && d_eps_out0 && d_w0 && d_b0 && N0 > 0 && K > 0, "bad layer 0");
    const int n_layers = d_w_mu1 ? 2 : 1;
    if (n_layers == 2)
        GYMRL_REQUIRE(d_w_sigma1 && d_b_mu1 && d_b_sigma1 && d_eps_in1 && d_eps_out1 && d_w1 && d_b1 && N1 > 0, "bad layer 1");
    NoisyLayerArgs l0{d_w_mu0, d_w_sigma0, d_b_mu0, d_b_sigma0, d_eps_in0, d_eps_out0, d_w0, d_b0, N0};
    NoisyLayerArgs l1{d_w_mu1, d_w_sigma1, d_b_mu1, d_b_sigma1, d_eps_in1, d_eps_out1, d_w1, d_b1, n_layers == 2 ? N1 : 0};
    noisy_refresh_kernel<<<1, 256, 0, as_stream(stream)>>>(l0, l1, n_layers, K, noisy, seed, entity_base, entity_stride, draw, d_draw_base,
                                                          counter_inc);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("noisy_refresh");
    return GYMRL_OK;
}

// the backward of the composition for the same pair of layers in one launch (blockIdx.y = layer)
struct NoisyBwdArgs {
    const float *dw, *db, *eps_in, *eps_out;
    float *dw_mu, *dw_sigma, *db_mu, *db_sigma;
    int N;
};
__global__ void noisy_backward2_kernel(const NoisyBwdArgs l0, const NoisyBwdArgs l1, int K, int accumulate) {
    const NoisyBwdArgs& L = blockIdx.y ? l1 : l0;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < L.N * K) {
        const int n = t / K, k = t % K;
        const float g = L.dw[t], gs = g * (L.eps_out[n] * L.eps_in[k]);
        L.dw_mu[t] = accumulate ? L.dw_mu[t] + g : g;
        L.dw_sigma[t] = accumulate ? L.dw_sigma[t] + gs : gs;
    }
    if (t < L.N) {
        const float g = L.db[t], gs = g * L.eps_out[t];
        L.db_mu[t] = accumulate ? L.db_mu[t] + g : g;
        L.db_sigma[t] = accumulate ? L.db_sigma[t] + gs : gs;
    }
}
extern "C" int gymrl_noisy_backward2(const float* d_dw0, const float* d_db0, const float* d_eps_in0, const float* d_eps_out0,
                                     float* d_dw_mu0, float* d_dw_sigma0, float* d_db_mu0, float* d_db_sigma0, int N0, const float* d_dw1,
                                     const float* d_db1, const float* d_eps_in1, const float* d_eps_out1, float* d_dw_mu1,
                                     float* d_dw_sigma1, float* d_db_mu1, float* d_db_sigma1, int N1, int K, int accumulate, void* stream) {
    GYMRL_REQUIRE(d_dw0 && d_db0 && d_eps_in0 && d_eps_out0 && d_dw_mu0 && d_dw_sigma0 && d_db_mu0 && d_db_sigma0 && N0 > 0 && K > 0, "bad layer 0");
    GYMRL_REQUIRE(d_dw1 && d_db1 && d_eps_in1 && d_eps_out1 && d_dw_mu1 && d_dw_sigma1 && d_db_mu1 && d_db_sigma1 && N1 > 0, "bad layer 1");
    NoisyBwdArgs l0{d_dw0, d_db0, d_eps_in0, d_eps_out0, d_dw_mu0, d_dw_sigma0, d_db_mu0, d_db_sigma0, N0};
    NoisyBwdArgs l1{d_dw1, d_db1, d_eps_in1, d_eps_out1, d_dw_mu1, d_dw_sigma1, d_db_mu1, d_db_sigma1, N1};
    const int nmax = N0 > N1 ? N0 : N1;
    noisy_backward2_kernel<<<dim3(ceil_div(nmax * K, 256), 2), 256, 0, as_stream(stream)>>>(l0, l1, K, accumulate);
    gymrl_count_launch();
    GYMRL_LAUNCH_CHECK("noisy_backward2");
    return GYMRL_OK;
}
