/* gymrl.h — C ABI of libgymrl_b200.so: the B200-native (sm_100a) vectorised rollout + update path
 * that sits underneath gymRL's Python trainer surface.
 *
 * The reference (Starlight0798/gymRL) is pure Python and has no FFI of its own, so
 * each entry point below cites the reference *Python* code whose arithmetic it replaces
 * (paths relative to the reference repo root).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer named d_* / documented "device" is a CUDA device pointer owned by the caller
 *    (the Python side passes torch tensors' data_ptr()); nothing here allocates after *_create;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises
 *    unless documented, and is CUDA-graph-capture safe;
 *  - return value: 0 (GYMRL_OK) or a negative GYMRL_E* code; the message is available from
 *    gymrl_last_error() (thread-local).  Nothing throws across the ABI;
 *  - a handle (gymrl_env, gymrl_per) is confined to one host thread at a time; distinct handles may
 *    be driven from distinct threads/processes (one process per GPU);
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails with GYMRL_ECUDA.
 */
#ifndef GYMRL_H
#define GYMRL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GYMRL_ABI_VERSION 1

#define GYMRL_OK 0
#define GYMRL_EINVAL (-1)
#define GYMRL_ECUDA (-2)
#define GYMRL_ENOMEM (-3)

int gymrl_version(void);
const char* gymrl_last_error(void);
/* Number of kernels this library has launched since load (bench.py's `gpu_launches` claim). */
uint64_t gymrl_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Environments.  Replaces gym.make(...).reset()/.step() at the call sites
 * algorithms/dqn_cartpole.py:174,180; rainbow_dqn_cartpole.py:367,373; sac_pendulum.py:273,280;
 * td3_pendulum.py:234,241; ppo_lunarlander.py:200,211,222; ppo_full_lunarlander.py:466,478,496.
 * The env arithmetic itself lives in third-party gymnasium/box2d-py (absent from the reference
 * tree); oracle/ restates it ("parity unpinned").
 *
 * N env copies are stepped in lockstep.  obs is [N][D] float32 row-major (D = 4 / 3 / 8), so a
 * rollout buffer [T][N][D] is written with fully coalesced 128 B lines and a minibatch row gather
 * touches one 32 B sector (D = 8).  Discrete actions are int32[N]; continuous are float32[N][A].
 * Auto-reset: when an env terminates or truncates, `obs` receives the first observation of the next
 * episode while `next_obs` (nullable) always receives the true post-step observation (what the
 * off-policy trainers store as next_state).
 * ---------------------------------------------------------------------------------------------- */
typedef struct gymrl_env gymrl_env;

#define GYMRL_ENV_CARTPOLE 0    /* CartPole-v1   D=4 A=2 (discrete) TimeLimit 500  */
#define GYMRL_ENV_PENDULUM 1    /* Pendulum-v1   D=3 A=1 (bound 2)  TimeLimit 200  */
#define GYMRL_ENV_LUNARLANDER 2 /* LunarLander-v3 D=8 A=4 (discrete) TimeLimit 1000 */

int gymrl_env_info(int kind, int* obs_dim, int* act_dim, int* n_actions, int* max_episode_steps,
                   float* action_bound, int* state_doubles);
/* Global env ids [first_env_id, first_env_id + n_envs) key the Philox streams, so a shard's
 * trajectories do not depend on how many GPUs the batch is split over. */
int gymrl_env_create(gymrl_env** out, int kind, int n_envs, uint64_t seed, uint64_t first_env_id);
int gymrl_env_destroy(gymrl_env* env);
/* mask: device uint8[N] (nullable = all).  Masked envs start a new episode. */
int gymrl_env_reset(gymrl_env* env, const uint8_t* d_mask, float* d_obs, void* stream);
/* d_done (nullable) receives terminated | truncated — the `done` flag every algorithms/ script stores
 * (SURVEY q11). */
int gymrl_env_step(gymrl_env* env, const void* d_actions, float* d_obs, float* d_next_obs,
                   float* d_reward, uint8_t* d_terminated, uint8_t* d_truncated, uint8_t* d_done,
                   void* stream);
/* Raw physics state snapshot, [N][state_doubles] float64 — used by the parity tests to
 * teacher-force the device env from the oracle (and vice versa). */
int gymrl_env_get_state(gymrl_env* env, double* d_state, void* stream);
int gymrl_env_set_state(gymrl_env* env, const double* d_state, void* stream);
/* Mean return/length over the last `last_k` finished episodes (all envs).  SYNCHRONOUS (one small
 * D2H) — call at log time only.  Replaces the deque(maxlen=100) bookkeeping at
 * algorithms/ppo_lunarlander.py:172,219-221. */
int gymrl_env_episode_stats(gymrl_env* env, int last_k, double* mean_return, double* mean_length,
                            uint64_t* total_episodes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Action selection
 * ---------------------------------------------------------------------------------------------- */
/* Categorical(logits).sample() == argmax_j softmax(logits)_j / E_j, E~Exp(1)   (SURVEY q3;
 * ActorCritic.get_action, algorithms/ppo_lunarlander.py:92-104; ppo_full_lunarlander.py:395-407).
 * d_noise: optional pre-drawn Exp(1) [N][A] (parity tests); NULL draws from Philox
 * (seed, first_id + i, draw + *d_draw_base); d_draw_base (nullable) is a device-resident counter so a
 * captured CUDA graph draws fresh noise on every replay (same convention for every RNG entry point).  logp = log_softmax(logits)[action]; entropy = -sum p log p.
 * deterministic != 0 -> action = argmax logits. */
/* d_value_in/ld_value_in -> d_value_out (nullable): copies the critic column of the fused head output
 * into the contiguous [N] value row of the rollout buffer in the same pass. */
int gymrl_sample_categorical(const float* d_logits, int ld_logits, const float* d_noise,
                             int32_t* d_action, float* d_logp, float* d_entropy,
                             const float* d_value_in, int ld_value_in, float* d_value_out, int n, int n_actions,
                             uint64_t seed, uint64_t first_id, uint32_t draw, const uint32_t* d_draw_base,
                             int deterministic, void* stream);
/* epsilon-greedy over Q-values (DQNTrainer.select_action, algorithms/dqn_cartpole.py:124-133).
 * Per env: u ~ U[0,1); if u < eps: uniform random action else argmax (first max). */
int gymrl_select_eps_greedy(const float* d_q, int ld_q, int32_t* d_action, int n, int n_actions,
                            float eps, uint64_t seed, uint64_t first_id, uint32_t draw,
                            const uint32_t* d_draw_base, void* stream);
/* tanh-Gaussian policy (Actor.sample / get_action, algorithms/sac_pendulum.py:76-98):
 * x = mean + exp(clamp(log_std)) * xi; a = tanh(x) * bound;
 * logp = sum_j [ N(x; mean, std).log_prob - log(bound * (1 - tanh(x)^2) + 1e-6) ].
 * d_noise: optional pre-drawn N(0,1) [N][A]; d_logp / d_pre_tanh nullable.  deterministic -> tanh(mean)*bound. */
int gymrl_sample_tanh_gaussian(const float* d_mean, const float* d_log_std, int ld, const float* d_noise,
                               float* d_action, float* d_logp, float* d_pre_tanh, int n, int act_dim,
                               float bound, float log_std_min, float log_std_max, uint64_t seed,
                               uint64_t first_id, uint32_t draw, const uint32_t* d_draw_base,
                               int deterministic, void* stream);
/* a = clip(mu + sigma * xi, -bound, bound)   (TD3Trainer.select_action, algorithms/td3_pendulum.py:157-170;
 * also the target smoothing noise at :194-204 with noise_clip > 0: a = clip(mu + clip(sigma*xi, +-noise_clip), +-bound)). */
int gymrl_add_gaussian_noise_clip(const float* d_mu, const float* d_noise, float* d_action, int n,
                                  int act_dim, float sigma, float noise_clip, float bound, uint64_t seed,
                                  uint64_t first_id, uint32_t draw, const uint32_t* d_draw_base, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GAE / returns
 * dialect 0 ("algorithms"): PPOTrainer.compute_gae, algorithms/ppo_lunarlander.py:179-196: done masks
 *   bootstrap and trace, V_{T} = v_last[N].  float64 recurrence; `dones` is a float32 array there, so under
 *   NumPy >= 2 the trace coefficient gamma*lam*(1-d) is rounded to float32 before use — reproduced.
 * dialect 2 ("ppo_full"): compute_advantages, ppo_full_lunarlander.py:507-535: same masking, decoupled
 *   lam_actor / lam_critic (ret = A(lam_critic) + V), float64 coefficient, but `values` are 0-dim float32
 *   tensors so gamma*V_{t+1} is a float32 product — reproduced.
 * dialect 1 ("utils"): ReplayBuffer_on_policy.compute_advantage, utils/buffer.py:21-35:
 *   bootstrap masked by dw (terminated), trace by done, per-step next values v_next[T][N].
 * All arrays are [T][N] (time-major); fp32 in/out, fp64 internal recurrence (the reference runs
 * this loop in fp64).
 * ---------------------------------------------------------------------------------------------- */
int gymrl_gae(const float* d_reward, const float* d_value, const float* d_v_last_or_next,
              const uint8_t* d_done, const uint8_t* d_dw, float* d_adv, float* d_ret, int T, int N,
              double gamma, double lam_actor, double lam_critic, int dialect, void* stream);

/* sums[0] += sum(x), sums[1] += sum(x^2)  (float64 accumulators on device; zero them first).
 * Multi-GPU runs all-reduce `sums` (+count) before normalising. */
int gymrl_sum_sumsq(const float* d_x, long long n, double* d_sums, void* stream);
/* x = (x - mean) / (std + eps) with mean/std from sums, count; ddof 0 = numpy
 * (algorithms/ppo_lunarlander.py:236), ddof 1 = torch (utils/buffer.py:33). */
int gymrl_normalize_inplace(float* d_x, long long n, const double* d_sums, double count, int ddof,
                            float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PPO losses: forward value + analytic gradient wrt logits and V in one pass.
 * ---------------------------------------------------------------------------------------------- */
#define GYMRL_PPO_DUALCLIP 0   /* algorithms/ppo_lunarlander.py:278-300                         */
#define GYMRL_PPO_FULL 1       /* algorithms/ppo_full_lunarlander.py:586-633 (ERC mask, clip-higher) */
#define GYMRL_PPO_VALUE_CLIP 4 /* flag, OR-ed in: algorithms/ppo_lstm_lunarlander.py:763-771    */

typedef struct gymrl_ppo_cfg {
    int mode;            /* GYMRL_PPO_* (| GYMRL_PPO_VALUE_CLIP)                       */
    float clip_eps_min;  /* lower clip: ratio >= 1 - clip_eps_min                      */
    float clip_eps_max;  /* upper clip: ratio <= 1 + clip_eps_max                      */
    float dual_clip;     /* 3.0                                                        */
    float value_coef;    /* 0.5                                                        */
    float entropy_coef;  /* 0.01 (ppo_full anneals it on the host)                     */
    float erc_low;       /* ERC: 1 - erc_low < H_new/(H_old+1e-8) < 1 + erc_high       */
    float erc_high;
    float vclip_eps_min; /* value-clip window (VALUE_CLIP flag)                        */
    float vclip_eps_max;
} gymrl_ppo_cfg;

/* metrics (device float[8], accumulated with += so zero before the first minibatch):
 * [0] policy_loss [1] value_loss (incl. coef) [2] entropy mean [3] clip_frac [4] approx_kl
 * [5] erc-clipped fraction [6] total loss [7] #minibatches accumulated.
 * row_index (nullable) gathers action/logp_old/adv/ret/(H_old,V_old) rows — the minibatch
 * permutation of ppo_lunarlander.py:262-272 — so no gathered copies are materialised. */
int gymrl_ppo_loss(const float* d_logits, int ld_logits, const float* d_value, int ld_value,
                   const int32_t* d_row_index, const int32_t* d_action, const float* d_logp_old,
                   const float* d_adv, const float* d_ret, const float* d_entropy_old,
                   const float* d_value_old, float* d_dlogits, int ld_dlogits, float* d_dvalue,
                   int ld_dvalue, float* d_metrics, int batch, int n_actions, const gymrl_ppo_cfg* cfg,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense layers: y = act(x W^T + b) with torch.nn.Linear's [out][in] weight layout.
 * fp32 storage and fp32-accurate accumulation.  `row_index` fuses the minibatch row gather into
 * the operand load (replaces states[mb_indices], ppo_lunarlander.py:268).
 * ---------------------------------------------------------------------------------------------- */
#define GYMRL_ACT_NONE 0
#define GYMRL_ACT_TANH 1
#define GYMRL_ACT_RELU 2

int gymrl_linear_forward(const float* d_x, int ldx, const int32_t* d_row_index, const float* d_w,
                         const float* d_b, float* d_y, int ldy, int M, int N, int K, int act,
                         void* stream);
/* dX[M][K] = (dY[M][N] W[N][K]) * act'(h_in) where h_in is the OUTPUT of the previous layer's
 * activation (tanh' = 1-h^2, relu' = h>0); h_in nullable.  accumulate != 0 adds into dX. */
int gymrl_linear_backward_input(const float* d_dy, int lddy, const float* d_w, const float* d_h_in,
                                int ldh, float* d_dx, int lddx, int M, int N, int K, int act_in,
                                int accumulate, void* stream);
/* dW[N][K] = dY^T X, db[N] = colsum(dY) (db nullable).  Deterministic split-M partial sums go
 * through `d_workspace` (>= gymrl_linear_backward_weight_workspace(M,N,K) bytes). */
size_t gymrl_linear_backward_weight_workspace(int M, int N, int K);
int gymrl_linear_backward_weight(const float* d_dy, int lddy, const float* d_x, int ldx,
                                 const int32_t* d_row_index, float* d_dw, float* d_db, int M, int N,
                                 int K, int accumulate, void* d_workspace, size_t workspace_bytes,
                                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser, clipping, target sync
 * ---------------------------------------------------------------------------------------------- */
/* *d_sumsq += sum(g^2) over a flat fp32 gradient buffer (float64 accumulator; zero it first). */
int gymrl_grad_sumsq(const float* d_grad, long long n, double* d_sumsq, void* stream);
/* torch.optim.Adam (no amsgrad / weight decay) on flat buffers, fused with
 *  - clip_grad_norm_(max_norm): g *= min(1, max_norm / (sqrt(*d_sumsq) + 1e-6)) when d_sumsq != NULL
 *    (algorithms/ppo_lunarlander.py:304-306, rainbow_dqn_cartpole.py:344), and/or
 *  - per-element clamp to [-clamp, clamp] when clamp > 0 (algorithms/dqn_cartpole.py:163-165).
 * d_lr: device float64 (host anneals it without re-capturing graphs); d_step: device int32, incremented
 * by the kernel (bias correction uses the incremented value, as torch does). grad_scale multiplies
 * g first (1/world_size after a sum all-reduce). */
int gymrl_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                    long long n, const double* d_lr, float beta1, float beta2, float eps,
                    int32_t* d_step, const double* d_sumsq, float max_norm, float clamp,
                    float grad_scale, void* stream);
/* target = tau * source + (1 - tau) * target   (rainbow_dqn_cartpole.py:347-352,
 * sac_pendulum.py:194-199, td3_pendulum.py:150-155); tau = 1 is the hard copy of dqn_cartpole.py:193. */
int gymrl_polyak(float* d_target, const float* d_source, long long n, float tau, void* stream);

/* perm[i] = a uniformly random permutation of [0,n) keyed by (seed, draw) — replaces
 * np.random.shuffle(indices) at algorithms/ppo_lunarlander.py:262 (bijective Feistel + cycle walk). */
int gymrl_random_permutation(int32_t* d_perm, int n, uint64_t seed, uint32_t draw,
                             const uint32_t* d_draw_base, void* stream);
/* *d_counter += inc (device-resident draw / minibatch counters so that captured CUDA graphs advance
 * their RNG streams on every replay). */
int gymrl_counter_add(uint32_t* d_counter, uint32_t inc, void* stream);
/* dst[i] = src[(*d_block_index) * n + i], i < n: the minibatch window indices[start:end] of
 * algorithms/ppo_lunarlander.py:264-266 with the window number resident on the device. */
int gymrl_slice_i32(int32_t* d_dst, const int32_t* d_src, int n, const uint32_t* d_block_index, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GYMRL_H */
