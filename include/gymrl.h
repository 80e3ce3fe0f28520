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
/* Diagnostic (LunarLander only; no reference counterpart): d_prof = device int64 [N][8], overwritten by every
 * subsequent step with {step cycles, collide, constraint setup, velocity iterations, position iterations (cycles),
 * touching manifolds, position iterations run, work slot}.  NULL switches it off (the default). */
int gymrl_env_set_profile(gymrl_env* env, long long* d_prof);
/* LunarLander only (no reference counterpart): which arrangement of the constraint-solver loops the step kernel runs.
 * 0 = the oracle's arrangement with the plain division; 2 = the position rows' divisions evaluated branch-free (the default:
 * never slower than 0); 3 = 2 + the velocity loop specialised on the joints' limit states (fastest under a fresh policy, slower
 * once the copies of a warp stop sharing their limit states - see env_lunar.cu).  All give the same results bit for
 * bit (tests/test_gpu_envs.py, tests/test_hostsim_lunar.py); the setter exists for A/B timing and for those tests.  A new env
 * starts with the library default (environment variable GYMRL_LL_SOLVER overrides it). */
int gymrl_env_set_solver(gymrl_env* env, int variant);
int gymrl_env_get_solver(gymrl_env* env, int* variant);
/* LunarLander keeps at most 8 touching manifolds per env copy (the device solver's contact slots; Box2D's contact list is
 * unbounded).  A ninth is dropped — in the CUDA env and in the CPU oracle alike — and counted here: the number of such events
 * since the env was created (synchronises the stream; 0 for the other envs).  Convergence runs report it (observed: 0). */
int gymrl_env_overflow_count(gymrl_env* env, uint64_t* count, void* stream);
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
/* The tail of a PPO rollout step in one launch (ActorCritic.get_action, algorithms/ppo_lunarlander.py:92-104, on the engine's
 * fused head-trunk activations d_h [n][2H] = (actor | critic)): logits = Wa h_a + ba (Wa [A][H]), V = Wc h_c + bc, then
 * gymrl_sample_categorical on those logits — bit-identical to gymrl_linear_forward x 2 + gymrl_sample_categorical (same
 * summation order, same Philox keys).  d_logp / d_entropy / d_value / d_lv_out ([n][8] = logits | V) are nullable. */
int gymrl_policy_heads_sample(const float* d_h, int ldh, const float* d_Wa, const float* d_ba, const float* d_Wc,
                              const float* d_bc, int H, int n_actions, int32_t* d_action, float* d_logp, float* d_entropy,
                              float* d_value, float* d_lv_out, int n, uint64_t seed, uint64_t first_id, uint32_t draw,
                              const uint32_t* d_draw_base, int deterministic, void* stream);
/* epsilon-greedy over Q-values (DQNTrainer.select_action, algorithms/dqn_cartpole.py:124-133).
 * Per env: u ~ U[0,1); if u < eps: uniform random action else argmax (first max). */
int gymrl_select_eps_greedy(const float* d_q, int ld_q, int32_t* d_action, int n, int n_actions,
                            float eps, uint64_t seed, uint64_t first_id, uint32_t draw,
                            const uint32_t* d_draw_base, void* stream);
/* Same, with epsilon read from a device float: the host writes the decayed value (dqn_cartpole.py:117-122) into the slot
 * before replaying a captured lockstep graph. */
int gymrl_select_eps_greedy_dev(const float* d_q, int ld_q, int32_t* d_action, int n, int n_actions, const float* d_eps,
                                uint64_t seed, uint64_t first_id, uint32_t draw, const uint32_t* d_draw_base, void* stream);
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
#define GYMRL_PPO_MASKED_MEAN 8 /* flag (with GYMRL_PPO_FULL): policy / value / entropy / clip_frac are masked_mean()s,
                                 * sum(x * mask) / mask.sum() and 0 when mask.sum() == 0 (ppo_lstm_lunarlander.py:646-655,
                                 * :757-771) instead of ppo_full's plain mean over the minibatch; gymrl_ppo_loss only */

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
    const float* d_entropy_coef; /* nullable device scalar overriding entropy_coef: lets a captured CUDA graph
                                  * follow ppo_full's per-update entropy anneal (:664-666)      */
    float* d_mask_count;         /* MASKED_MEAN: device float the call fills with mask.sum() of the minibatch */
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

/* GEMM engine: 1 (default) = tcgen05 3xTF32 tensor-core kernel (fp32-accurate: every operand split into
 * tf32 hi + lo, three MMAs per product, fp32 accumulation in TMEM) wherever the shape gate allows
 * (M >= 128, N % 64 == 0, K % 32 == 0, 16 B-aligned rows), fp32 FFMA tiles elsewhere; 0 = FFMA tiles only.
 * Also selectable with the environment variable GYMRL_GEMM=ffma. */
int gymrl_set_gemm_mode(int mode);
int gymrl_get_gemm_mode(void);

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

/* Whole backward of one dense layer in one call (what autograd does for nn.Linear, ppo_lunarlander.py:312
 * loss.backward()): dW, db and, when d_dx != NULL, dX = (dY W) * act'(x) where x — the layer input — is the
 * previous layer's activation output.  Policy / value / Q heads (N <= 8, K in {128, 256, 512}) run as ONE
 * bandwidth-bound sweep over x; other shapes are gymrl_linear_backward_weight + gymrl_linear_backward_input.
 * d_row_index (gathered input rows) requires d_dx == NULL.  Same workspace contract as backward_weight. */
int gymrl_linear_backward(const float* d_dy, int lddy, const float* d_x, int ldx,
                          const int32_t* d_row_index, const float* d_w, float* d_dw, float* d_db,
                          float* d_dx, int lddx, int M, int N, int K, int act_in, int accumulate,
                          void* d_workspace, size_t workspace_bytes, void* stream);

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

/* Deferred, merged fold of parameter-gradient partials.  Between gymrl_reduce_defer_begin() and gymrl_reduce_flush()
 * (per host thread) the backward entry points (gymrl_linear_backward, gymrl_linear_backward_weight,
 * gymrl_ppo_heads_fused) record their pending "dW/db = fixed-order sum of per-CTA partials" instead of launching one
 * small fold kernel each; the flush folds everything recorded in ONE launch.  Each recorded call must have been given
 * its own workspace, which has to stay untouched until the flush has executed (at most 24 pending sums per scope).
 * With d_sumsq_partials != NULL the flush also writes, per block, the float64 sum of squares of the final gradient
 * values it produced (*n_partials entries, <= capacity; *n_outputs = how many gradient elements they cover, so the
 * caller can check that the whole flat gradient went through the scope): the global norm of clip_grad_norm_
 * (algorithms/ppo_lunarlander.py:304-306) for gymrl_clip_adam_step, without another pass over the gradient. */
int gymrl_reduce_defer_begin(void);
int gymrl_reduce_flush(double* d_sumsq_partials, int capacity, int* n_partials, long long* n_outputs, void* stream);
/* gymrl_adam_step with clip_grad_norm_(max_norm) taken from gymrl_reduce_flush's partial sums of squares (folded in a
 * fixed order by every block), the step counter advanced by the last block to finish (d_done_counter: device uint32,
 * zero before the first call): one launch for algorithms/ppo_lunarlander.py:304-306. */
int gymrl_clip_adam_step(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, long long n,
                         const double* d_lr, float beta1, float beta2, float eps, int32_t* d_step,
                         const double* d_sumsq_partials, int n_partials, float max_norm, float grad_scale,
                         uint32_t* d_done_counter, void* stream);
/* ---- the one collective of the path (§8b `gymrl_comm_init / gymrl_allreduce_grads`) -----------------------------
 * The reference has no collective (single process).  The env-sharded engine sums the flat fp32 gradient over the ranks of
 * one node before every optimizer step.  gymrl_comm_* does that as a ONE-SHOT reduction over NVLink peer memory (cudaIpc
 * mappings of a staging buffer this library allocates), fused with the sum of squares that clip_grad_norm_
 * (algorithms/ppo_lunarlander.py:304-306) needs: every rank reads all W staging buffers and adds them in rank order, so all
 * ranks hold bit-identical sums; gymrl_clip_adam_step(d_reduced, d_sumsq_partials, gymrl_comm_n_partials(), grad_scale = 1/W)
 * finishes the step.  Host protocol (one process per GPU): create -> exchange every rank's gymrl_comm_get_handle() blob
 * (gymrl_comm_handle_bytes() bytes each, rank order) with any host-side all-gather -> gymrl_comm_open -> any number of
 * gymrl_comm_allreduce_sumsq calls, the same sequence on every rank (asynchronous on `stream`, CUDA-graph capturable: the
 * launch counter lives in device memory).  A peer that never arrives makes the kernel trap after a bounded spin. */
typedef struct gymrl_comm gymrl_comm;
int gymrl_comm_create(gymrl_comm** out, int rank, int world, long long n_floats, int n_blocks /* 0 = default */);
int gymrl_comm_handle_bytes(void);
int gymrl_comm_get_handle(gymrl_comm* comm, void* handle_out);
int gymrl_comm_open(gymrl_comm* comm, const void* handles /* world x handle_bytes, rank order */);
int gymrl_comm_n_partials(const gymrl_comm* comm);
int gymrl_comm_allreduce_sumsq(gymrl_comm* comm, const float* d_grad, float* d_reduced, double* d_sumsq_partials, void* stream);
int gymrl_comm_destroy(gymrl_comm* comm);

/* Pre-split tf32 images of weight matrices (csrc/wimages.cu): the TMA-fed B operand of the warp-specialised tensor-core GEMM.
 * d_images: 4 * n_floats floats [hi | lo | hi^T | lo^T]; mats: host int[n_mats][3] = {offset, rows, cols} of the matrices
 * (torch.nn.Linear [out][in] layout, offsets multiples of 4) whose dense layers should take that path.  After registration
 * gymrl_linear_forward / gymrl_linear_backward_input recognise those matrices by address + shape; gymrl_adam_step,
 * gymrl_clip_adam_step and gymrl_polyak re-split the images of a registered buffer after writing it.  Call
 * gymrl_weight_images_refresh after any other write to the parameters (initialisation, load_state_dict, broadcast) and
 * gymrl_weight_images_unregister before freeing the buffer. */
int gymrl_weight_images_register(const float* d_flat, long long n_floats, float* d_images, const int* mats, int n_mats);
int gymrl_weight_images_refresh(const float* d_flat, void* stream);
int gymrl_weight_images_unregister(const float* d_flat);

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

/* ------------------------------------------------------------------------------------------------
 * Replay memory.  All storage is caller-owned (torch tensors); the ring's
 * {cursor, size} live in a device int32[2] so stores/samples are CUDA-graph capturable.
 * ---------------------------------------------------------------------------------------------- */
/* idx[i] = i-th image of a keyed random bijection on [0, size): a uniform sample WITHOUT replacement
 * (random.sample(self.buffer, batch_size), algorithms/dqn_cartpole.py:75-77; same in sac/td3/ddpg). */
int gymrl_replay_sample_indices(int32_t* d_idx, int batch, const int32_t* d_ring_state, uint64_t seed,
                                uint32_t draw, const uint32_t* d_draw_base, void* stream);
/* dst[(cursor + i) % capacity][0:width] = src[i][0:width] for one field (4-byte elements, or uint8 -> float32
 * when src_is_u8) — ReplayBuffer.push (dqn_cartpole.py:72-73) for N lockstep transitions at once. */
int gymrl_replay_store(void* d_dst, const void* d_src, int n, int width, int src_is_u8, int capacity,
                       const int32_t* d_ring_state, void* stream);
/* cursor = (cursor + n) % capacity; size = min(capacity, size + n)  (deque(maxlen=capacity) semantics). */
int gymrl_replay_advance(int32_t* d_ring_state, int n, int capacity, void* stream);
/* dst[i] = [a[idx_a[i]][0:width_a], b[idx_b[i]][0:width_b]]: batch gather fused with
 * torch.cat([state, action], dim=1) (Critic.forward, algorithms/sac_pendulum.py:112). idx_*, b nullable. */
int gymrl_gather_concat(void* d_dst, int ld_dst, const void* d_a, int width_a, int ld_a, const int32_t* d_idx_a,
                        const void* d_b, int width_b, int ld_b, const int32_t* d_idx_b, int n, void* stream);
/* PrioritizedNStepBuffer.store_transition + _get_n_step_transition (rainbow_dqn_cartpole.py:179-218) for N
 * envs in lockstep: append to each env's n-slot window (planes [n][N][...]), and once the window is full
 * fold R = sum gamma^k r_k back-to-front (cut at done; s'/terminal from the earliest done) and write the
 * n-step transition into ring row (cursor + env).  Caller then calls gymrl_sumtree_store_new +
 * gymrl_replay_advance iff the window was full (pushed + 1 >= n_steps, known on the host).
 * d_trunc (nullable): time-limit flags; the stored terminal is d_term & !d_trunc (rainbow :376).  d_pushed: int32[2] =
 * {pushes so far, scratch}, both zero at the start; the call advances the count itself. */
int gymrl_nstep_push(float* w_obs, int32_t* w_act, float* w_rew, float* w_nobs, uint8_t* w_term, uint8_t* w_done,
                     const float* d_obs, const int32_t* d_act, const float* d_rew, const float* d_nobs,
                     const uint8_t* d_term, const uint8_t* d_trunc, const uint8_t* d_done, int n_envs, int obs_dim, int n_steps,
                     double gamma, int32_t* d_pushed, float* r_obs, int32_t* r_act, float* r_rew, float* r_nobs, float* r_term,
                     int capacity, const int32_t* d_ring_state, void* stream);
/* ReplayBuffer.push for a lockstep of n transitions (dqn_cartpole.py:75-76, sac_pendulum.py:140-141) in one launch: all five
 * fields at ring rows (cursor + i) % capacity, then {cursor, size} advanced (what gymrl_replay_store x 5 + gymrl_replay_advance
 * do in six).  action elements are 4 bytes (int32 or float32); d_done_ctr: one word, zero before the first call (left zero). */
int gymrl_replay_store_all(float* r_obs, float* r_next_obs, void* r_action, float* r_reward, float* r_done,
                           const float* d_obs, const float* d_next_obs, const void* d_action, const float* d_reward,
                           const uint8_t* d_done, int n, int obs_dim, int act_width, int capacity, int32_t* d_ring_state,
                           uint32_t* d_done_ctr, void* stream);
/* SumTree (rainbow_dqn_cartpole.py:116-152): float64 binary heap of 2*capacity-1 nodes, leaf i at
 * capacity-1+i — identical layout and tie rule, so results match the reference for any capacity.
 * update: batch of (data index, priority) with last-writer-wins for duplicates (update_priorities :258-261);
 * pass d_priority (float64) OR d_td_error (float32): priority = min(|td| + eps, clip_max?)^alpha evaluated in
 * float32 like the reference's NumPy expression (clip_max <= 0 disables the min; ddqn_per_cartpole.py:142-147
 * uses eps 1e-4, clip 1, alpha 0.6).
 * Every internal node is kept exactly fl(left + right): a batch = mark (last writer per leaf, touched leaves per subtree) +
 * set (the winners write their leaves; the warp that completes a 256-leaf subtree recomputes it; the last block rebuilds the
 * top 13 levels in shared memory) — no atomics on the tree, bitwise reproducible; against the reference's
 * `tree[parent] += change` only the summation order differs.
 * d_winner_scratch: int32[gymrl_sumtree_scratch_ints(capacity)], the first `capacity` entries -1 and the rest 0 before the
 * first call (every call leaves it so). */
int gymrl_sumtree_scratch_ints(int capacity);
int gymrl_sumtree_update(double* d_tree, int capacity, const int32_t* d_idx, const double* d_priority,
                         const float* d_td_error, int n, float eps, float alpha, float clip_max,
                         int32_t* d_winner_scratch, void* stream);
/* New transitions at ring rows [cursor, cursor+n) get priority max(all leaves) (1.0 while the tree is empty):
 * rainbow :201-202 with the O(capacity) max scan done as a device reduction.
 * d_max_scratch: one float64, zero before the first call (the call leaves it zero); d_winner_scratch as above. */
int gymrl_sumtree_store_new(double* d_tree, int capacity, int n, const int32_t* d_ring_state, double* d_max_scratch,
                            int32_t* d_winner_scratch, void* stream);
/* PrioritizedNStepBuffer.sample (rainbow :220-256): stratified v_i ~ U(seg*i, seg*(i+1)), root->leaf descent,
 * is_weight = (size * p/total)^(-beta) / max.  d_uniforms: optional pre-drawn U[0,1) float64[B] (parity).
 * return_tree_index bit 0: return tree indices like dialect B (ddqn_per_cartpole.py:94-106); bit 1: d_uniforms holds
 * raw prefix values v (a batch of SumTree.get_index(v) calls).  d_scratch_u32: two words, zero before the first call
 * (left zero): the running max of the weights and the count of finished blocks — the last block normalises. */
int gymrl_sumtree_sample(const double* d_tree, int capacity, int batch, const double* d_uniforms,
                         const int32_t* d_ring_state, const double* d_beta, int32_t* d_out_idx, float* d_out_is_weight,
                         double* d_out_priority, uint32_t* d_scratch_u32, int return_tree_index, uint64_t seed,
                         uint32_t draw, const uint32_t* d_draw_base, void* stream);

/* ------------------------------------------------------------------------------------------------
 * TD targets / losses.  Each writes dL/d(network outputs); the dense-layer backward
 * kernels take it from there.  Loss accumulators are device float[2] = {sum of losses, #calls}.
 * ---------------------------------------------------------------------------------------------- */
/* DQN (algorithms/dqn_cartpole.py:149-157): y = r + gamma*max_a' Qt(s')*(1-d), mse.  Pass d_qnext_online for
 * double-Q (rainbow_dqn_cartpole.py:319-338: a* = argmax online(s'), y = R + gamma^n (1-terminal) Qt(s')[a*],
 * loss = mean(w * td^2), td exported).  Pass the d_v* value streams for the dueling head
 * Q = V + A - mean(A) (rainbow :108-113); then d_q* are the advantage streams and d_dq/d_dv their gradients. */
int gymrl_dqn_loss(const float* d_q, int ld_q, const float* d_v, int ld_v, const float* d_qnext_target, int ld_qt,
                   const float* d_vnext_target, int ld_vt, const float* d_qnext_online, int ld_qo,
                   const float* d_vnext_online, int ld_vo, const int32_t* d_row_index, const int32_t* d_action,
                   const float* d_reward, const float* d_done, const float* d_is_weight, float* d_dq, int ld_dq,
                   float* d_dv, int ld_dv, float* d_td_error, float* d_loss_acc, int batch, int n_actions,
                   float gamma_n, void* stream);
/* y = r + gamma (1-d) (min(Q1t, Q2t) - alpha logpi')   (sac_pendulum.py:233-237; logp_next NULL -> TD3 :200-204) */
int gymrl_twin_q_target(const float* d_reward, const float* d_done, const int32_t* d_row_index, const float* d_q1t,
                        int ld_q1t, const float* d_q2t, int ld_q2t, const float* d_logp_next, const double* d_log_alpha,
                        float gamma, float* d_y, int batch, void* stream);
/* mse(q1, y) + mse(q2, y) and its gradients (sac :239-242, td3 :206-209) */
int gymrl_twin_q_loss(const float* d_q1, int ld_q1, const float* d_q2, int ld_q2, const float* d_y, float* d_dq1,
                      int ld_dq1, float* d_dq2, int ld_dq2, float* d_loss_acc, int batch, void* stream);
/* d/dq of -mean(min(q1, q2)) (sac :250-251, ties split like torch.min) or of -mean(q1) (q1_only; td3 :216);
 * *d_acc += that loss term (nullable). */
int gymrl_min_q_grad(const float* d_q1, int ld_q1, const float* d_q2, int ld_q2, float* d_dq1, int ld_dq1, float* d_dq2,
                     int ld_dq2, int batch, int q1_only, float* d_acc, void* stream);
/* SAC actor (sac :76-87, :248-251): gradient of mean(alpha*logpi - minQ) wrt (mean, log_std) through
 * x = mean + exp(clamp(log_std))*xi, a = tanh(x)*bound, given d(-minQ/B)/da from the critic's input gradient.
 * d_acc[0] += alpha*mean(logpi), d_acc[1] += sum(logpi) (consumed by gymrl_sac_alpha_step). */
int gymrl_sac_actor_grad(const float* d_pre_tanh, const float* d_noise, const float* d_log_std, int ld_log_std,
                         const float* d_dq_daction, int ld_dq, const double* d_log_alpha, float bound, float log_std_min,
                         float log_std_max, float* d_dmean, float* d_dlog_std, int ld_out, const float* d_logp,
                         float* d_acc, int batch, int act_dim, void* stream);
/* alpha_loss = -mean(log_alpha*(logpi + target_entropy)) and one Adam step (lr, default betas/eps) on the float64
 * scalar log_alpha (sac :257-263). d_adam_state = float64[3] {exp_avg, exp_avg_sq, step}. */
int gymrl_sac_alpha_step(double* d_log_alpha, double* d_adam_state, const float* d_acc, int batch, double target_entropy,
                         double lr, float* d_loss_out, void* stream);
/* Discrete SAC (algorithms/sac_cartpole.py): softmax policy p = softmax(logits), lp = log(p + 1e-8), H = -sum p lp.
 *  target      y = r + gamma (1 - d) (sum_a p_a min(Q1t, Q2t)_a + alpha H) with the actor's logits on s' (ref :164-176);
 *  critic_loss mse(Q1(s)[a], y), mse(Q2(s)[a], y): d_loss_acc[0], [1] += the two losses, gradients zero off the taken action (:178-189);
 *  actor_grad  d/dlogits of mean(-alpha H - sum_a p_a min(Q1, Q2)_a); d_acc[0] += that loss, d_acc[1] += sum_i H_i (:191-200);
 *  alpha_step  alpha_loss = mean(exp(log_alpha) (H - target_entropy).detach()) and one float32 Adam step on log_alpha,
 *              d_adam_state = float32[3] {exp_avg, exp_avg_sq, step} (:202-207).  log_alpha is a float32 device scalar here. */
int gymrl_sac_discrete_target(const float* d_logits_next, int ld_logits, const float* d_q1t, int ld_q1t, const float* d_q2t,
                              int ld_q2t, const float* d_reward, const float* d_done, const int32_t* d_row_index,
                              const float* d_log_alpha, float gamma, float* d_y, int batch, int n_actions, void* stream);
int gymrl_sac_discrete_critic_loss(const float* d_q1, int ld_q1, const float* d_q2, int ld_q2, const int32_t* d_action,
                                   const int32_t* d_row_index, const float* d_y, float* d_dq1, int ld_dq1, float* d_dq2,
                                   int ld_dq2, float* d_loss_acc, int batch, int n_actions, void* stream);
int gymrl_sac_discrete_actor_grad(const float* d_logits, int ld_logits, const float* d_q1, int ld_q1, const float* d_q2,
                                  int ld_q2, const float* d_log_alpha, float* d_dlogits, int ld_dlogits, float* d_acc,
                                  int batch, int n_actions, void* stream);
int gymrl_sac_discrete_alpha_step(float* d_log_alpha, float* d_adam_state, const float* d_acc, int batch, float target_entropy,
                                  float lr, float* d_loss_out, void* stream);
/* a = tanh(z)*bound (td3_pendulum.py:59-62) and its backward dz = da * bound * (1 - (a/bound)^2) (td3 :215-217). */
int gymrl_tanh_bound(const float* d_z, int ld_z, float* d_action, float bound, int batch, int act_dim, void* stream);
int gymrl_tanh_bound_grad(const float* d_action, const float* d_dq_daction, int ld_dq, float* d_dz, int ld_dz, float bound,
                          int batch, int act_dim, void* stream);
/* N(0,1) draws kept in a buffer (Normal.rsample / torch.randn_like) so forward and backward share them. */
int gymrl_fill_normal(float* d_out, int n, uint64_t seed, uint64_t entity0, uint32_t draw, const uint32_t* d_draw_base,
                      void* stream);
/* NoisyLinear (rainbow_dqn_cartpole.py:51-97): eps = sign(xi) sqrt|xi| (d_xi optional pre-drawn normals);
 * W = mu + sigma*outer(eps_out, eps_in), b = b_mu + b_sigma*eps_out; and the backward of that composition. */
int gymrl_noisy_sample(float* d_eps, const float* d_xi, int n, uint64_t seed, uint64_t entity, uint32_t draw,
                       const uint32_t* d_draw_base, void* stream);
int gymrl_noisy_compose(const float* d_w_mu, const float* d_w_sigma, const float* d_eps_in, const float* d_eps_out,
                        const float* d_b_mu, const float* d_b_sigma, float* d_w, float* d_b, int N, int K, void* stream);
int gymrl_noisy_backward(const float* d_dw, const float* d_db, const float* d_eps_in, const float* d_eps_out,
                         float* d_dw_mu, float* d_dw_sigma, float* d_db_mu, float* d_db_sigma, int N, int K,
                         int accumulate, void* stream);
/* reset_noise() + the composition for up to two NoisyLinear layers with a common input width K (the dueling head of
 * rainbow_dqn_cartpole.py:100-124: advantage [N0][K], value [N1][K]; layer 1 optional: d_w_mu1 = NULL) in ONE launch: draws
 * eps_in0, eps_out0, eps_in1, eps_out1 with the Philox keys gymrl_noisy_sample would use for entity_base + {0,1,2,3} *
 * entity_stride (noisy = 0: eps = 0, i.e. W = mu — a network in eval mode), composes W and b, and adds counter_inc to the
 * device draw counter d_draw_base (nullable) after every thread has read it. */
int gymrl_noisy_refresh(const float* d_w_mu0, const float* d_w_sigma0, const float* d_b_mu0, const float* d_b_sigma0,
                        float* d_eps_in0, float* d_eps_out0, float* d_w0, float* d_b0, int N0, const float* d_w_mu1,
                        const float* d_w_sigma1, const float* d_b_mu1, const float* d_b_sigma1, float* d_eps_in1,
                        float* d_eps_out1, float* d_w1, float* d_b1, int N1, int K, int noisy, uint64_t seed,
                        uint64_t entity_base, uint64_t entity_stride, uint32_t draw, uint32_t* d_draw_base, int counter_inc,
                        void* stream);
/* gymrl_noisy_backward for the same pair of layers in one launch. */
int gymrl_noisy_backward2(const float* d_dw0, const float* d_db0, const float* d_eps_in0, const float* d_eps_out0,
                          float* d_dw_mu0, float* d_dw_sigma0, float* d_db_mu0, float* d_db_sigma0, int N0, const float* d_dw1,
                          const float* d_db1, const float* d_eps_in1, const float* d_eps_out1, float* d_dw_mu1,
                          float* d_dw_sigma1, float* d_db_mu1, float* d_db_sigma1, int N1, int K, int accumulate, void* stream);

/* Fused output heads + PPO loss + heads backward over the head-trunk activations d_h [batch][2H] = (actor | critic):
 *   logits = Wa h_a + ba (Wa [4][H]), V = Wc h_c + bc (Wc [1][H]); the loss of gymrl_ppo_loss (same cfg, same metrics);
 *   d_dh [batch][2H] = dL/dh (times the tanh derivative 1 - h^2 when act_in = GYMRL_ACT_TANH);
 *   dWa, dba, dWc, dbc (= or += with `accumulate`).  d_lv_out (nullable) [batch][8] receives logits | V.
 * One launch instead of the last two Linear layers of ActorCritic (algorithms/ppo_lunarlander.py:83,86), the loss
 * (:278-300) and their backward; H in {128, 256}, 4 actions.  Workspace: gymrl_ppo_heads_workspace_bytes(H, 4). */
size_t gymrl_ppo_heads_workspace_bytes(int H, int n_actions);
int gymrl_ppo_heads_fused(const float* d_h, int ldh, const float* d_Wa, const float* d_ba, const float* d_Wc, const float* d_bc,
                          const int32_t* d_row_index, const int32_t* d_action, const float* d_logp_old, const float* d_adv,
                          const float* d_ret, const float* d_entropy_old, const float* d_value_old, float* d_dh, int lddh,
                          int act_in, float* d_dWa, float* d_dba, float* d_dWc, float* d_dbc, float* d_lv_out, float* d_metrics,
                          void* d_workspace, size_t workspace_bytes, int accumulate, int batch, int H, int n_actions,
                          const gymrl_ppo_cfg* cfg, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Recurrent PPO / PPG building blocks.
 * ---------------------------------------------------------------------------------------------- */
/* Sequence-minibatch gather: out[b][t][:] = src[seq_index[b] * seq_len + t][:], i.e. states.view(S, L, -1)[perm[start:end]]
 * of algorithms/ppo_lstm_lunarlander.py:682-707 (seq_len 8) for any per-step field of `width` floats. */
int gymrl_seq_gather(const float* d_src, int ld_src, const int32_t* d_seq_index, int n_seq, int seq_len, int width,
                     float* d_out, int ld_out, void* stream);
/* torch.nn.GRU's cell (ppo_rnn_lunarlander.py:124-139 MLPRNN.rnn, ppo_lstm_lunarlander.py:449-492 URNN, ppg_rnn :330-395):
 * gi = x W_ih^T + b_ih and gh = h W_hh^T + b_hh come from gymrl_linear_forward ([B][3H], gate order r | z | n);
 *   r = sigmoid(gi_r + gh_r), z = sigmoid(gi_z + gh_z), n = tanh(gi_n + r gh_n), h' = (1 - z) n + z h.
 * d_gates ([B][3H], nullable in inference) keeps r, z, n for the backward pass. */
int gymrl_gru_cell_forward(const float* d_gi, int ld_gi, const float* d_gh, int ld_gh, const float* d_h, int ld_h,
                           float* d_h_out, int ld_out, float* d_gates, int batch, int hidden, void* stream);
/* Given dL/dh', writes dL/dgi and dL/dgh ([B][3H]) and the direct path of dL/dh (= z dh'; += when accumulate_dh; nullable).
 * The caller completes BPTT with the dense-layer entry points: dx = dgi W_ih, dh += dgh W_hh, dW_ih += dgi^T x, ... */
int gymrl_gru_cell_backward(const float* d_dh_out, int ld_dh_out, const float* d_gates, const float* d_gh, int ld_gh,
                            const float* d_h, int ld_h, float* d_dgi, int ld_dgi, float* d_dgh, int ld_dgh, float* d_dh,
                            int ld_dh, int accumulate_dh, int batch, int hidden, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ppo_full network glue: manifold hyper-connection stages, RMSNorm, SiLU.
 * algorithms/ppo_full_lunarlander.py: sinkhorn_knopp_batched :76-103, ManifoldHyperConnectionFuse :106-194,
 * MHCBlock :197-229, MHCBackbone :232-267, RMSNorm :273-284, MLP :287-318.  mhc_rate n = 2, mhc_dim D in {128, 256}.
 * Rows are [2][D] float32 (row stride / branch stride given for inputs so that the backbone input x0 [M][D] can be
 * read as both branches: stride D / 0).  The Linear layers between the stages are gymrl_linear_* with ACT_NONE.
 * ---------------------------------------------------------------------------------------------- */
/* Bytes of workspace the *_backward* calls below need (per-block parameter-gradient partials). */
size_t gymrl_mhc_workspace_bytes(int D, int head_width, int head_groups);
/* One fused row-wise pass between two GEMMs of the backbone forward:
 *   (z_prev != NULL)  h_cur = depth_connection(prev stage) = post_i silu(z_prev) + sum_j P_ij h_prev_j, stored [M][2][D];
 *                     otherwise the row is h_prev itself (not stored);
 *   (g != NULL)       mapping() of the next stage on that row: coef_cur [M][24] = {pre0 pre1 post0 post1 | P00 P01 P10 P11 |
 *                     r_ s 0 0 | 0 0 0 0 | H0..H7} (the mapping is kept whole for the backward pass; coef_prev has the same layout),
 *                     h_pre [M][D] = sum_i pre_i h_i (the next GEMM's input);
 *   (final_weight)    feat [M][D] = RMSNorm(h_0 + h_1) * final_weight  (MHCBackbone.forward :263-267). */
int gymrl_mhc_stage_forward(const float* d_h_prev, int prev_row_stride, int prev_branch_stride, const float* d_z_prev,
                            const float* d_coef_prev, float* d_h_cur, const float* d_g, const float* d_w, const float* d_alpha,
                            const float* d_beta, float* d_coef_cur, float* d_h_pre, const float* d_final_weight, float* d_feat,
                            int M, int D, int sk_iters, float eps, void* stream);
/* Backward of one stage, split around its GEMM backward.  A: from d_dh_next = dL/d(stage output) [M][2][D], the saved stage
 * input h, the GEMM output z and the stage's coefficient rows d_coef [M][24] as the forward pass left them: dz [M][D] (the
 * GEMM's upstream gradient), dh_partial [M][2][D], and the six inner products dpost_i, dP_ij into d_coef[:, 10:16]. */
int gymrl_mhc_stage_backward_a(const float* d_h, int row_stride, int branch_stride, const float* d_z, const float* d_dh_next,
                               float* d_dz, float* d_dh_partial, float* d_coef, int M, int D, void* stream);
/* B (d_scratch = the d_coef rows after A): with d_dh_pre = dz @ W [M][D]: the full input gradient (d_dh [M][2][D], or its branch sum d_dx0 [M][D] when that is
 * non-NULL) and the parameter gradients of mhc.norm.weight (dg [2D]), mhc.w (dw [2D][8]), alpha [3], beta [8]. */
int gymrl_mhc_stage_backward_b(const float* d_h, int row_stride, int branch_stride, const float* d_dh_pre, const float* d_scratch,
                               const float* d_dh_partial, float* d_dh, float* d_dx0, const float* d_g, const float* d_w,
                               const float* d_alpha, float* d_dg, float* d_dw, float* d_dalpha, float* d_dbeta, void* d_workspace,
                               size_t workspace_bytes, int accumulate, int M, int D, void* stream);
/* y = a * rsqrt(mean(a^2) + eps) * weight per group of width W (128 | 256), `groups` groups per row;
 * a = silu(x) (silu = 1: the MLP heads' Linear -> SiLU -> RMSNorm, :305-308) | x | x[0:W] + x[W:2W] (sum2 = 1). */
int gymrl_rmsnorm_forward(const float* d_x, int ldx, int sum2, int silu, const float* d_weight, float* d_y, int ldy, int M, int W,
                          int groups, float eps, void* stream);
int gymrl_rmsnorm_backward(const float* d_x, int ldx, int sum2, int silu, const float* d_weight, const float* d_dy, int lddy,
                           float* d_dx, int lddx, float* d_dweight, void* d_workspace, size_t workspace_bytes, int accumulate,
                           int M, int W, int groups, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Running normalisation of the utils path: utils/normalization.py RunningMeanStd :4-22,
 * Normalization :25-35, RewardScaling :38-52 (applied per env step by utils/runner.py:112,125-126).
 * d_state = float64 [1 + 3 D] {n, mean[D], S[D], std[D]} (zero-initialised).  A call feeds the N rows of x in env
 * order: for N <= 32 with the reference's exact rule and precisions (first-sample quirk mean = std = x included),
 * so N = 1 is the reference; for larger N the batch moments are merged in float64.
 * ---------------------------------------------------------------------------------------------- */
int gymrl_running_stats_update(const float* d_x, int N, int D, double* d_state, void* stream);
/* y = (x - mean) / (std + 1e-8)  (center = 1, Normalization.__call__) or x / (std + 1e-8) (center = 0). */
int gymrl_running_normalize(const float* d_x, float* d_y, int N, int D, const double* d_state, int center, void* stream);
/* RewardScaling.__call__: R_i = gamma R_i + r_i (R_i zeroed first where d_reset[i], = RewardScaling.reset at an
 * episode start), statistic (d_state float64[4], D = 1) updated with the N values of R, out_i = r_i / (std + 1e-8). */
int gymrl_reward_scaling(const float* d_r, float* d_out, double* d_R, const uint8_t* d_reset, double gamma, double* d_state,
                         int N, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GYMRL_H */
