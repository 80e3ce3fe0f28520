// env.cuh — the gymrl_env handle shared by env.cu (classic control) and env_lunar.cu (LunarLander).
#pragma once
#include "common.cuh"

#define GYMRL_EP_RING 1024  // finished-episode ring (returns / lengths) kept on device

struct gymrl_env {
    int kind;
    int n;
    uint64_t seed;
    uint64_t first_id;
    int device;
    // classic-control state: [S][N] float64 SoA (CartPole S=4: x, x_dot, theta, theta_dot; Pendulum S=2: th, thdot)
    double* state;
    // LunarLander state: SoA planes, see env_lunar.cu
    float* ll_f;
    int32_t* ll_i;
    double* ll_d;
    // LunarLander spares: pre-built next episode per env (same plane layout), its first observation, and the refill queues
    float* ll_sf;
    int32_t* ll_si;
    double* ll_sd;
    float* spare_obs;        // [N][8]
    int32_t* spare_ready;    // [N]
    int32_t* refill_list;    // [3][N]
    int32_t* refill_count;   // [3] (+ tick at [3])
    int32_t* tick;           // step counter mod-3 addressing of the refill queues
    // LunarLander work scheduling: a step's solver chain is 4-8x longer for an env with ground contacts than for one
    // in free flight, so the step kernel runs the few heavy envs first and alone in their warps and packs the light
    // ones densely (results do not depend on the mapping).  cost = touching manifolds in the env's last step.
    int32_t* cost;           // [N]
    int32_t* order;          // [N] env ids, heavy first
    int32_t* order_cnt;      // [2] = {heavy count, heavy envs per warp}
    long long* prof;         // nullable diagnostic buffer [N][8], see gymrl_env_set_profile
    int solver;              // LunarLander solver loop variant (gymrl_env_set_solver; same results bit for bit)
    // shared bookkeeping, all [N]
    int32_t* elapsed;     // TimeLimit counter
    uint32_t* episode;    // episodes started so far (keys the reset draws)
    uint32_t* stepctr;    // env steps taken so far (keys the per-step noise draws)
    double* ep_return;    // running undiscounted return
    // finished-episode ring
    float* ring_ret;
    int32_t* ring_len;
    unsigned long long* ring_count;
};

// Push finished episodes into the ring with one atomic per warp (warp-ballot aggregated).
__device__ __forceinline__ void episode_ring_push(bool done, float ep_ret, int ep_len, float* ring_ret,
                                                  int32_t* ring_len, unsigned long long* ring_count) {
    const unsigned active = __activemask();
    const unsigned ballot = __ballot_sync(active, done);
    if (ballot == 0u) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(ballot) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(ring_count, (unsigned long long)__popc(ballot));
    base = __shfl_sync(active, base, leader);
    if (done) {
        const unsigned long long slot = (base + __popc(ballot & ((1u << lane) - 1u))) % GYMRL_EP_RING;
        ring_ret[slot] = ep_ret;
        ring_len[slot] = ep_len;
    }
}

// implemented in env_lunar.cu
int lunar_alloc(gymrl_env* e);
void lunar_free(gymrl_env* e);
int lunar_reset(gymrl_env* e, const uint8_t* mask, float* obs, cudaStream_t s);
int lunar_step(gymrl_env* e, const int32_t* actions, float* obs, float* next_obs, float* reward,
               uint8_t* terminated, uint8_t* truncated, uint8_t* done, cudaStream_t s);
int lunar_get_state(gymrl_env* e, double* state, cudaStream_t s);
int lunar_set_state(gymrl_env* e, const double* state, cudaStream_t s);
int lunar_default_solver();

void gymrl_count_launch(int n = 1);
