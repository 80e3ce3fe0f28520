"""ctypes binding of libgymrl_b200.so (the C ABI declared in include/gymrl.h).

This is the *only* place Python touches the native library.  torch is used as plumbing: it owns the
device memory and the streams; every call below passes raw device pointers + the current CUDA stream.
There is no CPU fallback: if the shared object is missing (or there is no CUDA device) the product
path raises — it never routes through oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libgymrl_b200.so"
_lib: Optional[C.CDLL] = None

c_void_p, c_int, c_float, c_double, c_u64, c_u32, c_ll, c_size_t = (
    C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_uint64, C.c_uint32, C.c_longlong, C.c_size_t)


class PPOCfg(C.Structure):
    """struct gymrl_ppo_cfg"""
    _fields_ = [("mode", c_int), ("clip_eps_min", c_float), ("clip_eps_max", c_float), ("dual_clip", c_float),
                ("value_coef", c_float), ("entropy_coef", c_float), ("erc_low", c_float), ("erc_high", c_float),
                ("vclip_eps_min", c_float), ("vclip_eps_max", c_float), ("d_entropy_coef", c_void_p),
                ("d_mask_count", c_void_p)]


PPO_DUALCLIP, PPO_FULL, PPO_VALUE_CLIP, PPO_MASKED_MEAN = 0, 1, 4, 8
ACT_NONE, ACT_TANH, ACT_RELU = 0, 1, 2
ENV_CARTPOLE, ENV_PENDULUM, ENV_LUNARLANDER = 0, 1, 2
ENV_KINDS = {"CartPole-v1": ENV_CARTPOLE, "Pendulum-v1": ENV_PENDULUM, "LunarLander-v3": ENV_LUNARLANDER}

# name -> (restype, argtypes); must cover every symbol in include/gymrl.h (tests/test_abi.py checks)
_P = c_void_p
SIGNATURES = {
    "gymrl_version": (c_int, []),
    "gymrl_last_error": (C.c_char_p, []),
    "gymrl_launch_count": (c_u64, []),
    "gymrl_env_info": (c_int, [c_int] + [_P] * 6),
    "gymrl_env_create": (c_int, [_P, c_int, c_int, c_u64, c_u64]),
    "gymrl_env_destroy": (c_int, [_P]),
    "gymrl_env_reset": (c_int, [_P, _P, _P, _P]),
    "gymrl_env_step": (c_int, [_P] * 9),
    "gymrl_env_get_state": (c_int, [_P, _P, _P]),
    "gymrl_env_set_state": (c_int, [_P, _P, _P]),
    "gymrl_env_set_profile": (c_int, [_P, _P]),
    "gymrl_env_overflow_count": (c_int, [_P, _P, _P]),
    "gymrl_env_set_solver": (c_int, [_P, c_int]),
    "gymrl_env_get_solver": (c_int, [_P, _P]),
    "gymrl_env_episode_stats": (c_int, [_P, c_int, _P, _P, _P, _P]),
    "gymrl_policy_heads_sample": (c_int, [_P, c_int, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, c_int, c_u64, c_u64, c_u32, _P, c_int, _P]),
    "gymrl_sample_categorical": (c_int, [_P, c_int, _P, _P, _P, _P, _P, c_int, _P, c_int, c_int, c_u64, c_u64, c_u32, _P, c_int, _P]),
    "gymrl_select_eps_greedy": (c_int, [_P, c_int, _P, c_int, c_int, c_float, c_u64, c_u64, c_u32, _P, _P]),
    "gymrl_select_eps_greedy_dev": (c_int, [_P, c_int, _P, c_int, c_int, _P, c_u64, c_u64, c_u32, _P, _P]),
    "gymrl_sample_tanh_gaussian": (c_int, [_P, _P, c_int, _P, _P, _P, _P, c_int, c_int, c_float, c_float, c_float,
                                           c_u64, c_u64, c_u32, _P, c_int, _P]),
    "gymrl_add_gaussian_noise_clip": (c_int, [_P, _P, _P, c_int, c_int, c_float, c_float, c_float, c_u64, c_u64, c_u32, _P, _P]),
    "gymrl_gae": (c_int, [_P] * 7 + [c_int, c_int, c_double, c_double, c_double, c_int, _P]),
    "gymrl_sum_sumsq": (c_int, [_P, c_ll, _P, _P]),
    "gymrl_normalize_inplace": (c_int, [_P, c_ll, _P, c_double, c_int, c_float, _P]),
    "gymrl_ppo_loss": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, c_int, _P, c_int, c_int, _P, _P]),
    "gymrl_set_gemm_mode": (c_int, [c_int]),
    "gymrl_get_gemm_mode": (c_int, []),
    "gymrl_linear_forward": (c_int, [_P, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "gymrl_linear_backward_input": (c_int, [_P, c_int, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "gymrl_linear_backward_weight_workspace": (c_size_t, [c_int, c_int, c_int]),
    "gymrl_linear_backward_weight": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "gymrl_linear_backward": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_size_t, _P]),
    "gymrl_grad_sumsq": (c_int, [_P, c_ll, _P, _P]),
    "gymrl_adam_step": (c_int, [_P, _P, _P, _P, c_ll, _P, c_float, c_float, c_float, _P, _P, c_float, c_float, c_float, _P]),
    "gymrl_reduce_defer_begin": (c_int, []),
    "gymrl_reduce_flush": (c_int, [_P, c_int, C.POINTER(c_int), C.POINTER(c_ll), _P]),
    "gymrl_clip_adam_step": (c_int, [_P, _P, _P, _P, c_ll, _P, c_float, c_float, c_float, _P, _P, c_int, c_float, c_float, _P, _P]),
    "gymrl_polyak": (c_int, [_P, _P, c_ll, c_float, _P]),
    "gymrl_weight_images_register": (c_int, [_P, c_ll, _P, _P, c_int]),
    "gymrl_weight_images_refresh": (c_int, [_P, _P]),
    "gymrl_weight_images_unregister": (c_int, [_P]),
    "gymrl_comm_create": (c_int, [_P, c_int, c_int, c_ll, c_int]),
    "gymrl_comm_handle_bytes": (c_int, []),
    "gymrl_comm_get_handle": (c_int, [_P, _P]),
    "gymrl_comm_open": (c_int, [_P, _P]),
    "gymrl_comm_n_partials": (c_int, [_P]),
    "gymrl_comm_allreduce_sumsq": (c_int, [_P, _P, _P, _P, _P]),
    "gymrl_comm_destroy": (c_int, [_P]),
    "gymrl_sac_discrete_target": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, _P, _P, _P, c_float, _P, c_int, c_int, _P]),
    "gymrl_sac_discrete_critic_loss": (c_int, [_P, c_int, _P, c_int, _P, _P, _P, _P, c_int, _P, c_int, _P, c_int, c_int, _P]),
    "gymrl_sac_discrete_actor_grad": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, _P, c_int, _P, c_int, c_int, _P]),
    "gymrl_sac_discrete_alpha_step": (c_int, [_P, _P, _P, c_int, c_float, c_float, _P, _P]),
    "gymrl_random_permutation": (c_int, [_P, c_int, c_u64, c_u32, _P, _P]),
    "gymrl_ppo_heads_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gymrl_ppo_heads_fused": (c_int, [_P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P,
                                      c_size_t, c_int, c_int, c_int, c_int, _P, _P]),
    "gymrl_running_stats_update": (c_int, [_P, c_int, c_int, _P, _P]),
    "gymrl_running_normalize": (c_int, [_P, _P, c_int, c_int, _P, c_int, _P]),
    "gymrl_reward_scaling": (c_int, [_P, _P, _P, _P, c_double, _P, c_int, _P]),
    "gymrl_mhc_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gymrl_mhc_stage_forward": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, _P]),
    "gymrl_mhc_stage_backward_a": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, c_int, c_int, _P]),
    "gymrl_mhc_stage_backward_b": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, c_int,
                                           c_int, c_int, _P]),
    "gymrl_rmsnorm_forward": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, c_int, c_float, _P]),
    "gymrl_rmsnorm_backward": (c_int, [_P, c_int, c_int, c_int, _P, _P, c_int, _P, c_int, _P, _P, c_size_t, c_int, c_int, c_int,
                                       c_int, c_float, _P]),
    "gymrl_seq_gather": (c_int, [_P, c_int, _P, c_int, c_int, c_int, _P, c_int, _P]),
    "gymrl_gru_cell_forward": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, _P]),
    "gymrl_gru_cell_backward": (c_int, [_P, c_int, _P, _P, c_int, _P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, c_int, _P]),
    "gymrl_counter_add": (c_int, [_P, c_u32, _P]),
    "gymrl_slice_i32": (c_int, [_P, _P, c_int, _P, _P]),
    "gymrl_replay_sample_indices": (c_int, [_P, c_int, _P, c_u64, c_u32, _P, _P]),
    "gymrl_replay_store": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    "gymrl_replay_advance": (c_int, [_P, c_int, c_int, _P]),
    "gymrl_gather_concat": (c_int, [_P, c_int, _P, c_int, c_int, _P, _P, c_int, c_int, _P, c_int, _P]),
    "gymrl_nstep_push": (c_int, [_P] * 13 + [c_int, c_int, c_int, c_double, _P] + [_P] * 5 + [c_int, _P, _P]),
    "gymrl_sumtree_update": (c_int, [_P, c_int, _P, _P, _P, c_int, c_float, c_float, c_float, _P, _P]),
    "gymrl_sumtree_scratch_ints": (c_int, [c_int]),
    "gymrl_sumtree_store_new": (c_int, [_P, c_int, c_int, _P, _P, _P, _P]),
    "gymrl_sumtree_sample": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_int, c_u64, c_u32, _P, _P]),
    "gymrl_dqn_loss": (c_int, [_P, c_int] * 6 + [_P] * 5 + [_P, c_int, _P, c_int, _P, _P, c_int, c_int, c_float, _P]),
    "gymrl_twin_q_target": (c_int, [_P, _P, _P, _P, c_int, _P, c_int, _P, _P, c_float, _P, c_int, _P]),
    "gymrl_twin_q_loss": (c_int, [_P, c_int, _P, c_int, _P, _P, c_int, _P, c_int, _P, c_int, _P]),
    "gymrl_min_q_grad": (c_int, [_P, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, c_int, _P, _P]),
    "gymrl_sac_actor_grad": (c_int, [_P, _P, _P, c_int, _P, c_int, _P, c_float, c_float, c_float, _P, _P, c_int, _P, _P, c_int, c_int, _P]),
    "gymrl_sac_alpha_step": (c_int, [_P, _P, _P, c_int, c_double, c_double, _P, _P]),
    "gymrl_tanh_bound": (c_int, [_P, c_int, _P, c_float, c_int, c_int, _P]),
    "gymrl_tanh_bound_grad": (c_int, [_P, _P, c_int, _P, c_int, c_float, c_int, c_int, _P]),
    "gymrl_fill_normal": (c_int, [_P, c_int, c_u64, c_u64, c_u32, _P, _P]),
    "gymrl_noisy_sample": (c_int, [_P, _P, c_int, c_u64, c_u64, c_u32, _P, _P]),
    "gymrl_noisy_compose": (c_int, [_P] * 8 + [c_int, c_int, _P]),
    "gymrl_noisy_backward": (c_int, [_P] * 8 + [c_int, c_int, c_int, _P]),
    "gymrl_noisy_refresh": (c_int, [_P] * 8 + [c_int] + [_P] * 8 + [c_int, c_int, c_int, c_u64, c_u64, c_u64, c_u32, _P, c_int, _P]),
    "gymrl_noisy_backward2": (c_int, [_P] * 8 + [c_int] + [_P] * 8 + [c_int, c_int, c_int, _P]),
    "gymrl_replay_store_all": (c_int, [_P] * 10 + [c_int, c_int, c_int, c_int, _P, _P, _P]),
}


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the native library (building nothing: run `python -m gymrl_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing. gymrl_b200 has no CPU fallback: build the CUDA library first "
            "(python -m gymrl_b200.build, or __graft_entry__.build()).")
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class GymrlError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = load().gymrl_last_error()
        raise GymrlError(f"gymrl error {rc}: {msg.decode() if msg else '?'}")


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("gymrl_b200 requires a CUDA device (B200 / sm_100a); there is no CPU fallback")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor], dtype: Optional[torch.dtype] = None) -> Optional[int]:
    """Device pointer of a tensor (None -> NULL), with the checks the ABI assumes."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GymrlError("expected a CUDA tensor")
    if dtype is not None and t.dtype != dtype:
        raise GymrlError(f"expected dtype {dtype}, got {t.dtype}")
    return t.data_ptr()


def launch_count() -> int:
    return int(load().gymrl_launch_count())
