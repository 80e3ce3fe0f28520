"""Host build of the CUDA LunarLander solver source (gymrl_b200/csrc/env_lunar.cu -DGYMRL_HOSTSIM).  TEST INFRASTRUCTURE.

The solver in env_lunar.cu is scalar code (one thread = one env copy) whose functions are __host__ __device__; nvcc compiles the
very same statements for the host here (kernels and CUDA glue dropped), so the arithmetic of the shipped CUDA source can be
checked bit for bit against oracle/lunar_lander.c on a box WITHOUT a GPU (tests/test_hostsim_lunar.py).  It is a checker of the
source, not a CPU fallback: nothing under gymrl_b200/ loads this library, and it is built outside gymrl_b200/lib.

Same rounding rules as the device build: -fmad=false for the device pass, -ffp-contract=off for the host pass (the only fused
multiply-adds are the explicit fmaf() calls), IEEE division / sqrt on both.
"""
import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np

_ROOT = Path(__file__).resolve().parent.parent
_SRC = _ROOT / "gymrl_b200" / "csrc" / "env_lunar.cu"
_BUILD = Path(__file__).resolve().parent / "_build"
_libs = {}
# host build -> (solver variant of env_lunar.cu, force the "division operand out of its window" flag on about half of the steps so
# that the repeat path runs).  Solver 0 = the oracle's arrangement, plain division; 2 = div_chain in the position iterations (the
# shipping default); 3 = 2 + the velocity loop specialised on the joints' limit states.
BUILDS = {0: (0, False), 1: (2, False), 2: (2, True), 3: (3, False), 4: (3, True)}


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    return None


def so_path(variant=0):
    return _BUILD / f"libhostsim_lunar_v{variant}.so"


def build(variant=0, force=False):
    """nvcc -DGYMRL_HOSTSIM -shared env_lunar.cu -> tests/_build/libhostsim_lunar_v<variant>.so (rebuilt when the source is newer)."""
    out = so_path(variant)
    deps = [_SRC, _SRC.parent / "env.cuh", _SRC.parent / "common.cuh"]
    if not force and out.exists() and all(out.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return out
    cc = nvcc()
    if cc is None:
        raise RuntimeError("nvcc not found: the host build of env_lunar.cu needs the CUDA toolkit (no GPU)")
    _BUILD.mkdir(exist_ok=True)
    solver, force_bad = BUILDS[variant]
    extra = ["-DLL_HOSTSIM_FORCE_BAD=1"] if force_bad else []
    cmd = [cc, "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-strict-aliasing", "-Xcompiler", "-ffp-contract=off",
           "-Xcompiler", "-mfma", "--expt-relaxed-constexpr", "-fmad=false", "-DGYMRL_HOSTSIM", f"-DLL_SOLVER_VARIANT={solver}", *extra,
           "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(out), str(_SRC)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("hostsim build failed:\n" + r.stdout + r.stderr)
    return out


def lib(variant=0):
    if variant not in _libs:
        L = C.CDLL(str(build(variant)))
        L.gymrl_hostsim_state_doubles.restype = C.c_int
        L.gymrl_hostsim_solver_variant.restype = C.c_int
        L.gymrl_hostsim_lunar_reset.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        L.gymrl_hostsim_lunar_reset.restype = None
        L.gymrl_hostsim_lunar_step.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64] + [C.c_void_p] * 6
        L.gymrl_hostsim_lunar_step.restype = None
        assert L.gymrl_hostsim_solver_variant() == BUILDS[variant][0]
        _libs[variant] = L
    return _libs[variant]


class HostSimLunarVec:
    """N env copies stepped through the host build of the CUDA solver; same surface as oracle.lunar.LunarLanderVec."""

    def __init__(self, num_envs, seed=0, first_env_id=0, variant=0):
        self.L = lib(variant)
        self.n, self.seed, self.first = int(num_envs), int(seed) & (2**64 - 1), int(first_env_id)
        self.sd = self.L.gymrl_hostsim_state_doubles()
        self.state = np.zeros((self.n, self.sd), np.float64)
        self.obs = np.zeros((self.n, 8), np.float32)
        self.next_obs = np.zeros((self.n, 8), np.float32)
        self.reward = np.zeros(self.n, np.float32)
        self.terminated = np.zeros(self.n, np.uint8)
        self.truncated = np.zeros(self.n, np.uint8)
        self.prof = np.zeros((self.n, 9), np.int32)

    def reset(self):
        for i in range(self.n):
            self.L.gymrl_hostsim_lunar_reset(self.state[i].ctypes.data, self.seed, self.first + i, self.obs[i].ctypes.data)
        return self.obs.copy()

    def step(self, action):
        a = np.asarray(action)
        for i in range(self.n):
            self.L.gymrl_hostsim_lunar_step(self.state[i].ctypes.data, int(a[i]), self.seed, self.first + i, self.obs[i].ctypes.data,
                                            self.next_obs[i].ctypes.data, self.reward[i:].ctypes.data, self.terminated[i:].ctypes.data,
                                            self.truncated[i:].ctypes.data, self.prof[i].ctypes.data)
        return self.obs.copy(), self.next_obs.copy(), self.reward.copy(), self.terminated.copy(), self.truncated.copy()

    def get_state(self):
        return self.state.copy()

    def set_state(self, s):
        self.state[:] = np.asarray(s, np.float64)
