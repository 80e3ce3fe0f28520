"""ctypes wrapper over oracle/lunar_lander.c (the LunarLander-v3 CPU restatement).  TEST INFRASTRUCTURE."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "liboracle_lunar.so"
_lib = None


def build():
    """make -C oracle (gcc only; the C file is a restatement, not reference source)."""
    subprocess.run(["make", "-C", str(_DIR)], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        L = C.CDLL(str(_SO))
        L.ll_create.restype = C.c_void_p
        L.ll_create.argtypes = [C.c_uint64, C.c_uint64]
        L.ll_destroy.argtypes = [C.c_void_p]
        L.ll_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.ll_step.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.ll_get_state.argtypes = [C.c_void_p, C.c_void_p]
        L.ll_set_state.argtypes = [C.c_void_p, C.c_void_p]
        L.ll_state_doubles.restype = C.c_int
        L.ll_mass_data.argtypes = [C.c_void_p]
        L.ll_vec_reset.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ll_vec_step.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        _lib = L
    return _lib


class LunarLanderVec:
    """N independent LunarLander-v3 copies stepped in a C loop; same vector semantics as include/gymrl.h."""
    obs_dim, n_actions, max_steps = 8, 4, 1000

    def __init__(self, num_envs, seed=0, first_env_id=0):
        L = lib()
        self.n = int(num_envs)
        self._envs = (C.c_void_p * self.n)(*[L.ll_create(seed & (2**64 - 1), first_env_id + i) for i in range(self.n)])
        self.obs = np.zeros((self.n, 8), np.float32)
        self.next_obs = np.zeros((self.n, 8), np.float32)
        self.reward = np.zeros(self.n, np.float32)
        self.terminated = np.zeros(self.n, np.uint8)
        self.truncated = np.zeros(self.n, np.uint8)
        self.sd = L.ll_state_doubles()

    def __del__(self):
        try:
            L = lib()
            for e in self._envs:
                L.ll_destroy(e)
        except Exception:
            pass

    def reset(self, mask=None):
        L = lib()
        if mask is None:
            L.ll_vec_reset(self._envs, self.n, self.obs.ctypes.data)
        else:
            for i in np.nonzero(mask)[0]:
                L.ll_reset(self._envs[i], self.obs[i].ctypes.data)
        return self.obs.copy()

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.int32)
        lib().ll_vec_step(self._envs, self.n, a.ctypes.data, self.obs.ctypes.data, self.next_obs.ctypes.data,
                          self.reward.ctypes.data, self.terminated.ctypes.data, self.truncated.ctypes.data)
        return self.obs.copy(), self.next_obs.copy(), self.reward.copy(), self.terminated.copy(), self.truncated.copy()

    def get_state(self):
        s = np.zeros((self.n, self.sd), np.float64)
        for i in range(self.n):
            lib().ll_get_state(self._envs[i], s[i].ctypes.data)
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        for i in range(self.n):
            lib().ll_set_state(self._envs[i], s[i].ctypes.data)


def mass_data():
    out = np.zeros((3, 4), np.float64)
    lib().ll_mass_data(out.ctypes.data)
    return out
