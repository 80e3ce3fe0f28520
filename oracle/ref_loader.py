"""Import the UNMODIFIED reference modules from /root/reference behind a stub `gymnasium`.

TEST INFRASTRUCTURE.  /root/reference exists only in the build container (never on the GPU box), so
this module is used by oracle/make_golden.py (fixtures are committed under tests/golden/) and by the
`not gpu` tests that are skipped when the reference tree is absent.  Nothing is copied from the
reference: files are executed where they lie via importlib.
"""
import importlib.util
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")


def available() -> bool:
    return (REFERENCE_ROOT / "algorithms" / "ppo_lunarlander.py").exists()


def install_gymnasium_stub(make=None):
    """Register a minimal `gymnasium` module (make / spaces.Box / ObservationWrapper / wrappers)."""
    if "gymnasium" in sys.modules and getattr(sys.modules["gymnasium"], "_gymrl_stub", False):
        if make is not None:
            sys.modules["gymnasium"].make = make
        return sys.modules["gymnasium"]
    gym = types.ModuleType("gymnasium")
    gym._gymrl_stub = True

    def _no_make(*a, **k):
        raise RuntimeError("stub gymnasium: no env backend installed (pass make=...)")

    gym.make = make or _no_make

    class Box:
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class ObservationWrapper:
        def __init__(self, env):
            self.env = env

    spaces = types.ModuleType("gymnasium.spaces")
    spaces.Box = Box
    wrappers = types.ModuleType("gymnasium.wrappers")
    wrappers.AtariPreprocessing = object
    gym.spaces, gym.wrappers, gym.ObservationWrapper = spaces, wrappers, ObservationWrapper
    sys.modules["gymnasium"] = gym
    sys.modules["gymnasium.spaces"] = spaces
    sys.modules["gymnasium.wrappers"] = wrappers
    return gym


def load(relpath: str, name: str = None):
    """Execute /root/reference/<relpath> as a module (e.g. 'algorithms/ppo_lunarlander.py')."""
    if not available():
        raise FileNotFoundError("reference tree not present (expected only in the build container)")
    install_gymnasium_stub()
    path = REFERENCE_ROOT / relpath
    name = name or ("ref_" + path.stem)
    if relpath.startswith("utils/"):
        # utils modules import each other as `utils.x`
        if "utils" not in sys.modules or not getattr(sys.modules["utils"], "_gymrl_ref", False):
            pkg = types.ModuleType("utils")
            pkg.__path__ = [str(REFERENCE_ROOT / "utils")]
            pkg._gymrl_ref = True
            sys.modules["utils"] = pkg
        name = "utils." + path.stem
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class FakeEnv:
    """Just enough of a gymnasium env for the reference Trainer constructors."""

    def __init__(self, obs_dim, n_actions=None, act_dim=None, bound=None, max_steps=500):
        import numpy as np
        self.observation_space = types.SimpleNamespace(shape=(obs_dim,))
        if n_actions is not None:
            self.action_space = types.SimpleNamespace(n=n_actions, sample=lambda: 0)
        else:
            self.action_space = types.SimpleNamespace(shape=(act_dim,), high=np.full(act_dim, bound, np.float32))
        self.spec = types.SimpleNamespace(max_episode_steps=max_steps)

    def reset(self, seed=None):
        raise RuntimeError("FakeEnv has no dynamics")

    def close(self):
        pass
