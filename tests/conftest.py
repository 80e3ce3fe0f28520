import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(GOLDEN / name, allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def native_lib():
    """The built C-ABI library (incremental nvcc build; cross-compiles without a GPU)."""
    from gymrl_b200 import build
    return build.build()


def _cu(x):
    import numpy as np
    import torch
    return torch.as_tensor(np.ascontiguousarray(x)).cuda()


def trainer_from_golden(g, use_graph, n_mb, seed=0):
    from gymrl_b200.algorithms import ppo_lunarlander as P
    B = g["states"].shape[0]
    cfg = P.Config()
    cfg.num_envs, cfg.num_steps, cfg.num_minibatches, cfg.num_epochs = 1, B, n_mb, 1
    cfg.seed, cfg.use_cuda_graph = seed, use_graph
    t = P.PPOTrainer(cfg)
    import torch
    sd = {k[3:]: torch.as_tensor(g[k]) for k in g.files if k.startswith("w0_")}
    t.model.load_state_dict(sd)
    t.net.fp.refresh_views()
    buf = t.buffer
    buf.obs[:B, 0].copy_(_cu(g["states"]))
    buf.action[:, 0].copy_(_cu(g["action"])); buf.log_prob[:, 0].copy_(_cu(g["logp_old"]))
    buf.value[:, 0].copy_(_cu(g["value_old"])); buf.reward[:, 0].copy_(_cu(g["reward"])); buf.done[:, 0].copy_(_cu(g["done"]))
    buf.filled = B
    return t




@pytest.fixture(scope="session")
def ops():
    from gymrl_b200 import ops as _ops
    return _ops
