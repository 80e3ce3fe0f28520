"""Generate tests/golden/ref_checkpoint.pth (+ ref_checkpoint_expect.npz): a checkpoint WRITTEN BY THE UNMODIFIED REFERENCE
`utils.model.ModelLoader.save_model` (/root/reference/utils/model.py:337-349) for an agent shaped like
legacy/LunarLander(PPO).py's (net = utils.model.MLP, optimizer = torch Adam, state_norm / reward_scaler = the reference's
utils.normalization objects pickled as plain attributes, learn_step).  tests/test_gpu_utils.py loads it through
gymrl_b200.utils.model.ModelLoader.  TEST INFRASTRUCTURE; build container only:
    python -m oracle.make_golden_checkpoint
"""
import os
import shutil
import tempfile
from pathlib import Path

import numpy as np
import torch

from . import ref_loader as rl

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def main():
    rl.load("utils/buffer.py")
    nz = rl.load("utils/normalization.py")
    model = rl.load("utils/model.py")
    torch.manual_seed(0)
    rng = np.random.default_rng(0)

    class Cfg:
        algo_name, env_name, device = "PPO", "LunarLander-v3", "cpu"

    class Agent(model.ModelLoader):
        def __init__(self, cfg):
            super().__init__(cfg)
            self.net = model.MLP([8, 32, 4])
            self.optimizer = torch.optim.Adam(self.net.parameters(), lr=3e-4, eps=1e-5)
            self.state_norm = nz.Normalization(shape=(8,))
            self.reward_scaler = nz.RewardScaling(shape=1, gamma=0.99)
            self.learn_step = 0

    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        cfg = Cfg()
        ag = Agent(cfg)
        x = rng.standard_normal((40, 8)).astype(np.float32) * 2 + 1
        for t in range(40):
            ag.state_norm(x[t])
            ag.reward_scaler(float(rng.standard_normal()))
        for _ in range(3):
            ag.optimizer.zero_grad()
            ag.net(torch.as_tensor(x)).pow(2).mean().backward()
            ag.optimizer.step()
            ag.learn_step += 1
        ag.save_model()
        shutil.copy(cfg.save_path, OUT / "ref_checkpoint.pth")
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)
    probe = rng.standard_normal(8).astype(np.float32)
    sd = ag.net.state_dict()
    np.savez_compressed(OUT / "ref_checkpoint_expect.npz", probe=probe,
                        probe_normalized=np.asarray(ag.state_norm(probe, update=False), np.float64),
                        norm_n=ag.state_norm.running_ms.n, norm_mean=np.asarray(ag.state_norm.running_ms.mean, np.float64),
                        norm_std=np.asarray(ag.state_norm.running_ms.std, np.float64), norm_S=np.asarray(ag.state_norm.running_ms.S, np.float64),
                        rs_R=np.asarray(ag.reward_scaler.R, np.float64), rs_std=np.asarray(ag.reward_scaler.running_ms.std, np.float64),
                        rs_n=ag.reward_scaler.running_ms.n, learn_step=ag.learn_step,
                        net_out=ag.net(torch.as_tensor(x[:4])).detach().numpy(), x4=x[:4],
                        adam_step=float(ag.optimizer.state_dict()["state"][0]["step"]),
                        adam_exp_avg0=ag.optimizer.state_dict()["state"][0]["exp_avg"].numpy(),
                        **{"w_" + k.replace(".", "_"): v.numpy() for k, v in sd.items()},
                        source="utils/model.py:337-349 ModelLoader.save_model; utils/normalization.py")
    print("wrote ref_checkpoint.pth", (OUT / "ref_checkpoint.pth").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
