"""Generate tests/golden/normalization.npz by running the UNMODIFIED reference classes
(/root/reference/utils/normalization.py: Normalization :25-35, RewardScaling :38-52) on seeded streams, fed exactly as
utils/runner.py:112,125-126 feeds them.  TEST INFRASTRUCTURE; build container only:  python -m oracle.make_golden_utils
"""
from pathlib import Path

import numpy as np

from . import ref_loader as rl

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def main():
    m = rl.load("utils/normalization.py")
    rng = np.random.default_rng(0)
    T, D = 300, 8
    x = (rng.standard_normal((T, D)) * np.array([1, 2, 0.5, 3, 1, 0.1, 1, 1]) + np.array([0, 1, -1, 0, 5, 0, 0, 0])).astype(np.float32)
    norm = m.Normalization(shape=(D,))
    y = np.stack([np.asarray(norm(x[t]), np.float64) for t in range(T)])
    y_eval = np.asarray(norm(x[0], update=False), np.float64)          # evaluation path: no update
    r = rng.standard_normal(T).astype(np.float64) * 3.0
    reset_at = np.zeros(T, np.uint8); reset_at[[0, 57, 58, 200]] = 1
    rs = m.RewardScaling(shape=1, gamma=0.99)
    out = np.zeros(T)
    for t in range(T):
        if reset_at[t]:
            rs.reset()
        out[t] = rs(float(r[t]))[0]
    np.savez_compressed(OUT / "normalization.npz", x=x, y=y, y_eval=y_eval, mean=np.asarray(norm.running_ms.mean, np.float64),
                        std=np.asarray(norm.running_ms.std, np.float64), S=np.asarray(norm.running_ms.S, np.float64), n=norm.running_ms.n,
                        r=r, reset_at=reset_at, r_scaled=out, gamma=0.99,
                        source="utils/normalization.py:4-52 driven as in utils/runner.py:109-126")
    print("wrote normalization.npz")


if __name__ == "__main__":
    main()
