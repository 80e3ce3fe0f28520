"""In-tree build of libgymrl_b200.so (sm_100a only) with nvcc — no torch.utils.cpp_extension, no JIT cache.

    python -m gymrl_b200.build            # incremental
    python -m gymrl_b200.build --force    # rebuild everything
    python -m gymrl_b200.build --ptxas    # print register / spill / smem usage (-Xptxas -v)

The shared object lands in gymrl_b200/lib/ (git-ignored, but shipped to the GPU box by gpurun).
Translation units that must round exactly like the CPU oracle (env physics, the utils-dialect GAE)
are compiled with -fmad=false.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB_DIR = ROOT / "lib"
OBJ_DIR = LIB_DIR / "obj"
LIB_PATH = LIB_DIR / "libgymrl_b200.so"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-strict-aliasing",
          "-Xcompiler", "-ffp-contract=off", "--expt-relaxed-constexpr"]
# sources -> extra flags
SOURCES = {
    "env.cu": ["-fmad=false"],
    "env_lunar.cu": ["-fmad=false"],
    "gae.cu": ["-fmad=false"],
    "sample.cu": [],
    "ppo_loss.cu": [],
    "linear.cu": [],
    "linear_tc.cu": [],
    "linear_skinny.cu": [],
    "optim.cu": [],
    "reduce.cu": [],
    "comm.cu": [],
    "recurrent.cu": [],
    "wimages.cu": [],
    "replay.cu": [],
    "qlearn.cu": [],
    "mhc.cu": [],
    "normalize.cu": ["-fmad=false"],
}


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: gymrl_b200 has no CPU fallback and cannot be built without the CUDA toolkit")


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "gymrl.h", Path(__file__)]
    return max(p.stat().st_mtime for p in hdrs if p.exists())


def build(force: bool = False, verbose: bool = False, ptxas: bool = False) -> Path:
    nvcc = nvcc_path()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    dep_m = _deps_mtime()
    jobs = []
    objs = []
    for src, extra in SOURCES.items():
        sp = CSRC / src
        if not sp.exists():
            continue
        op = OBJ_DIR / (sp.stem + ".o")
        objs.append(op)
        if force or ptxas or not op.exists() or op.stat().st_mtime < max(sp.stat().st_mtime, dep_m):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", str(sp), "-o", str(op)]
            if ptxas:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append((src, cmd))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r.returncode, r.stdout + r.stderr, " ".join(cmd)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            results = list(ex.map(run, jobs))
        failed = False
        for name, rc, out, cmdline in results:
            if verbose or ptxas or rc != 0:
                print(f"--- {name}: {cmdline}\n{out}", file=sys.stderr)
            if rc != 0:
                failed = True
        if failed:
            raise RuntimeError("nvcc compilation failed (see stderr)")
    if jobs or not LIB_PATH.exists():
        cmd = [nvcc, *ARCH, "-shared", "-o", str(LIB_PATH), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout + r.stderr, file=sys.stderr)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--ptxas", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose, ptxas=a.ptxas))
