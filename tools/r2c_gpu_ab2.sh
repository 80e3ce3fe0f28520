#!/bin/bash
# second A/B call: solver variants 0 / 1 / 2 — parity tests, then per-env cycle counters of each, twice (box noise).
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
t0=$(date +%s); stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/r2c2_timeline.log; }
stamp "env parity tests"
timeout 240 python -m pytest tests/test_gpu_envs.py -x -q > $O/r2c2_envs_pytest.log 2>&1; stamp "rc=$?"
for rep in a b; do for v in 0 2 3 1; do
  stamp "env_cycles solver=$v rep=$rep"
  GYMRL_LL_SOLVER=$v timeout 120 python tools/env_cycles.py > $O/r2c2_env_cycles_v${v}_$rep.log 2>&1; stamp "rc=$?"
done; done
for v in 0 2 3; do
  stamp "phase_times solver=$v"
  GYMRL_LL_SOLVER=$v timeout 150 python tools/phase_times.py > $O/r2c2_phase_times_v$v.log 2>&1; stamp "rc=$?"
done
stamp done
tail -2 $O/r2c2_envs_pytest.log; for f in $O/r2c2_env_cycles_v*; do echo $f; head -3 $f; done; tail -1 $O/r2c2_phase_times_v0.log; tail -1 $O/r2c2_phase_times_v2.log; tail -1 $O/r2c2_phase_times_v3.log
