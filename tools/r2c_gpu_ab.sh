#!/bin/bash
# Round 2, third session: one gpurun call that (1) holds solver variant 1 to the oracle, (2) A/Bs it against variant 0
# (per-env cycle counters, phase times), (3) runs the whole GPU suite and the bench with variant 1, (4) captures ncu.
# Every leg has its own timeout; outputs land in gpurun_out/r2c_*.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/r2c_timeline.log; }
stamp "env parity tests (both solver variants)"
timeout 240 python -m pytest tests/test_gpu_envs.py -x -q > $O/r2c_envs_pytest.log 2>&1; stamp "rc=$?"
for v in 0 1; do
  stamp "env_cycles solver=$v"
  GYMRL_LL_SOLVER=$v timeout 120 python tools/env_cycles.py > $O/r2c_env_cycles_v$v.log 2>&1; stamp "rc=$?"
done
for v in 0 1; do
  stamp "phase_times solver=$v"
  GYMRL_LL_SOLVER=$v timeout 150 python tools/phase_times.py > $O/r2c_phase_times_v$v.log 2>&1; stamp "rc=$?"
done
stamp "full GPU suite, solver=1"
GYMRL_LL_SOLVER=1 timeout 300 python -m pytest tests -m gpu -x -q > $O/r2c_pytest_gpu_v1.log 2>&1; stamp "rc=$?"
stamp "bench solver=1"
GYMRL_LL_SOLVER=1 timeout 200 python bench.py > $O/r2c_bench_v1.json 2> $O/r2c_bench_v1.err; stamp "rc=$?"
stamp "ncu lunar step solver=1"
GYMRL_LL_SOLVER=1 timeout 150 ncu --profile-from-start off --set full --import-source on -k regex:lunar_step -c 1 -o $O/r2c_lunar_step_v1 python tools/env_profile.py > $O/r2c_ncu.log 2>&1; stamp "rc=$?"
stamp done
tail -3 $O/r2c_envs_pytest.log; tail -3 $O/r2c_pytest_gpu_v1.log; head -3 $O/r2c_env_cycles_v0.log; head -3 $O/r2c_env_cycles_v1.log; cat $O/r2c_phase_times_v0.log | tail -2; cat $O/r2c_phase_times_v1.log | tail -2; cat $O/r2c_bench_v1.json
