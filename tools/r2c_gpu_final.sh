#!/bin/bash
# last call(s) of round 2: the whole GPU suite with the shipping default, A/B of the solver variants given as arguments,
# the default bench, an ncu capture of the shipping LunarLander step kernel.
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
T=${TAG:-r2c3}
t0=$(date +%s); stamp() { echo "[$(( $(date +%s) - t0 ))s] $*" | tee -a $O/${T}_timeline.log; }
stamp "full GPU suite (library default solver)"
timeout 300 python -m pytest tests -m gpu -x -q > $O/${T}_pytest_gpu.log 2>&1; stamp "rc=$?"
for v in "$@"; do
  stamp "env_cycles solver=$v"; GYMRL_LL_SOLVER=$v timeout 100 python tools/env_cycles.py > $O/${T}_env_cycles_v$v.log 2>&1; stamp "rc=$?"
  stamp "phase_times solver=$v"; GYMRL_LL_SOLVER=$v timeout 100 python tools/phase_times.py > $O/${T}_phase_times_v$v.log 2>&1; stamp "rc=$?"
done
stamp "bench (default)"
timeout 200 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; stamp "rc=$?"
if [ "${C5:-0}" = "1" ]; then
stamp "bench c5 (PPO-full shard)"
timeout 120 python bench.py --config c5 > $O/${T}_bench_c5.json 2> $O/${T}_bench_c5.err; stamp "rc=$?"
fi
stamp "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; stamp "rc=$?"
if [ "${NCU:-1}" = "1" ]; then
stamp "ncu lunar step (default solver)"
timeout 150 ncu --profile-from-start off --set full --import-source on -k regex:lunar_step -c 1 -o $O/${T}_lunar_step python tools/env_profile.py > $O/${T}_ncu.log 2>&1; stamp "rc=$?"
fi
stamp done
tail -3 $O/${T}_pytest_gpu.log; for v in "$@"; do head -3 $O/${T}_env_cycles_v$v.log | cut -c1-300; tail -1 $O/${T}_phase_times_v$v.log | cut -c1-160; done; cat $O/${T}_bench.json | cut -c1-400; [ -f $O/${T}_bench_c5.json ] && cut -c1-300 $O/${T}_bench_c5.json; tail -2 $O/${T}_smoke.log
