#!/bin/bash
# launch list of one training step (eager, so every kernel is visible to ncu), then a full capture of the named kernel
mkdir -p gpurun_out
export BPM_NO_GRAPH=1
CMD="python bench.py --steps 1 --warmup 3 --no-kernels --no-cpu ${PROF_ARGS}"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/prof_plain.log; exit 1; }
tail -1 gpurun_out/prof_plain.log | cut -c1-300
LPS=$(python -c "import json,sys; print(json.loads(open('gpurun_out/prof_plain.log').read().strip().splitlines()[-1])['launches_per_step'] or ${LPS_FALLBACK:-6886})" 2>/dev/null || echo ${LPS_FALLBACK:-6886})
echo "launches/step (eager count) = $LPS"
SKIP=${SKIP:-$((3 * ${LPS_FALLBACK:-6886}))}
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c ${COUNT:-${LPS_FALLBACK:-6886}} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/prof_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/prof_ncu.log | cut -c1-200
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
