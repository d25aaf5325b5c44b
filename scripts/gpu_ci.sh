#!/bin/bash
# Runs on the GPU box (via gpurun): GPU test suites (each in its own process + timeout), smoke, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
run t_ops 900 python -m pytest tests/test_ops_gpu.py -q -m gpu
run t_engine 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s
run t_trainer 900 python -m pytest tests/test_trainer_gpu.py -q -m gpu
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
if [ "${BENCH:-1}" = "1" ]; then TAILN=3 run bench 1500 python bench.py --steps ${STEPS:-3} --warmup 3 ${BENCH_ARGS}; fi
