#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -${TAILN:-12} gpurun_out/$name.log; }
run t_attn 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "xattn" -x
run t_engine 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -s
run t_trainer 900 python -m pytest tests/test_trainer_gpu.py -q -m gpu
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
