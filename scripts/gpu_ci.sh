#!/bin/bash
# Runs on the GPU box (via gpurun): isolated tcgen05 bring-up first (own process + timeout), then the GPU test suites.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== gemm (isolated)" ; timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "gemm" > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_gemm.log
echo "== ops (non-gemm)" ; timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "not gemm" > gpurun_out/t_ops.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_ops.log
echo "== engine" ; timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu > gpurun_out/t_engine.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_engine.log
