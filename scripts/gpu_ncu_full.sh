#!/bin/bash
# full ncu capture of the hot kernels in isolation (scripts/prof_kernel.py), after a plain run of the same command
# KERNELS="..." selects them; PB=<batch> overrides the batch (default 64); SUFFIX=_b8 names the reports
mkdir -p gpurun_out
for K in ${KERNELS:-attn_fwd attn_bwd gemm_q gemm_fc1 gemm_fc2 ln_fwd}; do
  python scripts/prof_kernel.py $K 3 $PB > gpurun_out/plain_$K.log 2>&1 || { echo "plain $K failed"; tail -5 gpurun_out/plain_$K.log; continue; }
  case $K in attn128_fwd*) RX=attn128_fwd_kernel;; attn128_bwd*) RX=attn128_bwd_kernel;; attn_fwd*) RX=attn_fwd_tc_kernel;; attn_bwd*) RX=attn_bwd_tc_kernel;; gemm*|wgrad*) RX=gemm_tc_kernel;; ln_fwd) RX=ln_fwd_kernel;; ln_bwd) RX=ln_bwd;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$RX -s 1 -c 1 -f -o gpurun_out/full_$K$SUFFIX python scripts/prof_kernel.py $K 3 $PB > gpurun_out/ncu_$K.log 2>&1
  echo "$K ncu rc=$?"; ls -la gpurun_out/full_$K$SUFFIX.ncu-rep 2>/dev/null | awk '{print $5}'
done
