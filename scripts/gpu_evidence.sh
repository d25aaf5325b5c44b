#!/bin/bash
# Round-2 evidence in one GPU call: `ncu --set full` of every kernel family bench.py reports (alone, at the bench shapes), summarised ON
# the box (the reports together exceed the 64 MiB that travel back), DRAM-traffic table for bench.py, and the launch list of one
# training step of cfg 2 and cfg 3.  Outputs: gpurun_out/r02_ncu_full_summary.txt, r02_traffic.json, r02_launches_step_cfg{2,3}.txt
mkdir -p gpurun_out
KERNELS="attn_fwd attn_bwd attn_fwd_drop attn_bwd_drop gemm_q gemm_fc1 gemm_fc2 ln_fwd ln_bwd wgrad_q wgrad_fc1 wgrad_fc2 wgrad_o wgrad_kv attn128_fwd attn128_bwd" bash scripts/gpu_ncu_full.sh 2>&1 | grep rc
KERNELS="attn128_fwd attn128_bwd" PB=8 SUFFIX=_b8 bash scripts/gpu_ncu_full.sh 2>&1 | grep rc
python scripts/ncu_summary.py "gpurun_out/full_*.ncu-rep" > gpurun_out/r02_ncu_full_summary.txt 2>&1
python scripts/ncu_traffic.py > gpurun_out/r02_traffic.log 2>&1 && cp profiles/r02_traffic.json gpurun_out/r02_traffic.json
# keep the two attention reports (source pages), drop the rest
mkdir -p gpurun_out/keep && mv gpurun_out/full_attn_bwd.ncu-rep gpurun_out/full_attn_fwd.ncu-rep gpurun_out/keep/ 2>/dev/null
rm -f gpurun_out/full_*.ncu-rep; mv gpurun_out/keep/* gpurun_out/ 2>/dev/null; rmdir gpurun_out/keep
for CFG in cfg2 cfg3; do
  LPS_FALLBACK=2400 COUNT=2400 SKIP=7200 PROF_ARGS="--config $CFG" bash scripts/gpu_profile.sh > gpurun_out/prof_$CFG.log 2>&1
  cp gpurun_out/launch_summary.txt gpurun_out/r02_launches_step_$CFG.txt; rm -f gpurun_out/launches.csv
done
tail -5 gpurun_out/r02_traffic.log; head -30 gpurun_out/r02_launches_step_cfg2.txt
