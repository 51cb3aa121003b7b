#!/bin/bash
# Launch-shape sweep of the HBM-bound kernels on the GPU box (tools/hbm_sweep.sh): every configuration in its own process
# (the knobs are read once per process). Output: gpurun_out/r2_hbm_sweep.txt
mkdir -p gpurun_out
out=gpurun_out/r2_hbm_sweep.txt; : > $out
for cfg in 12,2 8,2 6,2 4,2 8,3 6,4 4,4 3,8; do
  echo "== VB200_RMS_FWD_CFG=$cfg" >> $out
  VB200_RMS_FWD_CFG=$cfg timeout 100 python tools/microbench.py --only rmsnorm 2>&1 | grep "rmsnorm_fwd\[4096\|rror" | cut -c1-150 >> $out
done
for st in 6 4 3 2; do
  echo "== VB200_RMS_BWD_STAGES=$st" >> $out
  VB200_RMS_BWD_STAGES=$st timeout 100 python tools/microbench.py --only rmsnorm 2>&1 | grep "rmsnorm_bwd\|rror" | cut -c1-150 >> $out
done
for cfg in 2,2 4,2 2,0 4,0 2,4 4,4; do
  echo "== VB200_ROPE_CFG=$cfg" >> $out
  VB200_ROPE_CFG=$cfg timeout 100 python tools/microbench.py --only rope 2>&1 | grep "rope\|rror" | cut -c1-150 >> $out
done
VB200_RMS_FWD_CFG=4,4 VB200_RMS_BWD_STAGES=3 VB200_ROPE_CFG=4,0 timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "rmsnorm or rope or fused_add" 2>&1 | tail -3 >> $out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "rmsnorm or rope or fused_add" 2>&1 | tail -3 >> $out
cat $out
