#!/bin/bash
# Two-GPU validation (on the GPU box): multi-GPU parity worker + collective bandwidths, the default bench line at N=2, and the
# single-GPU checks of the kernels touched last (tests + microbench). Outputs under gpurun_out/.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29511 tests/multigpu_worker.py --bench > gpurun_out/r2_comm_n2b.log 2>&1
echo "worker ok ranks: $(grep -o 'WORKER OK rank [0-9]' gpurun_out/r2_comm_n2b.log | sort -u | wc -l)"; grep "ulysses_a2a\|Error\|error" gpurun_out/r2_comm_n2b.log | cut -c1-200 | head -8
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --skip-ab --skip-no-recompute > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err
echo "bench rc=$?"; tail -c 1200 gpurun_out/r2_bench_n2b.json; grep -i "error\|Traceback" gpurun_out/r2_bench_n2b.err | head -5
timeout 200 python -m pytest tests/test_ops_gpu.py tests/test_moe_gpu.py -q -k "rmsnorm or rope or fused_add or weight_grad" 2>&1 | tail -3
timeout 100 python tools/microbench.py --only rmsnorm,rope 2>&1 | cut -c1-150 | tee gpurun_out/r2_microbench_hbm_final.jsonl
echo "== bulk-staged forward for comparison"; VB200_RMS_FWD_CFG=12,2 timeout 100 python tools/microbench.py --only rmsnorm 2>&1 | grep "rmsnorm_fwd\[4096" | cut -c1-150
