#!/bin/bash
# Round-2 multi-GPU evidence run (on the GPU box, from the repo root): tools/run_multigpu_evidence.sh N
# 1. parity worker + collective bandwidths, 2. the default bench line (parity, comm, NCCL A/B inside) + torch profile,
# 3. BASELINE configs[2] (Ulysses SP4 @32k) and configs[3] (Qwen3-30B-A3B, EP = N) lines. Outputs under gpurun_out/.
N=${1:-8}
mkdir -p gpurun_out
export VB200_SYMM_BYTES=$((2 << 30))
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 280 $TR --master-port 29511 tests/multigpu_worker.py --bench > gpurun_out/r2_comm_n$N.log 2>&1
if ! grep -q "WORKER OK rank 0" gpurun_out/r2_comm_n$N.log; then
  echo "worker failed with CTA-pair clusters: retrying with VB200_COMM_CLUSTER=0"
  export VB200_COMM_CLUSTER=0
  timeout 280 $TR --master-port 29521 tests/multigpu_worker.py --bench > gpurun_out/r2_comm_n${N}_nocluster.log 2>&1
fi
timeout 330 $TR --master-port 29512 bench.py --gpus $N --steps 6 --warmup 3 --skip-no-recompute --torch-profile gpurun_out/r2_torchprof_n$N.txt \
  > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
timeout 300 $TR --master-port 29513 bench.py --gpus $N --workload ulysses32k --steps 3 --warmup 2 --skip-ab \
  > gpurun_out/r2_bench_ulysses32k_n$N.json 2> gpurun_out/r2_bench_ulysses32k_n$N.err
timeout 420 $TR --master-port 29514 bench.py --gpus $N --workload moe30b --steps 3 --warmup 2 --skip-ab \
  > gpurun_out/r2_bench_moe30b_n$N.json 2> gpurun_out/r2_bench_moe30b_n$N.err
grep -c "WORKER OK" gpurun_out/r2_comm_n$N.log
grep "copy-out\|push\|nccl\|ulysses_a2a" gpurun_out/r2_comm_n$N.log | cut -c1-160
for f in gpurun_out/r2_bench_n$N.json gpurun_out/r2_bench_ulysses32k_n$N.json gpurun_out/r2_bench_moe30b_n$N.json; do
  echo "== $f"; tail -c 900 $f; echo
done
tail -c 400 gpurun_out/r2_bench_ulysses32k_n$N.err; tail -c 400 gpurun_out/r2_bench_moe30b_n$N.err
