#!/bin/bash
# Round-2 single-GPU evidence run (on the GPU box, from the repo root): tools/run_n1_evidence.sh
# 1. the whole -m gpu suite + smoke, 2. the default bench line, 3. microbench (default + 2-CTA GroupGEMM), 4. the ncu launch
# list of the bench command and --set full captures of the dominant kernels. Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 420 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_n1.json
VB200_SKIP_QUACK=1 timeout 400 python tools/microbench.py --only rmsnorm,rope,swiglu,loss,fsdp,attention,moe > gpurun_out/r2_microbench.jsonl 2> gpurun_out/r2_microbench.err; echo "microbench rc=$?"
VB200_GG_2CTA=1 timeout 200 python tools/microbench.py --only moe > gpurun_out/r2_microbench_gg2cta.jsonl 2> gpurun_out/r2_microbench_gg2cta.err
echo "== 2-CTA GroupGEMM"; cat gpurun_out/r2_microbench_gg2cta.jsonl | cut -c1-220; tail -3 gpurun_out/r2_microbench_gg2cta.err
echo "== default GroupGEMM"; grep group_gemm gpurun_out/r2_microbench.jsonl | cut -c1-220
# ncu: launch list of the bench command (shares, not absolutes), then full captures of the top kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2_launches_bench_n1.csv \
  python bench.py --steps 1 --warmup 1 --skip-parity --skip-no-recompute > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu list rc=$?"
gzip -f gpurun_out/r2_launches_bench_n1.csv
for k in attn_bwd_dkdv attn_bwd_dq_n128 attn_fwd_tc; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r2_$k \
    python tools/microbench.py --only attention --iters 2 > gpurun_out/r2_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:group_gemm_swap -c 3 -f -o gpurun_out/r2_group_gemm_swap \
  python tools/microbench.py --only moe --iters 2 > gpurun_out/r2_ncu_gg.log 2>&1; echo "ncu gg rc=$?"
ls -la gpurun_out/*.ncu-rep
