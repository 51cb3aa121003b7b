#!/bin/bash
# Round-2 single-GPU evidence run (on the GPU box, from the repo root): tools/run_n1_evidence.sh
# 1. the whole -m gpu suite + smoke, 2. the default bench line (+ torch-profile with the stream-gap report), 3. the ncu launch
# list of the bench command (shares, not absolutes), 4. ONE `ncu --set full` pass over one launch of every dominant kernel
# (tools/microbench.py --profile-once), summarised on the box, 5. the microbenchmark timings. Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 150 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout 400 python bench.py --torch-profile gpurun_out/r2_torchprof_n1.txt > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/r2_bench_n1.json
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2_launches_bench_n1.csv \
  python bench.py --steps 1 --warmup 1 --skip-parity --skip-no-recompute --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/step_roofline.py gpurun_out/r2_launches_bench_n1.csv > gpurun_out/r2_step_roofline.txt 2>&1; tail -5 gpurun_out/r2_step_roofline.txt
python tools/launch_summary.py gpurun_out/r2_launches_bench_n1.csv > gpurun_out/r2_launches_summary.txt 2>&1
gzip -f gpurun_out/r2_launches_bench_n1.csv
K='attn_fwd_tc_kernel|attn_bwd_dq_n128|attn_bwd_dkdv_tc|group_gemm_swap_kernel|group_gemm_kernel|add_rmsnorm_fwd|rmsnorm_bwd_ring|qknorm_rope_fwd|qknorm_rope_bwd|cross_entropy_kernel|swiglu_fwd|moe_scatter|moe_gather'
VB200_SKIP_QUACK=1 timeout 420 ncu --set full --clock-control none --profile-from-start off -k "regex:$K" -c 40 -f -o gpurun_out/r2_topkernels \
  python tools/microbench.py --only moe,rmsnorm,rope,swiglu,loss,attention --profile-once > gpurun_out/r2_ncu_topkernels.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py gpurun_out/r2_topkernels.ncu-rep > gpurun_out/r2_topkernels_ncu.txt 2>&1; grep -c "^kernel:" gpurun_out/r2_topkernels_ncu.txt
ls -la gpurun_out/r2_topkernels.ncu-rep; [ $(stat -c %s gpurun_out/r2_topkernels.ncu-rep 2>/dev/null || echo 0) -gt 45000000 ] && rm -f gpurun_out/r2_topkernels.ncu-rep
VB200_SKIP_QUACK=1 timeout 300 python tools/microbench.py --only rmsnorm,rope,swiglu,loss,attention,moe > gpurun_out/r2_microbench.jsonl 2> gpurun_out/r2_microbench.err; echo "microbench rc=$?"
cut -c1-200 gpurun_out/r2_microbench.jsonl | grep -v "lib" | head -40
