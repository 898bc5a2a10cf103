#!/bin/bash
# final single-GPU evidence of the round: GPU tests, the default bench line, the reference arm, ncu launch list + full capture
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-400 | tail -8
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.log 2>&1; echo "bench default rc=$? wall=$(( $(date +%s) - T0 )) s"
grep '^{' gpurun_out/r2_bench_1gpu.log | cut -c1-600
T0=$(date +%s); timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.log 2>&1; echo "bench reference rc=$? wall=$(( $(date +%s) - T0 )) s"
grep '^{' gpurun_out/r2_bench_reference.log | cut -c1-400
HPSIM_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 36200 --csv --log-file gpurun_out/r2_launches.csv \
   python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/r2_ncu_list.log 2>&1; echo "ncu list rc=$?"
wc -l gpurun_out/r2_launches.csv
HPSIM_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_segment -s 1400 -c 2 \
   -f -o gpurun_out/r2c_16mi_ccl_k_segment python tools/prof_run.py 16777216 TBDB02 07DT > gpurun_out/ncu_full_k_segment.log 2>&1; echo "ncu full rc=$?"
du -sh gpurun_out
