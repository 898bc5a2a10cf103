#!/bin/bash
# round 2, first GPU job: new fixtures from the reference, reference timing sweep on the full linac, the new parity tests,
# the C4 bench line and ncu captures of the kick-chain kernels at 16 Mi particles
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 900 python tests/make_golden.py > gpurun_out/golden.log 2>&1; echo "golden rc=$?"; tail -8 gpurun_out/golden.log
cp gpurun_out/golden/* tests/golden/ 2>/dev/null
for n in 65536 262144; do
  timeout 900 python bench.py --impl reference --config c4 --ref-particles $n --steps 1 --warmup 1 >> gpurun_out/ref_sweep.log 2>&1
done
echo "ref sweep:"; cut -c1-400 gpurun_out/ref_sweep.log
timeout 1500 python -m pytest tests/test_gpu_full_linac.py -m gpu -q -s > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed|FULL_NOSC|SC_ON|SHARDED" gpurun_out/pytest_full.log | cut -c1-1500 | tail -30
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_c4.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_c4.log | cut -c1-3000
timeout 300 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2.log 2>&1; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_c2.log | cut -c1-1500
for K in k_segment k_moments k_deposit k_field; do
  HPSIM_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 300 -c 3 \
     -f -o gpurun_out/r2a_16mi_$K python tools/prof_run.py 16777216 TBDB02 02QM10 > gpurun_out/ncu_full_$K.log 2>&1
  echo "ncu full $K rc=$?"
done
du -sh gpurun_out
