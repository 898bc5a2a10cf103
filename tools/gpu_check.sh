#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, one bench line, the ncu launch list and one
# full capture of the dominant kernel.  Small artefacts land in gpurun_out/ (<= 64 MiB in total).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-600 | tail -30
fi
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.log 2>&1
  echo "bench rc=$?"; tail -3 gpurun_out/bench.log
fi
if [ "${SKIP_REF:-0}" != "1" ]; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
  echo "bench ref rc=$?"; tail -2 gpurun_out/bench_ref.log
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  # launch list of the whole run (lead-in, warm-up step, timed step, profiled step); tools/launch_summary.py
  # summarises the LAST C2 pass (1 626 launches)
  HPSIM_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-9000} --csv \
     --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-fast > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"
  for K in ${NCU_KERNELS:-k_segment}; do
    HPSIM_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${NCU_FULL_SKIP:-700} -c ${NCU_FULL_COUNT:-3} \
       -f -o gpurun_out/prof_$K python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-fast > gpurun_out/ncu_full_$K.log 2>&1
    echo "ncu full $K rc=$?"
  done
fi
du -sh gpurun_out; ls -la gpurun_out
