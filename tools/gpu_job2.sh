#!/bin/bash
# GPU job: full -m gpu suite, C4 + C2 bench lines (short), per-kernel split
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s -x ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed|FULL_NOSC|SC_ON|SHARDED|FUSED" gpurun_out/pytest_gpu.log | cut -c1-1200 | tail -40
if [ "${SKIP_BENCH:-0}" != "1" ]; then
timeout 900 python bench.py --steps ${STEPS:-2} --warmup 1 --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_c4.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_c4.log | cut -c1-3500
timeout 300 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c2.log 2>&1; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_c2.log | cut -c1-2500
fi
du -sh gpurun_out
