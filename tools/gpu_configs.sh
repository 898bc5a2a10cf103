#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed|ROW_PAST" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --config c3 --steps 3 --warmup 3 > gpurun_out/r2_bench_c3_1gpu.log 2>&1; echo "c3 rc=$?"; grep '^{' gpurun_out/r2_bench_c3_1gpu.log | cut -c1-300
timeout 600 python bench.py --config c1 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_c1_1gpu.log 2>&1; echo "c1 rc=$?"; grep '^{' gpurun_out/r2_bench_c1_1gpu.log | cut -c1-300
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_c2_1gpu.log 2>&1; echo "c2 rc=$?"; grep '^{' gpurun_out/r2_bench_c2_1gpu.log | cut -c1-300
