#!/bin/bash
set -u
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); r = d.get('roofline', {}); t = d['moment_chain']['sweep_timeline_us']
        print('  ms/step', round(d['ms_per_step'], 2), 'seg us', round(r.get('avg_launch_us', 0), 1), {k: round(v, 1) for k, v in r.get('by_kernel_ms', {}).items()},
              'sweep', round(t['sweep'], 1), 'fold', round(t['fold'], 1), 'fin', round(t['finalise'], 1), [round(x, 1) for x in t['finalise_phases']], 'pilot', round(t['pilot'], 1))
PY
}
timeout 600 python bench.py --particles 2097152 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_ef2_2097152.log 2>&1; echo "evict-first+deposit 2Mi rc=$?"; summ gpurun_out/bench_c4_ef2_2097152.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -6
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.log 2>&1; echo "bench default rc=$? wall=$(( $(date +%s) - T0 )) s"; summ gpurun_out/r2_bench_1gpu.log
