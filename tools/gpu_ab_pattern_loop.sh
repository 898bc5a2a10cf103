#!/bin/bash
# A/B of the pattern sweep's pair form (interleaved vs particle-major), fused-vs-chain diagnostic, GPU tests
set -u
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); r = d.get('roofline', {}); e = d.get('e2e') or {}
        print('  ms/step', round(d['ms_per_step'], 2), 'e2e', round(e.get('ms_per_step') or 0, 2), 'seg us', round(r.get('avg_launch_us', 0), 1), 'frac', round(r.get('frac', 0), 3),
              'step frac', round(r['whole_step']['frac'], 3), {k: round(v, 1) for k, v in r.get('by_kernel_ms', {}).items()}, d.get('moment_chain'))
PY
}
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_ab_pair.log 2>&1; echo "c4 interleaved rc=$?"; summ gpurun_out/bench_c4_ab_pair.log
HPSIM_B200_LIB=$PWD/hpsim_b200/csrc/_exp/libhpsim_b200_pm.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_ab_pm.log 2>&1; echo "c4 particle-major rc=$?"; summ gpurun_out/bench_c4_ab_pm.log
timeout 300 python tools/fused_vs_chain.py 2097152 > gpurun_out/fused_vs_chain_2mi.log 2>&1; echo "fvc rc=$?"; tail -25 gpurun_out/fused_vs_chain_2mi.log
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c2_n1.log 2>&1; echo "c2 n1 rc=$?"; summ gpurun_out/bench_c2_n1.log
[ "${SKIP_TESTS:-0}" = "1" ] || timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-400 | tail -12
