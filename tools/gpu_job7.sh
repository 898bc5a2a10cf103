#!/bin/bash
set -u
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); r = d.get('roofline', {}); e = d.get('e2e') or {}
        print('  ms/step', round(d['ms_per_step'], 2), 'e2e', round(e.get('ms_per_step') or 0, 2), 'seg us', round(r.get('avg_launch_us', 0), 1), 'frac', round(r.get('frac', 0), 3),
              'step frac', round(r['whole_step']['frac'], 3), {k: round(v, 1) for k, v in r.get('by_kernel_ms', {}).items()}, d['config']['run'], d.get('result'))
PY
}
for V in ${VARIANTS:-pat1:HPSIM_PATTERNS=1 pat2:HPSIM_PATTERNS=2}; do
  NAME=${V%%:*}; ENVS=${V#*:}
  echo "== c4 $NAME ($ENVS)"
  env $(echo $ENVS | tr ',' ' ') timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_$NAME.log 2>&1; echo "  rc=$?"; summ gpurun_out/bench_c4_$NAME.log
  echo "== c2 $NAME ($ENVS)"
  env $(echo $ENVS | tr ',' ' ') timeout 300 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c2_$NAME.log 2>&1; echo "  rc=$?"; summ gpurun_out/bench_c2_$NAME.log
done
grep -iE "error|Traceback" gpurun_out/bench_*.log | head -5
timeout 2400 python -m pytest tests -m gpu -q -s ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed|MONITOR_LLOSS" gpurun_out/pytest_gpu.log | cut -c1-1500 | tail -20
du -sh gpurun_out
