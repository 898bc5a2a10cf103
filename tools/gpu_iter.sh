#!/bin/bash
# One GPU iteration: A/B of the environment-switchable variants through bench.py (short runs), then the
# parity tests.  Usage under gpurun:  VARIANTS="name:ENV=V,ENV=V ..." tools/gpu_iter.sh
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
summ() {
python - "$1" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); r = d.get('roofline', {})
        e = d.get('e2e') or {}
        print('  ms/step', round(d['ms_per_step'], 2), 'e2e ms', round(e.get('ms_per_step') or 0, 2), 'seg us', round(r.get('avg_launch_us', 0), 1),
              'frac', round(r.get('frac', 0), 3), {k: round(v, 2) for k, v in r.get('by_kernel_ms', {}).items()}, 'launches', d['config']['run']['launches'],
              {k: round(v, 2) for k, v in r.get('scheff_phase_ms', {}).items()})
PY
}
for V in ${VARIANTS:-default:}; do
  NAME=${V%%:*}; ENVS=${V#*:}
  echo "== $NAME ($ENVS)"
  env $(echo $ENVS | tr ',' ' ') timeout ${BENCH_TIMEOUT:-300} python bench.py --steps ${STEPS:-3} --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_$NAME.log 2>&1
  echo "  rc=$?"; summ gpurun_out/bench_$NAME.log; grep -iE "error|Traceback" gpurun_out/bench_$NAME.log | head -3
done
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  env ${TEST_ENV:-} timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q -x ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?"
  grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-700 | tail -20
fi
du -sh gpurun_out
