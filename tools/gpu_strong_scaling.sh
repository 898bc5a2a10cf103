#!/bin/bash
# strong scaling of the headline configuration (C4, one 16 Mi-particle beam sharded over N GPUs) + sharded-result checks
set -u
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); r = d.get('roofline', {}); e = d.get('e2e') or {}
        print('  N', d['n_gpus'], 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(e.get('ms_per_step') or 0, 2), 'seg us', round(r.get('avg_launch_us', 0), 1),
              {k: round(v, 1) for k, v in r.get('by_kernel_ms', {}).items()}, d.get('exchange', {}).get('small_wait_us_avg'), d.get('exchange', {}).get('mesh_wait_us_avg'),
              (d.get('sharded_check') or {}).get('ok'), (d.get('sharded_check') or {}).get('short_range', {}).get('n_lloss'))
PY
}
for N in ${NLIST:-2 4 8}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_bench_${N}gpu.log 2>&1; echo "c4 n$N rc=$?"; summ gpurun_out/r2_bench_${N}gpu.log
done
grep -iE "HpsError|Traceback" gpurun_out/r2_bench_*gpu.log | head -5
