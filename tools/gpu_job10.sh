#!/bin/bash
set -u
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); r = d.get('roofline', {}); e = d.get('e2e') or {}
        print('  ms/step', round(d['ms_per_step'], 2), 'e2e', round(e.get('ms_per_step') or 0, 2), 'seg us', round(r.get('avg_launch_us', 0), 1), 'frac', round(r.get('frac', 0), 3),
              'step frac', round(r['whole_step']['frac'], 3), {k: round(v, 1) for k, v in r.get('by_kernel_ms', {}).items()}, d['config']['run'], d.get('moment_chain'), d.get('exchange'), d.get('sharded_check'))
PY
}
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_n1.log 2>&1; echo "c4 n1 rc=$?"; summ gpurun_out/bench_c4_n1.log
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c2_n1.log 2>&1; echo "c2 n1 rc=$?"; summ gpurun_out/bench_c2_n1.log
timeout 600 python bench.py --config c4 --particles 2097152 --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/bench_c4_2mi.log 2>&1; echo "c4 2Mi rc=$?"; summ gpurun_out/bench_c4_2mi.log
if [ "${NGPU:-1}" -gt 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $NGPU --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_c4_n$NGPU.log 2>&1; echo "c4 n$NGPU rc=$?"; summ gpurun_out/bench_c4_n$NGPU.log
fi
grep -iE "error|Traceback" gpurun_out/bench_c*.log | head -5
