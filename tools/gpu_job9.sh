#!/bin/bash
# 8 GPUs: monitor + sharded tests, strong-scaling C4 lines at N = 1, 2, 4, 8 (exchange wait times in the line)
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests -m gpu -q -s -k "monitor or sharded" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-600 | tail -12
summ() { python - "$1" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); r = d.get('roofline', {}); e = d.get('e2e') or {}
        print('  ms/step', round(d['ms_per_step'], 2), 'e2e', round(e.get('ms_per_step') or 0, 2), 'seg us', round(r.get('avg_launch_us', 0), 1), 'frac', round(r.get('frac', 0), 3),
              'step frac', round(r['whole_step']['frac'], 3), {k: round(v, 1) for k, v in r.get('by_kernel_ms', {}).items()}, d['config']['run'], d.get('result'), d.get('exchange'), d.get('sharded_check'))
PY
}
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_c4_n1.log 2>&1; echo "c4 n1 rc=$?"; summ gpurun_out/bench_c4_n1.log
for N in 2 4 8; do
  CHK="--no-check"; [ $N = 8 ] && CHK=""
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 2 --warmup 1 --no-cpu $CHK > gpurun_out/bench_c4_n$N.log 2>&1; echo "c4 n$N rc=$?"; summ gpurun_out/bench_c4_n$N.log
done
grep -iE "error|Traceback" gpurun_out/bench_c4_n*.log | head -5
du -sh gpurun_out
