#!/bin/bash
set -u
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import sys, json
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); r = d.get('roofline', {})
        print('  ms/step', round(d['ms_per_step'], 2), 'seg us', round(r.get('avg_launch_us', 0), 1), 'frac', round(r.get('frac', 0), 3),
              'step frac', round(r['whole_step']['frac'], 3), {k: round(v, 1) for k, v in r.get('by_kernel_ms', {}).items()}, d.get('result'))
PY
}
E=$PWD/hpsim_b200/csrc/_exp
HPSIM_B200_LIB=$E/libhpsim_b200_ppt1_3.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_ppt1_3.log 2>&1; echo "ppt1 x3 rc=$?"; summ gpurun_out/bench_c4_ppt1_3.log
HPSIM_PATTERNS=4 HPSIM_B200_LIB=$E/libhpsim_b200_ppt1_4.so timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_ppt1_4.log 2>&1; echo "ppt1 x4 rc=$?"; summ gpurun_out/bench_c4_ppt1_4.log
HPSIM_B200_LIB=$E/libhpsim_b200_ppt1_3.so timeout 600 python bench.py --particles 2097152 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_2mi_ppt1_3.log 2>&1; echo "2mi ppt1 x3 rc=$?"; summ gpurun_out/bench_c4_2mi_ppt1_3.log
timeout 600 python bench.py --particles 2097152 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_2mi.log 2>&1; echo "2mi default rc=$?"; summ gpurun_out/bench_c4_2mi.log
