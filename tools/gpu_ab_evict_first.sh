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
E=$PWD/hpsim_b200/csrc/_exp
for P in 2097152 16777216; do
  S=2; [ $P = 16777216 ] && S=1
  timeout 600 python bench.py --particles $P --steps $S --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_ef_$P.log 2>&1; echo "evict-first $P rc=$?"; summ gpurun_out/bench_c4_ef_$P.log
  HPSIM_B200_LIB=$E/libhpsim_b200_noef.so timeout 600 python bench.py --particles $P --steps $S --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_c4_noef_$P.log 2>&1; echo "plain $P rc=$?"; summ gpurun_out/bench_c4_noef_$P.log
done
