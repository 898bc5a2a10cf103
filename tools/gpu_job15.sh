#!/bin/bash
set -u
mkdir -p gpurun_out
NG=8
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $1 tools/sharded_bisect.py 16777216 > gpurun_out/$2 2>&1; echo "$2 rc=$?"; grep "SHARDED_BISECT\|probe\|first differing\|phi \|lloss \|<phi>\|one element\|HpsError" gpurun_out/$2 | tail -30; }
run 29512 sharded_bisect_8.log
