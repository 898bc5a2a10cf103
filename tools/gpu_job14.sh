#!/bin/bash
set -u
mkdir -p gpurun_out
NG=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/sharded_bisect.py ${NPART:-4194304} > gpurun_out/sharded_bisect_${NG}.log 2>&1; echo "bisect rc=$?"
grep -v "^W\|^\[W\|warn" gpurun_out/sharded_bisect_${NG}.log | tail -40
