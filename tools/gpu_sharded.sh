#!/bin/bash
# GPU box with N >= 2 GPUs: sharded-vs-single parity, then the sharded bench at N ranks (and replicas).
set -u
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for MODE in ${MODES:-p2p nccl}; do
  HPSIM_XR=$MODE timeout 600 $TR --master-port 29511 tests/sharded_check.py ${CHECK_N:-65536} TBDB02 ${CHECK_LAST:-01QM10} > gpurun_out/sharded_check_$MODE.log 2>&1
  echo "sharded_check $MODE rc=$?"; grep -E "^SHARDED|Error|error" gpurun_out/sharded_check_$MODE.log | cut -c1-1500 | head -5
done
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  for MODE in ${MODES:-p2p nccl}; do
    HPSIM_XR=$MODE timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-3} --warmup 3 --no-cpu > gpurun_out/bench_${N}gpu_$MODE.log 2>&1
    echo "bench $N gpus $MODE rc=$?"; tail -1 gpurun_out/bench_${N}gpu_$MODE.log | cut -c1-2500
  done
  timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps ${STEPS:-3} --warmup 3 --no-cpu --replicas > gpurun_out/bench_${N}gpu_replicas.log 2>&1
  echo "bench $N gpus replicas rc=$?"; tail -1 gpurun_out/bench_${N}gpu_replicas.log | cut -c1-1200
fi
du -sh gpurun_out
