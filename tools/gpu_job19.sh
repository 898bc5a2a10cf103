#!/bin/bash
set -u
mkdir -p gpurun_out
E=$PWD/hpsim_b200/csrc/_exp
HPSIM_B200_LIB=$E/libhpsim_b200_ppt1_3.so HPSIM_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:4408332 -s 40 -c 2 \
   -f -o gpurun_out/r2c_16mi_ccl_ppt1x3 python tools/prof_run.py 16777216 TBDB02 07DT > gpurun_out/ncu_full_ppt1x3.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_ppt1x3.log
