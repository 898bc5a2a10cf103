#!/bin/bash
set -u
mkdir -p gpurun_out
HPSIM_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:65804 -s 20 -c 2 \
   -f -o gpurun_out/r2c_16mi_drift_k_segment python tools/prof_run.py 16777216 TBDB02 spch_comp_7 > gpurun_out/ncu_full_k_segment_drift.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_k_segment_drift.log
