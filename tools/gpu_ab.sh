#!/bin/bash
# A/B of library builds / fused-kernel versions: device-resident bench lines (ms per step, fused stage ms)
# usage: gpu_ab.sh <lib|-> [FUSED version] ...   ("-" = the in-tree library)
while [ $# -gt 0 ]; do
  lib=$1; ver=${2:-3}; shift; shift
  for w in "c2" "c4 --cells-per-gpu 20000" "c3 --cells-per-gpu 20000"; do
    if [ "$lib" = "-" ]; then L=""; else L=$PWD/$lib; fi
    MRVI_B200_FUSED=$ver MRVI_B200_LIBRARY=$L timeout 100 python bench_old.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib v$ver $w', round(d['ms_per_step'],3), round(d['roofline']['stage_ms_last_step']['qz_chain'],3))"
  done
done
