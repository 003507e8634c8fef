#!/bin/bash
# timeline + device-resident bench of the fused kernel (run under gpurun); prints one line per workload
timeout 60 python tools/dbg/run_dbg_f3.py 128 0 > gpurun_out/dbg_f3_S128_grp.log 2>&1
timeout 60 python tools/dbg/run_dbg_f3.py 128 1 > gpurun_out/dbg_f3_S128_keep.log 2>&1
for w in "c2" "c4 --cells-per-gpu 20000" "c3 --cells-per-gpu 20000"; do
  timeout 100 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', round(d['ms_per_step'],3), d['roofline']['stage_ms_last_step'], round(d['roofline']['cuda_core_bound']['sm_clk_per_row'],1))"
done
