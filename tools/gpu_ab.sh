#!/bin/bash
# A/B of the device-resident bench (ms per step, fused stage, parity) under different environment switches.
# usage (under gpurun): [WL="c4 c2"] bash tools/gpu_ab.sh "" "MRVI_B200_LIBRARY=build_ab/x.so" ...     (one argument = one environment prefix)
for envs in "$@"; do
  for w in ${WL:-c4 c2 c3}; do
    env $envs timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-api 2>>gpurun_out/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d.get('parity') or {}
print('[$envs] $w ms/step', round(d['ms_per_step'],3), 'stages', {k: round(v,3) for k,v in d['roofline']['stage_ms_last_step'].items()}, 'clk/row', round(d['roofline']['cuda_core_bound']['sm_clk_per_row'],1), 'parity', p.get('max_norm_err'), p.get('p99_elementwise_rel'), p.get('ok'))"
  done
done
