#!/bin/bash
# One GPU-box pass that refreshes the evidence under gpurun_out/ (copied to profiles/ by hand afterwards):
#   GPU test suite, the default bench line, the ncu launch list of the bench command, one `ncu --set full` capture of the
#   dominant kernel on the benchmark launch, and (with FULL=1) BASELINE configs 3 and 5 at full size.
# usage (under gpurun): bash tools/gpu_round.sh [tag]
tag=${1:-r2}
o=gpurun_out
mkdir -p $o
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $o/${tag}_pytest_gpu.log
  tail -3 $o/${tag}_pytest_gpu.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > $o/${tag}_bench_c4.json 2> $o/${tag}_bench_c4.err; echo "bench exit $?"
head -c 1500 $o/${tag}_bench_c4.json; echo
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-api --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_c4.csv $B > $o/${tag}_ncu_launch.log 2>&1
echo "launch list exit $?"
if [ -z "$SKIP_NCU_FULL" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_fused -s 3 -c 1 -f -o $o/${tag}_fused_c4 $B > $o/${tag}_ncu_full_fused.log 2>&1
echo "ncu fused exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_encoder_tc -s 3 -c 1 -f -o $o/${tag}_encoder_c4 $B > $o/${tag}_ncu_full_enc.log 2>&1
echo "ncu encoder exit $?"
fi
if [ -n "$FULL" ]; then
  timeout 900 python tools/run_full_configs.py c3 c5 > $o/${tag}_full_configs.jsonl 2> $o/${tag}_full_configs.err; echo "full exit $?"
  cat $o/${tag}_full_configs.jsonl
fi
