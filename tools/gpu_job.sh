o=gpurun_out; mkdir -p $o
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-api --no-parity"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_fused -s 3 -c 1 -f -o $o/r2k_fused_c4 $B > $o/r2k_ncu_full_fused.log 2>&1
echo "ncu fused exit $?"
