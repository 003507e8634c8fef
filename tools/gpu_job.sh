o=gpurun_out; mkdir -p $o
timeout 120 ./tools/mma_microbench > $o/r2_mma_microbench.txt 2>&1; echo "microbench exit $?"
timeout 900 python -m pytest tests/test_gpu_de.py -x -q > $o/r2b_pytest_de.log 2>&1; echo "de exit $?"; tail -15 $o/r2b_pytest_de.log
timeout 600 python -m pytest tests/test_gpu_scale.py -x -q -k "C5 and tc" -s > $o/r2b_pytest_c5.log 2>&1; echo "c5 exit $?"; tail -5 $o/r2b_pytest_c5.log
timeout 1200 python tools/fuzz_parity.py 220 7 > $o/r2_fuzz_parity_220.log 2>&1; echo "fuzz exit $?"; tail -4 $o/r2_fuzz_parity_220.log
