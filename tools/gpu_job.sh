o=gpurun_out; mkdir -p $o
bash tools/gpu_ab.sh "A=1" 2>&1 | tee $o/r2h_ab.txt
timeout 1500 python -m pytest tests -x -q -m gpu -s > $o/r2h_pytest_gpu.log 2>&1; echo "gpu tests exit $?"; tail -3 $o/r2h_pytest_gpu.log; grep "\[scale\]" $o/r2h_pytest_gpu.log
