o=gpurun_out; mkdir -p $o
python tools/time_launch_overhead.py 2>&1 | tee $o/r2_time_launch_overhead.txt
