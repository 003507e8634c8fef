o=gpurun_out; mkdir -p $o
python -c "
import mrvi_b200
from mrvi_b200 import _capi
for S,L in ((128,30),(32,30),(256,50)):
    d=mrvi_b200.MrviDims(n_input=2000,n_sample=S,n_latent=L); h=_capi.Handle(d, mrvi_b200.init_params(d,seed=0), device=0, precision=_capi.PRECISION_TC); print(S, h.attention_table_info()); h.close()
"
bash tools/gpu_ab.sh "A=1" "MRVI_B200_NO_ATT_TABLE=1" 2>&1 | tee $o/r2c_ab_table.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu -k "tc" > $o/r2c_pytest_tc.log 2>&1; echo "tc tests exit $?"; tail -3 $o/r2c_pytest_tc.log
