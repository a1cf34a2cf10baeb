mkdir -p gpurun_out
J=gpurun_out/r2_j30
( AB_TIMING=1 timeout 300 python tools/relax_ab.py 4096 0 ) > ${J}_t4096.log 2>&1; grep "plan phases" ${J}_t4096.log | tail -n 3; grep -o "plan median [0-9.]* ms (min [0-9.]*)" ${J}_t4096.log
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 --timeout-method thread ) > ${J}_parity.log 2>&1; tail -n 4 ${J}_parity.log | cut -c1-300
