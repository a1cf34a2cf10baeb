mkdir -p gpurun_out
J=gpurun_out/r2_j26
( timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -q --timeout 300 --timeout-method thread ) > ${J}_batch.log 2>&1
tail -n 30 ${J}_batch.log | cut -c1-300
