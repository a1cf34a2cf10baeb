mkdir -p gpurun_out
J=gpurun_out/r2_j33
( timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_replay.py -m gpu -q --timeout 600 --timeout-method thread ) > ${J}_py.log 2>&1; tail -n 30 ${J}_py.log | cut -c1-300
