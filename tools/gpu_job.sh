mkdir -p gpurun_out
J=gpurun_out/r2_j27
( timeout 900 python -m pytest tests/test_gpu_views.py -m gpu -x -q --timeout 600 --timeout-method thread ) > ${J}_views.log 2>&1
tail -n 40 ${J}_views.log | cut -c1-400
