mkdir -p gpurun_out
J=gpurun_out/r2_j28
( timeout 300 ./tests/cpp/test_modules gpu ) > ${J}_modules.log 2>&1; grep -n "FAIL\|sampled\|failure" ${J}_modules.log | cut -c1-300 | head -40
( timeout 900 python -m pytest tests/test_plugin_modules.py tests/test_gpu_views.py -m gpu -q --timeout 600 --timeout-method thread ) > ${J}_py.log 2>&1; tail -n 5 ${J}_py.log | cut -c1-300
