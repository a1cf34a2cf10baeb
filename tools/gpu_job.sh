mkdir -p gpurun_out
J=gpurun_out/r2_j36
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread ) > ${J}_pytest.log 2>&1
tail -n 6 ${J}_pytest.log | cut -c1-300
