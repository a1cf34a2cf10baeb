mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_j1_smi.log 2>&1
( timeout 600 python tools/relax_ab.py 4096 0 1 ) > gpurun_out/r2_j1_ab4096.log 2>&1
( timeout 900 python tools/relax_ab.py 8192 0 1 ) > gpurun_out/r2_j1_ab8192.log 2>&1
( timeout 300 python tools/relax_ab.py 1024 0 1 ) > gpurun_out/r2_j1_ab1024.log 2>&1
( timeout 300 python tools/relax_ab.py 640 0 1 ) > gpurun_out/r2_j1_ab640.log 2>&1
( SIPP_RELAX_KERNEL=1 timeout 300 python tools/replan_bench.py 4096 ) > gpurun_out/r2_j1_replan4096_k1.log 2>&1
( SIPP_RELAX_KERNEL=0 timeout 300 python tools/replan_bench.py 4096 ) > gpurun_out/r2_j1_replan4096_k0.log 2>&1
( SIPP_RELAX_KERNEL=1 timeout 300 python tools/replan_bench.py 640 ) > gpurun_out/r2_j1_replan640_k1.log 2>&1
( SIPP_RELAX_KERNEL=0 timeout 300 python tools/replan_bench.py 640 ) > gpurun_out/r2_j1_replan640_k0.log 2>&1
( timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --timeout-method thread ) > gpurun_out/r2_j1_pytest.log 2>&1
tail -5 gpurun_out/r2_j1_ab4096.log gpurun_out/r2_j1_ab8192.log gpurun_out/r2_j1_pytest.log
