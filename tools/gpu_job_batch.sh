mkdir -p gpurun_out
J=gpurun_out/r2_j58
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_replay.py -x -q -m gpu > ${J}_tests.log 2>&1; echo "tests rc $?"; tail -n 2 ${J}_tests.log
SIPP_BATCH_THREADS=8 timeout 300 python bench.py --workload ensemble > ${J}_ensemble_n1.json 2> ${J}_ensemble_n1.err; echo "rc $?"; cut -c1-130 ${J}_ensemble_n1.json
SIPP_BATCH_THREADS=2 timeout 300 python bench.py --workload ensemble > ${J}_ensemble_n1_t2.json 2> ${J}_ensemble_n1_t2.err; echo "rc $?"; cut -c1-130 ${J}_ensemble_n1_t2.json
