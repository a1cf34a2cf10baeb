mkdir -p gpurun_out
J=gpurun_out/r2_j55
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q -m gpu > ${J}_tests.log 2>&1; echo "tests rc $?"; tail -n 3 ${J}_tests.log
timeout 600 python bench.py --no-cpu-baseline > ${J}_bench_n1.json 2> ${J}_bench_n1.err; echo "bench rc $?"; python -c "
import json; d=json.load(open('${J}_bench_n1.json')); print(d['cycle']['incremental'])"
SIPP_BATCH_THREADS=8 timeout 600 python bench.py --workload ensemble > ${J}_ensemble_n1.json 2> ${J}_ensemble_n1.err; echo "rc $?"; cut -c1-160 ${J}_ensemble_n1.json
SIPP_BENCH_ENSEMBLE_MAPS=1024 SIPP_BATCH_THREADS=8 timeout 600 python bench.py --workload ensemble > ${J}_ensemble_n1_m1024.json 2> ${J}_ensemble_n1_m1024.err; echo "rc $?"; cut -c1-160 ${J}_ensemble_n1_m1024.json
