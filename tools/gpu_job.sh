mkdir -p gpurun_out
J=gpurun_out/r2_j56
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q -m gpu > ${J}_batch_tests.log 2>&1; echo "tests rc $?"; tail -n 2 ${J}_batch_tests.log
for wp in 16 20; do
SIPP_ROAM_WARPS=$wp SIPP_BATCH_THREADS=8 timeout 600 python bench.py --workload ensemble > ${J}_ensemble_n1_w$wp.json 2> ${J}_ensemble_n1_w$wp.err; echo "warps $wp rc $?"; cut -c1-130 ${J}_ensemble_n1_w$wp.json
SIPP_ROAM_WARPS=$wp SIPP_BENCH_ENSEMBLE_MAPS=1024 SIPP_BATCH_THREADS=8 timeout 600 python bench.py --workload ensemble > ${J}_ensemble_n1_m1024_w$wp.json 2> ${J}_ensemble_n1_m1024_w$wp.err; echo "warps $wp m1024 rc $?"; cut -c1-130 ${J}_ensemble_n1_m1024_w$wp.json
done
