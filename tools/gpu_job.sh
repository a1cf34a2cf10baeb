mkdir -p gpurun_out
J=gpurun_out/r2_j57
SIPP_BATCH_THREADS=2 timeout 600 python bench.py --workload ensemble > ${J}_ensemble_n1_t2.json 2> ${J}_ensemble_n1_t2.err; echo "rc $?"; cut -c1-130 ${J}_ensemble_n1_t2.json
