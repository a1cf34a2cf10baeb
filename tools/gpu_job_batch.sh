mkdir -p gpurun_out
J=gpurun_out/r2_j59
for nb in 3 4; do
SIPP_BENCH_ENSEMBLE_BATCHES=$nb SIPP_BATCH_THREADS=8 timeout 200 python bench.py --workload ensemble > ${J}_ensemble_n1_b$nb.json 2> ${J}_ensemble_n1_b$nb.err; echo "batches $nb rc $?"; cut -c1-130 ${J}_ensemble_n1_b$nb.json
done
