mkdir -p gpurun_out
J=gpurun_out/r2_j54
timeout 1500 python -m pytest tests -m gpu -x -q > ${J}_tests.log 2>&1; echo "tests rc $?"; tail -n 3 ${J}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${J}_smoke.log 2>&1; echo "smoke rc $?"; tail -n 1 ${J}_smoke.log
timeout 600 python bench.py > ${J}_bench_n1.json 2> ${J}_bench_n1.err; echo "bench rc $?"; cut -c1-300 ${J}_bench_n1.json
SIPP_BATCH_THREADS=8 timeout 600 python bench.py --workload ensemble > ${J}_ensemble_n1.json 2> ${J}_ensemble_n1.err; echo "rc $?"; cut -c1-260 ${J}_ensemble_n1.json
SIPP_BENCH_ENSEMBLE_MAPS=1024 SIPP_BATCH_THREADS=8 timeout 600 python bench.py --workload ensemble > ${J}_ensemble_n1_m1024.json 2> ${J}_ensemble_n1_m1024.err; echo "rc $?"; cut -c1-260 ${J}_ensemble_n1_m1024.json
