# the round's verification job: bash tools/gpu_job.sh on a B200 box (gpurun) — GPU tests, smoke, the default bench line
mkdir -p gpurun_out
J=gpurun_out/r2_final
timeout 1500 python -m pytest tests -m gpu -x -q > ${J}_tests.log 2>&1; echo "tests rc $?"; tail -n 2 ${J}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${J}_smoke.log 2>&1; echo "smoke rc $?"; tail -n 1 ${J}_smoke.log
timeout 600 python bench.py > ${J}_bench_n1.json 2> ${J}_bench_n1.err; echo "bench rc $?"; cut -c1-200 ${J}_bench_n1.json
