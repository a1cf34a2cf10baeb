mkdir -p gpurun_out
J=gpurun_out/r2_j21
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread ) > ${J}_pytest.log 2>&1
tail -n 12 ${J}_pytest.log | cut -c1-300
( timeout 900 python bench.py > ${J}_bench.json 2> ${J}_bench.err ) ; echo "bench rc $?"; tail -c 1200 ${J}_bench.json; tail -n 3 ${J}_bench.err
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${J}_bench_ref.json 2> ${J}_bench_ref.err ); echo "ref rc $?"; tail -c 400 ${J}_bench_ref.json
( timeout 900 python tools/config_bench.py > ${J}_configs.json 2> ${J}_configs.err ); echo "configs rc $?"; tail -c 600 ${J}_configs.json
# launch list of the same bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${J}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline ) > ${J}_ncu_l.log 2>&1
python tools/launch_summary.py ${J}_launches.csv > ${J}_launches.txt 2>&1; head -n 30 ${J}_launches.txt | cut -c1-200
# one full-set capture of the top kernels on the short fixed workload
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gain|k_relax|k_backtrack|k_pred|k_inval|k_weights" -c 12 -o ${J}_full python tools/prof_run.py 4096 65536 3 ) > ${J}_ncu_f.log 2>&1
python tools/ncu_summary.py ${J}_full.ncu-rep > ${J}_full.txt 2>&1; head -n 40 ${J}_full.txt | cut -c1-160
