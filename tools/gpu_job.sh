mkdir -p gpurun_out
J=gpurun_out/r2_j50
SIPP_BATCH_ROAM=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file ${J}_ens_launches.csv python tools/ensemble_bench.py --mode batch --maps 256 --cycles 8 --batches 1 > ${J}_ncu.log 2>&1
echo rc=$?
