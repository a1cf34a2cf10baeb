mkdir -p gpurun_out
J=gpurun_out/r2_j20
( timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q --timeout 300 --timeout-method thread ) > ${J}_batch.log 2>&1
tail -n 3 ${J}_batch.log | cut -c1-400
for cfg in "128 1 8" "128 2 4" "128 4 4" "512 2 8" "512 4 4" "1024 4 4"; do
  set -- $cfg
  ( SIPP_BATCH_THREADS=$3 timeout 300 python tools/ensemble_bench.py --maps $1 --cycles 16 --mode batch --batches $2 --check 1 ) > ${J}_ens_m$1_b$2.json 2> ${J}_ens_m$1_b$2.err
  python -c "import json,sys; d=json.loads(open('${J}_ens_m$1_b$2.json').read().strip().splitlines()[-1]); print('maps $1 batches $2 threads $3:', round(d['value']), d['ms_per_map_cycle_in_ensemble'])" || tail -n 3 ${J}_ens_m$1_b$2.err
done
