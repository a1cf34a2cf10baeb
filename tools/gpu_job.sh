mkdir -p gpurun_out
J=gpurun_out/r2_j18
( timeout 600 python -m pytest tests/test_gpu_batch.py -m gpu -x -q --timeout 300 --timeout-method thread ) > ${J}_batch.log 2>&1
tail -n 5 ${J}_batch.log | cut -c1-400
if grep -q " passed" ${J}_batch.log && ! grep -q failed ${J}_batch.log; then
  for T in 1 4 8; do
    ( SIPP_BATCH_THREADS=$T timeout 300 python tools/ensemble_bench.py --maps 128 --cycles 20 --mode batch --check 1 ) > ${J}_ens_batch_t$T.json 2> ${J}_ens_batch_t$T.err
    python -c "import json,sys; d=json.loads(open('${J}_ens_batch_t$T.json').read().strip().splitlines()[-1]); print('threads $T', d['value'], d['ms_per_map_cycle_in_ensemble'], d['kernel_launches_per_batch_cycle'])"
  done
  ( SIPP_BATCH_THREADS=4 timeout 300 python tools/ensemble_bench.py --maps 512 --cycles 10 --mode batch --check 0 ) > ${J}_ens_batch_m512.json 2> ${J}_ens_batch_m512.err
  python -c "import json,sys; d=json.loads(open('${J}_ens_batch_m512.json').read().strip().splitlines()[-1]); print('m512', d['value'], d['ms_per_map_cycle_in_ensemble'])"
fi
