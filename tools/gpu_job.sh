mkdir -p gpurun_out
J=gpurun_out/r2_j40
for P in 0 1 0 1; do
  ( SIPP_RELAX_PAPE_BATCH=$P SIPP_BATCH_THREADS=4 timeout 300 python tools/ensemble_bench.py --maps 128 --cycles 16 --mode batch --batches 4 ) > ${J}_ens_p$P.json 2> ${J}_ens_p$P.err
  python -c "import json; d=json.loads(open('${J}_ens_p$P.json').read().strip().splitlines()[-1]); print('pape_batch $P:', round(d['value']), d['ms_per_map_cycle_in_ensemble'])" || tail -n 3 ${J}_ens_p$P.err
done
