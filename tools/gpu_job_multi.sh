# usage: bash tools/gpu_job_multi.sh N [what]     what: all (default) | config4 | ensemble
N=$1
WHAT=${2:-all}
mkdir -p gpurun_out
J=gpurun_out/r2_multi_n$N
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
if [ "$WHAT" = "all" ]; then
  ( run bench.py --gpus $N --steps 200 --warmup 10 > ${J}_bench.json 2> ${J}_bench.err ); echo "bench rc $?"; tail -c 700 ${J}_bench.json
fi
if [ "$WHAT" = "ensemble" ]; then
  ( SIPP_BATCH_THREADS=2 run bench.py --workload ensemble --gpus $N > ${J}_ensemble.json 2> ${J}_ensemble.err ); echo "ensemble rc $?"; cut -c1-400 ${J}_ensemble.json
  exit 0
fi
( SIPP_BENCH_WORKLOAD=config4 run bench.py --gpus $N --steps 100 --warmup 10 > ${J}_config4.json 2> ${J}_config4.err ); echo "config4 rc $?"; tail -c 500 ${J}_config4.json
if [ "$N" = "8" ] && [ "$WHAT" = "all" ]; then
  ( SIPP_BATCH_THREADS=2 run tools/ensemble_bench.py --maps 128 --cycles 16 --mode batch --batches 2 > ${J}_ens.jsonl 2> ${J}_ens.err ); echo "ens rc $?"
  python -c "
import json
v=[json.loads(l) for l in open('${J}_ens.jsonl') if l.startswith('{')]
print('ranks', len(v), 'sum map-cycles/s', sum(x['value'] for x in v), [round(x['value']) for x in v])"
fi
