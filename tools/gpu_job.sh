mkdir -p gpurun_out
J=gpurun_out/r2_j46
run() { # maps batches share threads
SIPP_BATCH_ROAM=1 SIPP_BATCH_SHARE=$3 SIPP_BATCH_THREADS=$4 timeout 300 python tools/ensemble_bench.py --mode batch --maps $1 --cycles 16 --batches $2 > ${J}_ens_m$1_b$2_s$3_t$4.json 2> ${J}_ens_m$1_b$2_s$3_t$4.err
echo "maps=$1 batches=$2 share=$3 threads=$4: $(python -c "import json,sys; d=json.loads(open('${J}_ens_m$1_b$2_s$3_t$4.json').read().strip().splitlines()[-1]); print(d['value'])" 2>&1 | tail -n 1)"
}
run 128 4 2 4
run 128 8 4 2
run 128 8 8 2
run 128 4 4 2
run 128 4 4 8
run 128 2 2 8
run 512 4 4 4
run 512 4 2 4
run 512 8 4 2
run 1024 4 4 4
