mkdir -p gpurun_out
J=gpurun_out/r2_j53
run() { # warps maps batches share threads
SIPP_ROAM_WARPS=$1 SIPP_BATCH_TIMING=1 SIPP_BATCH_SHARE=$4 SIPP_BATCH_THREADS=$5 timeout 300 python tools/ensemble_bench.py --mode batch --maps $2 --cycles 16 --batches $3 > ${J}_ens_w$1_m$2_b$3_s$4.json 2> ${J}_ens_w$1_m$2_b$3_s$4.err
echo "warps=$1 maps=$2 batches=$3 share=$4 threads=$5: $(python -c "import json,sys; d=json.loads(open('${J}_ens_w$1_m$2_b$3_s$4.json').read().strip().splitlines()[-1]); print(d['value'])" 2>&1 | tail -n 1)"
grep "batch of" ${J}_ens_w$1_m$2_b$3_s$4.err | tail -n 1 | cut -c150-300
}
run 16 512 2 2 8
run 20 512 2 2 8
run 24 512 2 2 8
run 16 512 1 1 8
run 24 512 1 1 8
run 24 128 2 2 8
run 16 128 2 2 8
