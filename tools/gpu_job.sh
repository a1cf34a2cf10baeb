mkdir -p gpurun_out
J=gpurun_out/r2_j43
for W in 2048 4096 8192; do
timeout 600 python tools/relax_ab.py $W 0 0,SIPP_RELAX_PREFETCH=1 0 0,SIPP_RELAX_PREFETCH=1 > ${J}_ab$W.log 2>&1; cut -c1-230 ${J}_ab$W.log
done
