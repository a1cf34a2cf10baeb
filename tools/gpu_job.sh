mkdir -p gpurun_out
J=gpurun_out/r2_j23
for W in 4096 8192 1024 640; do
  ( timeout 400 python tools/relax_ab.py $W "0,SIPP_RELAX_CHECK=0" "0,SIPP_RELAX_CHECK=1" ) > ${J}_ab$W.log 2>&1
  cut -c1-330 ${J}_ab$W.log
done
( SIPP_RELAX_CHECK=1 timeout 200 python tools/replan_bench.py 4096 ) > ${J}_replan4096.log 2>&1; tail -n 3 ${J}_replan4096.log | cut -c1-300
