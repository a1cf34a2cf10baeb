mkdir -p gpurun_out
J=gpurun_out/r2_j38
( timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread ) > ${J}_pytest.log 2>&1
tail -n 6 ${J}_pytest.log | cut -c1-300
for W in 8192 4096 2048; do
  ( timeout 400 python tools/relax_ab.py $W "0,SIPP_RELAX_PAPE=0" "0,SIPP_RELAX_PAPE=1" ) > ${J}_ab$W.log 2>&1
  grep -o "W [0-9]* \|PAPE=[01]\|plan median [0-9.]* ms (min [0-9.]*)\|'tile_activations': [0-9]*\|same_field=[A-Za-z]*" ${J}_ab$W.log | tr '\n' ' '; echo
done
( SIPP_RELAX_PAPE=1 timeout 200 python tools/replan_bench.py 4096 ) > ${J}_replan4096.log 2>&1; tail -n 1 ${J}_replan4096.log | cut -c1-300
( SIPP_RELAX_PAPE=0 timeout 200 python tools/replan_bench.py 4096 ) > ${J}_replan4096_0.log 2>&1; tail -n 1 ${J}_replan4096_0.log | cut -c1-300
