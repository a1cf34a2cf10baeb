mkdir -p gpurun_out
J=gpurun_out/r2_j39
for W in 4096 8192 640; do
  ( SIPP_LIB_PATH=$PWD/scouting-ipp_b200/libsipp_prev.so timeout 400 python tools/relax_ab.py $W 0 ) > ${J}_prev$W.log 2>&1
  ( AB_KEEP_REF=1 timeout 400 python tools/relax_ab.py $W 0 ) > ${J}_new$W.log 2>&1
  grep -o "W [0-9]* \|plan median [0-9.]* ms (min [0-9.]*)\|same_field=[A-Za-z]*" ${J}_prev$W.log ${J}_new$W.log | tr '\n' ' '; echo
done
