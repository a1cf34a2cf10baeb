mkdir -p gpurun_out
J=gpurun_out/r2_j12
export SIPP_BLK_RERUN=0
( timeout 300 python tools/relax_ab.py 4096 0 1 ) > ${J}_ab4096.log 2>&1
( timeout 400 python tools/relax_ab.py 8192 0 1 ) > ${J}_ab8192.log 2>&1
( timeout 200 python tools/relax_ab.py 1024 0 1 ) > ${J}_ab1024.log 2>&1
( timeout 200 python tools/relax_ab.py 640 0 1 ) > ${J}_ab640.log 2>&1
cat ${J}_ab4096.log ${J}_ab8192.log ${J}_ab1024.log ${J}_ab640.log | cut -c1-400
