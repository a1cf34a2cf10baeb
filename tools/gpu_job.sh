mkdir -p gpurun_out
J=gpurun_out/r2_j4
( timeout 300 python tools/blk_debug.py 130 256 640 1024 ) > ${J}_blk.log 2>&1
cat ${J}_blk.log
K=0
if ! grep -q "failed\|TIMEOUT\|same_field=False\|differs" ${J}_blk.log; then
  ( timeout 300 python tools/relax_ab.py 4096 0 1 ) > ${J}_ab4096.log 2>&1
  ( timeout 400 python tools/relax_ab.py 8192 0 1 ) > ${J}_ab8192.log 2>&1
  ( timeout 200 python tools/relax_ab.py 1024 0 1 ) > ${J}_ab1024.log 2>&1
  ( timeout 200 python tools/relax_ab.py 640 0 1 ) > ${J}_ab640.log 2>&1
  cat ${J}_ab4096.log ${J}_ab8192.log ${J}_ab1024.log ${J}_ab640.log
  if grep -q "same_field=True" ${J}_ab4096.log 2>/dev/null && grep -q "same_field=True" ${J}_ab8192.log; then K=1; fi
  if [ $K = 1 ]; then
    ( SIPP_RELAX_KERNEL=1 timeout 200 python tools/replan_bench.py 4096 ) > ${J}_replan4096_k1.log 2>&1
    ( SIPP_RELAX_KERNEL=1 timeout 200 python tools/replan_bench.py 640 ) > ${J}_replan640_k1.log 2>&1
    tail -n 3 ${J}_replan4096_k1.log ${J}_replan640_k1.log
  fi
fi
echo "pytest with SIPP_RELAX_KERNEL=$K"
( SIPP_RELAX_KERNEL=$K timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --timeout-method thread ) > ${J}_pytest.log 2>&1
tail -n 25 ${J}_pytest.log | cut -c1-400
( SIPP_RELAX_KERNEL=$K timeout 600 python bench.py > ${J}_bench.json 2> ${J}_bench.err ); tail -c 3000 ${J}_bench.json
( timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > ${J}_bench_ref.json 2> ${J}_bench_ref.err ); tail -c 1500 ${J}_bench_ref.json
