mkdir -p gpurun_out
J=gpurun_out/r2_j22
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > ${J}_smoke.log 2>&1; tail -n 3 ${J}_smoke.log | cut -c1-300
( timeout 300 ./tests/cpp/test_modules gpu ) > ${J}_modules.log 2>&1; grep -n "timing\|failure" ${J}_modules.log | cut -c1-250
