mkdir -p gpurun_out
timeout 170 ncu --set full --clock-control none --import-source on -k regex:k_relax_roam_g -s 3 -c 1 -f -o gpurun_out/r2_roam_full python tools/ensemble_bench.py --mode batch --maps 64 --cycles 4 --batches 1 > gpurun_out/r2_roam_full.log 2>&1
echo rc=$?; ls -la gpurun_out/r2_roam_full.ncu-rep
