mkdir -p gpurun_out
J=gpurun_out/r2_j32
( timeout 900 python bench.py > ${J}_bench.json 2> ${J}_bench.err ) ; echo "bench rc $?"
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${J}_bench_ref.json 2> ${J}_bench_ref.err ); echo "ref rc $?"
( timeout 900 python tools/config_bench.py > ${J}_configs.json 2> ${J}_configs.err ); echo "configs rc $?"
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${J}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline ) > ${J}_ncu_l.log 2>&1
python tools/launch_summary.py ${J}_launches.csv > ${J}_launches.txt 2>&1; head -n 8 ${J}_launches.txt | cut -c1-170
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gain|k_relax|k_backtrack|k_pred|k_inval|k_weights" -c 12 -o ${J}_full python tools/prof_run.py 4096 65536 3 ) > ${J}_ncu_f.log 2>&1
python tools/ncu_summary.py ${J}_full.ncu-rep > ${J}_full.txt 2>&1; grep -c "== kernel" ${J}_full.txt
python -c "
import json
d=json.loads(open('${J}_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cycle']['plan_ms'], d['cycle']['ms_per_full_cycle'], d['cycle']['incremental']['replan_ms_median'], d['roofline']['frac'])"
