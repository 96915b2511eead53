#!/bin/bash
# refresh of the headline workload's artefacts only (tools/measure_r02.sh does all four workloads)
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_hru_step|k_cell_output' --launch-skip 60 --launch-count 4 -o $O/r02_full_fe_hourly -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_ncu_fe.log 2>&1
export PROFILE_OUT=$O/prof
cp profiles/traffic.json $O/prof/traffic.json 2>/dev/null
python tools/summarize_profile.py $O/r02_launches.csv $O/r02_full_fe_hourly.ncu-rep r02 fe_hourly 10000 k_hru_step_nn3 > $O/r02_summarize_fe.log 2>&1
cp $O/prof/traffic.json profiles/traffic.json
python bench.py > $O/r02_bench_fe_hourly_full_year.json 2> $O/r02_bench_fe.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_fe_hourly_reference_arm.json 2> $O/r02_ref.err
for f in $O/r02_bench_fe_hourly_*.json; do echo $f; head -c 250 $f; echo; done
