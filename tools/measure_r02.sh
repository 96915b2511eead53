#!/bin/bash
# The round's measurements of record (run under gpurun): bench lines of every workload, the reference arm, the ncu launch list of the
# headline command and one ncu --set full capture per workload's step kernel.  Summaries: tools/summarize_profile.py -> profiles/.
O=gpurun_out
# launch list of the headline command (it ran clean without ncu in the calls before this one: tools/exp*.sh)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_ncu_launches.log 2>&1
# one full capture per workload: two step launches + two cell-output launches of the headline workload, one step launch of the others
ncu --set full --clock-control none --import-source on --kernel-name regex:'k_hru_step|k_cell_output' --launch-skip 60 --launch-count 4 -o $O/r02_full_fe_hourly -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_ncu_fe.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_hru_step --launch-skip 30 --launch-count 1 -o $O/r02_full_glacier -f python bench.py --workload glacier --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_ncu_glacier.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_hru_step --launch-skip 30 --launch-count 1 -o $O/r02_full_continental -f python bench.py --workload continental --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_ncu_cont.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_hru_step --launch-skip 30 --launch-count 1 -o $O/r02_full_frozen_bands -f python bench.py --workload frozen_bands --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_ncu_frozen.log 2>&1
# summarise on the box (gpurun brings back at most 64 MiB): keep the headline report, drop the others
export PROFILE_OUT=$O/prof
python tools/summarize_profile.py $O/r02_launches.csv $O/r02_full_fe_hourly.ncu-rep r02 fe_hourly 10000 k_hru_step_nn3 > $O/r02_summarize_fe.log 2>&1
python tools/summarize_profile.py - $O/r02_full_glacier.ncu-rep r02_glacier glacier 250000 k_hru_step_nn3 > $O/r02_summarize_glacier.log 2>&1
python tools/summarize_profile.py - $O/r02_full_continental.ncu-rep r02_continental continental 125000 k_hru_step_nn3 > $O/r02_summarize_cont.log 2>&1
python tools/summarize_profile.py - $O/r02_full_frozen_bands.ncu-rep r02_frozen_bands frozen_bands 100000 k_hru_step_nn10 > $O/r02_summarize_frozen.log 2>&1
rm -f $O/r02_full_glacier.ncu-rep $O/r02_full_continental.ncu-rep $O/r02_full_frozen_bands.ncu-rep
# the bench lines carry the per-launch DRAM traffic and FP64 flop counts of the captures above (profiles/traffic.json)
cp $O/prof/traffic.json profiles/traffic.json
python bench.py > $O/r02_bench_fe_hourly_full_year.json 2> $O/r02_bench_fe.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_fe_hourly_reference_arm.json 2> $O/r02_ref.err
python bench.py --workload glacier > $O/r02_bench_glacier.json 2> $O/r02_bench_glacier.err
python bench.py --workload continental > $O/r02_bench_continental_1gpu.json 2> $O/r02_bench_cont.err
python bench.py --workload frozen_bands > $O/r02_bench_frozen_bands.json 2> $O/r02_bench_frozen.err
for f in $O/r02_bench_*.json; do echo $f; head -c 300 $f; echo; done
du -sh $O
