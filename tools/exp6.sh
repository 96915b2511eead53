#!/bin/bash
O=gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name regex:k_hru_step --launch-skip 30 --launch-count 1 -o $O/r02_full_frozen_async -f python bench.py --workload frozen_bands --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $O/r02_ncu_frozen_async.log 2>&1
export PROFILE_OUT=$O/prof_async
python tools/summarize_profile.py - $O/r02_full_frozen_async.ncu-rep r02_frozen_async frozen_bands 100000 k_hru_step_nn10 > $O/r02_summarize_frozen_async.log 2>&1
rm -f $O/r02_full_frozen_async.ncu-rep
head -14 $O/prof_async/r02_frozen_async_k_hru_step_nn10_by_function.txt | cut -c1-150
grep -E "thread_inst_executed_per|gpu__time_duration|smsp__inst_executed.sum|warp_latency" $O/prof_async/r02_frozen_async_k_hru_step_nn10_ncu_raw_selected.csv
