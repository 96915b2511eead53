#!/bin/bash
# balanced blocks for single-wave domains
timeout 900 python -m pytest tests -m gpu -x -q -k "golden_case or launch_mode or thousand" > gpurun_out/exp9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp9_pytest.log
P="python tools/perf_probe.py --steps 10 --warmup 10"
{
VICGPU_BALANCE=0 $P --tag even
$P --tag bal28
VICGPU_BALANCE=20 $P --tag bal20
VICGPU_BALANCE=36 $P --tag bal36
VICGPU_BALANCE=44 $P --tag bal44
VICGPU_BALANCE=0 $P --start-day 180 --tag summer_even
$P --start-day 180 --tag summer_bal28
VICGPU_BALANCE=36 $P --start-day 180 --tag summer_bal36
VICGPU_BALANCE=0 $P --start-day 100 --tag spring_even
$P --start-day 100 --tag spring_bal28
$P --config glacier --tag glacier_bal28
VICGPU_BALANCE=0 $P --config glacier --tag glacier_even
} > gpurun_out/exp9.log 2>&1
tail -3 gpurun_out/exp9_pytest.log
grep -E "PROBE|rror" gpurun_out/exp9.log | cut -c1-230
