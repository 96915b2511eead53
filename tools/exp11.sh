#!/bin/bash
# ground-surface residual with by-value outputs
timeout 600 python -m pytest tests -m gpu -x -q -k "golden_case or thousand" > gpurun_out/exp11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp11_pytest.log
P="python tools/perf_probe.py --steps 10 --warmup 10"
{
$P --tag winter
$P --start-day 180 --tag summer
$P --start-day 100 --tag spring
$P --cells 125000 --steps 3 --warmup 2 --tag big
} > gpurun_out/exp11.log 2>&1
tail -3 gpurun_out/exp11_pytest.log
grep -E "PROBE|rror" gpurun_out/exp11.log | cut -c1-200
