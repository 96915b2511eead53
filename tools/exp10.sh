#!/bin/bash
# snow-melt season (days 100-120): re-sort period and finer snow keys
P="python tools/perf_probe.py --steps 10 --warmup 10 --start-day 100"
{
$P --tag spring
VICGPU_REBIN=12 $P --tag spring_rebin12
VICGPU_REBIN=6 $P --tag spring_rebin6
VICGPU_BINFINE=1 $P --tag spring_fine
VICGPU_BINFINE=1 VICGPU_REBIN=6 $P --tag spring_fine_rebin6
python tools/perf_probe.py --steps 10 --warmup 10 --start-day 280 --tag autumn
VICGPU_REBIN=6 python tools/perf_probe.py --steps 10 --warmup 10 --start-day 280 --tag autumn_rebin6
} > gpurun_out/exp10.log 2>&1
grep -E "PROBE|rror" gpurun_out/exp10.log | cut -c1-230
