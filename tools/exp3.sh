#!/bin/bash
# quick parity subset + zero-numerator divisions + rendezvous phase masks
P="python tools/perf_probe.py --steps 5 --warmup 2"
L=vic_b200/lib
timeout 900 python -m pytest tests -m gpu -x -q -k "golden_case or launch_mode or thousand" > gpurun_out/exp3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp3_pytest.log
{
$P --tag new
$P --start-day 180 --tag summer_new
for m in 0 1 2 3 4 5 6; do VICGPU_SYNCMASK=$m $P --tag mask$m; done
for m in 3 5 6; do VICGPU_SYNCMASK=$m $P --start-day 180 --tag summer_mask$m; done
$P --cells 125000 --steps 3 --tag big_new
VICGPU_SYNCMASK=3 $P --cells 125000 --steps 3 --tag big_mask3
VICGPU_SYNCMASK=0 $P --cells 125000 --steps 3 --tag big_mask0
} > gpurun_out/exp3.log 2>&1
tail -3 gpurun_out/exp3_pytest.log
grep -E "PROBE|rror" gpurun_out/exp3.log | cut -c1-230
