#!/bin/bash
# parity suite + A/B of the per-month aerodynamic table + ncu captures (winter and summer week)
P="python tools/perf_probe.py --steps 5 --warmup 2"
L=vic_b200/lib
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/exp2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp2_pytest.log
{
$P --tag aero
VICGPU_LIB=$L/libvicgpu_noaero.so $P --tag noaero
$P --start-day 180 --tag summer_aero
VICGPU_LIB=$L/libvicgpu_noaero.so $P --start-day 180 --tag summer_noaero
$P --start-day 100 --tag spring_aero
$P --cells 125000 --steps 3 --tag big_aero
VICGPU_LIB=$L/libvicgpu_noaero.so $P --cells 125000 --steps 3 --tag big_noaero
} > gpurun_out/exp2.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_hru_step --launch-skip 30 --launch-count 1 -o gpurun_out/prof_r02b_winter -f python tools/perf_probe.py --steps 1 --warmup 1 > gpurun_out/ncu_w.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:k_hru_step --launch-skip 30 --launch-count 1 -o gpurun_out/prof_r02b_summer -f python tools/perf_probe.py --steps 1 --warmup 1 --start-day 180 > gpurun_out/ncu_s.log 2>&1
tail -3 gpurun_out/exp2_pytest.log
grep -E "PROBE|rror" gpurun_out/exp2.log | cut -c1-230
