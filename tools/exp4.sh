#!/bin/bash
# transfer chunk size, block sizes, register budgets at scale
P="python tools/perf_probe.py --steps 5 --warmup 2"
L=vic_b200/lib
{
$P --tag base
VICGPU_LIB=$L/libvicgpu_x32.so $P --tag x32
VICGPU_LIB=$L/libvicgpu_x32.so $P --start-day 180 --tag summer_x32
for b in 128 256 384; do VICGPU_BLOCK=$b $P --tag block$b; done
VICGPU_BLOCK=256 $P --start-day 180 --tag summer_block256
$P --cells 125000 --steps 3 --tag big_base
VICGPU_LIB=$L/libvicgpu_x32.so $P --cells 125000 --steps 3 --tag big_x32
VICGPU_BLOCK=256 $P --cells 125000 --steps 3 --tag big_block256
VICGPU_LIB=$L/libvicgpu_r96.so VICGPU_BLOCK=640 $P --cells 125000 --steps 3 --tag big_r96
VICGPU_LIB=$L/libvicgpu_r96.so VICGPU_BLOCK=320 $P --cells 125000 --steps 3 --tag big_r96_b320
VICGPU_LIB=$L/libvicgpu_r80.so VICGPU_BLOCK=768 $P --cells 125000 --steps 3 --tag big_r80
VICGPU_LIB=$L/libvicgpu_r80.so VICGPU_BLOCK=384 $P --cells 125000 --steps 3 --tag big_r80_b384
VICGPU_LIB=$L/libvicgpu_r96.so VICGPU_BLOCK=640 $P --tag r96_10k
} > gpurun_out/exp4.log 2>&1
grep -E "PROBE|rror" gpurun_out/exp4.log | cut -c1-230
