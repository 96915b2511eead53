#!/bin/bash
# A/B set 1: cache-policy / staging variants, grid shapes, domain-size sweep
P="python tools/perf_probe.py --steps 5 --warmup 2"
L=vic_b200/lib
{
$P --tag base
VICGPU_LIB=$L/libvicgpu_cs.so $P --tag cs
VICGPU_LIB=$L/libvicgpu_pf.so $P --tag pf
VICGPU_LIB=$L/libvicgpu_fs.so $P --tag fs
VICGPU_LIB=$L/libvicgpu_all.so $P --tag all
VICGPU_L2PERSIST=24 $P --tag l2p24
VICGPU_L2PERSIST=24 VICGPU_LIB=$L/libvicgpu_all.so $P --tag all_l2p24
VICGPU_EVEN=1 $P --tag even
VICGPU_EVEN=1 VICGPU_BLOCK=384 $P --tag even384
VICGPU_EVEN=1 VICGPU_BLOCK=384 VICGPU_LIB=$L/libvicgpu_all.so $P --tag all_even384
VICGPU_SYNC=0 $P --tag nosync
PROBE_PHASE_TAX=1 $P --tag phasetax
for c in 1000 2000 5000 20000 50000; do $P --cells $c --tag size; done
$P --start-day 180 --tag summer_base
VICGPU_LIB=$L/libvicgpu_all.so $P --start-day 180 --tag summer_all
$P --cells 125000 --steps 3 --tag big_base
VICGPU_LIB=$L/libvicgpu_all.so $P --cells 125000 --steps 3 --tag big_all
VICGPU_LIB=$L/libvicgpu_cs.so $P --cells 125000 --steps 3 --tag big_cs
} > gpurun_out/exp1.log 2>&1
grep -E "PROBE|PHASETAX|rror" gpurun_out/exp1.log | cut -c1-250
