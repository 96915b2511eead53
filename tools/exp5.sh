#!/bin/bash
# frozen-soil sweeps: lane-asynchronous (default) against lock-step (libvicgpu_lockstep.so)
L=vic_b200/lib
timeout 900 python -m pytest tests -m gpu -x -q -k "(golden_case and frozen) or invalidated or (full_size and frozen) or (launch_mode)" > gpurun_out/exp5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp5_pytest.log
{
python tools/perf_probe.py --config frozen_bands --steps 2 --warmup 1 --tag async10k
VICGPU_LIB=$L/libvicgpu_lockstep.so python tools/perf_probe.py --config frozen_bands --steps 2 --warmup 1 --tag lock10k
python tools/perf_probe.py --config frozen_bands --cells 100000 --steps 1 --warmup 1 --tag async100k
VICGPU_LIB=$L/libvicgpu_lockstep.so python tools/perf_probe.py --config frozen_bands --cells 100000 --steps 1 --warmup 1 --tag lock100k
} > gpurun_out/exp5.log 2>&1
tail -3 gpurun_out/exp5_pytest.log
grep -E "PROBE|rror" gpurun_out/exp5.log | cut -c1-230
