#!/bin/bash
# IMPLICIT soil-temperature solver: parity on the GPU, throughput beside the explicit scheme; the 3-node kernel must be unchanged
timeout 1200 python -m pytest tests -m gpu -x -q -k "implicit or (golden_case and frozen) or launch_mode or invalidated" > gpurun_out/exp7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp7_pytest.log
{
python tools/perf_probe.py --steps 5 --warmup 2 --tag fe
python tools/perf_probe.py --config frozen_bands --steps 2 --warmup 1 --tag explicit10k
python tools/perf_probe.py --config frozen_bands --steps 2 --warmup 1 --set IMPLICIT=1 --tag implicit10k
python tools/perf_probe.py --config frozen_bands --cells 100000 --steps 1 --warmup 1 --set IMPLICIT=1 --tag implicit100k
} > gpurun_out/exp7.log 2>&1
tail -3 gpurun_out/exp7_pytest.log
grep -E "PROBE|rror" gpurun_out/exp7.log | cut -c1-230
