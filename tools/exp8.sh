#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "disagg or forcing_windows" > gpurun_out/exp8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp8_pytest.log
python tools/disagg_probe.py --cells 125000 --days 13 > gpurun_out/exp8.log 2>&1
python tools/disagg_probe.py --cells 125000 --days 365 >> gpurun_out/exp8.log 2>&1
python bench.py --workload continental --no-cpu-baseline > gpurun_out/exp8_bench_cont.json 2>> gpurun_out/exp8.log
tail -3 gpurun_out/exp8_pytest.log; grep DISAGG gpurun_out/exp8.log; python -c "
import json;d=json.loads(open('gpurun_out/exp8_bench_cont.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e'],d['config']['disagg'])"
VICGPU_WARPTIME=1 VICGPU_BLOCK=512 python tools/perf_probe.py --steps 2 --warmup 2 --tag warptime 2>&1 | grep -E "PROBE|WARPS|kind|warp starts|block" | cut -c1-400
VICGPU_WARPTIME=1 VICGPU_BLOCK=512 python tools/perf_probe.py --steps 2 --warmup 2 --start-day 180 --tag warptime_summer 2>&1 | grep -E "PROBE|WARPS|kind|warp starts" | cut -c1-400
for b in 416 448 480; do VICGPU_BLOCK=$b python tools/perf_probe.py --steps 5 --warmup 2 --tag block$b 2>&1 | grep PROBE | cut -c1-200; done
python tools/perf_probe.py --steps 5 --warmup 2 --tag block512 2>&1 | grep PROBE | cut -c1-200
