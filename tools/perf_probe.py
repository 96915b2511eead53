#!/usr/bin/env python
"""Quick device-time probe of the hot path on the bench domain (development tool; bench.py is the measurement of record).

    python tools/perf_probe.py [--cells C] [--steps K] [--warmup W] [--tag NAME]
Prints one line: tag, cells, ms per 24-record step, average step-kernel launch in us, cell-timesteps/s.
Tuning knobs are read by libvicgpu.so from the environment (VICGPU_BLOCK, VICGPU_NOBIN, ...)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vic_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=10000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--start-day", type=int, default=0)
ap.add_argument("--tag", default="")
ap.add_argument("--config", default="fe_hourly", help="fe_hourly | frozen_bands | glacier (bench_data/base_<config>.npz)")
ap.add_argument("--set", action="append", default=[], metavar="OPTION=VALUE", help="override an option of the base domain, e.g. --set IMPLICIT=1")
a = ap.parse_args()
dom = bench.build_domain(a.cells, 1, a.config)
if a.set:
    opt = api.parse_options(dom["options_raw"])
    for kv in a.set:
        k, v = kv.split("=")
        opt[k] = type(opt[k])(float(v))
    dom["options_raw"] = api.options_to_raw(opt)
g = api.VicGpu(dom["options_raw"], device=0)
g.set_veglib(dom["veglib"]); g.set_output_spec(dom["aggtype"]); g.set_cells(dom["cellpar"], dom["hrupar"]); g.set_state(dom["hrurec0"])
nd = a.warmup + a.steps
f = np.empty((nd * 24, a.cells, g.L.f_stride))
for d in range(nd):
    bench.forcing_day(dom, a.start_day + d, 1, f[d * 24:(d + 1) * 24])
g.set_forcing(0, f)
dmy = bench.make_dmy(nd * 24)
for s in range(a.warmup):
    g.step(s * 24, 24, dmy[s * 24:s * 24 + 25])
g.set_profiling(True)
ms = 0.0
for s in range(a.warmup, nd):
    g.step(s * 24, 24, dmy[s * 24:s * 24 + 25])
    ms += g.last_step_timing()[0]
kms, kn = g.kernel_profile()
env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("VICGPU_"))
print(f"PROBE {a.tag} {a.config} [{env}] cells={a.cells} hrus={g.nhru} ms_per_step={ms / a.steps:.3f} hru_kernel_us={kms / max(kn, 1) * 1e3:.1f} "
      f"cell_steps_per_s={a.cells * 24 * a.steps / (ms / 1e3):.4g} state_sum={np.nansum(g.get_state()):.17g}", flush=True)
if os.environ.get("PROBE_PHASE_TAX"):
    print("PHASETAX", a.config, "cells", a.cells, " ".join(f"nframe={n}:{g.phase_tax(n):.1f}us" for n in (0, 64, 128, 192)), flush=True)
if os.environ.get("VICGPU_WARPTIME"):
    t0, t1, kind = g.warp_times()
    dur = (t1 - t0) / 1e3
    B = int(os.environ.get("VICGPU_BLOCK", "384")) // 32
    nb = (len(dur) + B - 1) // B
    blk = np.array([t1[b * B:(b + 1) * B].max() - t0[b * B:(b + 1) * B].min() for b in range(nb)]) / 1e3
    print(f"WARPS n={len(dur)} dur us: min {dur.min():.0f} mean {dur.mean():.0f} p50 {np.median(dur):.0f} p90 {np.percentile(dur, 90):.0f} max {dur.max():.0f}; "
          f"kernel span {(t1.max() - t0.min()) / 1e3:.0f}; blocks n={nb} mean {blk.mean():.0f} max {blk.max():.0f} min {blk.min():.0f}")
    print(f"  warp starts us: p50 {np.median(t0)/1e3:.0f} p90 {np.percentile(t0,90)/1e3:.0f} p99 {np.percentile(t0,99)/1e3:.0f} max {t0.max()/1e3:.0f}; ends: p50 {np.median(t1)/1e3:.0f} p90 {np.percentile(t1,90)/1e3:.0f} max {t1.max()/1e3:.0f}")
    bs = np.array([t0[b * B:(b + 1) * B].min() for b in range(nb)]) / 1e3
    print("  block starts us (sorted, every 10th):", np.round(np.sort(bs)[::10]).astype(int).tolist())
    for kd in np.unique(kind):
        m = kind == kd
        print(f"  kind {int(kd):8d}: warps {m.sum():4d} mean {dur[m].mean():6.0f} max {dur[m].max():6.0f} us")
g.close()
