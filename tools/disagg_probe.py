#!/usr/bin/env python
"""Throughput of the forcing disaggregation (vicgpu_disagg: initialize_atmos + MTCLIM 4.3) on the bench domain (development tool).

    python tools/disagg_probe.py [--cells C] [--days D]
daily PREC / TMAX / TMIN / WIND -> hourly forcing for C cells x D days, device-resident (no copy back); prints cell-days/s."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from vic_b200 import api  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=10000)
ap.add_argument("--days", type=int, default=365)
a = ap.parse_args()
dom = bench.build_domain(a.cells, 1)
gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "fe_hourly_winter.npz")))
raw = gold["disagg_raw"].copy()
raw[5] = a.days
opt = api.parse_options(dom["options_raw"])
opt["nrecs"] = a.days * 24
g = api.VicGpu(api.options_to_raw(opt), device=0)
g.set_veglib(dom["veglib"]); g.set_cells(dom["cellpar"], dom["hrupar"])
rng = np.random.default_rng(5)
doy = np.arange(a.days)
tmean = dom["avg_temp"][:, None] + 12.0 * np.sin(2 * np.pi * (doy[None, :] - 105) / 365.0) + rng.normal(0, 2, (a.cells, a.days))
dtr = rng.uniform(6, 12, (a.cells, a.days))
daily = np.stack([np.where(rng.uniform(size=(a.cells, a.days)) < 0.4, rng.gamma(0.6, 6.0, (a.cells, a.days)), 0.0),
                  tmean + dtr / 2, tmean - dtr / 2, rng.uniform(1, 5, (a.cells, a.days))], axis=2)
g.disagg(raw, daily, want_host=False)  # warm-up (allocations)
import torch
torch.cuda.synchronize()
t0 = time.perf_counter()
g.disagg(raw, daily, want_host=False)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"DISAGG cells={a.cells} days={a.days} seconds={dt:.3f} (incl. H2D of {daily.nbytes / 1e6:.0f} MB daily input) cell_days_per_s={a.cells * a.days / dt:.4g} "
      f"cell_hours_per_s={a.cells * a.days * 24 / dt:.4g}", flush=True)
g.close()
