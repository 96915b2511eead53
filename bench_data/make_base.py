"""Writes bench_data/base_<config>.npz: the flat C-ABI parameter and initial-state tables of a 16x16-cell
synthetic domain, produced by the reference's OWN readers and initialisation code (read_soilparam,
read_vegparam, read_snowband, initialize_model_state) through oracle/_ref/vic_ref_harness --no-run.

Run in the development container (needs /root/reference built through oracle/Makefile):
    python bench_data/make_base.py fe_hourly
bench.py tiles these 256 cells to the benchmark's domain size and generates hourly forcing itself, so that
nothing under oracle/ or /root/reference is needed to set the benchmark up on the GPU box."""
import dataclasses
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vic_b200 import synth  # noqa: E402
from vic_b200.casefile import read_case  # noqa: E402

SEED = 20260
NLAT = NLON = 16


def make(cfgname):
    cfg = dataclasses.replace(synth.CONFIGS[cfgname], ndays=2, out_step=24 if synth.CONFIGS[cfgname].dt < 24 else 0)
    with tempfile.TemporaryDirectory() as d:
        r = synth.generate(d, cfg, NLAT, NLON, SEED)
        case = os.path.join(d, "case.bin")
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "vic_ref_harness"), "-g", r["global_file"], "-o", case, "--no-run"], check=True,
                       stdout=subprocess.DEVNULL)
        c = read_case(case)
    keep = {k: c[k] for k in ("options_raw", "veglib", "cellpar", "hrupar", "hrurec0", "aggtype", "valid0")}
    out = os.path.join(ROOT, "bench_data", f"base_{cfgname}.npz")
    np.savez_compressed(out, **keep)
    print(out, os.path.getsize(out) // 1024, "KiB", "ncell", c["meta"][0], "nhru", c["meta"][1])


if __name__ == "__main__":
    for n in (sys.argv[1:] or ["fe_hourly"]):
        make(n)
