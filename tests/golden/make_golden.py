"""Regenerates the golden case files in this directory FROM THE REFERENCE ITSELF.

Run in the development container (needs /root/reference, built through oracle/Makefile):
    python tests/golden/make_golden.py
For every configuration it writes synthetic inputs with vic_b200.synth (fixed seed) and runs the reference's
unmodified physics (oracle/_ref/vic_ref_harness: the reference's own sources, linked against the platform's glibc 2.39 libm).
The CUDA library must agree BIT FOR BIT (state, outputs, disaggregated forcing, counters): its elementary functions are
the operation-by-operation restatement of that libm (vic_b200/csrc/vic_glibm.cuh).
Each file keeps, compressed:
  the flat C-ABI inputs (options_raw, veglib, cellpar, hrupar, hrurec0, aggtype, valid0, dmy, forcing; disagg_raw + daily:
  the daily PREC/TMAX/TMIN/WIND the reference read, of which `forcing` is ITS disaggregation = the answer for vicgpu_disagg)
  the reference's answers: hrurec_ref at dump_recs, agg_ref at agg_recs (daily aggregates of all 184
  output variables), balance_ref, status_ref and out_ref for the first and last 24 records.
"""
import dataclasses
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vic_b200 import synth  # noqa: E402
from vic_b200.casefile import read_case  # noqa: E402

HARNESS = {"": os.path.join(ROOT, "oracle", "_ref", "vic_ref_harness")}

# name -> (config name, overrides, nlat, nlon, seed, dump_every)
GOLDEN = {
    "fe_hourly_winter": ("fe_hourly", dict(ndays=12, out_step=24), 2, 2, 11, 48),
    "fe_hourly_summer": ("fe_hourly", dict(ndays=10, out_step=24, startday=182), 2, 2, 12, 48),
    "wb_daily": ("wb_daily", dict(ndays=120), 2, 2, 13, 30),
    "frozen_bands": ("frozen_bands", dict(ndays=8, out_step=24), 2, 2, 14, 48),
    "glacier": ("glacier", dict(ndays=10, out_step=24), 2, 2, 15, 48),
    # IMPLICIT TRUE: Newton-Raphson soil-temperature solver with the explicit scheme as fallback (third oracle patch, oracle/Makefile)
    "frozen_implicit": ("frozen_implicit", dict(ndays=8, out_step=24), 2, 2, 16, 48),
}


def make(name, flavour=""):
    cfgname, over, nlat, nlon, seed, dump_every = GOLDEN[name]
    cfg = dataclasses.replace(synth.CONFIGS[cfgname], **over)
    with tempfile.TemporaryDirectory() as d:
        r = synth.generate(d, cfg, nlat, nlon, seed)
        case = os.path.join(d, "case.bin")
        subprocess.run([HARNESS[flavour], "-g", r["global_file"], "-o", case, "--dump-every", str(dump_every)], check=True, stdout=subprocess.DEVNULL)
        c = read_case(case)
        # the daily forcing the reference read (ASCII PREC TMAX TMIN WIND, one file per cell): input of the disaggregation tests
        from vic_b200.layout import TABLES
        lat = c["cellpar"][:, TABLES["cpar"].index("CP_lat")]
        lng = c["cellpar"][:, TABLES["cpar"].index("CP_lng")]
        ndays = int(c["disagg_raw"][5])
        c["daily"] = np.stack([np.loadtxt(os.path.join(d, "forc", f"f_{la:.5f}_{lo:.5f}"))[:ndays] for la, lo in zip(lat, lng)])
    nrec = c["out_ref"].shape[0]
    keep = {k: c[k] for k in ("options_raw", "meta", "veglib", "cellpar", "hrupar", "hrurec0", "aggtype", "valid0", "dmy", "forcing",
                              "dump_recs", "disagg_raw", "daily", "hrurec_ref", "agg_recs", "agg_ref", "balance_ref", "status_ref")}
    keep["out_ref_head"] = c["out_ref"][:24]
    keep["out_ref_tail"] = c["out_ref"][nrec - 24:]
    # uninitialised aggdata of the reference's very first output step shows up as denormal garbage; not part of the contract
    keep["agg_ref"] = np.where(np.abs(keep["agg_ref"]) < 1e-300, 0.0, keep["agg_ref"])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), name + flavour + ".npz")
    np.savez_compressed(out, **keep)
    print(name + flavour, "nrec", nrec, "ncell", c["meta"][0], "nhru", c["meta"][1], os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    for n in (sys.argv[1:] or GOLDEN):
        for fl in HARNESS:
            make(n, fl)
