"""Parity fuzzer (CPU): random draws over the options, calendar and domain shapes the fixed test configurations never vary, each run through
the reference build (oracle/_ref/vic_ref_harness) and through the host build of the kernels' headers (oracle/_ref/vicport: the step;
oracle/_ref/disaggport: the forcing disaggregation), compared bit for bit.  A draw the library refuses (an option combination
vicgpu_create rejects) is reported as such, not as a failure.

  python tools/parity_fuzz.py --trials 40 --seed 1 [--jobs 8] [--keep]
"""
import argparse
import concurrent.futures as cf
import dataclasses
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vic_b200 import synth  # noqa: E402
from vic_b200.casefile import read_case, write_case  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
OPTIONS = {
    "SNOW_ALBEDO": ["USACE", "SUN1999"],
    "SNOW_DENSITY": ["DENS_BRAS", "DENS_SNTHRM"],
    "AERO_RESIST_CANSNOW": ["AR_406", "AR_406_LS", "AR_406_FULL", "AR_410"],
    "GRND_FLUX_TYPE": ["GF_406", "GF_410", "GF_FULL"],
    "TEMP_TH_TYPE": ["VIC_412", "KIENZLE"],
    "BASEFLOW": ["ARNO", "NIJSSEN2001"],
    "MOISTFRACT": ["FALSE", "TRUE"],
    "ALMA_OUTPUT": ["FALSE", "TRUE"],
    "LW_TYPE": ["LW_TVA", "LW_ANDERSON", "LW_BRUTSAERT", "LW_SATTERLUND", "LW_IDSO", "LW_PRATA"],
    "LW_CLOUD": ["LW_CLOUD_BRAS", "LW_CLOUD_DEARDORFF"],
    "VP_ITER": ["VP_ITER_NONE", "VP_ITER_ALWAYS", "VP_ITER_ANNUAL", "VP_ITER_CONVERGE"],
    "VP_INTERP": ["TRUE", "FALSE"],
    "MTCLIM_SWE_CORR": ["TRUE", "FALSE"],
    "PLAPSE": ["TRUE", "FALSE"],
    "TFALLBACK": ["TRUE", "FALSE"],
}


def draw(rng):
    base = rng.choice(["fe_hourly", "fe_hourly", "wb_daily", "glacier", "frozen_bands", "fe_blowing", "fe_corrprec", "glacier_dyn", "treeline", "frozen_implicit",
                       "glacier_blowing", "glacier_multi"])
    cfg = synth.CONFIGS[str(base)]
    over = {}
    if cfg.frozen_soil:  # the soil-thermal options (QUICK_SOLVE with IMPLICIT is refused: schedule-dependent in the reference)
        if rng.random() < 0.4:
            over["exp_trans"] = True
        if rng.random() < 0.4:
            over["noflux"] = True
        if not cfg.implicit and rng.random() < 0.3:
            over["quick_solve"] = True
    if cfg.frozen_soil:
        over["ndays"] = int(rng.integers(3, 9))
    elif base == "glacier_blowing":
        over["ndays"] = int(rng.integers(5, 15))
    else:
        over["ndays"] = int(rng.integers(20, 90))
    over["startyear"] = int(rng.choice([2001, 2003, 2004]))  # 2004: leap year
    if base != "treeline":
        over["startday"] = int(rng.choice([1, 1, 45, 100, 182, 275, 330]))
    if base in ("fe_hourly", "fe_corrprec") and rng.random() < 0.4:
        dt = int(rng.choice([3, 6]))
        over["dt"], over["snow_step"] = dt, dt
    if base == "wb_daily" and rng.random() < 0.5:
        over["snow_step"] = int(rng.choice([1, 3, 6]))
    if base in ("fe_hourly", "glacier", "fe_blowing") and rng.random() < 0.4:
        over["nbands"] = int(rng.choice([2, 3]))
    if rng.random() < 0.3:
        over["ntiles"] = int(rng.choice([2, 3, 7]))
    if rng.random() < 0.3:
        over["out_step"] = 24
    nopt = int(rng.integers(1, 6))
    extra = list(cfg.extra_global)
    for k in rng.choice(sorted(OPTIONS), size=nopt, replace=False):
        extra.append(f"{k} {rng.choice(OPTIONS[str(k)])}")
    over["extra_global"] = extra
    # where on the globe: the latitude drives the solar geometry of the disaggregation (68 N: polar night and midnight sun; the southern
    # hemisphere), the longitude the offset between the forcing's local days and the model's time zone (time_zone_lng -120: -2 .. +2 hours)
    lat0 = float(rng.choice([48.03125, 48.03125, 35.03125, 60.03125, 68.03125, -35.03125]))
    lon0 = float(rng.choice([-121.96875, -121.96875, -150.03125, -135.03125, -100.03125, -90.03125]))
    # the weather: the generator's mid-latitude climate shifted and scaled (arctic cold, heat, drought, deluge, calm, gale)
    climate = (float(rng.choice([0.0, 0.0, -25.0, -10.0, 10.0, 20.0])), float(rng.choice([1.0, 1.0, 0.0, 0.2, 5.0])), float(rng.choice([1.0, 1.0, 0.05, 4.0])))
    return str(base), dataclasses.replace(cfg, **over), int(rng.integers(1, 1 << 30)), int(rng.integers(2, 4)), int(rng.integers(2, 4)), lat0, lon0, climate


def run_trial(t):
    idx, base, cfg, seed, nlat, nlon, lat0, lon0, climate, keep = t
    d = tempfile.mkdtemp(prefix=f"fuzz{idx}_")
    label = (f"[{d}] " if keep else "") + ("(binned, roles) " if idx % 2 else "") + f"#{idx} {base} {nlat}x{nlon} at {lat0:.2f} {lon0:.2f} climate {climate[0]:+.0f}C x{climate[1]:g} prec x{climate[2]:g} wind seed {seed} days {cfg.ndays} start {cfg.startyear}/{cfg.startday} dt {cfg.dt}/{cfg.snow_step} bands {cfg.nbands} tiles {cfg.ntiles} " \
            f"out_step {cfg.out_step} | " + ", ".join(cfg.extra_global)
    try:
        try:
            r = synth.generate(os.path.join(d, "in"), cfg, nlat, nlon, seed, lat0=lat0, lon0=lon0)
        except Exception as e:
            return label, "skipped", f"generator: {e}"
        if climate != (0.0, 1.0, 1.0):  # rewrite the daily forcing files: PREC TMAX TMIN WIND
            fdir = os.path.join(r["dir"], "forc")
            for fn in os.listdir(fdir):
                a = np.loadtxt(os.path.join(fdir, fn), ndmin=2)
                a[:, 0] *= climate[1]
                a[:, 1:3] += climate[0]
                a[:, 3] *= climate[2]
                np.savetxt(os.path.join(fdir, fn), a, fmt="%.4f")
        sb = os.path.join(r["dir"], "snowband.txt")
        if idx % 3 == 0 and os.path.exists(sb):  # every third banded draw: some bands without area (their share goes to the first band)
            rows = [ln.split() for ln in open(sb).read().splitlines() if ln.strip()]
            nb = cfg.nbands
            zr = np.random.default_rng(seed)
            for row in rows:
                frac = [float(x) for x in row[1:1 + nb]]
                for b in range(1, nb):
                    if zr.random() < 0.35:
                        frac[0] += frac[b]
                        frac[b] = 0.0
                row[1:1 + nb] = [f"{x:.6f}" for x in frac]
            open(sb, "w").write("\n".join(" ".join(row) for row in rows) + "\n")
        case, out, fout = os.path.join(d, "case.bin"), os.path.join(d, "res.bin"), os.path.join(d, "forc.bin")
        h = subprocess.run([os.path.join(REF, "vic_ref_harness"), "-g", r["global_file"], "-o", case, "--dump-every", "240"], capture_output=True, text=True)
        if h.returncode != 0:
            return label, "skipped", f"the reference refuses the draw (rc {h.returncode})"
        # every other draw in the device's binned row order with the three-thread cell output emulated (vicport --binned --roles)
        devlike = ["--binned", "--roles"] if idx % 2 else []
        p = subprocess.run([os.path.join(REF, "vicport"), case, out, *devlike], capture_output=True, text=True)
        if p.returncode != 0:
            msg = (p.stderr or p.stdout).strip().splitlines()[-1:] or ["?"]
            if "not implemented" in msg[0] or "unsupported" in msg[0].lower():
                return label, "refused", msg[0]
            return label, "FAIL", f"vicport rc {p.returncode}: {msg[0]}"
        c, res = read_case(case), read_case(out)
        # cells the reference invalidates (an ERROR return with TFALLBACK FALSE): rows are compared up to the failing record -- the reference
        # writes that record from a half-stepped cell (dist_prec.c:159-171), the library keeps the last good row (tests/test_cpu.py
        # check_until_invalid) -- and their state and balance rows are left out
        ok_cell = c["status_ref"] == 0
        bad = [] if np.array_equal(res["status"], c["status_ref"]) else ["status"]
        ref_out, out_rows = c["out_ref"], res["out"]
        for cell in range(ref_out.shape[1]):
            f = ref_out.shape[0]
            if not ok_cell[cell]:
                changed = [k for k in range(1, f) if not np.array_equal(ref_out[k, cell], ref_out[k - 1, cell], equal_nan=True)]
                f = changed[-1] if changed else 0
            if not np.array_equal(out_rows[:f, cell], ref_out[:f, cell], equal_nan=True):
                bad.append(f"out (cell {cell})")
                break
        from vic_b200.layout import TABLES
        hru_ok = ok_cell[c["hrupar"][:, list(TABLES["hpar"]).index("HP_cell")].astype(int)]
        if not np.array_equal(res["hrurec"][:, hru_ok], c["hrurec_ref"][:, hru_ok], equal_nan=True):
            bad.append("hrurec")
        if not np.array_equal(res["balance"][ok_cell], c["balance_ref"][ok_cell], equal_nan=True):
            bad.append("balance")
        # the forcing disaggregation: daily PREC / TMAX / TMIN / WIND as the reference read them
        ncell = int(c["meta"][0])
        nd = int(c["disagg_raw"][5])
        lat, lng = [c["cellpar"][:, k] for k in (synth_cp("CP_lat"), synth_cp("CP_lng"))]
        daily = np.stack([np.loadtxt(os.path.join(r["dir"], "forc", f"f_{la:.5f}_{lo:.5f}"))[:nd] for la, lo in zip(lat, lng)])
        dcase = os.path.join(d, "dcase.bin")
        write_case(dcase, {"options_raw": c["options_raw"], "disagg_raw": c["disagg_raw"], "meta": c["meta"], "cellpar": c["cellpar"], "daily": daily})
        q = subprocess.run([os.path.join(REF, "disaggport"), dcase, fout], capture_output=True, text=True)
        if q.returncode != 0:
            msg = (q.stderr or q.stdout).strip().splitlines()[-1:] or ["?"]
            if "COMPUTE_TREELINE" in msg[0] or "not implemented" in msg[0]:
                pass
            else:
                bad.append(f"disaggport rc {q.returncode}: {msg[0]}")
        elif not np.array_equal(read_case(fout)["forcing"], c["forcing"]):
            bad.append("forcing")
        invalid = int((c["status_ref"] != 0).sum())
        return label, ("FAIL" if bad else "ok"), (", ".join(bad) if bad else f"{ncell} cells, {c['out_ref'].shape[0]} records, {invalid} cells invalidated by the reference")
    finally:
        if not keep:
            shutil.rmtree(d, ignore_errors=True)


def synth_cp(name):
    from vic_b200.layout import TABLES
    return list(TABLES["cpar"]).index(name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=24)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--keep", action="store_true", help="keep the trial directories (printed with the label)")
    ap.add_argument("--only", type=int, default=None, help="run this trial of the sequence alone")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    trials = [(i, *draw(rng), a.keep) for i in range(a.trials)]
    if a.only is not None:
        trials = [t for t in trials if t[0] == a.only]
    counts = {}
    with cf.ProcessPoolExecutor(a.jobs) as ex:
        for label, verdict, detail in ex.map(run_trial, trials):
            counts[verdict] = counts.get(verdict, 0) + 1
            print(f"{verdict:8s} {label}\n         -> {detail}", flush=True)
    print("summary:", counts)
    return 1 if counts.get("FAIL") else 0


if __name__ == "__main__":
    sys.exit(main())
