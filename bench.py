#!/usr/bin/env python
"""Benchmark of the hot path: cell-timesteps/s of the VIC cell loop on B200s, next to the reference's own CPU build.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--cells C] [--impl ours|reference]

Workloads (BASELINE.json configs; one bench "step" = one day = 24 hourly records over every cell of the rank):
  fe_hourly     configs[1]  FULL_ENERGY, QUICK_FLUX, 3 layers, 3 thermal nodes, 5 vegetation tiles (2 overstory + 3 short, + the automatic
                            bare-soil HRU where the tiles do not cover the cell), 1 snow band, 10,000 cells per GPU.  THE DEFAULT: the
                            configuration the metric is quoted on; K + W = 365 steps is its full year.
  frozen_bands  configs[2]  FROZEN_SOIL, QUICK_FLUX FALSE, 10 thermal nodes, 5 snow bands, 100,000 cells per GPU
  glacier       configs[3]  PCIC glacier mass-balance mode (glacier tiles in the upper bands, 5 bands), 250,000 cells per GPU
  continental   configs[4]  daily PREC/TMAX/TMIN/WIND in -> mtclim disaggregation on the device (vicgpu_disagg) -> full-energy hourly
                            steps -> daily float32 aggregates out; 125,000 cells per GPU (1,000,000 over 8 GPUs), one gather of the
                            last day's aggregates over NCCL at the end, timed separately

Parameters and initial state: bench_data/base_<config>.npz (256 cells produced by the reference's own readers and initialisation
code, see bench_data/make_base.py) tiled to the domain size with perturbed infiltration / baseflow / conductivity parameters; forcing:
synthetic weather generated here (seeded).  Cells are independent, so with N GPUs every rank owns its own cells (weak scaling) and
there is no collective on the data path.

--impl reference: the reference's own CPU build (oracle/_ref/vic_ref_harness: its unmodified sources, OpenMP cell loop of
vicNl.c:514-517, -O3) on the box's host cores, timed by the harness's own loop timer exactly as vicNl.c:501, 614-623 times it; for
fe_hourly on 10,000 synthetic cells of the same generator.

Lines printed (rank 0): one JSON object, see the keys at the bottom of main().
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vic_b200.layout import TABLES, layout_from_options, parse_options  # noqa: E402

RECS_PER_STEP = 24
BASE_SEED = 20260            # seed of the 16x16 base domains (bench_data/make_base.py)
SIGMA = 5.6696e-8

# algorithmic bytes per cell-timestep, B = H (2 S_hru + P_hru) + P_cell + F_cell + O_cell:
#   fe_hourly    H = 5, Nn = 3:            5 (960 + 40) + 1900 + 81 + 13  = 7.0 KB
#   frozen_bands Nn = 10, 5 bands:          5 (1520 + 40) + 2400 + 81      = 10.3 KB (the survey's figure)
#   glacier      H = 6.6, Nn = 3, 5 bands:  6.6 (960 + 40) + 2060 + 94     = 8.8 KB
WORKLOADS = {
    "fe_hourly": dict(base="fe_hourly", cells=10000, steps=362, warmup=3, algo_bytes=7.0e3, kernel="k_hru_step_nn3", config=1, ref_side=100,
                      desc="fe_hourly: FULL_ENERGY hourly, QUICK_FLUX, 3 layers, 3 nodes, 5 veg tiles, 1 band"),
    "frozen_bands": dict(base="frozen_bands", cells=100000, steps=2, warmup=3, algo_bytes=10.3e3, kernel="k_hru_step_nn10", config=2, ref_side=16,
                         desc="frozen_bands: FROZEN_SOIL, QUICK_FLUX FALSE, 10 thermal nodes, 5 snow bands, hourly"),
    # the same domain with the reference's own speed option for frozen soils: the surface-temperature search runs on the nodes above the thaw
    # depth, the whole profile is solved once at the accepted temperature (QUICK_SOLVE, calc_surf_energy_bal.c:289-314, 400-475)
    "frozen_quick_solve": dict(base="frozen_bands", synth="frozen_quick_solve", set={"QUICK_SOLVE": 1}, cells=100000, steps=5, warmup=3, algo_bytes=10.3e3,
                               kernel="k_hru_step_nn10", config=2, ref_side=16,
                               desc="frozen_quick_solve: frozen_bands with QUICK_SOLVE TRUE (FROZEN_SOIL, QUICK_FLUX FALSE, 10 thermal nodes, 5 snow bands, hourly)"),
    "glacier": dict(base="glacier", cells=250000, steps=10, warmup=3, algo_bytes=8.8e3, kernel="k_hru_step_nn3", config=3, ref_side=32,
                    desc="glacier: PCIC glacier mass-balance mode (surface_fluxes_glac), 5 snow bands, hourly"),
    "continental": dict(base="fe_hourly", cells=125000, steps=10, warmup=3, algo_bytes=7.0e3, kernel="k_hru_step_nn3", config=4, ref_side=100, disagg=True,
                        desc="continental: mtclim daily->hourly disaggregation on the device + fe_hourly physics, float32 daily aggregates out"),
}


def col(table, name):
    return TABLES[table].index(name)


# ---------------------------------------------------------------------------------------------- domain
def build_domain(ncell, seed, base="fe_hourly"):
    """tile the 256-cell base domain to ncell cells; tile 0 is the base domain itself"""
    b = dict(np.load(os.path.join(ROOT, "bench_data", f"base_{base}.npz")))
    nb = b["cellpar"].shape[0]
    cell_of_hru = b["hrupar"][:, col("hpar", "HP_cell")].astype(np.int64)
    h0 = np.searchsorted(cell_of_hru, np.arange(nb + 1))
    src = np.arange(ncell) % nb
    cellpar = b["cellpar"][src].copy()
    rng = np.random.default_rng([seed, 7])
    tiled = np.arange(ncell) >= nb
    # perturb parameters that no derived table depends on (runoff.c, arno_evap.c, penman.c read them directly)
    for name, lo, hi in (("CP_b_infilt", 0.05, 0.4), ("CP_Ds", 0.001, 0.1), ("CP_Dsmax", 5.0, 30.0), ("CP_Ws", 0.6, 0.95)):
        c = col("cpar", name)
        cellpar[tiled, c] = rng.uniform(lo, hi, int(tiled.sum()))
    nscal = len(TABLES["cpar"])
    k0 = nscal + col("cpar_layer", "CL_Ksat") * 3
    cellpar[tiled, k0:k0 + 3] *= np.exp(rng.uniform(-0.5, 0.5, (int(tiled.sum()), 3)))
    hru_idx = np.concatenate([np.arange(h0[s], h0[s + 1]) for s in src])
    counts = (h0[1:] - h0[:-1])[src]
    hrupar = b["hrupar"][hru_idx].copy()
    hrupar[:, col("hpar", "HP_cell")] = np.repeat(np.arange(ncell), counts)
    hrurec0 = b["hrurec0"][hru_idx].copy()
    elev = cellpar[:, col("cpar", "CP_elevation")]
    lat = cellpar[:, col("cpar", "CP_lat")]
    avg_temp = cellpar[:, col("cpar", "CP_avg_temp")]
    return dict(options_raw=b["options_raw"], veglib=b["veglib"], aggtype=b["aggtype"], cellpar=cellpar, hrupar=hrupar, hrurec0=hrurec0,
                elev=elev, lat=lat, avg_temp=avg_temp)


def svp_pa(t):
    s = 0.61078 * np.exp(17.269 * t / (237.3 + t))
    s = np.where(t < 0, s * (1.0 + 0.00972 * t + 0.000042 * t * t), s)
    return s * 1000.0


def forcing_day(dom, day, seed, out):
    """hourly forcing of one day for every cell -> out [24][ncell][11] (column order of VICGPU_FORCING)"""
    n = dom["elev"].shape[0]
    rng = np.random.default_rng([seed, 11, day])
    doy = day % 365
    tmean = dom["avg_temp"] + 2.0 + 12.0 * np.sin(2 * np.pi * (doy - 105) / 365.0) - 0.004 * (dom["elev"] - 1000.0) + rng.normal(0.0, 2.0, n)
    dtr = rng.uniform(6.0, 12.0, n)
    wet = rng.uniform(size=n) < 0.4
    amount = np.where(wet, rng.gamma(0.6, 6.0, n), 0.0)
    start = rng.integers(0, 18, n)
    wind = rng.uniform(1.0, 5.0, n)
    cloud = np.where(wet, 0.8, rng.uniform(0.0, 0.4, n))
    hours = np.arange(24)[:, None]
    t = tmean[None, :] + 0.5 * dtr[None, :] * np.cos(2 * np.pi * (hours - 15) / 24.0)
    tmin = tmean - 0.5 * dtr
    pressure = 101325.0 * np.exp(-dom["elev"] / 8434.5)
    es = svp_pa(t)
    vp = np.minimum(svp_pa(tmin)[None, :], es)
    decl = -0.4092797 * np.cos(2 * np.pi * (doy + 10) / 365.0)
    latr = np.deg2rad(dom["lat"])[None, :]
    cosz = np.sin(latr) * np.sin(decl) + np.cos(latr) * np.cos(decl) * np.cos(2 * np.pi * (hours - 12) / 24.0)
    sw = np.maximum(cosz, 0.0) * 1000.0 * (0.75 - 0.5 * cloud[None, :])
    prec = np.where((hours >= start[None, :]) & (hours < start[None, :] + 6), amount[None, :] / 6.0, 0.0)
    f = {"FV_air_temp": t, "FV_density": pressure[None, :] / (287.0 * (t + 273.15)), "FV_longwave": (0.7 + 0.25 * cloud[None, :]) * SIGMA * (t + 273.15) ** 4,
         "FV_prec": prec, "FV_pressure": np.broadcast_to(pressure[None, :], t.shape), "FV_shortwave": sw, "FV_tskc": np.broadcast_to(cloud[None, :], t.shape),
         "FV_vp": vp, "FV_vpd": es - vp, "FV_wind": np.broadcast_to(wind[None, :], t.shape), "FV_snowflag": ((prec > 0) & (t < 8.0)).astype(np.float64)}
    for k, name in enumerate(TABLES["forcing"]):
        out[:, :, k] = f[name]


def make_dmy(nrec, year=2001):
    """integer calendar exactly as make_dmy() builds it for an hourly run starting 1 Jan, 00h (make_dmy.c:105-127)"""
    mdays = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    d = np.zeros((nrec + 1, 5), dtype=np.int32)
    day, month, hour, diy, yr = 1, 1, 0, 1, year
    for r in range(nrec + 1):
        d[r] = (day, diy, hour, month, yr)
        hour += 1
        if hour == 24:
            hour = 0
            day += 1
            diy += 1
            leap = (yr % 4 == 0 and yr % 100 != 0) or yr % 400 == 0
            if day > mdays[month - 1] + (1 if (month == 2 and leap) else 0):
                day = 1
                month += 1
                if month > 12:
                    month, diy = 1, 1
                    yr += 1
    return d


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples = []
        self.stop = threading.Event()
        self.t = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in o.strip().split(",")]
                if len(p) >= 6:
                    self.samples.append((float(p[0]), float(p[1]), p[2:6]))
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median([s[0] for s in self.samples])), "sm_max_mhz": self.samples[0][1], "reasons": reasons}


# ---------------------------------------------------------------------------------------------- reference arm
def run_reference(workload, side, ndays, warm_days, threads, seed=1):
    """The reference's own CPU implementation (oracle/_ref/vic_ref_harness = its unmodified sources, OpenMP cell loop of
    vicNl.c:514-517) on side x side synthetic cells of the workload's configuration (vic_b200/synth.py, the generator the base
    domains of bench_data/ come from) with the bench's hourly forcing generator.  Returns (cell_steps_per_s, seconds, ncell): the
    harness's own timer around the record loop, exactly what vicNl.c:501, 614-623 report as "Model execution time"; the first
    warm_days days are run but not timed."""
    import dataclasses
    from vic_b200 import synth
    from vic_b200.casefile import write_case
    harness = os.path.join(ROOT, "oracle", "_ref", "vic_ref_harness")
    if not os.path.exists(harness):
        raise FileNotFoundError(harness)
    wl = WORKLOADS[workload]
    n = side * side
    with tempfile.TemporaryDirectory() as d:
        cfg = dataclasses.replace(synth.CONFIGS[wl.get("synth", wl["base"])], ndays=ndays, out_step=24)
        r = synth.generate(d, cfg, side, side, BASE_SEED, forcing=False, threads=threads)
        dom = {k: np.array([c[k] for c in r["cells"]]) for k in ("elev", "lat", "avg_temp")}
        f = np.empty((ndays * 24, n, len(TABLES["forcing"])))
        for day in range(ndays):
            forcing_day(dom, day, seed, f[day * 24:(day + 1) * 24])
        fb = os.path.join(d, "forcing.bin")
        write_case(fb, {"forcing": f})
        del f
        o = subprocess.run([harness, "-g", r["global_file"], "--forcing-bin", fb, "--time-only", "--threads", str(threads), "--time-from", str(warm_days * 24)],
                           capture_output=True, text=True, check=True).stdout
    line = [x for x in o.splitlines() if x.startswith("run_seconds")][0].split()
    return float(line[7]), float(line[1]), n


def cpu_baseline(workload, host_threads):
    """bounded samples of the workload on the host cores: all threads and one thread"""
    wl = WORKLOADS[workload]
    side = wl["ref_side"]
    if workload in ("fe_hourly", "continental"):
        v, secs, n = run_reference(workload, side, 5, 1, host_threads)
        v1, secs1, n1 = run_reference(workload, 32, 3, 1, 1)
        sample = (f"{n} cells x 96 hourly records (after 24 untimed), {host_threads} OpenMP threads, {secs:.1f} s; one thread: {n1} cells x 48 records, {secs1:.1f} s")
    else:
        days = 2 if workload.startswith("frozen") else 3
        v, secs, n = run_reference(workload, side, days, 1, host_threads)
        v1, secs1, n1 = run_reference(workload, 8, days, 1, 1)
        sample = (f"{n} cells x {(days - 1) * 24} hourly records (after 24 untimed), {host_threads} OpenMP threads, {secs:.1f} s; one thread: {n1} cells, {secs1:.1f} s")
    return {"value": v, "unit": "cell-timesteps/s", "cores": host_threads, "kind": "reference", "value_1thread": v1,
            "sample": sample + "; reference build oracle/_ref/vic_ref_harness (-O3), synthetic cells of the workload's configuration, the bench's forcing generator, "
                               "timed by the harness's record-loop timer (vicNl.c:501, 614-623)"}


NC_VARS = ["pr", "tasmax", "tasmin", "wind"]


def synth_daily(dom, ndays, seed):
    """daily PREC [mm], TMAX, TMIN [C], WIND [m/s] per cell: [ncell][ndays][4] (the generator of vic_b200/synth.py, vectorised)"""
    n = dom["elev"].shape[0]
    rng = np.random.default_rng([seed, 13])
    doy = np.arange(ndays)[None, :]
    tmean = dom["avg_temp"][:, None] + 2.0 + 12.0 * np.sin(2 * np.pi * (doy - 105) / 365.0) - 0.004 * (dom["elev"][:, None] - 1000.0) + rng.normal(0.0, 2.0, (n, ndays))
    dtr = rng.uniform(6.0, 12.0, (n, ndays))
    wet = rng.uniform(size=(n, ndays)) < 0.4
    out = np.empty((n, ndays, 4))
    out[:, :, 0] = np.round(np.where(wet, rng.gamma(0.6, 6.0, (n, ndays)), 0.0), 4)
    out[:, :, 1] = np.round(tmean + dtr / 2, 4)
    out[:, :, 2] = np.round(tmean - dtr / 2, 4)
    out[:, :, 3] = np.round(rng.uniform(1.0, 5.0, (n, ndays)), 4)
    return out


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="fe_hourly", choices=sorted(WORKLOADS))
    ap.add_argument("--cells", type=int, default=None, help="cells per GPU (default: the workload's BASELINE size)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--forcing-nc", action="store_true",
                    help="continental only: the e2e leg takes its daily forcing from a NetCDF (classic) file through vicgpu_nc_read_slab + vicgpu_disagg_tm")
    a = ap.parse_args()
    wl = WORKLOADS[a.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K = a.steps if a.steps is not None else wl["steps"]
    W = max(a.warmup if a.warmup is not None else wl["warmup"], 0)
    cells = a.cells if a.cells is not None else wl["cells"]
    host_threads = os.cpu_count() or 1
    workload = f"{wl['desc']}, {cells} cells/GPU (BASELINE configs[{wl['config']}]), 1 step = 24 hourly records"

    if a.impl == "reference":
        if rank != 0:
            return
        # at most 30 days per run: the forcing of 10,000 cells travels to the harness as one file (211 MB per 10 days)
        side = wl["ref_side"]
        Kr = min(K, 30 - min(W, 3))
        Wr = min(W, 3)
        v, secs, n = run_reference(a.workload, side, Wr + Kr, Wr, host_threads, seed=1)
        sample = (f"{n} synthetic cells of the workload's configuration x {Kr * 24} hourly records ({Kr} of the {K} steps asked for; {Wr} untimed warm-up days), "
                  f"reference CPU build (-O3), OpenMP cell loop, {host_threads} threads, harness record-loop timer {secs:.2f} s")
        line = {"impl": "reference", "metric": "cell-timesteps/s", "value": v, "unit": "cell-timesteps/s", "n_gpus": a.gpus, "steps": K, "warmup": W,
                "ms_per_step": secs / Kr * 1e3 * (cells / n), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "sample": sample},
                "cpu_baseline": {"value": v, "unit": "cell-timesteps/s", "cores": host_threads, "kind": "reference", "sample": sample},
                "e2e": {"value": v, "unit": "cell-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from vic_b200 import api
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    seed = 1 + rank
    dom = build_domain(cells, seed, wl["base"])
    if wl.get("set"):  # option overrides of the base domain
        opt = api.parse_options(dom["options_raw"])
        opt.update(wl["set"])
        dom["options_raw"] = api.options_to_raw(opt)
    ndays = W + K
    nrec = ndays * RECS_PER_STEP
    dmy = make_dmy(nrec)
    disagg = bool(wl.get("disagg"))
    options_raw = dom["options_raw"]
    if disagg:
        # the disaggregation fills the forcing window of the whole run: the run is the W + K days of this benchmark
        opt = parse_options(options_raw)
        opt["nrecs"] = nrec
        options_raw = api.options_to_raw(opt)
    g = api.VicGpu(options_raw, device=local_rank)
    L = g.L
    g.set_veglib(dom["veglib"])
    g.set_output_spec(dom["aggtype"])
    out_dtype = np.float32 if disagg else np.float64   # continental: the daily aggregates leave as the NetCDF writer stores them
    obuf = torch.empty((1, cells, L.nout), dtype=torch.float32 if disagg else torch.float64, pin_memory=True)
    onp = obuf.numpy()

    def reset():
        g.set_cells(dom["cellpar"], dom["hrupar"])
        g.set_state(dom["hrurec0"])

    disagg_s = None
    if disagg:
        gd = dict(np.load(os.path.join(ROOT, "tests", "golden", "fe_hourly_winter.npz")))
        disagg_raw = gd["disagg_raw"].copy()   # the mtclim options of the synthetic global file (vic_b200/synth.py)
        disagg_raw[1:6] = (0, 2001, 1, 1, ndays)  # starthour, startyear, startmonth, startday, Ndays
        dbuf = torch.empty((cells, ndays, 4), dtype=torch.float64, pin_memory=True)
        daily = dbuf.numpy()
        daily[:] = synth_daily(dom, ndays, seed)
        fnp = None
        nc_path = None
        if a.forcing_nc:
            # the same daily forcing as a (time, lat, lon) NetCDF file of doubles (written once, untimed, by an independent implementation of
            # the format); cell c sits at grid point (c // nlon, c % nlon) of a grid with a few unmodelled points at the end
            from scipy.io import netcdf_file
            nlon = int(np.ceil(np.sqrt(cells)))
            nlat = (cells + nlon - 1) // nlon
            glat, glon = 30.0 + 0.0625 * np.arange(nlat), -130.0 + 0.0625 * np.arange(nlon)
            nc_path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"vicgpu_bench_daily_{os.getpid()}.nc")
            f = netcdf_file(nc_path, "w", version=2)
            f.createDimension("time", None)
            f.createDimension("lat", nlat)
            f.createDimension("lon", nlon)
            for name, vals in (("time", np.arange(ndays, dtype=np.float64)), ("lat", glat), ("lon", glon)):
                v = f.createVariable(name, "d", (name,))
                v[:] = vals
            grid = np.zeros((ndays, nlat * nlon))
            for k, name in enumerate(NC_VARS):
                grid[:, :cells] = daily[:, :, k].T
                v = f.createVariable(name, "d", ("time", "lat", "lon"))
                v[:] = grid.reshape(ndays, nlat, nlon)
            f.close()
            cell_lat, cell_lon = glat[np.arange(cells) // nlon], glon[np.arange(cells) % nlon]
            tmbuf = torch.empty((ndays, 4, cells), dtype=torch.float64, pin_memory=True)
            daily_tm = tmbuf.numpy()
    else:
        # pinned host buffer: the run's hourly forcing
        fbuf = torch.empty((nrec, cells, L.f_stride), dtype=torch.float64, pin_memory=True)
        fnp = fbuf.numpy()
        for day in range(ndays):
            forcing_day(dom, day, seed, fnp[day * 24:(day + 1) * 24])

    # ---- leg 1: inputs resident in HBM, device-side daily aggregation, nothing copied back
    reset()
    if disagg:
        barrier()
        t0 = time.perf_counter()
        g.disagg(disagg_raw, daily, want_host=False)
        barrier()
        disagg_s = time.perf_counter() - t0
    else:
        g.set_forcing(0, fnp)
    for s in range(W):
        g.step(s * 24, 24, dmy[s * 24:s * 24 + 25])
    g.set_profiling(True)
    launches = 0
    barrier()
    with ClockSampler(local_rank) as cs:
        t0 = time.perf_counter()
        dev_ms = 0.0
        for s in range(W, W + K):
            g.step(s * 24, 24, dmy[s * 24:s * 24 + 25])
            ms, nl = g.last_step_timing()
            dev_ms += ms
            launches += nl
        barrier()
        wall = time.perf_counter() - t0
    hru_ms, hru_n = g.kernel_profile()
    g.set_profiling(False)
    status = g.cell_status()
    # ---- leg 2: end to end through the C-ABI with host buffers.  Hourly workloads: per step the H2D of the day's forcing (the upload of
    # day d + 1 is queued before the step over day d: two device windows) + the D2H of the daily aggregates.  continental: the H2D of
    # the daily input and its disaggregation are inside the timed region, then per step the D2H of the float32 daily aggregates.
    e2e_wall = 0.0
    if not a.no_e2e:
        reset()
        if disagg:
            barrier()
            t0 = time.perf_counter()
            if nc_path:
                t1 = time.perf_counter()
                with api.NcForcing(nc_path) as nc:
                    nc.read_slab(NC_VARS, 0, ndays, cell_lat, cell_lon, out=daily_tm)
                nc_read_s = time.perf_counter() - t1
                g.disagg_tm(disagg_raw, daily_tm, want_host=False)
            else:
                g.disagg(disagg_raw, daily, want_host=False)
            for s in range(W + K):
                g.step(s * 24, 24, dmy[s * 24:s * 24 + 25], None, onp)
            barrier()
            e2e_wall = (time.perf_counter() - t0) * K / (W + K)   # the region covers W + K days; K of them are the timed steps' share
        else:
            g.set_forcing(0, fnp[0:24])
            for s in range(W):
                g.set_forcing((s + 1) * 24, fnp[(s + 1) * 24:(s + 2) * 24])
                g.step(s * 24, 24, dmy[s * 24:s * 24 + 25], None, onp)
            barrier()
            t0 = time.perf_counter()
            for s in range(W, W + K):
                if s + 1 < W + K:
                    g.set_forcing((s + 1) * 24, fnp[(s + 1) * 24:(s + 2) * 24])
                g.step(s * 24, 24, dmy[s * 24:s * 24 + 25], None, onp)
            barrier()
            e2e_wall = time.perf_counter() - t0
    if disagg and nc_path:
        os.remove(nc_path)
    # ---- the one collective of a multi-GPU run: the gather of the last day's aggregates at the end (not in any timed region above)
    gather_ms = None
    if world > 1 and not a.no_e2e:
        from vic_b200.shard import gather_cells
        barrier()
        t0 = time.perf_counter()
        allrows = gather_cells(onp[0], 0, cells * world)
        barrier()
        gather_ms = (time.perf_counter() - t0) * 1e3
        if rank == 0:
            assert allrows.shape[0] == cells * world
    # max over ranks (vic_b200/shard.py; covered by the world_size-2 gloo test)
    from vic_b200.shard import max_over_ranks
    dev_s, wall_s, e2e_s, dis_s = max_over_ranks([dev_ms / 1e3, wall, e2e_wall, disagg_s or 0.0], device="cuda")
    units = cells * RECS_PER_STEP * K * world
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6544.7))
        per_launch_bytes = wl["algo_bytes"] * cells
        hru_avg_s = hru_ms / 1e3 / max(hru_n, 1)
        achieved = per_launch_bytes / hru_avg_s / 1e9
        traffic, fp64_flop, prof_cells = None, None, 10000
        try:  # DRAM bytes and FP64 flops per launch of the same kernel from the committed ncu --set full capture (profiles/)
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[f"{wl['kernel']}:{a.workload}"]
            prof_cells = prof.get("cells", 10000)
            traffic = prof["dram_bytes_per_launch"] * (cells / prof_cells)
            fp64_flop = prof.get("fp64_flop_per_launch")
        except Exception:
            pass
        fp64 = None
        if fp64_flop:
            # SURVEY 8(d) asks for the FP64 fraction beside the HBM one: executed DFMA x 2 + DADD + DMUL per launch (ncu) over the live
            # launch time, against the DFMA throughput measured on this device just now
            try:
                peak_tf = api.measure_fp64_peak(local_rank)
                ach_tf = fp64_flop * (cells / prof_cells) / hru_avg_s / 1e12
                fp64 = {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                        "flop_per_launch": fp64_flop * (cells / prof_cells),
                        "peak_source": "vicgpu_measure_fp64_peak (register-resident DFMA loop, best of 5, this run)"}
            except Exception as e:
                fp64 = {"bound": "fp64", "unavailable": str(e)}
        state_mb = g.nhru * L.hr_stride * 8 / 1e6
        cfgd = {"workload": workload, "cells_per_gpu": cells, "hrus_per_gpu": int(g.nhru), "records_per_step": RECS_PER_STEP,
                "l2": f"no flush: what a record touches (two {state_mb:.0f} MB state buffers, the step kernel's thread-local stack, output rows, forcing) exceeds the "
                      "126 MB L2" + ("" if state_mb * 2 > 126 else " together") + "; the persistent model state is legitimately cache/HBM resident between records",
                "timing": "value: sum of CUDA-event device time of the K timed vicgpu_step calls (max over ranks); wall for the same region "
                          f"{wall_s:.3f} s", "invalid_cells": int((status != 0).sum())}
        if disagg:
            cfgd["disagg"] = (f"vicgpu_disagg of {ndays} days x {cells} cells incl. the H2D of the daily input: {dis_s:.3f} s = "
                              f"{cells * ndays / max(dis_s, 1e-9) / 1e6:.2f} M cell-days/s per GPU (outside `value`, inside `e2e`)")
        if disagg and a.forcing_nc and not a.no_e2e:
            assert np.array_equal(daily_tm, daily.transpose(1, 2, 0))  # what came out of the file is what went in
            cfgd["forcing_nc"] = (f"e2e leg: daily forcing read from a NetCDF classic (CDF-2) file of doubles, {ndays} x 4 grids of {nlat} x {nlon}, by vicgpu_nc_read_slab "
                                  f"into a pinned [day][variable][cell] buffer ({nc_read_s:.3f} s = {cells * ndays / max(nc_read_s, 1e-9) / 1e6:.1f} M cell-days/s, one host "
                                  "thread, file in the page cache), then vicgpu_disagg_tm; both inside the e2e region")
        if gather_ms is not None:
            cfgd["gather"] = f"end-of-run gather of one day's aggregates ({cells * world} cells x {L.nout} x {onp.itemsize} B) to rank 0 over NCCL: {gather_ms:.1f} ms (not in any timed region)"
        line = {"metric": "cell-timesteps/s", "value": units / dev_s, "unit": "cell-timesteps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_s / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfgd,
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "kernel": wl["kernel"], "avg_launch_us": hru_avg_s * 1e6, "launches_timed": int(hru_n),
                             "algorithmic_bytes_per_launch": per_launch_bytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6544.7",
                             "kernel_share_of_step": (hru_ms / 1e3) / dev_s if dev_s > 0 else None,
                             "note": "latency bound (thread-local memory, dependent FP64 chains), not bandwidth bound: see profiles/r02_summary.md and DESIGN.md section 6"},
                "roofline_fp64": fp64,
                "clocks": cs.summary()}
        if not a.no_e2e:
            h2d = cells * 4 * 8 if disagg else 24 * cells * L.f_stride * 8
            line["e2e"] = {"value": units / e2e_s, "unit": "cell-timesteps/s", "h2d_bytes_per_step": int(h2d),
                           "d2h_bytes_per_step": int(cells * L.nout * onp.itemsize), "seconds": e2e_s}
        if world == 1 and not a.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(a.workload, host_threads)
            except Exception as e:  # the reference binary is test infrastructure; its absence must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "cell-timesteps/s", "cores": host_threads, "kind": "reference", "sample": f"unavailable: {e}"}
        print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
