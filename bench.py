#!/usr/bin/env python
"""Benchmark of the hot path: cell-timesteps/s of the full-energy hourly VIC cell loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--cells C] [--impl ours|reference]

Workload (BASELINE.json configs[1]): FULL_ENERGY=TRUE, QUICK_FLUX=TRUE, 3 soil layers, 3 thermal nodes, 5 vegetation
tiles (2 overstory + 3 short, + the automatic bare-soil HRU where the tiles do not cover the cell), 1 snow band,
hourly step, 10,000 cells per GPU.  One bench "step" = one day = 24 hourly records over every cell of the rank
(240,000 cell-timesteps per GPU); the default K + W = 365 steps is the configuration's full year.

Parameters and initial state: bench_data/base_fe_hourly.npz (256 cells produced by the reference's own readers and
initialisation code, see bench_data/make_base.py) tiled to the domain size with perturbed infiltration / baseflow /
conductivity parameters; forcing: synthetic hourly weather generated here (seeded).  Cells are independent, so with
N GPUs every rank owns its own 10,000 cells (weak scaling) and there is no collective on the data path.

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
BASE_SEED = 20260            # seed of the 16x16 base domain (bench_data/make_base.py)
# algorithmic bytes per cell-timestep of this configuration (H=5, Nn=3, Tb=1)
ALGO_BYTES_PER_CELL_STEP = 7.0e3
SIGMA = 5.6696e-8


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
def run_reference(ncell_sample, ndays, warm_days, threads, seed):
    """the reference's own CPU implementation (oracle/_ref/vic_ref_harness = its unmodified sources, OpenMP cell loop of
    vicNl.c:514-517) on the base domain's first cells with the same forcing generator; returns (cell_steps_per_s, seconds)"""
    import dataclasses
    from vic_b200 import synth
    from vic_b200.casefile import write_case
    harness = os.path.join(ROOT, "oracle", "_ref", "vic_ref_harness")
    if not os.path.exists(harness):
        raise FileNotFoundError(harness)
    side = 16
    assert ncell_sample == side * side
    dom = build_domain(ncell_sample, seed)
    with tempfile.TemporaryDirectory() as d:
        cfg = dataclasses.replace(synth.CONFIGS["fe_hourly"], ndays=ndays, out_step=24)
        r = synth.generate(d, cfg, side, side, BASE_SEED, forcing=False, threads=threads)
        L = layout_from_options(parse_options(dom["options_raw"]))
        f = np.empty((ndays * 24, ncell_sample, L.f_stride))
        for day in range(ndays):
            forcing_day(dom, day, seed, f[day * 24:(day + 1) * 24])
        fb = os.path.join(d, "forcing.bin")
        write_case(fb, {"forcing": f})
        o = subprocess.run([harness, "-g", r["global_file"], "--forcing-bin", fb, "--time-only", "--threads", str(threads), "--time-from", str(warm_days * 24)],
                           capture_output=True, text=True, check=True).stdout
    line = [x for x in o.splitlines() if x.startswith("run_seconds")][0].split()
    return float(line[7]), float(line[1])


# ---------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=362)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cells", type=int, default=10000, help="cells per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = a.steps, max(a.warmup, 0)
    host_threads = os.cpu_count() or 1
    workload = f"fe_hourly: FULL_ENERGY hourly, QUICK_FLUX, 3 layers, 3 nodes, 5 veg tiles, 1 band, {a.cells} cells/GPU, 1 step = 24 hourly records"

    if a.impl == "reference":
        if rank != 0:
            return
        ncs = 256
        v, secs = run_reference(ncs, W + K, W, host_threads, seed=1)
        line = {"impl": "reference", "metric": "cell-timesteps/s", "value": v, "unit": "cell-timesteps/s", "n_gpus": a.gpus, "steps": K, "warmup": W,
                "ms_per_step": secs / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "sample": f"{ncs} cells (the base domain) x {K * 24} hourly records per run; reference CPU build, OpenMP cell loop"},
                "cpu_baseline": {"value": v, "unit": "cell-timesteps/s", "cores": host_threads, "kind": "reference",
                                 "sample": f"{ncs} cells x {K * 24} records, {host_threads} OpenMP threads"},
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
    dom = build_domain(a.cells, seed)
    g = api.VicGpu(dom["options_raw"], device=local_rank)
    L = g.L
    g.set_veglib(dom["veglib"])
    g.set_output_spec(dom["aggtype"])
    ndays = W + K
    nrec = ndays * RECS_PER_STEP
    dmy = make_dmy(nrec)
    # pinned host buffers: the year's forcing and one day's aggregated output
    fbuf = torch.empty((nrec, a.cells, L.f_stride), dtype=torch.float64, pin_memory=True)
    fnp = fbuf.numpy()
    for day in range(ndays):
        forcing_day(dom, day, seed, fnp[day * 24:(day + 1) * 24])
    obuf = torch.empty((1, a.cells, L.nout), dtype=torch.float64, pin_memory=True)
    onp = obuf.numpy()

    def reset():
        g.set_cells(dom["cellpar"], dom["hrupar"])
        g.set_state(dom["hrurec0"])

    # ---- leg 1: inputs resident in HBM, device-side daily aggregation, nothing copied back
    reset()
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
    # ---- leg 2: end to end through the C-ABI with host buffers: per step H2D of the day's forcing + D2H of the daily output
    e2e = None
    if not a.no_e2e:
        reset()
        for s in range(W):
            g.set_forcing(s * 24, fnp[s * 24:s * 24 + 24])
            g.step(s * 24, 24, dmy[s * 24:s * 24 + 25], None, onp)
        barrier()
        t0 = time.perf_counter()
        for s in range(W, W + K):
            g.set_forcing(s * 24, fnp[s * 24:s * 24 + 24])
            g.step(s * 24, 24, dmy[s * 24:s * 24 + 25], None, onp)
        barrier()
        e2e_wall = time.perf_counter() - t0
    # max over ranks (vic_b200/shard.py; covered by the world_size-2 gloo test)
    from vic_b200.shard import max_over_ranks
    dev_s, wall_s, e2e_s = max_over_ranks([dev_ms / 1e3, wall, e2e_wall if not a.no_e2e else 0.0], device="cuda")
    units = a.cells * RECS_PER_STEP * K * world
    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6544.7))
        recs_per_launch = (K * RECS_PER_STEP) / max(hru_n, 1)   # 1 unless VICGPU_RECBLOCK > 1
        per_launch_bytes = ALGO_BYTES_PER_CELL_STEP * a.cells * recs_per_launch
        hru_avg_s = hru_ms / 1e3 / max(hru_n, 1)
        achieved = per_launch_bytes / hru_avg_s / 1e9
        traffic, fp64_flop = None, None
        try:  # DRAM bytes and FP64 flops per launch of the same kernel from the committed ncu --set full capture (profiles/)
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_hru_step_nn3"]
            traffic = prof["dram_bytes_per_launch"]
            fp64_flop = prof.get("fp64_flop_per_launch")
        except Exception:
            pass
        fp64 = None
        if fp64_flop:
            # SURVEY 8(d) asks for the FP64 fraction beside the HBM one: executed DFMA x 2 + DADD + DMUL per launch (ncu, per 10,000 cells)
            # over the live launch time, against the DFMA throughput measured on this device just now
            try:
                peak_tf = api.measure_fp64_peak(local_rank)
                ach_tf = fp64_flop * (a.cells / 10000.0) * recs_per_launch / hru_avg_s / 1e12
                fp64 = {"bound": "fp64", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                        "flop_per_launch": fp64_flop * (a.cells / 10000.0) * recs_per_launch,
                        "peak_source": "vicgpu_measure_fp64_peak (register-resident DFMA loop, best of 5, this run)"}
            except Exception as e:
                fp64 = {"bound": "fp64", "unavailable": str(e)}
        line = {"metric": "cell-timesteps/s", "value": units / dev_s, "unit": "cell-timesteps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_s / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "cells_per_gpu": a.cells, "hrus_per_gpu": int(g.nhru), "records_per_step": RECS_PER_STEP,
                           "l2": "no flush: what a record touches (two 76 MB state buffers, 380 MB of thread-local stack, 33 MB of output rows, forcing) exceeds the 126 MB L2; "
                                 "the persistent model state is legitimately cache/HBM resident between records",
                           "timing": "value: sum of CUDA-event device time of the K timed vicgpu_step calls (max over ranks); wall for the same region "
                                     f"{wall_s:.3f} s", "invalid_cells": int((status != 0).sum())},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "kernel": "k_hru_step_nn3", "avg_launch_us": hru_avg_s * 1e6, "launches_timed": int(hru_n),
                             "algorithmic_bytes_per_launch": per_launch_bytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6544.7",
                             "kernel_share_of_step": (hru_ms / 1e3) / dev_s if dev_s > 0 else None,
                             "note": "latency bound (thread-local memory, dependent FP64 chains), not bandwidth bound: see profiles/r01c_summary.md and DESIGN.md section 6"},
                "roofline_fp64": fp64,
                "clocks": cs.summary()}
        if not a.no_e2e:
            line["e2e"] = {"value": units / e2e_s, "unit": "cell-timesteps/s", "h2d_bytes_per_step": int(24 * a.cells * L.f_stride * 8),
                           "d2h_bytes_per_step": int(a.cells * L.nout * 8), "seconds": e2e_s}
        if world == 1 and not a.no_cpu_baseline:
            try:
                sample_days = 30
                v, secs = run_reference(256, sample_days, 0, host_threads, seed=1)
                line["cpu_baseline"] = {"value": v, "unit": "cell-timesteps/s", "cores": host_threads, "kind": "reference",
                                        "sample": f"first 256 cells of the domain x {sample_days * 24} hourly records, same forcing generator; "
                                                  f"reference build (oracle/_ref/vic_ref_harness), {host_threads} OpenMP threads, {secs:.2f} s"}
            except Exception as e:  # the reference binary is test infrastructure; its absence must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "cell-timesteps/s", "cores": host_threads, "kind": "reference", "sample": f"unavailable: {e}"}
        print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
