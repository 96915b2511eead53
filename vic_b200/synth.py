"""Synthetic VIC input generator (seeded, deterministic).

Writes the *unchanged* reference input file formats so that the reference's own
readers (read_soilparam.c:203-885, read_veglib.c:46-107, read_vegparam.c:117-340,
read_snowband.c:45-116, read_atmos_data.c:419-446 ASCII branch) consume them:

  <dir>/global.txt     global parameter file (get_global_param.c:308-922 key list)
  <dir>/soil.txt       one line per cell, PCIC glacier soil-file layout
  <dir>/veglib.txt     vegetation library (11 UMD-like classes + glacier class 22)
  <dir>/vegparam.txt   PCIC layout: one HRU per line incl. its band index
  <dir>/snowband.txt   only when nbands > 1
  <dir>/forc/f_<lat>_<lon>   daily ASCII forcing PREC TMAX TMIN WIND

The five BASELINE.json configurations are exposed through `CONFIGS`.
Ranges follow SURVEY.md section 8(d).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

GLACIER_ID = 22


@dataclass
class Config:
    name: str
    full_energy: bool = True
    frozen_soil: bool = False
    quick_flux: bool = True
    nodes: int = 3
    dt: int = 1
    snow_step: int = 1
    nbands: int = 1
    glacier: bool = False
    glacier_tiles: int = 1  # glacier HRUs per cell, each in its own band (>= 3: the quadratic mass-balance fit has something to fit)
    output_force: bool = False
    ntiles: int = 5
    startyear: int = 2001  # leap-free year
    ndays: int = 365
    implicit: bool = False
    exp_trans: bool = False
    noflux: bool = False
    quick_solve: bool = False
    corrprec: bool = False
    blowing: bool = False  # BLOWING TRUE: every vegetation tile carries sigma_slope, lag_one, fetch (read_vegparam.c:172-190)
    glacier_dynamics: bool = False  # GLACIER_DYNAMICS TRUE, and the glacier tile of every other cell has zero area (a placeholder the glacier model may grow)
    out_step: int = 0  # OUT_STEP [h]; 0 = every model step
    startday: int = 1  # day of January the run (and the forcing files) start on
    extra_global: list = field(default_factory=list)


CONFIGS = {
    # BASELINE.json configs[0]: water-balance mode, daily step, SNOW_STEP 3 (NF=8)
    "wb_daily": Config("wb_daily", full_energy=False, dt=24, snow_step=3),
    # configs[1]: full energy hourly, QUICK_FLUX, 5 veg tiles, 1 band  (headline bench workload)
    "fe_hourly": Config("fe_hourly"),
    # configs[2]: frozen soil + 5 bands, QUICK_FLUX=FALSE, 10 nodes
    "frozen_bands": Config("frozen_bands", frozen_soil=True, quick_flux=False, nodes=10, nbands=5),
    # the same with the implicit (Newton-Raphson, tridiagonal) soil-temperature solver, explicit scheme as its fallback
    "frozen_implicit": Config("frozen_implicit", frozen_soil=True, quick_flux=False, nodes=10, nbands=5, implicit=True),
    # full energy with 5 bands and COMPUTE_TREELINE: the reference marks the bands whose mean July air temperature is below 10 C
    # (compute_treeline.c) and put_data leaves the overstory tiles of those bands out of the cell averages (put_data.c:205, 290)
    "treeline": Config("treeline", nbands=5, startday=182, extra_global=["COMPUTE_TREELINE 10"]),
    # full energy with the gauge-undercatch correction of the precipitation (CORRPREC, correct_precip.c)
    "fe_corrprec": Config("fe_corrprec", corrprec=True),
    # full energy with sublimation of blowing snow (BLOWING, CalcBlowingSnow.c)
    "fe_blowing": Config("fe_blowing", blowing=True),
    "glacier_blowing": Config("glacier_blowing", glacier=True, nbands=5, blowing=True),
    "frozen_quick_solve": Config("frozen_quick_solve", frozen_soil=True, quick_flux=False, nodes=10, nbands=5, quick_solve=True),
    # configs[3]: PCIC glacier mass-balance mode
    "glacier": Config("glacier", glacier=True, nbands=5),
    # the same with glacier HRUs in four bands of every cell: exercises accumulateGlacierMassBalance's quadratic fit
    "glacier_multi": Config("glacier_multi", glacier=True, nbands=5, glacier_tiles=4),
    # glacier mode with GLACIER_DYNAMICS: zero-area glacier tiles are stepped too (full_energy.c:220, 389)
    "glacier_dyn": Config("glacier_dyn", glacier=True, nbands=5, glacier_dynamics=True),
    # configs[4] (first half): OUTPUT_FORCE disaggregation only
    "disagg": Config("disagg", output_force=True),
}

# (class, overstory, rarc, rmin, LAI_max, LAI_min, albedo, rough, displ, wind_h, RGL, rad_atten, wind_atten, trunk_ratio)
_VEG = [
    (1, 1, 60.0, 250.0, 3.4, 3.4, 0.12, 1.476, 8.04, 50.0, 30.0, 0.5, 0.5, 0.2),   # evergreen needleleaf
    (2, 1, 60.0, 250.0, 3.4, 3.4, 0.12, 1.476, 8.04, 50.0, 30.0, 0.5, 0.5, 0.2),   # evergreen broadleaf
    (3, 1, 60.0, 150.0, 5.0, 1.5, 0.18, 1.23, 6.70, 50.0, 30.0, 0.5, 0.5, 0.2),    # deciduous needleleaf
    (4, 1, 60.0, 150.0, 5.0, 1.5, 0.18, 1.23, 6.70, 50.0, 30.0, 0.5, 0.5, 0.2),    # deciduous broadleaf
    (5, 1, 60.0, 200.0, 5.0, 1.5, 0.18, 1.23, 6.70, 50.0, 50.0, 0.5, 0.5, 0.2),    # mixed
    (6, 1, 60.0, 200.0, 4.0, 1.5, 0.18, 1.23, 6.70, 50.0, 50.0, 0.5, 0.5, 0.2),    # woodland
    (7, 0, 40.0, 125.0, 3.5, 0.7, 0.19, 0.495, 1.0, 10.0, 75.0, 0.5, 0.5, 0.2),    # wooded grassland
    (8, 0, 50.0, 135.0, 3.0, 0.5, 0.19, 0.495, 1.0, 10.0, 75.0, 0.5, 0.5, 0.2),    # closed shrub
    (9, 0, 50.0, 135.0, 2.5, 0.3, 0.19, 0.495, 1.0, 10.0, 75.0, 0.5, 0.5, 0.2),    # open shrub
    (10, 0, 25.0, 120.0, 3.0, 0.3, 0.20, 0.0738, 0.402, 10.0, 100.0, 0.5, 0.5, 0.2),  # grassland
    (11, 0, 25.0, 120.0, 4.5, 0.0, 0.10, 0.006, 1.005, 10.0, 100.0, 0.5, 0.5, 0.2),   # cropland (LAI 0 in winter)
    (GLACIER_ID, 0, 0.0, 0.0, 0.0, 0.0, 0.30, 0.002, 0.0, 10.0, 100.0, 0.0, 0.0, 0.0),  # glacier
]


def _monthly(vmax, vmin):
    # northern-hemisphere seasonal cycle peaking in July
    out = []
    for m in range(12):
        w = 0.5 * (1.0 - math.cos(2.0 * math.pi * (m - 0.5) / 12.0))
        out.append(vmin + (vmax - vmin) * w)
    return out


def write_veglib(path):
    with open(path, "w") as f:
        f.write("#Class OvrStry Rarc Rmin JAN-LAI..DEC-LAI ALB x12 ROU x12 DIS x12 WIND_H RGL rad_atn wnd_atn trnk_rat COMMENT\n")
        for (cls, over, rarc, rmin, lmax, lmin, alb, rough, displ, wind_h, rgl, ra, wa, tr) in _VEG:
            lai = _monthly(lmax, lmin)
            if cls == 11:
                lai = [0.0 if (m < 3 or m > 9) else v for m, v in enumerate(lai)]
            vals = [f"{cls:d}", f"{over:d}", f"{rarc:.4f}", f"{rmin:.4f}"]
            vals += [f"{v:.4f}" for v in lai]
            vals += [f"{alb:.4f}"] * 12
            vals += [f"{rough:.4f}"] * 12
            # displacement must be > 0 where LAI > 0 (read_veglib.c:78-82)
            vals += [f"{displ:.4f}"] * 12
            vals += [f"{wind_h:.4f}", f"{rgl:.4f}", f"{ra:.4f}", f"{wa:.4f}", f"{tr:.4f}", f"class{cls}"]
            f.write(" ".join(vals) + "\n")


def _fmt_coord(v, dec):
    return f"{v:.{dec}f}"


def generate(outdir, config: Config | str, nlat=4, nlon=4, seed=1234, threads=1,
             lat0=48.03125, lon0=-121.96875, res=0.0625, grid_decimal=5, forcing=True,
             result_dir=None):
    """Write all input files for `nlat*nlon` cells; returns dict with paths + metadata."""
    if isinstance(config, str):
        config = CONFIGS[config]
    cfg = config
    rng = np.random.default_rng(seed)
    os.makedirs(outdir, exist_ok=True)
    fdir = os.path.join(outdir, "forc")
    os.makedirs(fdir, exist_ok=True)
    rdir = result_dir or os.path.join(outdir, "results")
    os.makedirs(rdir, exist_ok=True)
    ncell = nlat * nlon
    nl = 3

    write_veglib(os.path.join(outdir, "veglib.txt"))

    lats = [lat0 + i * res for i in range(nlat)]
    lons = [lon0 + j * res for j in range(nlon)]
    soil_lines, veg_lines, band_lines = [], [], []
    cells = []
    cid = 0
    for la in lats:
        for lo in lons:
            cid += 1
            b_infilt = rng.uniform(0.05, 0.4)
            Ds = rng.uniform(0.001, 0.1)
            Dsmax = rng.uniform(5.0, 30.0)
            Ws = rng.uniform(0.6, 0.95)
            c = 2.0
            expt = rng.uniform(8.0, 20.0, nl)
            Ksat = np.exp(rng.uniform(math.log(50.0), math.log(2000.0), nl))
            phi_s = np.full(nl, -999.0)
            depth = np.array([0.1, rng.uniform(0.2, 0.5), rng.uniform(0.5, 2.0)])
            bulk = rng.uniform(1300.0, 1600.0, nl)
            sdens = np.full(nl, 2650.0)
            poros = 1.0 - bulk / sdens
            max_moist = depth * poros * 1000.0
            init_moist = 0.6 * max_moist
            elev = rng.uniform(200.0, 3000.0)
            avg_temp = rng.uniform(-2.0, 8.0)
            dp = 4.0
            bubble = rng.uniform(5.0, 30.0, nl)
            quartz = rng.uniform(0.2, 0.8, nl)
            off_gmt = -8.0
            wcr = rng.uniform(0.4, 0.6, nl)
            wpwp = rng.uniform(0.15, 0.3, nl)
            rough, snow_rough = 0.01, 0.001
            annual_prec = rng.uniform(300.0, 2000.0)
            resid = np.full(nl, 0.02)
            fs_active = 1 if cfg.frozen_soil else 0
            glac = [0.85, 0.94, 0.58, 0.82, 0.46,  # NEW_SNOW_ALB, ACCUM_A, ACCUM_B, THAW_A, THAW_B
                    rng.uniform(0.5, 1.5),          # MIN_RAIN_TEMP == TT under KIENZLE
                    rng.uniform(8.0, 13.0),         # MAX_SNOW_TEMP == TR under KIENZLE
                    1.0, 1.0, 6.5,
                    0.0003 if cfg.nbands > 1 else 0.0,  # PGRAD
                    100.0, 91.7, 0.01, 0.24, 20.0, 0.3, 0.002]
            vals = [1, cid, la, lo, b_infilt, Ds, Dsmax, Ws, c]
            vals += list(expt) + list(Ksat) + list(phi_s) + list(init_moist)
            vals += [elev] + list(depth) + [avg_temp, dp]
            vals += list(bubble) + list(quartz) + list(bulk) + list(sdens)
            vals += [off_gmt] + list(wcr) + list(wpwp) + [rough, snow_rough, annual_prec]
            vals += list(resid) + [fs_active]
            vals += glac
            toks = []
            for k, v in enumerate(vals):
                if k in (0, 1) or (isinstance(v, (int, np.integer)) and not isinstance(v, bool)):
                    toks.append(str(int(v)))
                elif k in (2, 3):
                    toks.append(_fmt_coord(v, grid_decimal))
                else:
                    toks.append(f"{float(v):.6f}")
            soil_lines.append(" ".join(toks))

            # --- elevation bands
            if cfg.nbands > 1:
                af = rng.dirichlet(np.full(cfg.nbands, 4.0))
                af = np.round(af, 4)
                af[-1] = round(1.0 - af[:-1].sum(), 4)
                spread = rng.uniform(300.0, 900.0)
                be = elev + np.linspace(-0.5, 0.5, cfg.nbands) * spread
                be = np.maximum(be, 10.0)
                band_lines.append(" ".join([str(cid)] + [f"{x:.4f}" for x in af] + [f"{x:.2f}" for x in be]))

            # --- vegetation tiles: 2 overstory + 3 short, Cv summing to 1 (config 2) or <1 (bare soil HRU added)
            over_cls = rng.choice([1, 3, 4, 5, 6], size=2, replace=False)
            short_cls = rng.choice([7, 8, 9, 10, 11], size=cfg.ntiles - 2, replace=False)
            classes = list(over_cls) + list(short_cls)
            tiles = []
            if cfg.glacier:
                classes = classes[:-1] + [GLACIER_ID] * cfg.glacier_tiles
            w = rng.dirichlet(np.full(len(classes), 3.0))
            # leave bare soil in ~1/4 of the cells so the artificial bare-soil HRU path is exercised
            tot = 1.0 if rng.uniform() > 0.25 else rng.uniform(0.7, 0.95)
            cv = np.round(w * tot, 4)
            if tot == 1.0:
                cv[0] = round(1.0 - cv[1:].sum(), 4)
            nglac = 0
            for k, cl in enumerate(classes):
                if cfg.nbands > 1:
                    if cl == GLACIER_ID and cfg.glacier_tiles > 1:
                        band = cfg.nbands - 1 - nglac  # one glacier HRU per band, from the top down
                        nglac += 1
                    elif cl == GLACIER_ID:
                        band = cfg.nbands - 1 - int(rng.integers(0, 2))
                    else:
                        band = int(rng.integers(0, cfg.nbands))
                else:
                    band = 0
                if cl <= 6:
                    zones = [(0.10, 0.10), (0.60, 0.50), (1.50, 0.40)]
                else:
                    zones = [(0.10, 0.30), (0.50, 0.50), (1.00, 0.20)]
                tiles.append((int(cl), float(cv[k]), zones, band))
            if cfg.glacier_dynamics and cfg.glacier and (len(cells) % 2 == 0):
                # zero-area glacier tile: its area goes to the first tile
                gi = [k for k, tl in enumerate(tiles) if tl[0] == GLACIER_ID][0]
                tiles[0] = (tiles[0][0], round(tiles[0][1] + tiles[gi][1], 4), tiles[0][2], tiles[0][3])
                tiles[gi] = (tiles[gi][0], 0.0, tiles[gi][2], tiles[gi][3])
            veg_lines.append(f"{cid} {len(tiles)}")
            for (cl, cvv, zones, band) in tiles:
                z = " ".join(f"{d:.2f} {fr:.2f}" for d, fr in zones)
                bl = f" {rng.uniform(0.02, 0.12):.4f} {rng.uniform(0.80, 0.98):.3f} {rng.uniform(200.0, 3000.0):.1f}" if cfg.blowing else ""
                veg_lines.append(f"  {cl} {cvv:.4f} {z}{bl} {band}")
            cells.append(dict(id=cid, lat=la, lon=lo, elev=elev, avg_temp=avg_temp))

            # --- daily forcing
            if forcing:
                nd = cfg.ndays
                doy = np.arange(nd) + (cfg.startday - 1)
                tmean = avg_temp + 2.0 - 0.0065 * 0.0 + 12.0 * np.sin(2 * np.pi * (doy - 105) / 365.0) \
                    - 0.004 * (elev - 1000.0) + rng.normal(0.0, 2.0, nd)
                dtr = rng.uniform(6.0, 12.0, nd)
                tmax = tmean + dtr / 2
                tmin = tmean - dtr / 2
                wet = rng.uniform(size=nd) < 0.4
                prec = np.where(wet, rng.gamma(0.6, 6.0, nd), 0.0)
                wind = rng.uniform(1.0, 5.0, nd)
                fn = os.path.join(fdir, f"f_{_fmt_coord(la, grid_decimal)}_{_fmt_coord(lo, grid_decimal)}")
                with open(fn, "w") as f:
                    for d in range(nd):
                        f.write(f"{prec[d]:.4f} {tmax[d]:.4f} {tmin[d]:.4f} {wind[d]:.4f}\n")

    with open(os.path.join(outdir, "soil.txt"), "w") as f:
        f.write("\n".join(soil_lines) + "\n")
    with open(os.path.join(outdir, "vegparam.txt"), "w") as f:
        f.write("\n".join(veg_lines) + "\n")
    if cfg.nbands > 1:
        with open(os.path.join(outdir, "snowband.txt"), "w") as f:
            f.write("\n".join(band_lines) + "\n")

    # --- end date
    import datetime
    d0 = datetime.date(cfg.startyear, 1, 1) + datetime.timedelta(days=cfg.startday - 1)
    d1 = d0 + datetime.timedelta(days=cfg.ndays - 1)

    def tf(b):
        return "TRUE" if b else "FALSE"

    g = []
    g += [f"NLAYER 3", f"NODES {cfg.nodes}", f"TIME_STEP {cfg.dt}", f"SNOW_STEP {cfg.snow_step}",
          f"STARTYEAR {d0.year}", f"STARTMONTH {d0.month}", f"STARTDAY {d0.day}", "STARTHOUR 0",
          f"ENDYEAR {d1.year}", f"ENDMONTH {d1.month}", f"ENDDAY {d1.day}",
          f"FULL_ENERGY {tf(cfg.full_energy)}", f"FROZEN_SOIL {tf(cfg.frozen_soil)}",
          f"QUICK_FLUX {tf(cfg.quick_flux)}", f"NO_FLUX {tf(cfg.noflux)}",
          f"IMPLICIT {tf(cfg.implicit)}", f"EXP_TRANS {tf(cfg.exp_trans)}",
          f"QUICK_SOLVE {tf(cfg.quick_solve)}", "SNOW_ALBEDO USACE", "SNOW_DENSITY DENS_BRAS", f"BLOWING {tf(cfg.blowing)}",
          "DIST_PRCP FALSE", f"CORRPREC {tf(cfg.corrprec)}", "MIN_WIND_SPEED 0.1", "CONTINUEONERROR TRUE",
          "TFALLBACK TRUE", "COMPUTE_TREELINE FALSE", "EQUAL_AREA FALSE", f"RESOLUTION {res}",
          "AERO_RESIST_CANSNOW AR_406_FULL", "GRND_FLUX_TYPE GF_410", "PLAPSE TRUE",
          "MTCLIM_SWE_CORR TRUE", "VP_ITER VP_ITER_ALWAYS", "VP_INTERP TRUE", "LW_TYPE LW_TVA",
          "LW_CLOUD LW_CLOUD_DEARDORFF", f"PARALLEL_THREADS {threads}"]
    if cfg.glacier:
        g += [f"GLACIER_ID {GLACIER_ID}", f"GLACIER_DYNAMICS {tf(cfg.glacier_dynamics)}",
              f"GLACIER_ACCUM_START_YEAR {d0.year}", "GLACIER_ACCUM_START_MONTH 1",
              "GLACIER_ACCUM_START_DAY 2", "GLACIER_ACCUM_INTERVAL 1"]
    else:
        g += [f"GLACIER_ID {GLACIER_ID}"]
    g += [f"FORCING1 {fdir}/f_", "FORCE_FORMAT ASCII", "FORCE_ENDIAN LITTLE", "N_TYPES 4",
          "FORCE_TYPE PREC", "FORCE_TYPE TMAX", "FORCE_TYPE TMIN", "FORCE_TYPE WIND",
          "FORCE_DT 24", f"FORCEYEAR {d0.year}", f"FORCEMONTH {d0.month}", f"FORCEDAY {d0.day}", "FORCEHOUR 0",
          f"GRID_DECIMAL {grid_decimal}", "WIND_H 10.0", "MEASURE_H 2.0", "ALMA_INPUT FALSE"]
    if cfg.output_force:
        g += ["OUTPUT_FORCE TRUE"]
    g += [f"SOIL {outdir}/soil.txt", "ARC_SOIL FALSE", "BASEFLOW ARNO", "JULY_TAVG_SUPPLIED FALSE",
          "ORGANIC_FRACT FALSE", f"VEGLIB {outdir}/veglib.txt", f"VEGPARAM {outdir}/vegparam.txt",
          "ROOT_ZONES 3", "VEGPARAM_LAI FALSE", "LAI_SRC LAI_FROM_VEGLIB"]
    if cfg.nbands > 1:
        g += [f"SNOW_BAND {cfg.nbands} {outdir}/snowband.txt"]
    else:
        g += ["SNOW_BAND 1"]
    g += [f"RESULT_DIR {rdir}", f"OUT_STEP {cfg.out_step}", "SKIPYEAR 0", "COMPRESS FALSE", "OUTPUT_FORMAT BINARY",
          "ALMA_OUTPUT FALSE", "MOISTFRACT FALSE", "PRT_HEADER FALSE", "PRT_SNOW_BAND FALSE"]
    g += cfg.extra_global
    gpath = os.path.join(outdir, "global.txt")
    with open(gpath, "w") as f:
        f.write("\n".join(g) + "\n")
    return dict(global_file=gpath, dir=outdir, ncell=ncell, cells=cells, config=cfg)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("outdir")
    ap.add_argument("--config", default="fe_hourly", choices=sorted(CONFIGS))
    ap.add_argument("--nlat", type=int, default=4)
    ap.add_argument("--nlon", type=int, default=4)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--ndays", type=int, default=None)
    ap.add_argument("--no-forcing", action="store_true")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    if a.ndays:
        import dataclasses
        cfg = dataclasses.replace(cfg, ndays=a.ndays)
    r = generate(a.outdir, cfg, a.nlat, a.nlon, a.seed, forcing=not a.no_forcing)
    print(r["global_file"])
