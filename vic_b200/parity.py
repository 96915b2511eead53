"""Parity metrics between a result (CUDA library or host port) and the reference's answers of a case file.

Tolerances follow BASELINE.json north_star: per-step state 1e-9 relative, annual totals 1e-6 relative.
Relative error of a column is |a-b| / max(|b|, floor) with a per-column floor of 1e-3 x the column's largest
magnitude (so an entry that happens to be near zero must still agree to 1e-12 of the column's scale), so that values that are sums of cancelling terms (balance errors, fluxes near zero)
are judged against the size of their terms and not against their own near-zero value.
"""
from __future__ import annotations

import numpy as np

from .layout import layout_from_options, parse_options


def rel_err(a, b, floor):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(a - b)
    d[both_nan] = 0.0
    d[np.isnan(d)] = np.inf  # NaN on one side only
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    d[same_inf] = 0.0
    den = np.maximum(np.abs(b), floor)
    den = np.where(np.isfinite(den), den, 1.0)
    return d / den


# Columns that ARE residuals of a balance (root-finder residuals, closure errors): their own magnitude is
# rounding noise of the terms they are the sum of, so they are judged against the size of those terms
# (energy fluxes O(100) W/m2 -> 1 W/m2 is a conservative unit; water storages O(100) mm -> 1 mm; SWE -> 1e-3 m).
# Runoff is formed as inflow - top_max_moist + top_moist + top_max_moist*basis^(1+b) (runoff.c:808-809), a difference of
# O(100 mm) storages, and the DEL* outputs are differences of two storages: 1e-3 mm is 1e-5 of the terms.
RESIDUAL_FLOORS = {"E_error": 1.0, "ENERGY_ERROR": 1.0, "WATER_ERROR": 1.0, "S_mass_error": 1e-3, "S_Qnet": 1.0, "G_Qnet": 1.0,
                   "E_AtmosError": 1.0, "RUNOFF": 1e-3, "C_runoff": 1e-3, "DELSOILMOIST": 1e-3, "DELSWE": 1e-3, "DELINTERCEPT": 1e-3,
                   "DELSURFSTOR": 1e-3}


def column_report(got, ref, names, scale_floor=1e-3, abs_floor=1e-9):
    """got/ref: [..., ncol]; returns list of (name, max_rel_err, argmax index) sorted worst first."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    ncol = ref.shape[-1]
    g = got.reshape(-1, ncol)
    r = ref.reshape(-1, ncol)
    out = []
    for c in range(ncol):
        fin = r[:, c][np.isfinite(r[:, c])]
        mag = np.max(np.abs(fin)) if fin.size else 0.0
        floor = max(mag * scale_floor, abs_floor, RESIDUAL_FLOORS.get(names[c].split("[")[0], 0.0))
        e = rel_err(g[:, c], r[:, c], floor)
        k = int(np.argmax(e)) if e.size else 0
        out.append((names[c], float(e[k]) if e.size else 0.0, k))
    out.sort(key=lambda x: -x[1])
    return out


def row_errors(got, ref, names, scale_floor=1e-3, abs_floor=1e-9):
    """got/ref: [nrec, nrow, ncol] -> [nrec, nrow] largest relative error of each row, same floors as column_report"""
    got = np.asarray(got)
    ref = np.asarray(ref)
    worst = np.zeros(ref.shape[:-1])
    for c in range(ref.shape[-1]):
        fin = ref[..., c][np.isfinite(ref[..., c])]
        mag = np.max(np.abs(fin)) if fin.size else 0.0
        floor = max(mag * scale_floor, abs_floor, RESIDUAL_FLOORS.get(names[c].split("[")[0], 0.0))
        worst = np.maximum(worst, rel_err(got[..., c], ref[..., c], floor))
    return worst


def compare_case(case, result, keys=(("out", "out_ref"), ("hrurec", "hrurec_ref"), ("agg", "agg_ref"))):
    """Returns dict key -> column_report list, plus integer checks."""
    opt = parse_options(case["options_raw"])
    L = layout_from_options(opt)
    rep = {}
    for k, kr in keys:
        if k not in result or kr not in case:
            continue
        got, ref = result[k], case[kr]
        n = min(got.shape[0], ref.shape[0])
        names = L.out_names if k in ("out", "agg") else L.hru_names
        rep[k] = column_report(got[:n], ref[:n], names)
    if "status" in result and "status_ref" in case:
        rep["status_equal"] = bool(np.array_equal(result["status"], case["status_ref"]))
    return rep, L


INT_SUBSTR = ("fbflag", "fbcount")
INT_EXACT = ("S_last_snow", "S_MELTING", "S_snow", "S_store_snow", "E_Nfrost", "E_Nthaw", "E_frozen")


def integer_mismatches(got, ref, names):
    """bit-exactness of integer bookkeeping columns (counters, flags, last_snow)"""
    bad = {}
    ncol = ref.shape[-1]
    g = got.reshape(-1, ncol)
    r = ref.reshape(-1, ncol)
    for c, n in enumerate(names):
        if any(t in n for t in INT_SUBSTR) or n in INT_EXACT:
            m = int(np.sum(~((g[:, c] == r[:, c]) | (np.isnan(g[:, c]) & np.isnan(r[:, c])))))
            if m:
                bad[n] = m
    return bad
