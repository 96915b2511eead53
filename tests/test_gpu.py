"""GPU parity tests (-m gpu): the CUDA library, called through its C-ABI, against (1) the committed golden vectors generated
from the reference and (2) the reference build itself (oracle/_ref/vic_ref_harness travels to the GPU box) on freshly generated,
larger synthetic domains.

ONE bar: BIT-EXACT against the reference linked with the platform's own glibc -- every state column, every output variable of
every record, the disaggregated forcing, counters, balance errors, cell status -- for any run length (full years are run below).
The kernels' elementary functions are the operation-by-operation restatement of that glibc (vic_b200/csrc/vic_glibm.cuh), so
there is no tolerance to state; the north_star bars (1e-9 per step, 1e-6 on annual totals) are asserted on top, per cell."""
import dataclasses
import os
import subprocess

import numpy as np
import pytest

from vic_b200 import api, synth
from vic_b200.casefile import read_case
from vic_b200.layout import TABLES, layout_from_options, parse_options
from vic_b200.parity import column_report, integer_mismatches, row_errors

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-9
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def _check(res, ref, keys, L, tol=TOL_STEP):
    """tol == 0: bit-exact (NaN == NaN: the reference's INVALID sentinel)"""
    for k, kr, names in keys:
        n = min(res[k].shape[0], ref[kr].shape[0])
        worst = column_report(res[k][:n], ref[kr][:n], names)[:3]
        if tol == 0:
            assert np.array_equal(res[k][:n], ref[kr][:n], equal_nan=True), (k, worst)
        else:
            assert worst[0][1] < tol, (k, worst)


def _tol(name):
    return 0


@pytest.mark.parametrize("name", GOLDEN)
def test_golden_case(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    L = layout_from_options(parse_options(g["options_raw"]))
    nrec = None
    res = api.run_case(g, device=0, nrec=nrec)
    tol = _tol(name)
    _check(res, g, (("hrurec", "hrurec_ref", L.hru_names),), L, tol)
    # the reference's first aggregate holds uninitialised heap memory (see tests/test_cpu.py): compare from the second on
    na = res["agg"].shape[0]
    if na > 1:
        _check({"agg": res["agg"][1:]}, {"agg_ref": g["agg_ref"][1:na]}, (("agg", "agg_ref", L.out_names),), L, tol)
    _check({"out": res["out"][:24]}, {"o": g["out_ref_head"]}, (("out", "o", L.out_names),), L, tol)
    nd = res["hrurec"].shape[0]
    assert integer_mismatches(res["hrurec"], g["hrurec_ref"][:nd], L.hru_names) == {}
    assert np.array_equal(res["status"], g["status_ref"])
    if nrec is None:
        _check({"out": res["out"][-24:]}, {"o": g["out_ref_tail"]}, (("out", "o", L.out_names),), L, tol)
        # balance errors: cumulative sums of per-step residuals that are ~1e-13 each; compare absolutely
        if tol == 0:
            assert np.array_equal(res["balance"], g["balance_ref"], equal_nan=True)
        else:
            assert np.nanmax(np.abs(res["balance"][:, 1:] - g["balance_ref"][:, 1:])) < 1e-6


def _reference_case(harness, cfgname, nlat, nlon, ndays, seed, tmp_path):
    cfg = dataclasses.replace(synth.CONFIGS[cfgname], ndays=ndays)
    r = synth.generate(str(tmp_path / "in"), cfg, nlat, nlon, seed)
    case = str(tmp_path / "case.bin")
    # (the cells are independent: the reference's OpenMP cell loop over all host threads computes the same bits as one thread)
    subprocess.run([harness, "-g", r["global_file"], "-o", case, "--dump-every", "240", "--threads", str(os.cpu_count() or 1)], check=True,
                   stdout=subprocess.DEVNULL)
    return read_case(case)


@pytest.mark.parametrize("cfgname,nlat,nlon,ndays,seed", [("fe_hourly", 6, 6, 365, 201), ("wb_daily", 5, 5, 365, 202), ("glacier", 4, 4, 365, 203), ("frozen_bands", 2, 3, 365, 204), ("frozen_implicit", 2, 3, 45, 206), ("treeline", 3, 3, 40, 901), ("fe_corrprec", 3, 3, 40, 555), ("glacier_dyn", 3, 3, 40, 666), ("fe_blowing", 3, 3, 30, 777), ("glacier_blowing", 3, 3, 10, 777)])
def test_bit_exact_against_reference_build(cfgname, nlat, nlon, ndays, seed, ref_harness, tmp_path):
    """the reference's own CPU build (its sources, the platform's glibc) over a FULL YEAR, frozen soil with 10 thermal nodes and
    5 snow bands included (freeze-up, winter, thaw; explicit scheme, and the IMPLICIT Newton-Raphson scheme with its explicit fallback): every record's 184 outputs, the state at every 240th record, balance errors and
    status are bit-identical; annual runoff / baseflow / SWE / glacier mass balance per cell therefore too (north_star asks 1e-6)"""
    from test_cpu import annual_totals_match
    c = _reference_case(ref_harness, cfgname, nlat, nlon, ndays, seed, tmp_path)
    L = layout_from_options(parse_options(c["options_raw"]))
    res = api.run_case(c, device=0)
    _check(res, c, (("out", "out_ref", L.out_names), ("hrurec", "hrurec_ref", L.hru_names)), L, 0)
    assert np.array_equal(res["status"], c["status_ref"])
    assert np.array_equal(res["balance"], c["balance_ref"], equal_nan=True)
    annual_totals_match(res["out"], c["out_ref"], list(L.out_names))


@pytest.mark.parametrize("over", [dict(exp_trans=True, noflux=True), dict(implicit=True, exp_trans=True, noflux=True), dict(quick_solve=True)],
                         ids=lambda o: "+".join(sorted(o)))
def test_soil_thermal_options_against_reference_build(over, ref_harness, tmp_path):
    """EXP_TRANS + NO_FLUX with the explicit and with the IMPLICIT soil-temperature scheme, and QUICK_SOLVE (frozen soil, ten nodes, five
    bands), 4 winter days"""
    cfg = dataclasses.replace(synth.CONFIGS["frozen_bands"], ndays=4, **over)
    r = synth.generate(str(tmp_path / "in"), cfg, 2, 2, 333)
    case = str(tmp_path / "case.bin")
    subprocess.run([ref_harness, "-g", r["global_file"], "-o", case, "--dump-every", "48", "--threads", "4"], check=True, stdout=subprocess.DEVNULL)
    c = read_case(case)
    res = api.run_case(c, device=0)
    assert np.array_equal(res["out"], c["out_ref"], equal_nan=True)
    assert np.array_equal(res["hrurec"], c["hrurec_ref"], equal_nan=True)
    assert np.array_equal(res["status"], c["status_ref"])


def test_thousand_cells_against_reference_build(ref_harness, tmp_path):
    """1,024 cells (5,400 HRUs: every land-cover kind, many warps, re-binning by snow state) x 60 days against the reference
    build: daily aggregates of all 184 variables, the state, balance errors and status bit-identical"""
    cfg = dataclasses.replace(synth.CONFIGS["fe_hourly"], ndays=60)
    r = synth.generate(str(tmp_path / "in"), cfg, 32, 32, 205)
    case = str(tmp_path / "case.bin")
    subprocess.run([ref_harness, "-g", r["global_file"], "-o", case, "--dump-every", "720", "--agg-only", "--threads", str(os.cpu_count() or 1)],
                   check=True, stdout=subprocess.DEVNULL)
    c = read_case(case)
    res = api.run_case(c, device=0, want_out=False)
    assert np.array_equal(res["agg"][1:], c["agg_ref"][1:], equal_nan=True)
    assert np.array_equal(res["hrurec"], c["hrurec_ref"], equal_nan=True)
    assert np.array_equal(res["status"], c["status_ref"])
    assert np.array_equal(res["balance"], c["balance_ref"], equal_nan=True)


def test_glacier_mass_balance_fit(ref_harness, tmp_path):
    """the per-cell quadratic mass-balance curve at the end of an accumulation interval (accumulateGlacierMassBalance,
    GraphingEquation.c:35-126): four glacier HRUs in four bands per cell, 367 days; bit-identical to the reference"""
    c = _reference_case(ref_harness, "glacier_multi", 2, 2, 367, 303, tmp_path)
    res = api.run_case(c, device=0, want_out=False)
    assert np.all(c["gmb_ref"][:, 2] != 0)
    assert np.array_equal(res["gmb"], c["gmb_ref"])
    assert np.array_equal(res["hrurec"], c["hrurec_ref"], equal_nan=True)


@pytest.mark.parametrize("name", GOLDEN)
def test_disagg_golden(name):
    """vicgpu_disagg against the reference's initialize_atmos() output for the same daily inputs; then the model is
    stepped from the device-resident disaggregated forcing and must land on the reference's state"""
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    gp = api.VicGpu(g["options_raw"])
    L = gp.L
    gp.set_veglib(g["veglib"]); gp.set_cells(g["cellpar"], g["hrupar"]); gp.set_output_spec(g["aggtype"]); gp.set_state(g["hrurec0"])
    f = gp.disagg(g["disagg_raw"], g["daily"])
    names = [f"{v}[{s}]" for v in TABLES["forcing"] for s in range(L.f_nslot)]
    worst = column_report(f, g["forcing"], names)[:3]
    assert np.array_equal(f, g["forcing"]), worst
    nrec = min(int(g["dump_recs"][1]) + 1, f.shape[0])
    gp.step(0, nrec, g["dmy"][:nrec + 1])
    k = 1 if nrec == int(g["dump_recs"][1]) + 1 else None
    if k is not None:
        st = gp.get_state()
        assert np.array_equal(st, g["hrurec_ref"][k], equal_nan=True)
    gp.close()


def test_state_roundtrip_and_restart():
    """set_state/get_state are exact, and stepping 48 records at once == 24 + get_state/set_state + 24"""
    g = dict(np.load(os.path.join(GOLDEN_DIR, "fe_hourly_winter.npz")))
    a = api.run_case({**g, "dump_recs": np.array([47], dtype=np.int32)}, nrec=48)
    b = api.run_case({**g, "dump_recs": np.array([23, 47], dtype=np.int32)}, nrec=48)
    assert np.array_equal(a["hrurec"][-1], b["hrurec"][-1], equal_nan=True)
    assert np.array_equal(a["out"], b["out"], equal_nan=True)
    gp = api.VicGpu(g["options_raw"])
    gp.set_veglib(g["veglib"]); gp.set_cells(g["cellpar"], g["hrupar"]); gp.set_state(g["hrurec0"])
    assert np.array_equal(gp.get_state(), g["hrurec0"], equal_nan=True)
    gp.close()


def test_float32_outputs_are_the_narrowed_doubles():
    """vicgpu_step_f32 (float32 narrowing on the device, double-buffered device-to-host staging): the float32 rows must
    be exactly float32(double rows) for per-record data and daily aggregates (WriteOutputNetCDF.c:279 applies the same conversion)"""
    g = dict(np.load(os.path.join(GOLDEN_DIR, "fe_hourly_winter.npz")))
    nrec = 72
    res = {}
    for dt in (np.float64, np.float32):
        gp = api.VicGpu(g["options_raw"])
        gp.set_veglib(g["veglib"]); gp.set_cells(g["cellpar"], g["hrupar"]); gp.set_output_spec(g["aggtype"]); gp.set_state(g["hrurec0"])
        gp.set_forcing(0, g["forcing"][:nrec])
        out = np.zeros((nrec, gp.ncell, gp.L.nout), dtype=dt)
        agg = np.zeros((nrec // 24, gp.ncell, gp.L.nout), dtype=dt)
        gp.step(0, nrec, g["dmy"][:nrec + 1], out, agg)
        gp.close()
        res[dt] = (out, agg)
    with np.errstate(over="ignore"):
        assert np.array_equal(res[np.float32][0], res[np.float64][0].astype(np.float32), equal_nan=True)
        assert np.array_equal(res[np.float32][1], res[np.float64][1].astype(np.float32), equal_nan=True)
    assert np.array_equal(res[np.float64][0][:24], g["out_ref_head"], equal_nan=True)


def test_forcing_windows_overlap_upload_and_step():
    """two device-resident forcing windows: uploading block b + 1 before stepping over block b gives the same bits as one window
    holding everything"""
    g = dict(np.load(os.path.join(GOLDEN_DIR, "fe_hourly_winter.npz")))
    a = api.run_case(g, nrec=96)
    gp = api.VicGpu(g["options_raw"])
    gp.set_veglib(g["veglib"]); gp.set_cells(g["cellpar"], g["hrupar"]); gp.set_output_spec(g["aggtype"]); gp.set_state(g["hrurec0"])
    out = np.zeros((96, gp.ncell, gp.L.nout))
    gp.set_forcing(0, g["forcing"][0:24])
    for b in range(4):
        if b < 3:
            gp.set_forcing((b + 1) * 24, g["forcing"][(b + 1) * 24:(b + 2) * 24])
        gp.step(b * 24, 24, g["dmy"][b * 24:b * 24 + 25], out[b * 24:(b + 1) * 24], None)
    with pytest.raises(api.VicGpuError):
        gp.step(0, 24, g["dmy"][:25])  # block 0 has been replaced by block 2
    gp.close()
    assert np.array_equal(out, a["out"], equal_nan=True)


def test_call_order_errors():
    g = dict(np.load(os.path.join(GOLDEN_DIR, "fe_hourly_winter.npz")))
    gp = api.VicGpu(g["options_raw"])
    with pytest.raises(api.VicGpuError):
        gp._chk(gp.lib.vicgpu_set_state(gp.h, g["hrurec0"].ctypes.data_as(api.C.POINTER(api.C.c_double))))  # before set_cells
    gp.set_veglib(g["veglib"]); gp.set_cells(g["cellpar"], g["hrupar"]); gp.set_state(g["hrurec0"])
    with pytest.raises(api.VicGpuError):
        gp.step(0, 4, g["dmy"][:5])  # no forcing resident
    gp.close()


@pytest.mark.parametrize("base,ncell,nrec", [("fe_hourly", 10000, 48), ("fe_hourly", 100000, 24), ("glacier", 10000, 48), ("frozen_bands", 10000, 24)])
def test_full_size_domain_properties(base, ncell, nrec):
    """BASELINE-size domains (configs[1]: 10,000 cells; configs[2]-size: 100,000), through properties that do not need the reference
    to run at that size: (1) cells are independent, so the first 256 cells of the big domain -- the unperturbed base domain the
    reference itself is timed on -- must give bit-identical outputs and state whether they are advanced alone or inside the big domain
    (different row binning, different block placement, different re-sort decisions); (2) the reference's own closure check holds in every
    cell: |water balance error| of a step < 1e-5 mm (calc_water_energy_balance_errors.c:33-43); (3) no cell is flagged invalid."""
    import bench
    dom = bench.build_domain(ncell, 1, base)
    nb = 256
    h_nb = int(np.searchsorted(dom["hrupar"][:, TABLES["hpar"].index("HP_cell")], nb))

    def run(nc, nh):
        g = api.VicGpu(dom["options_raw"])
        g.set_veglib(dom["veglib"]); g.set_output_spec(dom["aggtype"])
        g.set_cells(dom["cellpar"][:nc], dom["hrupar"][:nh]); g.set_state(dom["hrurec0"][:nh])
        g.set_forcing(0, np.ascontiguousarray(fbig[:, :nc]))
        out = np.zeros((nrec, nc, g.L.nout))
        g.step(0, nrec, bench.make_dmy(nrec), out, None)
        st, status, bal = g.get_state(), g.cell_status(), g.balance_errors()
        g.close()
        return out, st, status, bal

    L = layout_from_options(parse_options(dom["options_raw"]))
    fbig = np.empty((nrec, ncell, L.f_stride))
    for d in range(nrec // 24):
        bench.forcing_day(dom, d, 1, fbig[d * 24:(d + 1) * 24])
    out_big, st_big, status_big, bal_big = run(ncell, dom["hrupar"].shape[0])
    out_small, st_small, status_small, _ = run(nb, h_nb)
    assert np.array_equal(out_big[:, :nb], out_small, equal_nan=True)
    assert np.array_equal(st_big[:h_nb], st_small, equal_nan=True)
    assert not status_big.any() and not status_small.any()
    werr = out_big[:, :, L.out_names.index("WATER_ERROR")]
    assert np.nanmax(np.abs(werr)) < 1e-5, np.nanmax(np.abs(werr))
    assert np.all(np.abs(bal_big[:, 2]) < 1e-5 + 1e-12)  # CellBalanceErrors::water_max_error


@pytest.mark.parametrize("env", [{"VICGPU_NOOVERLAP": "1"}, {"VICGPU_NOBIN": "1"}, {"VICGPU_REBIN": "0"}, {"VICGPU_REBIN": "1"}, {"VICGPU_SYNC": "50000"},
                                 {"VICGPU_SYNC": "0"}, {"VICGPU_BLOCK": "128"}, {"VICGPU_BLOCK": "512"}, {"VICGPU_OUTBLOCK": "128"}, {"VICGPU_EVEN": "1"}, {"VICGPU_BINFINE": "1"}, {"VICGPU_PDLWAIT": "0"}, {"VICGPU_BINCOST": "1"}, {"VICGPU_BINCOST": "2", "VICGPU_REBIN": "1"},
                                 {"VICGPU_BALANCE": "0"}, {"VICGPU_BALANCE": "100", "VICGPU_REBIN": "2"}, {"VICGPU_SYNCMASK": "4"}],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_every_launch_mode_gives_the_same_bits(env, monkeypatch):
    """the tuning / A-B knobs of libvicgpu.so (read from the environment at vicgpu_create) change how the work is laid out and launched --
    row order, records per launch, streams, block sizes -- never the arithmetic: every mode must reproduce the reference bit for bit"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for name in ("fe_hourly_winter", "glacier"):
        g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
        res = api.run_case(g, device=0, nrec=96)
        n = res["hrurec"].shape[0]
        assert n >= 1
        assert np.array_equal(res["hrurec"], g["hrurec_ref"][:n], equal_nan=True)
        assert np.array_equal(res["out"][:24], g["out_ref_head"], equal_nan=True)
        assert np.array_equal(res["agg"][1:], g["agg_ref"][1:res["agg"].shape[0]], equal_nan=True)


def test_cells_invalidated_like_the_reference(ref_harness, tmp_path):
    """error behaviour: without TFALLBACK a failed solve invalidates the cell (vicNl.c:545-559).  Nine frozen-soil cells drop out at
    nine different records of the first day while the others carry on: same cells invalid, every row up to a cell's failing record
    bit-identical, the row frozen from there on -- in every launch mode that treats records differently"""
    from test_cpu import _failing_case, check_until_invalid
    _, c = _failing_case(ref_harness, tmp_path)
    for env in ({}, {"VICGPU_NOOVERLAP": "1"}, {"VICGPU_REBIN": "1"}):
        for k, v in env.items():
            os.environ[k] = v
        try:
            res = api.run_case(c, device=0)
        finally:
            for k in env:
                del os.environ[k]
        assert np.array_equal(res["status"], c["status_ref"]) and (c["status_ref"] != 0).all(), env
        fails = check_until_invalid(res["out"], c)
        assert len(set(fails)) >= 4, fails
