"""The drop-in itself: oracle/_ref/vicNl is the reference's own executable (its main, its readers, its
runModel), oracle/_ref/vicNl_gpu the same objects with runModel() replaced by vic_b200/host/vicNl_gpu.cpp + libvicgpu.so.  Both are
run on the same global parameter file (the reference's own input formats, written by vic_b200/synth.py); the output stream
(OutputData::aggdata of every cell and variable at every output step, as raw doubles) and the state file written by the reference's
write_model_state() (write_model_state.c:107-371) -- for the drop-in: from structs refilled through vicgpu_get_state +
vicgpu_unpack_hrurec -- must be identical byte for byte."""
import dataclasses
import os
import subprocess

import numpy as np
import pytest

from vic_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _run(exe, cfgname, tmp_path, tag, ndays, stateday, seed, extra=(), stream=True):
    p = os.path.join(REF, exe)
    if not os.path.exists(p):
        pytest.skip(f"{p} not built (oracle/Makefile)")
    res = tmp_path / f"res_{tag}"
    res.mkdir()
    cfg = dataclasses.replace(synth.CONFIGS[cfgname], ndays=ndays, out_step=24,
                              extra_global=["STATENAME state", "STATEYEAR 2001", "STATEMONTH 1", f"STATEDAY {stateday}", "STATE_FORMAT ASCII", *extra])
    r = synth.generate(str(tmp_path / f"in_{tag}"), cfg, 2, 3, seed, result_dir=str(res))
    # get_global_param.c:1142 builds the state file name with an overlapping sprintf: the file lands in the working directory as
    # "_<yyyy>-<mm>-<dd>" whatever STATENAME says
    o = subprocess.run([p, "-g", r["global_file"]], cwd=str(res), capture_output=True, text=True)
    assert o.returncode == 0, o.stderr[-2000:]
    out = np.fromfile(res / "results.nc.f64", dtype=np.float64) if stream else None
    state = (res / f"_2001-01-{stateday:02d}").read_bytes()
    return out, state, o.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("cfgname,ndays,stateday,seed", [("fe_hourly", 5, 3, 501), ("glacier", 4, 4, 502), ("frozen_bands", 2, 2, 503), ("wb_daily", 40, 31, 504)])
def test_vicNl_gpu_matches_stock_vicNl(cfgname, ndays, stateday, seed, tmp_path):
    out_cpu, state_cpu, _ = _run("vicNl", cfgname, tmp_path, "cpu", ndays, stateday, seed)
    out_gpu, state_gpu, err = _run("vicNl_gpu", cfgname, tmp_path, "gpu", ndays, stateday, seed)
    assert "Model execution time (GPU)" in err
    assert out_cpu.size == out_gpu.size and out_cpu.size > 0
    nsteps = ndays if cfgname != "wb_daily" else ndays  # one output step per day
    per_step = out_cpu.size // nsteps
    assert per_step * nsteps == out_cpu.size
    # the reference never initialises aggdata before its first output step (output_list_utils.c:20-24): compare from the second on
    a, b = out_cpu.reshape(nsteps, per_step)[1:], out_gpu.reshape(nsteps, per_step)[1:]
    assert np.array_equal(a, b, equal_nan=True), np.argwhere(~((a == b) | ((a != a) & (b != b))))[:5]
    assert len(state_cpu) > 1000 and state_cpu == state_gpu


def test_stock_vicNl_runs_and_writes_state(tmp_path):
    """CPU-only part: the reference's own executable builds from its sources, runs the synthetic case and writes output + state"""
    out, state, err = _run("vicNl", "fe_hourly", tmp_path, "cpu", 2, 2, 505)
    assert "VIC model run done" in err
    assert out.size > 0 and len(state) > 1000


# ---- the drop-in's loader (vic_b200/host/vicgpu_fastread.h) ------------------------------------------------
def _records(path, lines_after_header):
    """split a per-cell parameter file into its records: header line + lines_after_header(header) lines"""
    lines = open(path).read().splitlines(keepends=True)
    recs, i = [], 0
    while i < len(lines):
        if not lines[i].strip():
            i += 1
            continue
        n = 1 + lines_after_header(lines[i])
        recs.append(lines[i:i + n])
        i += n
    return recs


def _shuffle_parameter_files(indir, seed, with_decoys):
    """vegetation and snow-band records in another order than the soil file's cells; optionally a second, different record of an
    existing cell number further down (the reference takes the FIRST match, read_vegparam.c:117) and a snow-band file that lacks one
    cell (read_snowband.c:49-54: warning, one band)"""
    rng = np.random.default_rng(seed)
    glob = open(os.path.join(indir, "global.txt")).read()
    per_tile = 2 if "VEGPARAM_LAI TRUE" in glob.replace("\t", " ") else 1
    vp = os.path.join(indir, "vegparam.txt")
    recs = _records(vp, lambda h: int(h.split()[1]) * per_tile)
    order = rng.permutation(len(recs))
    out = [recs[k] for k in order]
    if with_decoys:
        a, b = recs[0], recs[1]
        out.append([a[0].split()[0] + " " + " ".join(b[0].split()[1:]) + "\n"] + b[1:])  # cell of record 0 with record 1's tiles
    open(vp, "w").write("".join("".join(r) for r in out))
    sb = os.path.join(indir, "snowband.txt")
    if os.path.exists(sb):
        recs = _records(sb, lambda h: 0)
        order = rng.permutation(len(recs))
        out = [recs[k] for k in order]
        if with_decoys:
            out = out[:-1] + [[out[0][0].split()[0] + " " + " ".join(out[1][0].split()[1:]) + "\n"]]
        open(sb, "w").write("".join("".join(r) for r in out))


@pytest.mark.parametrize("cfgname,decoys", [("frozen_bands", False), ("frozen_bands", True), ("fe_blowing", False), ("glacier", True)])
def test_indexed_parameter_lookups_match_the_reference_scans(cfgname, decoys, tmp_path):
    """read_vegparam / read_snowband behind the one-pass index and the O(N log N) initGrid against the reference's own scans
    (oracle/_ref/readercheck): HRU lists, band tables and grid geometry of every cell bit-identical, records in shuffled order"""
    exe = os.path.join(REF, "readercheck")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (oracle/Makefile)")
    cfg = dataclasses.replace(synth.CONFIGS[cfgname], ndays=1)
    r = synth.generate(str(tmp_path / "in"), cfg, 7, 9, 808)
    _shuffle_parameter_files(r["dir"], 5, decoys)
    o = subprocess.run([exe, "-g", r["global_file"]], capture_output=True, text=True)
    assert o.returncode == 0 and "identical" in o.stdout, o.stdout[-2000:]
    assert "ncell 63 " in o.stdout and "compared 63 " in o.stdout


def test_reference_physics_behind_the_indexed_loader(tmp_path):
    """the reference's executable with the drop-in's loader linked in (oracle/_ref/vicNl_fastread: read_vegparam, read_snowband and
    initGrid replaced, everything else -- its runModel on the CPU included -- untouched) against the stock vicNl on shuffled
    parameter files: output stream and state file identical byte for byte; VICGPU_STOCK_READERS=1 switches the scans back"""
    def run(exe, tag, env=None):
        p = os.path.join(REF, exe)
        if not os.path.exists(p):
            pytest.skip(f"{p} not built (oracle/Makefile)")
        res = tmp_path / f"res_{tag}"
        res.mkdir()
        cfg = dataclasses.replace(synth.CONFIGS["frozen_bands"], ndays=2, out_step=24,
                                  extra_global=["STATENAME state", "STATEYEAR 2001", "STATEMONTH 1", "STATEDAY 2", "STATE_FORMAT ASCII"])
        r = synth.generate(str(tmp_path / f"in_{tag}"), cfg, 2, 3, 606, result_dir=str(res))
        _shuffle_parameter_files(r["dir"], 6, False)
        o = subprocess.run([p, "-g", r["global_file"]], cwd=str(res), capture_output=True, text=True, env=dict(os.environ, **(env or {})))
        assert o.returncode == 0, o.stderr[-2000:]
        return np.fromfile(res / "results.nc.f64", dtype=np.float64), (res / "_2001-01-02").read_bytes()

    out0, st0 = run("vicNl", "stock")
    out1, st1 = run("vicNl_fastread", "fast")
    out2, st2 = run("vicNl_fastread", "fast_stockscan", {"VICGPU_STOCK_READERS": "1"})
    half = out0.size // 2  # (first output step: uninitialised aggdata in the reference, see above)
    assert out0.size > 0 and np.array_equal(out0[half:], out1[half:], equal_nan=True) and np.array_equal(out0[half:], out2[half:], equal_nan=True)
    assert len(st0) > 1000 and st0 == st1 == st2


def test_indexed_lookup_with_lai_lines_in_the_vegetation_file(tmp_path):
    """VEGPARAM_LAI TRUE: every tile carries a second line of twelve monthly LAI values (read_vegparam.c:104, 200-228), header lines with
    trailing blanks, blank lines at the end of the file, shuffled records and decoys: identical to the reference's scan"""
    exe = os.path.join(REF, "readercheck")
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (oracle/Makefile)")
    cfg = dataclasses.replace(synth.CONFIGS["frozen_bands"], ndays=1, extra_global=["VEGPARAM_LAI TRUE", "LAI_SRC LAI_FROM_VEGPARAM"])
    r = synth.generate(str(tmp_path / "in"), cfg, 6, 7, 5)
    vp = os.path.join(r["dir"], "vegparam.txt")
    lines, out, i, rng = open(vp).read().splitlines(), [], 0, np.random.default_rng(1)
    while i < len(lines):
        n = int(lines[i].split()[1])
        out.append(lines[i] + "   ")
        for k in range(n):
            out += [lines[i + 1 + k], "  " + " ".join(f"{x:.2f}" for x in rng.uniform(0.5, 5, 12))]
        i += n + 1
    open(vp, "w").write("\n".join(out) + "\n\n\n")
    _shuffle_parameter_files(r["dir"], 3, True)
    o = subprocess.run([exe, "-g", r["global_file"]], capture_output=True, text=True)
    assert o.returncode == 0 and "identical" in o.stdout and "ncell 42 " in o.stdout, o.stdout[-2000:]


# ---- OUTPUT_FORCE TRUE: the meteorological disaggregator (BASELINE configs[4], first half; vicNl.c:445-490) ---------------------------
def _run_output_force(exe, tmp_path, tag, ndays, seed, extra=()):
    p = os.path.join(REF, exe)
    if not os.path.exists(p):
        pytest.skip(f"{p} not built (oracle/Makefile)")
    res = tmp_path / f"res_{tag}"
    res.mkdir()
    cfg = dataclasses.replace(synth.CONFIGS["disagg"], ndays=ndays, extra_global=list(extra))
    r = synth.generate(str(tmp_path / f"in_{tag}"), cfg, 2, 3, seed, result_dir=str(res))
    o = subprocess.run([p, "-g", r["global_file"]], cwd=str(res), capture_output=True, text=True)
    assert o.returncode == 0, o.stderr[-2000:]
    return np.fromfile(res / "results.nc.force.f64", dtype=np.float64), o.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("ndays,seed,extra", [(25, 611, ()), (12, 612, ("DISAGG_WRITE_CHUNK_SIZE 7", "ALMA_OUTPUT TRUE", "LW_TYPE LW_PRATA"))])
def test_vicNl_gpu_disaggregator_mode_matches_stock_vicNl(ndays, seed, extra, tmp_path):
    """OUTPUT_FORCE TRUE: the stock vicNl disaggregates every cell's daily forcing on the CPU (initialize_atmos) and writes it through
    write_forcing_file / write_data_one_cell; vicNl_gpu reads the same files with the reference's reader, disaggregates on the device
    (vicgpu_disagg) and writes through the same two routines: the written stream (every variable of every record of every cell) is
    identical byte for byte"""
    cpu, _ = _run_output_force("vicNl", tmp_path, "cpu", ndays, seed, extra)
    gpu, err = _run_output_force("vicNl_gpu", tmp_path, "gpu", ndays, seed, extra)
    assert "Execution time (GPU)" in err
    assert cpu.size == 6 * ndays * 24 * 204 and cpu.size == gpu.size
    assert np.array_equal(cpu, gpu, equal_nan=True), np.argwhere(~((cpu == gpu) | ((cpu != cpu) & (gpu != gpu))))[:5]


def _out_names():
    """the 204 output columns (184 variables, the multi-element ones expanded) of a three-layer, one-band configuration"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "fe_hourly_winter.npz"))
    from vic_b200.layout import layout_from_options, parse_options
    return layout_from_options(parse_options(g["options_raw"])).out_names


def test_stock_vicNl_disaggregator_mode_writes_the_forcing_stream(tmp_path):
    """CPU-only part: the reference's executable in OUTPUT_FORCE mode writes one record of 204 values per cell and hour, and the values
    are its hourly forcing (air temperature within the day's TMIN..TMAX, precipitation summing to the daily input)"""
    out, err = _run_output_force("vicNl", tmp_path, "cpu", 10, 613)
    assert "disaggregated forcings generation done" in err
    rows = out.reshape(6, 240, 204)
    names = list(_out_names())
    prec, tair = rows[:, :, names.index("PREC")], rows[:, :, names.index("AIR_TEMP")]
    assert np.all(prec >= 0) and prec.sum() > 0 and np.all(np.abs(tair) < 60)


@pytest.mark.gpu
@pytest.mark.parametrize("cfgname,ndays,seed", [("fe_hourly", 4, 521), ("frozen_bands", 2, 522)])
def test_vicNl_gpu_netcdf_output_holds_the_stock_outputs_narrowed(cfgname, ndays, seed, tmp_path):
    """VICGPU_NC_OUTPUT: vicNl_gpu writes its daily aggregates through the library's NetCDF writer (float32 rows of vicgpu_step_f32, one
    record per output step, metadata from the reference's own output_mapping and grid).  Every variable of the file, at every modelled
    cell and step, is the stock vicNl's aggregate narrowed to float32 (what WriteOutputNetCDF.c:412 stores); coordinates, time axis and
    the unused depth planes are as the reference's writer lays them out."""
    from scipy.io import netcdf_file
    out_cpu, _, _ = _run("vicNl", cfgname, tmp_path, "cpu", ndays, ndays, seed)
    nc_path = str(tmp_path / "out.nc")
    os.environ["VICGPU_NC_OUTPUT"] = nc_path
    try:
        _, _, err = _run("vicNl_gpu", cfgname, tmp_path, "gpu", ndays, ndays, seed, stream=False)  # (the stub's stream is not written in this mode)
    finally:
        del os.environ["VICGPU_NC_OUTPUT"]
    assert "Model execution time (GPU)" in err
    ncell = 6
    per_cell = out_cpu.size // (ndays * ncell)
    ref = out_cpu.reshape(ndays, ncell, per_cell)
    # the stream's columns: the 184 variables in enum order, multi-element ones expanded (layout of the case's options)
    import dataclasses as dc
    from vic_b200.casefile import read_case
    cfg = dc.replace(synth.CONFIGS[cfgname], ndays=1)
    r = synth.generate(str(tmp_path / "in_layout"), cfg, 2, 3, seed)
    case = str(tmp_path / "layout.bin")
    subprocess.run([os.path.join(REF, "vic_ref_harness"), "-g", r["global_file"], "-o", case, "--no-run"], check=True, stdout=subprocess.DEVNULL)
    from vic_b200.layout import layout_from_options, parse_options
    names = list(layout_from_options(parse_options(read_case(case)["options_raw"])).out_names)
    assert len(names) == per_cell
    f = netcdf_file(nc_path, "r", mmap=False)
    assert f.dimensions["time"] is None and f.dimensions["lat"] == 2 and f.dimensions["lon"] == 3 and f.dimensions["depth"] == 30
    assert np.array_equal(f.variables["time"][:], np.arange(ndays, dtype=np.float32)) and f.variables["time"].units == b"days since 2001-1-1"
    checked = 0
    for name, var in f.variables.items():
        if name in ("lat", "lon", "time", "depth"):
            continue
        vic = var.internal_vic_name.decode().replace("OUT_", "", 1)
        cols = [k for k, n in enumerate(names) if n == vic or n.startswith(vic + "[")]
        assert cols, vic
        data = var[:].reshape(ndays, -1, ncell)  # (time, depth or 1, lat * lon): the six cells fill the 2 x 3 grid in order
        with np.errstate(over="ignore"):  # (unset band slots hold huge doubles: both narrowings give inf)
            want = ref[:, :, cols].astype(np.float32).transpose(0, 2, 1)
        assert np.array_equal(data[1:, :len(cols)], want[1:], equal_nan=True), vic  # (first step: uninitialised aggdata in the reference)
        if data.shape[1] > len(cols):
            assert np.all(data[:, len(cols):] == np.float32(1e20))
        checked += 1
    f.close()
    assert checked >= 20


def test_indexed_loader_fails_like_the_reference_when_a_cell_is_missing(tmp_path):
    """a cell of the soil file without a record in the vegetation file: the indexed lookup falls back to the reference's own scan, which
    ends the run with its own message and exit code (read_vegparam.c:130-133)"""
    outs = {}
    for exe in ("vicNl", "vicNl_fastread"):
        p = os.path.join(REF, exe)
        if not os.path.exists(p):
            pytest.skip(f"{p} not built (oracle/Makefile)")
        res = tmp_path / f"res_{exe}"
        res.mkdir()
        cfg = dataclasses.replace(synth.CONFIGS["fe_hourly"], ndays=1)
        r = synth.generate(str(tmp_path / f"in_{exe}"), cfg, 2, 2, 77, result_dir=str(res))
        vp = os.path.join(r["dir"], "vegparam.txt")
        recs = _records(vp, lambda h: int(h.split()[1]))
        open(vp, "w").write("".join("".join(rec) for rec in recs[:2] + recs[3:]))  # cell 3 has no record
        o = subprocess.run([p, "-g", r["global_file"]], cwd=str(res), capture_output=True, text=True)
        outs[exe] = (o.returncode, [ln for ln in o.stderr.splitlines() if "not found" in ln])
    assert outs["vicNl"] == outs["vicNl_fastread"] and outs["vicNl"][0] == 99 and outs["vicNl"][1], outs
