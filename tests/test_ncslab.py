"""NetCDF forcing ingestion: vicgpu_nc_* (include/vicgpu.h; vic_b200/host/vicgpu_ncslab.h) reads a (time, lat, lon)
forcing file as time-major slabs [time][variable][cell].  The reference's NetCDF branch (read_atmos_data.c:109-338) needs libnetcdf,
which this image does not have, so it cannot be run as the oracle here; the expected values below are its documented conversions
applied with numpy to the arrays the files were written from (files written by scipy.io.netcdf_file, an independent implementation
of the classic format)."""
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

from vic_b200 import api

VARS = ["pr", "tasmax", "tasmin", "wind"]


def _write(path, version, record_time, nt=7, nlat=3, nlon=5, seed=0, extra_int=False, dup_lat=False):
    rng = np.random.default_rng(seed)
    lat = (48.03125 + 0.0625 * np.arange(nlat)).astype(np.float64)
    if dup_lat:
        lat[-1] = lat[0]
    lon = (-121.96875 + 0.0625 * np.arange(nlon)).astype(np.float32)  # a float coordinate variable
    data = {
        "pr": rng.integers(0, 4000, (nt, nlat, nlon)).astype(np.int16),
        "tasmax": rng.integers(-3000, 3500, (nt, nlat, nlon)).astype(np.int16),
        "tasmin": rng.normal(0, 10, (nt, nlat, nlon)).astype(np.float32),
        "wind": rng.gamma(2.0, 2.0, (nt, nlat, nlon)).astype(np.float64),
    }
    f = netcdf_file(path, "w", version=version)
    f.createDimension("time", None if record_time else nt)
    f.createDimension("lat", nlat)
    f.createDimension("lon", nlon)
    v = f.createVariable("time", "d", ("time",))
    v[:] = np.arange(nt, dtype=np.float64)
    v = f.createVariable("lat", "d", ("lat",))
    v[:] = lat
    v = f.createVariable("lon", "f", ("lon",))
    v[:] = lon
    v = f.createVariable("pr", "h", ("time", "lat", "lon"))
    v[:] = data["pr"]
    v.scale_factor = np.float32(0.025)
    v.units = "mm"
    v = f.createVariable("tasmax", "h", ("time", "lat", "lon"))
    v[:] = data["tasmax"]
    v.inverse_scale_factor = np.float64(100.0)
    v.scale_factor = np.float32(7.0)
    v = f.createVariable("tasmin", "f", ("time", "lat", "lon"))
    v[:] = data["tasmin"]
    v = f.createVariable("wind", "d", ("time", "lat", "lon"))
    v[:] = data["wind"]
    if extra_int:
        v = f.createVariable("count", "i", ("time", "lat", "lon"))
        v[:] = np.zeros((nt, nlat, nlon), dtype=np.int32)
    f.close()
    return lat, lon.astype(np.float64), data


def _expected(data, t0, nt, ii, jj):
    """read_atmos_data.c:226-310, value for value"""
    sl = (slice(t0, t0 + nt), ii, jj)
    return np.stack([
        data["pr"][sl].astype(np.float64) * np.float64(np.float32(0.025)),     # scale_factor, as float
        data["tasmax"][sl].astype(np.float64) / np.float64(np.float32(100.0)),  # inverse_scale_factor wins when both are present
        data["tasmin"][sl].astype(np.float64),
        data["wind"][sl],
    ], axis=1)


@pytest.mark.parametrize("version", [1, 2])
@pytest.mark.parametrize("record_time", [True, False])
def test_time_major_slab_matches_the_reference_conversions(version, record_time, tmp_path):
    path = str(tmp_path / "forcing.nc")
    lat, lon, data = _write(path, version, record_time, seed=version)
    # modelled cells: a subset of the grid in another order than the file's
    ii = np.array([2, 0, 1, 2, 0, 1, 1])
    jj = np.array([4, 0, 3, 1, 2, 2, 0])
    with api.NcForcing(path) as nc:
        assert (nc.ntime, nc.nlat, nc.nlon) == (7, 3, 5)
        got = nc.read_slab(VARS, 0, 7, lat[ii], lon[jj])
        assert got.shape == (7, 4, 7) and np.array_equal(got, _expected(data, 0, 7, ii, jj))
        # a window that starts at the reference's skip_recs, variables in another order
        got = nc.read_slab(VARS[::-1], 2, 3, lat[ii], lon[jj])
        assert np.array_equal(got, _expected(data, 2, 3, ii, jj)[:, ::-1])


def test_first_exact_match_of_a_repeated_coordinate(tmp_path):
    path = str(tmp_path / "forcing.nc")
    lat, lon, data = _write(path, 1, True, dup_lat=True)
    with api.NcForcing(path) as nc:
        got = nc.read_slab(["wind"], 0, 7, lat[[2]], lon[[1]])  # lat[2] == lat[0]: the reference stops at index 0 (:176-181)
    assert np.array_equal(got[:, 0, 0], data["wind"][:, 0, 1])


def test_errors_are_the_references(tmp_path):
    path = str(tmp_path / "forcing.nc")
    lat, lon, _ = _write(path, 2, True, extra_int=True)
    with api.NcForcing(path) as nc:
        with pytest.raises(api.VicGpuError, match="no exactly matching grid point"):
            nc.read_slab(VARS, 0, 7, np.array([lat[0] + 1e-9]), lon[:1])
        with pytest.raises(api.VicGpuError, match="type not supported"):
            nc.read_slab(["count"], 0, 7, lat[:1], lon[:1])
        with pytest.raises(api.VicGpuError, match="the file has 7"):
            nc.read_slab(VARS, 5, 3, lat[:1], lon[:1])
        with pytest.raises(api.VicGpuError, match="no variable 'snowfall'"):
            nc.read_slab(["snowfall"], 0, 1, lat[:1], lon[:1])
        with pytest.raises(api.VicGpuError, match="not .time, lat, lon."):
            nc.read_slab(["lat"], 0, 1, lat[:1], lon[:1])
    hdf = str(tmp_path / "nc4.nc")
    open(hdf, "wb").write(b"\x89HDF\r\n\x1a\n" + bytes(64))
    with pytest.raises(api.VicGpuError) as e:
        api.NcForcing(hdf)
    assert e.value.code == -3 and "NetCDF-4" in str(e.value)
    with pytest.raises(api.VicGpuError):
        api.NcForcing(str(tmp_path / "missing.nc"))


def test_large_grid_in_file_order(tmp_path):
    """every cell of an 89 x 121 grid over 40 days, record variables with padded (odd-sized short) slabs"""
    path = str(tmp_path / "forcing.nc")
    lat, lon, data = _write(path, 2, True, nt=40, nlat=89, nlon=121, seed=3)
    ii, jj = [a.ravel() for a in np.meshgrid(np.arange(89), np.arange(121), indexing="ij")]
    with api.NcForcing(path) as nc:
        got = nc.read_slab(VARS, 0, 40, lat[ii], lon[jj])
    assert np.array_equal(got, _expected(data, 0, 40, ii, jj))


@pytest.mark.gpu
def test_disagg_from_a_netcdf_file_gives_the_reference_forcing(tmp_path):
    """daily PREC / TMAX / TMIN / WIND of the golden case written to a NetCDF file, read back as a time-major slab and handed to
    vicgpu_disagg_tm: the hourly forcing is the reference's initialize_atmos() output bit for bit (the cell-major vicgpu_disagg too)"""
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fe_hourly_winter.npz")))
    daily = g["daily"]  # [ncell][Ndays][4]
    ncell, ndays, _ = daily.shape
    path = str(tmp_path / "daily.nc")
    lat = 40.0 + 0.5 * np.arange(ncell)
    f = netcdf_file(path, "w", version=2)
    f.createDimension("time", None)
    f.createDimension("lat", ncell)
    f.createDimension("lon", 1)
    v = f.createVariable("time", "d", ("time",))
    v[:] = np.arange(ndays, dtype=np.float64)
    v = f.createVariable("lat", "d", ("lat",))
    v[:] = lat
    v = f.createVariable("lon", "d", ("lon",))
    v[:] = np.array([-120.0])
    for k, name in enumerate(VARS):
        v = f.createVariable(name, "d", ("time", "lat", "lon"))
        v[:] = daily[:, :, k].T[:, :, None]
    f.close()
    with api.NcForcing(path) as nc:
        slab = nc.read_slab(VARS, 0, ndays, lat, np.full(ncell, -120.0))
    assert np.array_equal(slab, daily.transpose(1, 2, 0))
    gp = api.VicGpu(g["options_raw"])
    gp.set_veglib(g["veglib"])
    gp.set_cells(g["cellpar"], g["hrupar"])
    gp.set_output_spec(g["aggtype"])
    gp.set_state(g["hrurec0"])
    f_tm = gp.disagg_tm(g["disagg_raw"], slab)
    f_cm = gp.disagg(g["disagg_raw"], daily)
    gp.close()
    assert np.array_equal(f_tm, g["forcing"]) and np.array_equal(f_cm, g["forcing"])


# ---- the output side: vicgpu_ncout_* (WriteOutputNetCDF.c:163-299, 386-452) --------------------------------------------------------
def test_output_file_holds_the_reference_writers_content(tmp_path):
    """float32 rows of a few output steps -> one record per step; read back with scipy (an independent reader): dimensions, coordinate
    values and attributes, per-variable attributes and fill value, every modelled cell's values at its (lat, lon) position, the fill
    value at the unmodelled grid points and in the unused depth planes, as the reference's writer lays them out"""
    rng = np.random.default_rng(4)
    nlat, nlon, ncell, nstep = 5, 7, 23, 4
    pos = rng.permutation(nlat * nlon)[:ncell]
    pos.sort()  # cells in grid order, as the soil file must list them (vicNl.c readSoilData)
    lat_i, lon_i = pos // nlon, pos % nlon
    variables = [
        dict(name="RUNOFF", nelem=1, long_name="surface runoff", units="mm", standard_name="runoff_amount", cell_methods="time: sum", internal_vic_name="OUT_RUNOFF", category="fluxes"),
        dict(name="SOIL_MOIST", nelem=3, long_name="soil moisture", units="mm", standard_name="soil_moisture", cell_methods="time: point", internal_vic_name="OUT_SOIL_MOIST", category="fluxes"),
        dict(name="SWE", nelem=1, long_name="snow water equivalent", units="mm", standard_name="swe", cell_methods="time: point", internal_vic_name="OUT_SWE", category="snow"),
    ]
    columns = [4, 9, 0]  # where each variable starts in a row of 16 values
    rows = rng.normal(size=(nstep, ncell, 16)).astype(np.float32)
    path = str(tmp_path / "results.nc")
    with api.NcOutput(path, nlat, 48.03125, 0.0625, nlon, -121.96875, 0.0625, "hours since 2001-1-1 0:00", 3, variables, columns, lat_i, lon_i, depth=30,
                      global_text=[("title", "VIC model run output."), ("Conventions", "CF-1.6")], global_int=[("model_start_year", 2001), ("model_end_day", 31)]) as w:
        for s in range(nstep):
            w.write_step(rows[s])
    f = netcdf_file(path, "r", mmap=False)
    assert f.version_byte == 2
    assert {k: (v if v is not None else None) for k, v in f.dimensions.items()} == {"lat": nlat, "lon": nlon, "bnds": 2, "time": None, "depth": 30}
    assert f.title == b"VIC model run output." and f.Conventions == b"CF-1.6" and f.model_start_year == 2001 and f.model_end_day == 31
    assert np.array_equal(f.variables["lat"][:], 48.03125 + np.arange(nlat) * 0.0625) and f.variables["lat"].units == b"degrees_north"
    assert np.array_equal(f.variables["lon"][:], -121.96875 + np.arange(nlon) * 0.0625) and f.variables["lon"].axis == b"X"
    assert np.array_equal(f.variables["depth"][:], np.arange(30, dtype=np.float32))
    t = f.variables["time"]
    assert np.array_equal(t[:], np.arange(nstep, dtype=np.float32) * 3) and t.units == b"hours since 2001-1-1 0:00" and t.calendar == b"gregorian"
    fill = np.float32(1e20)
    for v, c0 in zip(variables, columns):
        var = f.variables[v["name"]]
        assert var.dimensions == (("time", "depth", "lat", "lon") if v["nelem"] > 1 else ("time", "lat", "lon"))
        assert var.units == v["units"].encode() and var.internal_vic_name == v["internal_vic_name"].encode() and var.category == v["category"].encode()
        assert np.float32(var._FillValue) == fill
        data = var[:].reshape(nstep, -1, nlat, nlon)
        want = np.full(data.shape, fill, dtype=np.float32)
        for e in range(v["nelem"]):
            want[:, e, lat_i, lon_i] = rows[:, :, c0 + e]
        assert np.array_equal(data, want)
    f.close()


def test_output_writer_refuses_bad_requests(tmp_path):
    ok = dict(name="X", nelem=1)
    with pytest.raises(api.VicGpuError, match="outside the grid"):
        api.NcOutput(str(tmp_path / "a.nc"), 2, 0, 1, 2, 0, 1, "days since 2001-1-1", 1, [ok], [0], [2], [0])
    with pytest.raises(api.VicGpuError, match="nelem outside"):
        api.NcOutput(str(tmp_path / "b.nc"), 2, 0, 1, 2, 0, 1, "days since 2001-1-1", 1, [dict(name="X", nelem=31)], [0], [0], [0])
    with pytest.raises(api.VicGpuError, match="cannot create"):
        api.NcOutput(str(tmp_path / "no_such_dir" / "c.nc"), 2, 0, 1, 2, 0, 1, "days since 2001-1-1", 1, [ok], [0], [0], [0])


@pytest.mark.parametrize("seed", range(12))
def test_reader_on_random_file_structures(seed, tmp_path):
    """random classic files: both versions, record or fixed time dimension, forcing variables of the three served types created in random order
    between unrelated variables (other dimensions, characters, bytes, odd sizes that need padding), random grid sizes: the slab is what scipy reads"""
    rng = np.random.default_rng(100 + seed)
    nt, nlat, nlon = int(rng.integers(1, 9)), int(rng.integers(1, 7)), int(rng.integers(1, 9))
    record = bool(rng.integers(0, 2))
    path = str(tmp_path / "f.nc")
    f = netcdf_file(path, "w", version=int(rng.choice([1, 2])))
    f.createDimension("time", None if record else nt)
    f.createDimension("lat", nlat)
    f.createDimension("lon", nlon)
    f.createDimension("odd", 3)
    f.createDimension("strlen", 5)
    lat = np.sort(rng.uniform(-80, 80, nlat)).astype(np.float32).astype(np.float64)
    lon = np.sort(rng.uniform(-170, 170, nlon))
    makers, expect = [], {}

    def coord(name, vals, typ):
        def make():
            v = f.createVariable(name, typ, (name,))
            v[:] = vals
        return make
    makers += [coord("time", np.arange(nt, dtype=np.float64), rng.choice(["d", "f"])), coord("lat", lat, "f"), coord("lon", lon, "d")]
    for k in range(int(rng.integers(1, 5))):
        name, typ = f"v{k}", str(rng.choice(["h", "f", "d"]))
        if typ == "h":
            data = rng.integers(-30000, 30000, (nt, nlat, nlon)).astype(np.int16)
            inv = bool(rng.integers(0, 2))
            scale = np.float32(rng.choice([10.0, 100.0, 0.1, 0.025]))
            expect[name] = data.astype(np.float64) / np.float64(scale) if inv else data.astype(np.float64) * np.float64(scale)
        else:
            data = rng.normal(0, 50, (nt, nlat, nlon)).astype(np.float32 if typ == "f" else np.float64)
            inv, scale = False, None
            expect[name] = data.astype(np.float64)

        def make(name=name, typ=typ, data=data, inv=inv, scale=scale):
            v = f.createVariable(name, typ, ("time", "lat", "lon"))
            v[:] = data
            if scale is not None:
                setattr(v, "inverse_scale_factor" if inv else "scale_factor", scale)
            v.long_name = "x" * int(rng.integers(0, 9))  # (attribute text of any length: header padding)
        makers.append(make)

    def junk1():
        v = f.createVariable("station", "c", ("odd", "strlen"))
        v[:] = np.array([list("abcde"), list("fghij"), list("klmno")], dtype="S1")
    def junk2():
        v = f.createVariable("flags", "b", ("odd",))
        v[:] = np.array([1, 2, 3], dtype=np.int8)
    def junk3():
        v = f.createVariable("perstep", "h", ("time", "odd"))
        v[:] = rng.integers(0, 9, (nt, 3)).astype(np.int16)
    makers += [junk1, junk2, junk3]
    for i in rng.permutation(len(makers)):
        makers[i]()
    f.title = "random"
    f.close()
    ii, jj = rng.integers(0, nlat, 11), rng.integers(0, nlon, 11)
    names = sorted(expect)
    t0 = int(rng.integers(0, nt))
    with api.NcForcing(path) as nc:
        assert (nc.ntime, nc.nlat, nc.nlon) == (nt, nlat, nlon)
        got = nc.read_slab(names, t0, nt - t0, lat[ii], lon[jj])
    want = np.stack([expect[n][t0:, ii, jj] for n in names], axis=1)
    assert np.array_equal(got, want)
