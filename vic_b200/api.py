"""ctypes binding of libvicgpu.so (include/vicgpu.h) for the tests and bench.py.

The product's host side is C (the reference's own vicNl + vic_b200/host/vicgpu_pack.h, see
INTEGRATION.md); this module only lets Python drive the same C-ABI.  There is no fallback of
any kind: if the shared library is missing, or no CUDA device is present, it raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .layout import OPTION_INT_FIELDS, layout_from_options, parse_options

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VICGPU_LIB") or os.path.join(_HERE, "lib", "libvicgpu.so")  # VICGPU_LIB: an alternative build, for A/B measurements

N_OUTVARS = 184

ERRORS = {0: "OK", -1: "EINVAL", -2: "ENODEV", -3: "EUNSUPPORTED", -4: "ECUDA", -5: "ESTATE"}

# every symbol include/vicgpu.h declares
SYMBOLS = [
    "vicgpu_abi_version", "vicgpu_last_error", "vicgpu_create", "vicgpu_destroy", "vicgpu_get_layout",
    "vicgpu_set_veglib", "vicgpu_set_cells", "vicgpu_set_output_spec", "vicgpu_set_cell_status", "vicgpu_set_state",
    "vicgpu_get_state", "vicgpu_set_forcing", "vicgpu_step", "vicgpu_step_f32", "vicgpu_get_cell_status", "vicgpu_get_balance_errors",
    "vicgpu_get_last_step_timing", "vicgpu_measure_phase_tax", "vicgpu_set_profiling", "vicgpu_get_kernel_profile", "vicgpu_disagg", "vicgpu_disagg_tm", "vicgpu_nc_open", "vicgpu_nc_close", "vicgpu_nc_dims", "vicgpu_nc_read_slab", "vicgpu_ncout_create", "vicgpu_ncout_write_step", "vicgpu_ncout_close", "vicgpu_ice_melt", "vicgpu_get_warp_times", "vicgpu_get_glacier_fit", "vicgpu_measure_fp64_peak",
]


class NcOutVar(C.Structure):  # vicgpu_ncout_var
    _fields_ = [("name", C.c_char_p), ("nelem", C.c_int), ("long_name", C.c_char_p), ("units", C.c_char_p), ("standard_name", C.c_char_p),
                ("cell_methods", C.c_char_p), ("internal_vic_name", C.c_char_p), ("category", C.c_char_p)]


class NcOutSpec(C.Structure):  # vicgpu_ncout_spec
    _fields_ = [("nlat", C.c_int), ("nlon", C.c_int), ("depth", C.c_int), ("lat0", C.c_double), ("dlat", C.c_double), ("lon0", C.c_double), ("dlon", C.c_double),
                ("time_units", C.c_char_p), ("time_step", C.c_double), ("nvar", C.c_int), ("vars", C.POINTER(NcOutVar)),
                ("ntext", C.c_int), ("text_keys", C.POINTER(C.c_char_p)), ("text_values", C.POINTER(C.c_char_p)),
                ("nint", C.c_int), ("int_keys", C.POINTER(C.c_char_p)), ("int_values", C.POINTER(C.c_int)),
                ("ncell", C.c_int), ("lat_index", C.POINTER(C.c_int)), ("lon_index", C.POINTER(C.c_int))]


class VicGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvicgpu: {ERRORS.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load_library(path=LIB_PATH):
    """dlopen libvicgpu.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a)")
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    vp = C.c_void_p
    lib.vicgpu_abi_version.restype = C.c_int
    lib.vicgpu_last_error.restype = C.c_char_p
    lib.vicgpu_create.argtypes = [C.POINTER(vp), vp, C.c_int]
    lib.vicgpu_destroy.argtypes = [vp]
    lib.vicgpu_get_layout.argtypes = [vp, vp]
    lib.vicgpu_set_veglib.argtypes = [vp, C.c_int, dp]
    lib.vicgpu_set_cells.argtypes = [vp, C.c_int, dp, C.c_int, dp]
    lib.vicgpu_set_output_spec.argtypes = [vp, ip]
    lib.vicgpu_set_cell_status.argtypes = [vp, ip]
    lib.vicgpu_set_state.argtypes = [vp, dp]
    lib.vicgpu_get_state.argtypes = [vp, dp]
    lib.vicgpu_set_forcing.argtypes = [vp, C.c_int, C.c_int, dp]
    lib.vicgpu_step.argtypes = [vp, C.c_int, C.c_int, ip, dp, dp]
    lib.vicgpu_step_f32.argtypes = [vp, C.c_int, C.c_int, ip, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.vicgpu_get_cell_status.argtypes = [vp, ip]
    lib.vicgpu_get_balance_errors.argtypes = [vp, dp]
    lib.vicgpu_get_last_step_timing.argtypes = [vp, dp, C.POINTER(C.c_longlong)]
    lib.vicgpu_disagg.argtypes = [vp, vp, dp, dp]
    lib.vicgpu_disagg_tm.argtypes = [vp, vp, dp, dp]
    lib.vicgpu_ice_melt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, dp, dp]
    lib.vicgpu_ncout_create.argtypes = [C.POINTER(vp), C.c_char_p, C.POINTER(NcOutSpec)]
    lib.vicgpu_ncout_write_step.argtypes = [vp, C.POINTER(C.c_float), C.c_longlong, ip]
    lib.vicgpu_ncout_close.argtypes = [vp]
    lib.vicgpu_nc_open.argtypes = [C.POINTER(vp), C.c_char_p]
    lib.vicgpu_nc_close.argtypes = [vp]
    lib.vicgpu_nc_dims.argtypes = [vp] + [C.POINTER(C.c_longlong)] * 3
    lib.vicgpu_nc_read_slab.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.c_longlong, C.c_longlong, C.c_int, dp, dp, dp]
    if hasattr(lib, "vicgpu_measure_phase_tax") or not os.environ.get("VICGPU_LIB"):  # (an older A/B build may lack it)
        lib.vicgpu_measure_phase_tax.argtypes = [vp, C.c_int, C.c_int, dp]
    lib.vicgpu_set_profiling.argtypes = [vp, C.c_int]
    lib.vicgpu_get_kernel_profile.argtypes = [vp, dp, C.POINTER(C.c_longlong)]
    lib.vicgpu_get_warp_times.argtypes = [vp, dp, dp, C.c_int]
    lib.vicgpu_get_glacier_fit.argtypes = [vp, dp]
    lib.vicgpu_measure_fp64_peak.argtypes = [C.c_int, dp]
    for s in SYMBOLS:
        if not os.environ.get("VICGPU_LIB"):
            getattr(lib, s)
    _lib = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _as_f64(a):
    """contiguous float64 view without copying when already so (torch pinned tensors come in through .numpy())"""
    a = np.asarray(a)
    if a.dtype != np.float64 or not a.flags["C_CONTIGUOUS"]:
        a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def options_to_raw(opt: dict) -> np.ndarray:
    """dict -> int32 view of struct vicgpu_options"""
    n = len(OPTION_INT_FIELDS)
    npad = n + n % 2
    raw = np.zeros(npad + 4, dtype=np.int32)
    for i, k in enumerate(OPTION_INT_FIELDS):
        raw[i] = int(opt[k])
    raw[npad:npad + 4] = np.array([opt["wind_h"], opt["MIN_WIND_SPEED"]], dtype=np.float64).view(np.int32)
    return raw


class VicGpu:
    """One model domain resident on one GPU (mirrors the calls a patched runModel() makes)."""

    def __init__(self, options_raw, device=0):
        self.lib = load_library()
        self.options_raw = np.ascontiguousarray(options_raw, dtype=np.int32)
        self.opt = parse_options(self.options_raw)
        self.L = layout_from_options(self.opt)
        self.h = C.c_void_p()
        self._chk(self.lib.vicgpu_create(C.byref(self.h), self.options_raw.ctypes.data_as(C.c_void_p), int(device)))
        self.ncell = self.nhru = 0

    def _chk(self, rc):
        if rc != 0:
            raise VicGpuError(rc, (self.lib.vicgpu_last_error() or b"").decode())

    def close(self):
        if self.h:
            self.lib.vicgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_veglib(self, veglib):
        v = _as_f64(veglib)
        self._chk(self.lib.vicgpu_set_veglib(self.h, int(v.shape[0]), _dptr(v)))

    def set_cells(self, cellpar, hrupar):
        c, p = _as_f64(cellpar), _as_f64(hrupar)
        assert c.shape[1] == self.L.cp_stride, (c.shape, self.L.cp_stride)
        self.ncell, self.nhru = int(c.shape[0]), int(p.shape[0])
        self._chk(self.lib.vicgpu_set_cells(self.h, self.ncell, _dptr(c), self.nhru, _dptr(p)))

    def set_output_spec(self, aggtype=None):
        if aggtype is None:
            self._chk(self.lib.vicgpu_set_output_spec(self.h, None))
        else:
            a = np.ascontiguousarray(aggtype, dtype=np.int32)
            assert a.size == N_OUTVARS
            self._chk(self.lib.vicgpu_set_output_spec(self.h, _iptr(a)))

    def set_cell_status(self, status):
        s = np.ascontiguousarray(status, dtype=np.int32)
        self._chk(self.lib.vicgpu_set_cell_status(self.h, _iptr(s)))

    def set_state(self, hrurec):
        r = _as_f64(hrurec)
        assert r.shape == (self.nhru, self.L.hr_stride), (r.shape, self.nhru, self.L.hr_stride)
        self._chk(self.lib.vicgpu_set_state(self.h, _dptr(r)))

    def get_state(self):
        r = np.empty((self.nhru, self.L.hr_stride), dtype=np.float64)
        self._chk(self.lib.vicgpu_get_state(self.h, _dptr(r)))
        return r

    def set_forcing(self, rec0, forcing):
        f = _as_f64(forcing)
        assert f.shape[1:] == (self.ncell, self.L.f_stride), f.shape
        # the upload is asynchronous: the array must outlive it (the library keeps two windows)
        self._keep = (getattr(self, "_keep", ()) + (f,))[-2:]
        self._chk(self.lib.vicgpu_set_forcing(self.h, int(rec0), int(f.shape[0]), _dptr(f)))

    def disagg(self, disagg_raw, daily, want_host=True):
        """daily [ncell][Ndays][4] (PREC, TMAX, TMIN, WIND) -> fills the device forcing window [0, nrecs); returns the
        hourly / sub-daily forcing [nrecs][ncell][f_stride] when want_host (vicgpu_disagg, include/vicgpu.h)."""
        raw = np.ascontiguousarray(disagg_raw, dtype=np.int32)
        d = _as_f64(daily)
        assert d.shape == (self.ncell, int(raw[5]), 4), d.shape
        out = np.empty((self.opt["nrecs"], self.ncell, self.L.f_stride), dtype=np.float64) if want_host else None
        self._chk(self.lib.vicgpu_disagg(self.h, raw.ctypes.data_as(C.c_void_p), _dptr(d), _dptr(out) if want_host else None))
        return out

    def disagg_tm(self, disagg_raw, daily_tm, want_host=True):
        """the same from time-major input daily_tm [Ndays][4][ncell] (what NcForcing.read_slab returns): vicgpu_disagg_tm"""
        raw = np.ascontiguousarray(disagg_raw, dtype=np.int32)
        d = _as_f64(daily_tm)
        assert d.shape == (int(raw[5]), 4, self.ncell), d.shape
        out = np.empty((self.opt["nrecs"], self.ncell, self.L.f_stride), dtype=np.float64) if want_host else None
        self._chk(self.lib.vicgpu_disagg_tm(self.h, raw.ctypes.data_as(C.c_void_p), _dptr(d), _dptr(out) if want_host else None))
        return out

    def n_output_steps(self, nrec, step_count0=0):
        return (step_count0 + nrec) // self.opt["out_step_ratio"]

    def step(self, rec0, nrec, dmy, out_data=None, out_agg=None):
        """dmy: int32 [nrec+1][5].  out_data / out_agg: preallocated float64 (vicgpu_step) or float32 (vicgpu_step_f32: narrowed on
        the device as the reference's NetCDF writer does) arrays, or None."""
        d = np.ascontiguousarray(dmy, dtype=np.int32)
        assert d.shape[0] >= nrec + 1 and d.shape[1] == 5
        dts = {a.dtype for a in (out_data, out_agg) if a is not None}
        assert len(dts) <= 1 and dts <= {np.dtype(np.float64), np.dtype(np.float32)}, dts
        if dts == {np.dtype(np.float32)}:
            fp = C.POINTER(C.c_float)
            od = out_data.ctypes.data_as(fp) if out_data is not None else None
            oa = out_agg.ctypes.data_as(fp) if out_agg is not None else None
            self._chk(self.lib.vicgpu_step_f32(self.h, int(rec0), int(nrec), _iptr(d), od, oa))
            return
        od = _dptr(out_data) if out_data is not None else None
        oa = _dptr(out_agg) if out_agg is not None else None
        self._chk(self.lib.vicgpu_step(self.h, int(rec0), int(nrec), _iptr(d), od, oa))

    def cell_status(self):
        s = np.empty(self.ncell, dtype=np.int32)
        self._chk(self.lib.vicgpu_get_cell_status(self.h, _iptr(s)))
        return s

    def balance_errors(self):
        e = np.empty((self.ncell, 5), dtype=np.float64)
        self._chk(self.lib.vicgpu_get_balance_errors(self.h, _dptr(e)))
        return e

    def glacier_fit(self):
        """[ncell][4]: b0, b1, b2, fitError of the glacier mass-balance curve of the last completed accumulation interval"""
        g = np.empty((self.ncell, 4), dtype=np.float64)
        self._chk(self.lib.vicgpu_get_glacier_fit(self.h, _dptr(g)))
        return g

    def set_profiling(self, on=True):
        self._chk(self.lib.vicgpu_set_profiling(self.h, 1 if on else 0))

    def kernel_profile(self):
        ms = C.c_double()
        n = C.c_longlong()
        self._chk(self.lib.vicgpu_get_kernel_profile(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def warp_times(self):
        """(start_ns, end_ns, kind) per warp of the last profiled launch of the step kernel"""
        nw = (self.nhru + 31) // 32
        t = np.zeros((nw, 2))
        k = np.zeros(nw)
        n = self.lib.vicgpu_get_warp_times(self.h, _dptr(t), _dptr(k), nw)
        if n < 0:
            self._chk(n)
        return t[:, 0], t[:, 1], k

    def phase_tax(self, nframe, reps=20):
        """device time [us] of one record-in / frame-in-out / record-out pass over the domain (vicgpu_measure_phase_tax)"""
        v = C.c_double()
        self._chk(self.lib.vicgpu_measure_phase_tax(self.h, int(nframe), int(reps), C.byref(v)))
        return v.value

    def last_step_timing(self):
        ms = C.c_double()
        n = C.c_longlong()
        self._chk(self.lib.vicgpu_get_last_step_timing(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def measure_fp64_peak(device=0):
    """FP64 FMA throughput of the device in TFLOP/s (vicgpu_measure_fp64_peak)"""
    lib = load_library()
    v = C.c_double()
    rc = lib.vicgpu_measure_fp64_peak(int(device), C.byref(v))
    if rc != 0:
        raise VicGpuError(rc, (lib.vicgpu_last_error() or b"").decode())
    return v.value


def run_case(case, device=0, nrec=None, want_out=True, block=None):
    """Run a case file's inputs through the CUDA library; returns arrays named like the host port's result file."""
    g = VicGpu(case["options_raw"], device)
    try:
        L = g.L
        g.set_veglib(case["veglib"])
        g.set_cells(case["cellpar"], case["hrupar"])
        g.set_output_spec(case["aggtype"])
        if "valid0" in case:
            g.set_cell_status(np.where(case["valid0"] != 0, 0, -999).astype(np.int32))
        g.set_state(case["hrurec0"])
        ntot = int(case["forcing"].shape[0])
        nrec = ntot if nrec is None else min(nrec, ntot)
        g.set_forcing(0, case["forcing"][:nrec])
        dmy = case["dmy"]
        dump_recs = [int(r) for r in case.get("dump_recs", []) if r < nrec]
        ratio = g.opt["out_step_ratio"]
        out = np.zeros((nrec, g.ncell, L.nout)) if want_out else None
        agg = np.zeros((nrec // ratio, g.ncell, L.nout))
        hru = []
        # advance from dump record to dump record so that the state can be read back in between
        cuts = sorted(set([r + 1 for r in dump_recs] + [nrec]))
        r0 = 0
        nagg = 0
        for c in cuts:
            n = c - r0
            if n <= 0:
                continue
            na = ((r0 % ratio) + n) // ratio
            g.step(r0, n, dmy[r0:r0 + n + 1], out[r0:c] if want_out else None, agg[nagg:nagg + na] if na else None)
            nagg += na
            if (c - 1) in dump_recs:
                hru.append(g.get_state())
            r0 = c
        res = {"agg": agg, "hrurec": np.array(hru), "balance": g.balance_errors(), "status": g.cell_status(), "gmb": g.glacier_fit()}
        if want_out:
            res["out"] = out
        return res
    finally:
        g.close()


# columns of vicgpu_ice_melt's records (VICGPU_ICE_IN / VICGPU_ICE_OUT, include/vicgpu.h; tests/test_lakeice.py checks the order)
ICE_IN = ("z2 aero_resist latent_heat_Le Z0 rainfall snowfall wind Tcutoff air_temp net_short longwave density pressure vpd vp swq surf_temp pack_temp "
          "pack_water surf_water vapor_flux surface_flux surf_temp_fbflag surf_temp_fbcount ice_water_eq areai hice volume").split()
ICE_OUT = ("rc aero_resist_used melt advection deltaCC SnowFlux latent sensible Qnet refreeze_energy LWnet swq surf_temp pack_temp pack_water surf_water "
           "vapor_flux blowing_flux surface_flux surf_temp_fbflag surf_temp_fbcount coverage mass_error coldcontent ice_water_eq volume").split()


def ice_melt(columns, delta_t, tfallback=True, device=0):
    """the reference's ice_melt() (ice_melt.c:30-585) for a batch of lake-ice columns on the device: columns [n][len(ICE_IN)] ->
    [n][len(ICE_OUT)] (vicgpu_ice_melt, include/vicgpu.h)"""
    lib = load_library()
    a = _as_f64(columns)
    assert a.ndim == 2 and a.shape[1] == len(ICE_IN), a.shape
    out = np.empty((a.shape[0], len(ICE_OUT)), dtype=np.float64)
    rc = lib.vicgpu_ice_melt(int(device), int(a.shape[0]), int(delta_t), int(bool(tfallback)), _dptr(a), _dptr(out))
    if rc != 0:
        raise VicGpuError(rc, (lib.vicgpu_last_error() or b"").decode())
    return out


class NcForcing:
    """A NetCDF (classic CDF-1 / CDF-2) forcing file read as time-major slabs through the library's host-side reader
    (vicgpu_nc_*, include/vicgpu.h; stands in for read_atmos_data.c:109-338).  No device needed."""

    def __init__(self, path):
        self.lib = load_library()
        self.h = C.c_void_p()
        rc = self.lib.vicgpu_nc_open(C.byref(self.h), os.fsencode(path))
        if rc != 0:
            raise VicGpuError(rc, self.lib.vicgpu_last_error().decode())
        n = [C.c_longlong() for _ in range(3)]
        rc = self.lib.vicgpu_nc_dims(self.h, *[C.byref(x) for x in n])
        if rc != 0:
            msg = self.lib.vicgpu_last_error().decode()
            self.close()
            raise VicGpuError(rc, msg)
        self.ntime, self.nlat, self.nlon = (int(x.value) for x in n)

    def read_slab(self, varnames, t0, nt, lat, lng, out=None):
        """-> float64 [nt][len(varnames)][ncell] (into `out` when given, e.g. a pinned buffer); lat / lng: the cells' (double)(float) coordinates"""
        la, lo = _as_f64(lat), _as_f64(lng)
        assert la.shape == lo.shape and la.ndim == 1
        shape = (int(nt), len(varnames), la.shape[0])
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        assert out.shape == shape and out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
        names = (C.c_char_p * len(varnames))(*[v.encode() for v in varnames])
        rc = self.lib.vicgpu_nc_read_slab(self.h, len(varnames), names, int(t0), int(nt), int(la.shape[0]), _dptr(la), _dptr(lo), _dptr(out))
        if rc != 0:
            raise VicGpuError(rc, self.lib.vicgpu_last_error().decode())
        return out

    def close(self):
        if self.h:
            self.lib.vicgpu_nc_close(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class NcOutput:
    """The model output as a NetCDF file, one record per output step, through the library's host-side writer (vicgpu_ncout_*,
    include/vicgpu.h; stands in for WriteOutputNetCDF.c:163-299, 386-452).  No device needed.
    variables: list of dicts {name, nelem, long_name, units, standard_name, cell_methods, internal_vic_name, category};
    columns: first column of each variable in the float32 rows handed to write_step."""

    def __init__(self, path, nlat, lat0, dlat, nlon, lon0, dlon, time_units, time_step, variables, columns, lat_index, lon_index, depth=30,
                 global_text=(), global_int=()):
        self.lib = load_library()
        enc = lambda x: (x or "").encode()  # noqa: E731
        vs = (NcOutVar * len(variables))(*[NcOutVar(enc(v["name"]), int(v.get("nelem", 1)), enc(v.get("long_name")), enc(v.get("units")), enc(v.get("standard_name")),
                                                     enc(v.get("cell_methods")), enc(v.get("internal_vic_name")), enc(v.get("category"))) for v in variables])
        tk = (C.c_char_p * max(1, len(global_text)))(*[k.encode() for k, _ in global_text])
        tv = (C.c_char_p * max(1, len(global_text)))(*[v.encode() for _, v in global_text])
        ik = (C.c_char_p * max(1, len(global_int)))(*[k.encode() for k, _ in global_int])
        iv = (C.c_int * max(1, len(global_int)))(*[int(v) for _, v in global_int])
        li = np.ascontiguousarray(lat_index, dtype=np.int32)
        lo = np.ascontiguousarray(lon_index, dtype=np.int32)
        assert li.shape == lo.shape and li.ndim == 1
        self.ncell = li.shape[0]
        self.cols = np.ascontiguousarray(columns, dtype=np.int32)
        assert self.cols.shape == (len(variables),)
        spec = NcOutSpec(int(nlat), int(nlon), int(depth), float(lat0), float(dlat), float(lon0), float(dlon), time_units.encode(), float(time_step), len(variables), vs,
                         len(global_text), tk, tv, len(global_int), ik, iv, self.ncell, _iptr(li), _iptr(lo))
        self.h = C.c_void_p()
        rc = self.lib.vicgpu_ncout_create(C.byref(self.h), os.fsencode(path), C.byref(spec))
        if rc != 0:
            raise VicGpuError(rc, (self.lib.vicgpu_last_error() or b"").decode())

    def write_step(self, rows):
        """rows: float32 [ncell][row_stride] (one output step of vicgpu_step_f32's out_agg)"""
        r = np.asarray(rows)
        assert r.dtype == np.float32 and r.ndim == 2 and r.shape[0] == self.ncell and r.flags["C_CONTIGUOUS"], (r.dtype, r.shape)
        rc = self.lib.vicgpu_ncout_write_step(self.h, r.ctypes.data_as(C.POINTER(C.c_float)), int(r.shape[1]), _iptr(self.cols))
        if rc != 0:
            raise VicGpuError(rc, (self.lib.vicgpu_last_error() or b"").decode())

    def close(self):
        if self.h:
            rc = self.lib.vicgpu_ncout_close(self.h)
            self.h = C.c_void_p()
            if rc != 0:
                raise VicGpuError(rc, (self.lib.vicgpu_last_error() or b"").decode())

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
