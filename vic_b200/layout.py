"""Column layout of the flat C-ABI records, derived by parsing the X-macro tables of
include/vicgpu_fields.h and include/vicgpu.h (so Python cannot drift from the C side).

Mirrors vicgpu_layout_init() (include/vicgpu.h).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field

_INC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")

NLAYER, NFRONTS, NPET, NZWT = 3, 3, 6, 11


def _macro_body(text, name):
    m = re.search(r"#define\s+" + name + r"\s*\([^)]*\)\s*\\\n((?:.*\\\n)*.*\n)", text)
    if not m:
        raise KeyError(name)
    return m.group(1)


def _entries(text, name):
    body = _macro_body(text, name)
    out = []
    for m in re.finditer(r"X\(\s*(?:P##)?(\w+)\s*,", body):
        out.append(m.group(1))
    return out


def _load():
    with open(os.path.join(_INC, "vicgpu_fields.h")) as f:
        fields = f.read()
    with open(os.path.join(_INC, "vicgpu.h")) as f:
        api = f.read()
    t = {}
    t["energy"] = _entries(fields, "VICGPU_HRU_ENERGY")
    t["snow"] = _entries(fields, "VICGPU_HRU_SNOW")
    t["cell"] = _entries(fields, "VICGPU_HRU_CELL")
    t["veg"] = _entries(fields, "VICGPU_HRU_VEG")
    t["glac"] = _entries(fields, "VICGPU_HRU_GLAC")
    t["layer"] = _entries(fields, "VICGPU_HRU_LAYER")
    t["front"] = _entries(fields, "VICGPU_HRU_FRONT")
    t["node"] = _entries(fields, "VICGPU_HRU_NODE")
    t["hpar"] = _entries(fields, "VICGPU_HPAR_SCALARS")
    t["cpar"] = _entries(fields, "VICGPU_CPAR_SCALARS")
    t["cpar_layer"] = _entries(fields, "VICGPU_CPAR_LAYER")
    t["cpar_node"] = _entries(fields, "VICGPU_CPAR_NODE")
    t["cpar_zwt"] = _entries(fields, "VICGPU_CPAR_ZWT")
    t["cpar_band"] = _entries(fields, "VICGPU_CPAR_BAND")
    t["veglib"] = _entries(fields, "VICGPU_VEGLIB_SCALARS")
    t["veglib_monthly"] = _entries(fields, "VICGPU_VEGLIB_MONTHLY")
    t["forcing"] = _entries(fields, "VICGPU_FORCING")
    body = _macro_body(api, "VICGPU_OUTVARS")
    t["outvars"] = [(m.group(1), m.group(2), m.group(3)) for m in re.finditer(r"X\(\s*(\w+)\s*,\s*(\w+)\s*,\s*(\w+)\s*\)", body)]
    return t


TABLES = _load()


@dataclass
class Layout:
    nnode: int
    nbands: int
    frozen_soil: bool
    nf: int
    hru_names: list = field(default_factory=list)
    out_names: list = field(default_factory=list)
    out_off: dict = field(default_factory=dict)
    out_nelem: dict = field(default_factory=dict)
    out_agg: dict = field(default_factory=dict)

    def __post_init__(self):
        t = TABLES
        names = ["E_" + n for n in t["energy"]] + ["S_" + n for n in t["snow"]] + ["C_" + n for n in t["cell"]]
        names += ["V_" + n for n in t["veg"]] + ["G_" + n for n in t["glac"]] + ["H_mu"]
        self.hr_nscalar = len(names)
        self.hr_layer0 = len(names)
        for f in t["layer"]:
            names += [f"L_{f}[{i}]" for i in range(NLAYER)]
        self.hr_front0 = len(names)
        for f in t["front"]:
            names += [f"F_{f}[{i}]" for i in range(NFRONTS)]
        self.hr_pet0 = len(names)
        names += [f"pot_evap[{i}]" for i in range(NPET)]
        self.hr_node0 = len(names)
        for f in t["node"]:
            names += [f"N_{f}[{i}]" for i in range(self.nnode)]
        self.hru_names = names
        self.hr_stride = len(names)
        self.cp_stride = len(t["cpar"]) + len(t["cpar_layer"]) * NLAYER + len(t["cpar_node"]) * self.nnode \
            + len(t["cpar_zwt"]) * (NLAYER + 2) * NZWT + len(t["cpar_band"]) * self.nbands
        self.f_nslot = self.nf + 1 if self.nf > 1 else 1
        self.f_stride = len(t["forcing"]) * self.f_nslot
        nel = {"1": 1, "L": NLAYER, "N": self.nnode, "B": self.nbands, "F": NFRONTS if self.frozen_soil else 1}
        off = 0
        for (n, e, a) in t["outvars"]:
            self.out_off[n] = off
            self.out_nelem[n] = nel[e]
            self.out_agg[n] = a
            for i in range(nel[e]):
                self.out_names.append(n if nel[e] == 1 else f"{n}[{i}]")
            off += nel[e]
        self.nout = off

    def hru_col(self, name):
        return self.hru_names.index(name)

    def out_col(self, name, elem=0):
        return self.out_off[name] + elem


# field order of struct vicgpu_options (include/vicgpu.h); all int32 except the two trailing doubles
OPTION_INT_FIELDS = [
    "abi_version", "Nlayer", "Nnode", "Nbands", "dt", "SNOW_STEP", "NR", "NF", "nrecs", "out_step_ratio",
    "FULL_ENERGY", "FROZEN_SOIL", "QUICK_FLUX", "QUICK_SOLVE", "IMPLICIT", "EXP_TRANS", "NOFLUX",
    "GRND_FLUX_TYPE", "AERO_RESIST_CANSNOW", "SNOW_ALBEDO", "SNOW_DENSITY", "TEMP_TH_TYPE",
    "TFALLBACK", "BLOWING", "DIST_PRCP", "CORRPREC", "LAKES", "COMPUTE_TREELINE", "GLACIER_ID", "GLACIER_DYNAMICS",
    "MOISTFRACT", "ALMA_OUTPUT", "NVegLibTypes", "glacierAccumStartYear", "glacierAccumStartMonth", "glacierAccumStartDay",
    "glacierAccumInterval",
]


def parse_options(raw_i32):
    """options_raw (int32 view of struct vicgpu_options) -> dict"""
    import numpy as np
    raw = np.asarray(raw_i32, dtype=np.int32)
    d = {k: int(raw[i]) for i, k in enumerate(OPTION_INT_FIELDS)}
    n = len(OPTION_INT_FIELDS)
    n += n % 2  # doubles are 8-byte aligned
    dbl = raw[n:n + 4].view(np.float64)
    d["wind_h"] = float(dbl[0])
    d["MIN_WIND_SPEED"] = float(dbl[1])
    return d


def layout_from_options(opt):
    return Layout(nnode=opt["Nnode"], nbands=opt["Nbands"], frozen_soil=bool(opt["FROZEN_SOIL"]), nf=opt["NF"])
