"""Reader/writer for the tagged-array "case file" container (oracle/casefile.h).

A case file carries the flat C-ABI inputs (options, veglib, cellpar, hrupar, hrurec0, dmy,
forcing) and, when written by the oracle harness, the reference's answers (hrurec_ref,
out_ref, agg_ref, balance_ref, status_ref).
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"VICCASE1"


def read_case(path):
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path}: not a VICCASE1 file")
        while True:
            hdr = f.read(32 + 4 + 4 + 32)
            if len(hdr) < 72:
                break
            name = hdr[:32].split(b"\0", 1)[0].decode()
            dtype, ndim = struct.unpack("<ii", hdr[32:40])
            dims = struct.unpack("<4q", hdr[40:72])[:ndim]
            n = int(np.prod(dims)) if ndim else 1
            dt = np.float64 if dtype == 0 else np.int32
            arr = np.fromfile(f, dtype=dt, count=n)
            if arr.size != n:
                raise ValueError(f"{path}: truncated record {name}")
            out[name] = arr.reshape(dims)
    return out


def write_case(path, arrays):
    with open(path, "wb") as f:
        f.write(MAGIC)
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            if a.dtype == np.float64:
                dtype = 0
            elif a.dtype == np.int32:
                dtype = 1
            else:
                raise TypeError(f"{name}: unsupported dtype {a.dtype}")
            dims = list(a.shape) + [1] * (4 - a.ndim)
            f.write(name.encode()[:31].ljust(32, b"\0"))
            f.write(struct.pack("<ii", dtype, a.ndim))
            f.write(struct.pack("<4q", *dims))
            f.write(a.tobytes())
