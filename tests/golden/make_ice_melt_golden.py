"""Golden vectors of the lake-ice operator (vicgpu_ice_melt): inputs and the answers of the REFERENCE's own ice_melt()
(oracle/_ref/icemeltcheck links the objects compiled from /root/reference) for seeded random lake-ice columns, one file per
(time step, TFALLBACK).  Run in the development container: python tests/golden/make_ice_melt_golden.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vic_b200.casefile import read_case  # noqa: E402

exe = os.path.join(ROOT, "oracle", "_ref", "icemeltcheck")
cases = {}
for tag, n, seed, dt, tfb in (("dt1", 700, 11, 1, 1), ("dt3", 500, 12, 3, 0)):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "case.bin")
        o = subprocess.run([exe, "-n", str(n), "--seed", str(seed), "--dt", str(dt), "--tfallback", str(tfb), "-o", path], capture_output=True, text=True)
        assert o.returncode == 0 and "identical" in o.stdout, o.stdout
        c = read_case(path)
    cases[f"{tag}_in"], cases[f"{tag}_out_ref"], cases[f"{tag}_meta"] = c["in"], c["out_ref"], c["meta"]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ops", "ice_melt.npz"), **cases)
print({k: v.shape for k, v in cases.items()})
