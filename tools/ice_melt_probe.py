"""Throughput of the lake-ice operator (vicgpu_ice_melt) on the device, end to end through the C-ABI with host buffers (allocation, H2D, kernel,
D2H), on the golden columns repeated to N; checks every repetition against the reference's golden answers.  python tools/ice_melt_probe.py [N]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vic_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
g = np.load(os.path.join(ROOT, "tests", "golden", "ops", "ice_melt.npz"))
a, ref = g["dt1_in"], g["dt1_out_ref"]
reps = n // a.shape[0] + 1
big = np.ascontiguousarray(np.tile(a, (reps, 1))[:n])
api.ice_melt(big[:1000], 1, True)  # context, module load
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    out = api.ice_melt(big, 1, True)
    best = min(best, time.perf_counter() - t0)
want = np.tile(ref, (reps, 1))[:n]
ok = want[:, 0] == 0
same = np.array_equal(out[ok], want[ok], equal_nan=True)
print(f"ICE_MELT_PROBE columns={n} seconds={best:.4f} columns_per_s={n / best:.4e} bytes_in={big.nbytes} bytes_out={out.nbytes} bit_identical_to_reference={same}")
