"""Lake-ice surface solve: ice_melt() with its residual IceEnergyBalance::calculate and icerad(), restated in
vic_b200/csrc/vic_lakeice.cuh and served as the batch operator vicgpu_ice_melt.  Oracle: the reference's own ice_melt(), compiled from
its sources (oracle/_ref/icemeltcheck); bar: bit-exact on every output of every column (NaN -- the reference's INVALID surface
temperature of a thin pack -- equals NaN)."""
import os
import re
import subprocess

import numpy as np
import pytest

from vic_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "oracle", "_ref", "icemeltcheck")
GOLD = os.path.join(ROOT, "tests", "golden", "ops", "ice_melt.npz")


def test_record_columns_are_the_headers():
    hdr = open(os.path.join(ROOT, "include", "vicgpu.h")).read()
    for macro, names in (("VICGPU_ICE_IN", api.ICE_IN), ("VICGPU_ICE_OUT", api.ICE_OUT)):
        body = re.search(r"#define " + macro + r"\(X\)((?:.*\\\n)*.*)\n", hdr).group(1)
        assert re.findall(r"X\((\w+)\)", body) == list(names)


@pytest.mark.parametrize("seed,dt,tfallback", [(21, 1, 1), (22, 3, 0), (23, 24, 1)])
def test_host_build_matches_the_references_ice_melt(seed, dt, tfallback):
    """60,000 seeded random columns per case through the reference's ice_melt() and the host build of vic_lakeice.cuh: identical;
    all three regimes are hit (balance closed at 0 C, surface temperature solved by root_brent, thin pack with INVALID surface)"""
    if not os.path.exists(CHECK):
        pytest.skip(f"{CHECK} not built (oracle/Makefile)")
    o = subprocess.run([CHECK, "-n", "60000", "--seed", str(seed), "--dt", str(dt), "--tfallback", str(tfallback)], capture_output=True, text=True)
    assert o.returncode == 0 and "identical" in o.stdout, o.stdout[-3000:]
    m = re.search(r"balance at 0 C (\d+), surface solved (\d+), thin pack \(INVALID surface\) (\d+)", o.stdout)
    assert all(int(x) > 0 for x in m.groups()), o.stdout


def test_golden_vectors_are_the_references():
    """the committed fixture is what icemeltcheck writes today (tests/golden/make_ice_melt_golden.py)"""
    if not os.path.exists(CHECK):
        pytest.skip(f"{CHECK} not built (oracle/Makefile)")
    from vic_b200.casefile import read_case
    g = np.load(GOLD)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "c.bin")
        subprocess.run([CHECK, "-n", "700", "--seed", "11", "--dt", "1", "--tfallback", "1", "-o", path], check=True, capture_output=True)
        c = read_case(path)
    assert np.array_equal(c["in"], g["dt1_in"]) and np.array_equal(c["out_ref"], g["dt1_out_ref"], equal_nan=True)


def _same(out, ref):
    ok = ref[:, 0] == 0  # after an ERROR return the reference's outputs are undefined
    assert np.array_equal(out[:, 0], ref[:, 0])
    assert np.array_equal(out[ok], ref[ok], equal_nan=True), np.argwhere(~((out[ok] == ref[ok]) | ((out[ok] != out[ok]) & (ref[ok] != ref[ok]))))[:5]


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["dt1", "dt3"])
def test_operator_matches_the_reference_golden(tag):
    g = np.load(GOLD)
    dt, tfb = (int(x) for x in g[f"{tag}_meta"])
    _same(api.ice_melt(g[f"{tag}_in"], dt, bool(tfb)), g[f"{tag}_out_ref"])


@pytest.mark.gpu
def test_operator_at_size_and_ragged_batches():
    """300,001 columns (the golden columns repeated; not a multiple of the 128-row block): every repetition gives the golden bits;
    one column and an empty batch work"""
    g = np.load(GOLD)
    a, ref = g["dt1_in"], g["dt1_out_ref"]
    reps = 300001 // a.shape[0] + 1
    big = np.tile(a, (reps, 1))[:300001]
    out = api.ice_melt(big, 1, True)
    _same(out, np.tile(ref, (reps, 1))[:300001])
    _same(api.ice_melt(a[:1], 1, True), ref[:1])
    assert api.ice_melt(np.empty((0, len(api.ICE_IN))), 1, True).shape == (0, len(api.ICE_OUT))


def test_operator_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    g = np.load(GOLD)
    with pytest.raises(api.VicGpuError) as e:
        api.ice_melt(g["dt1_in"], 1, True)
    assert e.value.code == -2
