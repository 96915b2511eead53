"""Multi-rank path on the CPU (gloo, world_size 2): the domain partition, the max-over-ranks timing reduction and the end-of-run
gather.  Each rank advances its own cells with the host build of the physics headers (oracle/_ref/vicport -- the checker, there
being no GPU here) and rank 0 compares the gathered outputs with the unsharded run bit for bit: cutting the domain changes nothing."""
import os
import subprocess
import sys

import numpy as np
import pytest

from vic_b200.shard import cell_ranges, cell_ranges_by_weight, hrus_per_cell, shard_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "fe_hourly_winter.npz")
KEYS = ("options_raw", "meta", "veglib", "cellpar", "hrupar", "hrurec0", "aggtype", "valid0", "dmy", "forcing", "dump_recs")


def test_cell_ranges_cover_the_domain_exactly():
    for ncell in (1, 7, 16, 10000, 1000003):
        for world in (1, 2, 3, 8):
            r = cell_ranges(ncell, world)
            assert r[0][0] == 0 and r[-1][1] == ncell
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_weighted_ranges_balance_the_hrus():
    """contiguous cuts by HRU count: cover the domain in order, leave no rank empty, and no rank is further from the even share
    than one cell's weight"""
    rng = np.random.default_rng(7)
    for ncell in (1, 2, 9, 1000, 40000):
        w = rng.integers(1, 31, ncell)  # 1..30 HRUs per cell (tiles x bands)
        for world in (1, 2, 3, 8):
            r = cell_ranges_by_weight(w, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == ncell
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            if ncell >= world:
                assert all(b > a for a, b in r)
                sums = np.array([w[a:b].sum() for a, b in r])
                assert np.abs(sums - w.sum() / world).max() <= w.max() * 1.0 + 1e-9, (ncell, world, sums)
    # equal weights reduce to the equal-count partition up to where the odd cells go
    sizes = [b - a for a, b in cell_ranges_by_weight(np.ones(10), 3)]
    assert sorted(sizes) == [3, 3, 4]
    # a heavy head: the first rank gets few cells
    w = np.r_[np.full(10, 30), np.ones(300)]
    r = cell_ranges_by_weight(w, 2)
    assert abs(w[r[0][0]:r[0][1]].sum() - w.sum() / 2) <= 30 and r[0][1] == 10


def test_shard_case_with_hru_balanced_ranges():
    g = dict(np.load(GOLDEN))
    ranges = cell_ranges_by_weight(hrus_per_cell(g), 2)
    parts = [shard_case(g, r, 2, ranges=ranges) for r in range(2)]
    assert [p["_cells"] for p in parts] == ranges
    assert np.array_equal(np.concatenate([p["hrurec0"] for p in parts]), g["hrurec0"], equal_nan=True)
    assert np.array_equal(np.concatenate([p["forcing"] for p in parts], axis=1), g["forcing"])
    assert sum(p["meta"][1] for p in parts) == g["hrupar"].shape[0]


def test_shard_case_renumbers_cells_and_keeps_hru_order():
    g = dict(np.load(GOLDEN))
    parts = [shard_case(g, r, 2) for r in range(2)]
    assert sum(p["cellpar"].shape[0] for p in parts) == g["cellpar"].shape[0]
    assert np.array_equal(np.concatenate([p["hrurec0"] for p in parts]), g["hrurec0"], equal_nan=True)
    for p in parts:
        cells = p["hrupar"][:, 0]
        assert cells[0] == 0 and np.all(np.diff(cells) >= 0) and cells[-1] == p["cellpar"].shape[0] - 1


WORKER = r'''
import os, sys, subprocess, tempfile
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from vic_b200.casefile import read_case, write_case
from vic_b200.shard import shard_case, max_over_ranks, gather_cells, cell_ranges_by_weight, hrus_per_cell
root, golden, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = dict(np.load(golden))
keys = ("options_raw", "meta", "veglib", "cellpar", "hrupar", "hrurec0", "aggtype", "valid0", "dmy", "forcing", "dump_recs")
w = hrus_per_cell(g).astype(float)
w[0] += 7  # (an uneven weight so that the cut differs from the equal-count one)
ranges = cell_ranges_by_weight(w, world)
assert ranges != [(0, 2), (2, 4)] or g["cellpar"].shape[0] != 4
mine = shard_case(g, rank, world, ranges=ranges)
with tempfile.TemporaryDirectory() as d:
    write_case(os.path.join(d, "c.bin"), {k: mine[k] for k in keys})
    subprocess.run([os.path.join(root, "oracle", "_ref", "vicport"), os.path.join(d, "c.bin"), os.path.join(d, "r.bin")], check=True)
    res = read_case(os.path.join(d, "r.bin"))
t = max_over_ranks([1.0 + rank, 5.0 - rank])
full = gather_cells(res["out"], 1, g["cellpar"].shape[0], ranges=ranges)
if rank == 0:
    np.savez(out_path, out=full, tmax=np.array(t))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_ranks_gloo_sharded_run_equals_unsharded(tmp_path, vicport):
    g = dict(np.load(GOLDEN))
    from vic_b200.casefile import read_case, write_case
    write_case(str(tmp_path / "all.bin"), {k: g[k] for k in KEYS})
    subprocess.run([vicport, str(tmp_path / "all.bin"), str(tmp_path / "all_res.bin")], check=True)
    ref = read_case(str(tmp_path / "all_res.bin"))["out"]
    (tmp_path / "worker.py").write_text(WORKER)
    out = str(tmp_path / "gathered.npz")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(tmp_path / "worker.py"), ROOT, GOLDEN, out], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    z = np.load(out)
    assert z["out"].shape == ref.shape
    assert np.array_equal(z["out"], ref, equal_nan=True)
    assert list(z["tmax"]) == [2.0, 5.0]
