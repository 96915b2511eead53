"""Domain partition for multi-GPU runs: cells are independent columns, so a domain is cut into contiguous ranges of cells, one
range per rank (one process per GPU), with no exchange on the data path; at most one gather of the outputs at the end
(BASELINE.json north_star).  The reference's own parallelism is the OpenMP loop over cells of vicNl.c:514-517; this is the same
decomposition across devices.

Everything here is host-side bookkeeping on the flat C-ABI arrays (include/vicgpu_fields.h) and works with any
torch.distributed backend: nccl on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

from .layout import TABLES


def cell_ranges(ncell: int, world: int):
    """[(c0, c1)] per rank: contiguous, sizes differ by at most one cell (integer bookkeeping, exact)"""
    base, extra = divmod(ncell, world)
    out, c = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((c, c + n))
        c += n
    return out


def cell_ranges_by_weight(weight, world: int):
    """[(c0, c1)] per rank: contiguous ranges whose summed weight -- HRUs per cell (the device's unit of work is the HRU), or a
    measured cost per cell -- is as even as a contiguous cut allows: rank r ends where the running sum is closest to (r + 1) / world of the
    total.  Every rank gets at least one cell when there are at least `world` cells.  Integer weights give an exact, reproducible cut."""
    w = np.asarray(weight, dtype=np.float64)
    ncell = w.shape[0]
    if world <= 1 or ncell == 0:
        return [(0, ncell)] + [(ncell, ncell)] * (world - 1)
    csum = np.cumsum(w)
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(csum, target, side="left"))  # first cell whose running sum reaches the target
        # cutting after cell k or before it: whichever running sum is closer to the target
        if k < ncell and k > 0 and abs(csum[k - 1] - target) <= abs(csum[k] - target):
            k -= 1
        c = k + 1
        c = max(c, cuts[-1] + 1) if ncell >= world else max(c, cuts[-1])  # no empty rank, ranges in order
        c = min(c, ncell - (world - r)) if ncell >= world else min(c, ncell)
        cuts.append(c)
    cuts.append(ncell)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def hrus_per_cell(case: dict):
    """number of HRUs of every cell of a case (hrupar rows are grouped by cell)"""
    cell_of_hru = case["hrupar"][:, TABLES["hpar"].index("HP_cell")].astype(np.int64)
    return np.bincount(cell_of_hru, minlength=case["cellpar"].shape[0])


def shard_case(case: dict, rank: int, world: int, ranges=None) -> dict:
    """the rank's part of a case (arrays named as in oracle/casefile.h): its cells, their HRUs renumbered from cell 0, its forcing.
    ranges: the partition to use (default: equal cell counts; cell_ranges_by_weight(hrus_per_cell(case), world) balances the HRUs)"""
    ncell = case["cellpar"].shape[0]
    c0, c1 = (ranges if ranges is not None else cell_ranges(ncell, world))[rank]
    hp = case["hrupar"]
    cell_of_hru = hp[:, TABLES["hpar"].index("HP_cell")].astype(np.int64)
    h0, h1 = np.searchsorted(cell_of_hru, [c0, c1])
    out = dict(case)
    out["cellpar"] = case["cellpar"][c0:c1]
    out["hrupar"] = hp[h0:h1].copy()
    out["hrupar"][:, TABLES["hpar"].index("HP_cell")] -= c0
    out["hrurec0"] = case["hrurec0"][h0:h1]
    out["forcing"] = case["forcing"][:, c0:c1]
    if "valid0" in case:
        out["valid0"] = case["valid0"][c0:c1]
    if "meta" in case:
        m = np.array(case["meta"]).copy()
        m[0], m[1] = c1 - c0, h1 - h0
        out["meta"] = m
    out["_cells"] = (c0, c1)
    out["_hrus"] = (int(h0), int(h1))
    return out


def max_over_ranks(values, device=None):
    """element-wise MAX of a list of floats over all ranks (the timing rule: a multi-GPU time is the slowest rank's)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_cells(local: np.ndarray, cell_axis: int, ncell_total: int, dst: int = 0, ranges=None):
    """the single end-of-run gather: rank `dst` gets the array of all cells (concatenated along cell_axis), the others None.
    ranges: the partition the ranks were cut with (default: cell_ranges(ncell_total, world))"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    if ranges is None:
        ranges = cell_ranges(ncell_total, world)
    x = torch.from_numpy(np.ascontiguousarray(np.moveaxis(local, cell_axis, 0)))
    if dist.get_backend() == "nccl":
        x = x.cuda()
    parts = None
    if rank == dst:
        parts = [torch.empty((c1 - c0,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device) for c0, c1 in ranges]
    # gather with unequal first dimensions: one send/recv pair per rank keeps it backend-neutral
    if rank == dst:
        parts[dst].copy_(x)
        for r in range(world):
            if r != dst:
                dist.recv(parts[r], src=r)
        full = torch.cat(parts, dim=0).cpu().numpy()
        return np.moveaxis(full, 0, cell_axis)
    dist.send(x, dst=dst)
    return None
