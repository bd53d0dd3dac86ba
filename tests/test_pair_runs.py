"""Host logic of the real-space pair kernels: the static candidate list (build_pair_runs in csrc/pair.cu, exposed
through the host-only entry conp_plan_pair_runs) must cover, for every fixed point, every sort cell -- seen
through every periodic image -- that holds a position within the search radius.  The list replaces the half
neighbour list walked by blist_coul_cal / alist_coul_cal (fix_conp.cpp:1225-1365); since round 2 each row of cells
is trimmed to the chord of the cut-off sphere, which is exactly what a brute-force distance test can check."""
import itertools

import numpy as np
import pytest

from conp_b200 import abi


def cell_bounds(lo, prd, nc, periodic, a, k):
    """[low, high) of cell k along axis a; edge cells of a non-periodic axis are unbounded outwards (they hold the
    clamped positions)."""
    w = prd[a] / nc[a]
    lo_k, hi_k = lo[a] + k * w, lo[a] + (k + 1) * w
    if not periodic[a]:
        if k == 0:
            lo_k = -np.inf
        if k == nc[a] - 1:
            hi_k = np.inf
    return lo_k, hi_k


def brute_force(lo, prd, periodic, rc, nc, p):
    """Set of (cell, sx, sy, sz) whose box, shifted by the image, comes closer than rc to point p."""
    need = set()
    smax = [int(np.ceil(rc / prd[a])) + 1 if periodic[a] else 0 for a in range(3)]
    for sz, sy, sx in itertools.product(*[range(-smax[a], smax[a] + 1) for a in (2, 1, 0)]):
        sh = (sx * prd[0], sy * prd[1], sz * prd[2])
        # per axis: distance from p to the (shifted) cell interval, for every cell index
        d2 = []
        for a in range(3):
            row = np.empty(nc[a])
            for k in range(nc[a]):
                l, h = cell_bounds(lo, prd, nc, periodic, a, k)
                l, h = l + sh[a], h + sh[a]
                row[k] = 0.0 if l <= p[a] < h else min(abs(p[a] - l), abs(p[a] - h)) ** 2
            d2.append(row)
        tot = d2[2][:, None, None] + d2[1][None, :, None] + d2[0][None, None, :]
        for cz, cy, cx in zip(*np.nonzero(tot < rc * rc * (1 - 1e-9))):
            need.add(((cz * nc[1] + cy) * nc[0] + cx, sx, sy, sz))
    return need


@pytest.mark.parametrize("periodic,rc,prd", [
    ((1, 1, 0), 12.0, (34.0, 29.5, 80.0)),     # slab geometry of the bench workloads
    ((1, 1, 1), 9.0, (21.0, 25.0, 30.0)),      # fully periodic (fix conp with p p p)
    ((1, 1, 0), 12.0, (17.0, 19.0, 60.0)),     # cut-off beyond half the box: several images of the same cell
])
def test_pair_runs_cover_every_cell_within_the_cutoff(periodic, rc, prd):
    rng = np.random.default_rng(7)
    lo = np.array([-3.0, 2.0, -40.0])
    prd = np.array(prd)
    pts = lo + rng.random((24, 3)) * prd
    pts[0] = lo + 1e-9                                    # corners and faces
    pts[1] = lo + prd * (1 - 1e-12)
    if not periodic[2]:
        pts[2, 2] = lo[2] - 5.0                           # outside a non-periodic axis: binned into the edge cell
    nc, run_start, runs = abi.plan_pair_runs(lo, prd, periodic, rc, pts)
    assert run_start[0] == 0 and run_start[-1] == len(runs) and np.all(np.diff(run_start) >= 0)
    ncells = int(nc[0]) * int(nc[1]) * int(nc[2])
    trimmed = 0
    for i, p in enumerate(pts):
        mine = runs[run_start[i]:run_start[i + 1]]
        assert np.all(mine[:, 0] < mine[:, 1]) and np.all(mine[:, 0] >= 0) and np.all(mine[:, 1] <= ncells)
        # (a run may span several x-rows of cells: rows that are reached over their whole length are merged)
        have = set()
        for c0, c1, sx, sy, sz in mine:
            for c in range(c0, c1):
                key = (c, sx, sy, sz)
                assert key not in have, "a (cell, image) pair listed twice would double-count its charges"
                have.add(key)
        need = brute_force(lo, prd, periodic, rc, nc, p)
        missing = need - have
        assert not missing, f"point {i}: {len(missing)} reachable (cell, image) pairs are not in its runs"
        trimmed += len(have) - len(need)
    # the list is tight: on average less than one surplus cell per run end
    assert trimmed <= 2 * len(runs)
