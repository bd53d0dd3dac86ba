"""CPU: the tile plan of the owner-computes PPPM spread (conp_plan_spread, host-only entry point).
Every charge whose stencil (pppm_conp.cpp:146-148, 199-217) touches a tile must sit in one of the sort
cells the plan lists for that tile -- checked by brute force on random positions, for slab and periodic
boxes, odd and even orders, small meshes (whole-axis tiles with halo) and a rank's z-slab."""
import os

import numpy as np
import pytest

from conp_b200 import abi

OFFSET = 16384


@pytest.fixture(scope="module", autouse=True)
def _lib():
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return abi.load_library()


def _check(mesh, order, prd, periodic, slab, rc, nranks=1, rank=0, seed=0, n_atoms=4000):
    mesh = np.asarray(mesh)
    boxlo = np.array([-1.5, 2.0, -0.5 * prd[2]])
    prd = np.asarray(prd, dtype=float)
    prd_slab = prd * np.array([1.0, 1.0, slab])
    delinv = mesh / prd_slab
    nlower = -((order - 1) // 2)
    shift = OFFSET + 0.5 if order % 2 else float(OFFSET)
    # plane pruning as conp_pppm_setup does it
    if periodic[2]:
        zin_lo, nzi = 0, int(mesh[2])
    else:
        base_hi = int(prd[2] * delinv[2] + shift) - OFFSET
        zin_lo, hi = nlower - 2, base_hi + order // 2 + 2
        nzi = hi - zin_lo + 1
        if nzi >= mesh[2]:
            zin_lo, nzi = 0, int(mesh[2])
    zs_lo = (nzi * rank) // nranks
    zs_n = (nzi * (rank + 1)) // nranks - zs_lo
    plan = abi.plan_spread(mesh, order, shift, boxlo, prd, periodic, slab, rc, zin_lo, nzi, zs_lo, zs_n)
    nc = np.array([plan["ncx"], plan["ncy"], plan["ncz"]])
    rng = np.random.default_rng(seed)
    x = boxlo + rng.random((n_atoms, 3)) * prd
    x[: n_atoms // 20, 2] = boxlo[2] + prd[2] * rng.choice([0.0, 1.0 - 1e-12], n_atoms // 20)   # box faces
    x[n_atoms // 20: n_atoms // 10, 0] = boxlo[0] + prd[0] * rng.choice([0.0, 1.0 - 1e-12], n_atoms // 10 - n_atoms // 20)
    f = (x - boxlo) * delinv
    n0 = (f + shift).astype(np.int64) - OFFSET
    cell = np.clip(np.floor((x - boxlo) * (nc / prd)).astype(np.int64), 0, nc - 1)
    cid = (cell[:, 2] * nc[1] + cell[:, 1]) * nc[0] + cell[:, 0]
    T = np.array([plan["tx"], plan["ty"], plan["tz"]])
    NT = np.array([plan["ntx"], plan["nty"], plan["ntz"]])
    rs, runs = plan["run_start"], plan["runs"]
    covered = {}
    for t in range(plan["ntiles"]):
        cells = set()
        for c0, c1 in runs[rs[t]:rs[t + 1]]:
            cells.update(range(c0, c1))
        covered[t] = cells
    k = np.arange(order)
    # mesh index sets per axis: x, y global (mod mesh); z compact slab plane t = wrap(n+nlower+k-zin_lo) - zs_lo
    gx = (n0[:, 0, None] + nlower + k) % mesh[0]
    gy = (n0[:, 1, None] + nlower + k) % mesh[1]
    tz = (n0[:, 2, None] + nlower + k - zin_lo) % mesh[2] - zs_lo
    missing = 0
    for i in range(n_atoms):
        txs = set((gx[i] // T[0]).tolist())
        tys = set((gy[i] // T[1]).tolist())
        tzs = set(int(v) // T[2] for v in tz[i] if 0 <= v < zs_n)
        for iz in tzs:
            for iy in tys:
                for ix in txs:
                    if ix >= NT[0] or iy >= NT[1] or iz >= NT[2]:
                        continue
                    t = (iz * NT[1] + iy) * NT[0] + ix
                    if int(cid[i]) not in covered[t]:
                        missing += 1
    assert missing == 0, f"{missing} (charge, tile) overlaps not covered by the plan"
    # the plan is a real restriction, not "all cells for every tile"
    if plan["ntiles"] >= 8:
        mean_cells = np.mean([len(c) for c in covered.values()])
        assert mean_cells < 0.6 * nc.prod()
    return plan


@pytest.mark.parametrize("order", [2, 3, 4, 5, 7])
def test_slab_capacitor_mesh(order):
    p = _check((64, 108, 300), order, (61.5, 106.5, 100.0), (1, 1, 0), 3.0, 12.0, seed=order)
    assert p["halo_x"] == 0 and p["halo_y"] == 0 and p["halo_z"] == 0 and p["ntiles"] > 100


def test_small_periodic_mesh_uses_whole_axis_tiles_with_halo():
    p = _check((20, 24, 36), 5, (20.0, 24.0, 36.0), (1, 1, 1), 1.0, 4.0)
    assert p["ntx"] == 1 and p["halo_x"] == 4          # 20 <= 32 + 4: one tile spans x, stencils wrap inside it
    assert p["nty"] == 3 and p["halo_y"] == 0


def test_dilute_reference_meshes():
    _check((27, 24, 432), 5, (26.9, 23.3, 100.0), (1, 1, 0), 3.0, 4.0)
    _check((27, 24, 144), 5, (26.9, 23.3, 100.0), (1, 1, 1), 1.0, 4.0)


@pytest.mark.parametrize("rank", [0, 1, 2, 3])
def test_rank_slab_of_four(rank):
    _check((64, 108, 300), 5, (61.5, 106.5, 100.0), (1, 1, 0), 3.0, 12.0, nranks=4, rank=rank, seed=10 + rank)


# ---------------------------------------------------------------------------------------------------------------
# z-sweep spread (plan_pppm_sweep through conp_plan_sweep): columns of 8 rows x 32 mesh columns, swept upward in
# segments of planes with a window of `order` planes; charges are binned by (column, origin plane)
# ---------------------------------------------------------------------------------------------------------------
SW_FY, SW_FX, SW_MAXS = 8, 32, 64


@pytest.mark.parametrize("mesh,order,nzi", [((125, 216, 1215), 5, 414), ((64, 108, 1000), 5, 350),
                                            ((54, 48, 288), 5, 288), ((40, 48, 900), 4, 300),
                                            ((54, 48, 864), 7, 288)])
@pytest.mark.parametrize("nranks", [1, 2, 8])
def test_sweep_plan_covers_every_plane_once_and_bins_every_origin(mesh, order, nzi, nranks):
    from conp_b200 import abi
    nx, ny, nz = mesh
    per = -(-nzi // nranks)
    for rank in range(nranks):
        zs_lo = min(rank * per, nzi)
        zs_n = max(0, min(per, nzi - zs_lo))
        p = abi.plan_sweep(mesh, order, nzi, zs_lo, zs_n)
        if zs_n == 0:
            assert not p["usable"]
            continue
        assert p["usable"] == 1
        ncx, ncy = -(-nx // SW_FX), -(-ny // SW_FY)
        assert (p["ncolx"], p["ncoly"]) == (ncx, ncy) and p["nbins"] == ncx * ncy * p["npz"]
        # every (column, slab plane) belongs to exactly one work item; a segment plus its warm-up fits the kernel
        cover = np.zeros((ncx * ncy, zs_n), dtype=np.int32)
        for col, t0, t1 in p["items"]:
            assert 0 <= t0 < t1 <= zs_n and (t1 - t0) + (order - 1) <= SW_MAXS
            cover[col, t0:t1] += 1
        assert np.all(cover == 1)
        assert 1 <= p["grid"] <= min(len(p["items"]), 148 * 12)
        # every origin plane whose stencil (planes o .. o + order - 1) touches the slab has a bin
        wrap = nzi == nz
        assert p["wrap_z"] == int(wrap)
        for o in range(-(order - 1), nzi):
            planes = [(o + k) % nz if wrap else o + k for k in range(order)]
            if not wrap and (o < 0 or o + order > nzi):
                continue                                     # "Out of range atoms": rejected before binning
            if any(zs_lo <= z < zs_lo + zs_n for z in planes):
                t = (o % nz if wrap else o) - p["pz_lo"]
                if wrap:
                    t %= nz
                assert 0 <= t < p["npz"], (o, p["pz_lo"], p["npz"])


@pytest.mark.parametrize("mesh,order", [((36, 48, 100), 5),     # one column + a stencil would wrap inside it
                                        ((66, 48, 100), 5),     # last column (2 wide) narrower than order - 1
                                        ((64, 12, 100), 5)])    # too few rows
def test_sweep_plan_refuses_meshes_whose_stencils_would_wrap_inside_a_column(mesh, order):
    from conp_b200 import abi
    assert not abi.plan_sweep(mesh, order, 100, 0, 100)["usable"]
