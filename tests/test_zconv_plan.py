"""Host logic of the windowed z-convolution (plan_pppm_zconv in csrc/pppm.cu through the host-only entry
conp_plan_zconv).  The kernel replaces the z-part of FFT -> greensfn -> FFT (pppm_conp.cpp:230-267) by a circular
convolution that, for a (kx,ky) column, only visits the input planes within the column's kernel radius of an
output plane.  A plane missing from a group's staging list would silently drop its charge from the potential, so
the plan is checked against a brute-force statement of "which planes matter" for slabs of 1, 2 and 8 ranks."""
import numpy as np
import pytest

from conp_b200 import abi

ZC_COLS = 8
SMEM_LIMIT = 44 * 1024


def ring_dist(a, b, nz):
    d = abs(a - b)
    return min(d, nz - d)


@pytest.mark.parametrize("nz,nzi,zin_lo,zout", [
    (1215, 414, 400, list(range(398, 406)) + list(range(808, 816))),   # cfg5-like slab: two thin groups of planes
    (288, 288, 0, [10, 11, 12, 13, 150, 151, 152]),                    # periodic z: every plane can hold charge
    (96, 40, 90, [88, 89, 90, 30, 31, 32]),                            # occupied range wraps around the ring
])
@pytest.mark.parametrize("nranks", [1, 2, 8])
def test_zconv_plan_stages_exactly_the_planes_within_reach(nz, nzi, zin_lo, zout, nranks):
    rng = np.random.default_rng(3)
    ncol = 203                                   # not a multiple of the group width
    krad = rng.integers(2, 60, ncol).astype(np.int32)
    krad[:3] = nz                                # k_xy ~ 0: the kernel reaches across the whole mesh
    krad[100:104] = [0, 1, nz // 2, nz // 2 - 1]
    per = -(-nzi // nranks)
    seen_cols_total = None
    for rank in range(nranks):
        zs_lo = min(rank * per, nzi)
        nzl = max(0, min(per, nzi - zs_lo))
        plan = abi.plan_zconv(ncol, nz, nzi, zs_lo, nzl, zin_lo, krad, zout, real_kernel=True)
        aout = [(z - zin_lo) % nz for z in zout]
        assert list(plan["aout"]) == aout
        # distance of every compact input plane to the nearest output plane, on the ring
        dist = [min(ring_dist(a, zi, nz) for a in aout) for zi in range(nzi)]
        cols = set(int(c) for c in plan["wide"])
        assert len(cols) == len(plan["wide"])
        for g in plan["narrow"]:
            members = [c for c in range(g["c0"], min(g["c0"] + ZC_COLS, ncol))]
            assert g["c0"] % ZC_COLS == 0 and not (cols & set(members))
            cols |= set(members)
            rb = max(int(krad[c]) for c in members)
            assert g["rblock"] == rb <= plan["rcap"] and 2 * rb + 1 < nz
            # the intervals are disjoint, ascending, compacted without gaps ...
            staged, row = [], 0
            for lo, hi, base in g["intervals"]:
                assert 0 <= lo < hi <= nzl and base == row and (not staged or lo > staged[-1])
                staged += list(range(lo, hi))
                row += hi - lo
            assert g["np"] == row <= plan["npcap"] and len(g["intervals"]) <= 8
            # ... and hold exactly the slab planes within the group's radius of an output plane
            want = [t for t in range(nzl) if dist[zs_lo + t] <= rb]
            assert staged == want
        assert cols == set(range(ncol))          # every column exactly once, narrow or wide
        # the narrow path fits the shared memory it was sized for (double2 planes x 8 columns + real kernel rows)
        assert 16 * plan["npcap"] * ZC_COLS + 8 * ZC_COLS * (2 * plan["rcap"] + 1) <= SMEM_LIMIT or not plan["narrow"]
        # whole-mesh columns can never be narrow
        assert {0, 1, 2} <= set(int(c) for c in plan["wide"]) or 2 * nz + 1 < nz
        seen_cols_total = cols
    assert seen_cols_total == set(range(ncol))


def test_zconv_plan_rejects_bad_arguments():
    with pytest.raises(RuntimeError):
        abi.plan_zconv(16, 64, 80, 0, 10, 0, np.ones(16, dtype=np.int32), [1, 2])      # nzi > nz
    with pytest.raises(RuntimeError):
        abi.plan_zconv(16, 64, 32, 30, 10, 0, np.ones(16, dtype=np.int32), [1, 2])     # slab beyond the occupied planes
