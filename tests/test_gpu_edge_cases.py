"""Edge cases of the hot path, CUDA (through the C ABI) vs the oracle:
empty / uncharged electrolyte, charges that left the periodic box between
reneighbourings, cut-off larger than half the box (several periodic images
per pair), minimal electrodes, other PPPM stencil orders, ragged row counts."""
import numpy as np
import pytest

import conp_oracle as O
from cases import dilute, synthetic
from conp_b200 import MockLammps, load_reference_case
from conp_b200.fix_conp import make_fix
from conp_b200.system import System

pytestmark = pytest.mark.gpu


def both(lmp_factory, arg_extra=()):
    lmp, arg = lmp_factory()
    lmp2, arg2 = lmp_factory()
    fix = make_fix(lmp, list(arg) + list(arg_extra))
    ref = O.OracleFixConp(lmp2, list(arg2) + list(arg_extra))
    fix.setup()
    ref.setup()
    return fix, ref, fix.pre_force(), ref.pre_force()


def close(q, qr):
    assert np.abs(q - qr).max() <= 1e-9 * np.abs(qr).max() + 1e-12
    assert abs(q.sum()) < 1e-12


@pytest.mark.parametrize("pppm", [False, True])
def test_no_charged_electrolyte(pppm):
    """b == 0: the charges are dV * S.d exactly (fix_conp.cpp:1156)."""
    def case():
        lmp, arg = dilute(2, pppm=pppm)
        ele = np.isin(lmp.system.mol, (81, 82))
        lmp.system.q[~ele] = 0.0
        return lmp, arg
    fix, ref, q, qr = both(case)
    close(q, qr)
    b, bk = fix.ctx.get_b()
    assert np.abs(b).max() == 0.0 and np.abs(bk).max() == 0.0
    fix.close()


def test_electrode_only_system():
    def case():
        s = load_reference_case("dilute")
        keep = np.isin(s.mol, (81, 82))
        s2 = System(s.boxlo, s.boxhi, s.id[keep], s.mol[keep], s.type[keep], s.q[keep], s.x[keep], s.ntypes)
        lmp = MockLammps(s2, "p p p")
        lmp.pair_style_coul_long(4.0)
        lmp.kspace("pppm", 1e-6, 0.77236341)
        lmp.group_molecule("eleleft", 81)
        lmp.group_molecule("eleright", 82)
        return lmp, "e eleleft conp 1 eleright 1.979 1.0 log ffield".split()
    fix, ref, q, qr = both(case)
    close(q, qr)
    fix.close()


@pytest.mark.parametrize("pppm", [False, True])
def test_charges_outside_the_periodic_box(pppm):
    """LAMMPS remaps atoms only at reneighbouring; images must give identical results."""
    def case(shifted):
        lmp, arg = dilute(2, pppm=pppm)
        if shifted:
            s = lmp.system
            oth = np.nonzero(~np.isin(s.mol, (81, 82)))[0]
            rng = np.random.default_rng(1)
            k = rng.integers(-1, 2, (len(oth), 3))
            s.x[oth] += k * s.prd[None, :]
        return lmp, arg
    fix, ref, q, qr = both(lambda: case(True))
    close(q, qr)
    _, _, q0, _ = both(lambda: case(False))
    assert np.abs(q - q0).max() <= 2e-9 * np.abs(q0).max() + 1e-12
    fix.close()


def test_cutoff_beyond_half_box_sums_all_images():
    """cut 9 A in the 9.8 x 8.5 A dilute cell: every pair has several images (and self images in A)."""
    def case():
        lmp, arg = dilute(5)
        lmp.pair_style_coul_long(9.0)
        return lmp, arg
    fix, ref, q, qr = both(case)
    assert np.abs(fix.ctx.get_matrix() - ref.S).max() <= 1e-10 * np.abs(ref.S).max()
    b, _ = fix.ctx.get_b()
    assert np.abs(b - ref.bbb_all).max() <= 5e-12 * max(1.0, np.abs(ref.bbb_all).max())
    close(q, qr)
    fix.close()


def test_two_atom_electrodes():
    def case():
        lmp, arg = dilute(2)
        s = lmp.system
        left = np.nonzero(s.mol == 81)[0]
        right = np.nonzero(s.mol == 82)[0]
        s.mol[left[1:]] = 999
        s.mol[right[1:]] = 998
        lmp.group_molecule("eleleft", 81)
        lmp.group_molecule("eleright", 82)
        return lmp, "e eleleft conp 1 eleright 1.979 1.0 log_conp ffield".split()
    fix, ref, q, qr = both(case)
    assert fix.N == 2
    close(q, qr)
    assert abs(q[0] + q[1]) < 1e-14
    fix.close()


@pytest.mark.parametrize("order", [3, 4, 7])
def test_other_pppm_orders(order):
    def case():
        lmp, arg = dilute(2, pppm=True)
        lmp.kspace("pppm/conp", 1e-6, 0.77236341, mesh=(27, 24, 144), order=order)
        return lmp, arg
    fix, ref, q, qr = both(case)
    b, bk = fix.ctx.get_b()
    assert np.abs(bk - ref.b_kspace).max() <= 5e-12 * max(1.0, np.abs(ref.b_kspace).max())
    close(q, qr)
    fix.close()


def test_gemv_row_tails_many_sizes():
    """S.b through the TMA GEMV for row/column counts that are not multiples of the tile
    (ragged strips, partial column chunks): totsetq = sum_left (S.d) vs numpy."""
    from conp_b200 import abi
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 15, 17, 148, 149, 511, 513, 1031, 2500):
        ctx = abi.Context()
        ctx.set_cell([0, 0, -50], [60, 100, 100], [1, 1, 0], 1, 3.0, 0)
        ctx.set_ewald(0.26, 1e-2, 1000.0, 1000)
        ctx.set_pair(0, 1.979, 12.0, 1, np.full((2, 2), 144.0))
        side = np.where(np.arange(n) % 2 == 0, 1, -1)
        ctx.set_electrodes(np.arange(1, n + 1), np.ones(n), side, rng.uniform(0, 50, (n, 3)))
        S = rng.standard_normal((n, n))
        ctx.load_matrix(S, True)
        tot = ctx.set_unit_voltage(0.0694)
        ref = (S @ (-0.5 * 0.0694 * side))[side == 1].sum()
        assert abs(tot - ref) <= 1e-12 * max(1.0, np.abs(S).sum() * 0.0694)
        ctx.close()


def _matvec_ctx(n, rng):
    from conp_b200 import abi
    ctx = abi.Context()
    ctx.set_cell([0, 0, -50], [60, 100, 100], [1, 1, 0], 1, 3.0, 0)
    ctx.set_ewald(0.26, 1e-2, 1000.0, 1000)
    ctx.set_pair(0, 1.979, 12.0, 1, np.full((2, 2), 144.0))
    side = np.where(np.arange(n) % 2 == 0, 1, -1)
    ctx.set_electrodes(np.arange(1, n + 1), np.ones(n), side, rng.uniform(0, 50, (n, 3)))
    return ctx


def test_symmetric_matvec_many_sizes():
    """S.v through the half-band symmetric kernel (symv_tma_kernel) for sizes that exercise even/odd N
    (the distance-N/2 tie rule), band wrap-around, ragged strips and partial column chunks; a
    symmetric matrix must take the symmetric path, and give S.v to rounding."""
    rng = np.random.default_rng(1)
    for n in (64, 65, 191, 192, 513, 1031, 1200, 2500, 4099):
        ctx = _matvec_ctx(n, rng)
        A = rng.standard_normal((n, n))
        S = A + A.T
        ctx.load_matrix(S, True)
        info = ctx.info()
        assert info.symmetric_matvec == 1 and info.asymmetry == 0.0, n
        for _ in range(2):
            v = rng.standard_normal(n)
            out = ctx.matvec(v)
            ref = S @ v
            assert np.abs(out - ref).max() <= 1e-13 * np.abs(S).sum(axis=1).max() * np.abs(v).max(), n
        # same answer, bit for bit, on a second call (fixed summation order)
        assert np.array_equal(ctx.matvec(v), out)
        ctx.close()


def test_asymmetric_matrix_takes_the_general_kernel():
    """A loaded matrix that is not symmetric to the last bit is used exactly as given (GEMV)."""
    rng = np.random.default_rng(2)
    n = 300
    ctx = _matvec_ctx(n, rng)
    A = rng.standard_normal((n, n))
    S = A + A.T
    S[5, 17] += 1e-13
    ctx.load_matrix(S, True)
    info = ctx.info()
    assert info.symmetric_matvec == 0 and info.asymmetry > 0
    v = rng.standard_normal(n)
    assert np.abs(ctx.matvec(v) - S @ v).max() <= 1e-13 * np.abs(S).sum(axis=1).max() * np.abs(v).max()
    ctx.close()


def test_symmetric_and_general_paths_agree(monkeypatch):
    """The same deck solved with the symmetric product and with CONP_NO_SYMV=1 (full GEMV)."""
    q = {}
    for flag in ("0", "1"):
        if flag == "1":
            monkeypatch.setenv("CONP_NO_SYMV", "1")
        lmp, arg = synthetic("small", mode="pppm")
        fix = make_fix(lmp, arg)
        fix.setup()
        q[flag] = fix.pre_force().copy()
        assert fix.ctx.info().symmetric_matvec == (1 if flag == "0" else 0)
        fix.close()
    assert np.abs(q["0"] - q["1"]).max() <= 1e-9 * np.abs(q["1"]).max() + 1e-12


def test_unsymmetric_influence_function_takes_the_complex_kernel():
    """greensfn is an input at the ABI.  LAMMPS' table is even in every k component, which makes the
    z-convolution kernel K(kx,ky;d) real; a table that is not even in kz (here: scaled by
    1 + 0.25 sin(2 pi kz/nz)) gives a complex K and exercises zconv_kernel<false>.  The oracle
    consumes the same table with full complex 3-D FFTs (pppm_conp.cpp:235-266)."""
    def factory():
        lmp, arg = synthetic("tiny", h=1.5)
        orig = lmp.pppm_tables

        def skewed():
            t = orig()
            nx, ny, nz = t.mesh
            kz = np.arange(nz)
            f = 1.0 + 0.25 * np.sin(2.0 * np.pi * kz / nz)
            t.greensfn = (t.greensfn.reshape(nz, ny, nx) * f[:, None, None]).reshape(-1).copy()
            return t
        lmp.pppm_tables = skewed
        return lmp, arg
    fix, ref, q, qr = both(factory)
    b, bk = fix.ctx.get_b()
    assert np.abs(bk - ref.b_kspace).max() <= 5e-12 * max(1.0, np.abs(ref.b_kspace).max())
    close(q, qr)
    fix.close()


@pytest.mark.parametrize("kernel", ["atomic", "atomic_sorted", "smem", "mma", "sweep"])
@pytest.mark.parametrize("which", ["small_slab", "dilute_periodic", "dilute_slab_order7", "small_order4"])
def test_all_spread_kernels_give_the_oracle_density(kernel, which, monkeypatch):
    """elyte_make_rho (pppm_conp.cpp:172-228) through each of the spread kernels -- red.global (reading the
    charges as packed, the default, or cell-sorted), shared-memory tiles (whole-axis tiles with halo on the small
    meshes), FP64 tensor-core tiles, FP64 tensor-core z-sweep -- against the oracle's brick, incl. even / high
    orders, periodic z and charges straddling every tile border."""
    monkeypatch.setenv("CONP_SPREAD", kernel.split("_")[0])
    if kernel == "atomic_sorted":
        monkeypatch.setenv("CONP_SPREAD_UNSORTED", "0")

    def case():
        if which == "small_slab":
            return synthetic("small", h=0.5, accuracy=1e-4)          # 40 x 48 x ~1000 mesh: tensor-core shape
        if which == "small_order4":
            lmp, arg = synthetic("small", h=0.5, accuracy=1e-4)
            lmp.kspace("pppm/conp", 1e-4, 0.26, slab=3.0, mesh=lmp.mesh, order=4)
            return lmp, arg
        if which == "dilute_periodic":
            lmp, arg = dilute(2, pppm=True)
            lmp.kspace("pppm/conp", 1e-6, 0.77236341, mesh=(54, 48, 288), order=5)
            return lmp, arg
        lmp, arg = dilute(0, pppm=True)
        lmp.kspace("pppm/conp", 1e-6, 0.77236341, slab=3.0, mesh=(54, 48, 864), order=7)
        return lmp, arg
    fix, ref, q, qr = both(case)
    rho = fix.ctx.get_density(0)
    assert np.abs(rho - ref.elyte_density).max() <= 1e-12 * np.abs(ref.elyte_density).max()
    b, bk = fix.ctx.get_b()
    assert np.abs(bk - ref.b_kspace).max() <= 5e-12 * max(1.0, np.abs(ref.b_kspace).max())
    close(q, qr)
    # the owner-computes kernels are deterministic for a given order of the sorted charges; all kernels must
    # reproduce the brick to rounding on a second solve
    fix.pre_force()
    assert np.abs(fix.ctx.get_density(0) - rho).max() <= 1e-15 * np.abs(rho).max()
    fix.close()
