"""BASELINE configs[3] at full size (synthetic 10 000 electrode atoms / 100 000 charges, conp, PPPM,
slab) through the C ABI.  The oracle cannot finish the O(N^2 K) setup at this size in test time, so
the checks are the size-independent properties of SURVEY 8(c): electroneutrality, S.e = 0, linearity
of the charges in the applied voltage, conq inverting conp, bit-reproducibility of the matvec, and the
symmetric half-band product against a dense product with the same S."""
import numpy as np
import pytest

from cases import synthetic
from conp_b200.fix_conp import make_fix

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg4():
    lmp, arg = synthetic("cfg4", mode="pppm", accuracy=1e-4)
    fix = make_fix(lmp, arg)
    fix.setup()
    yield fix
    fix.close()


def test_uses_the_symmetric_kernel_and_is_neutral(cfg4):
    info = cfg4.ctx.info()
    assert info.n_ele == 10000 and info.n_elyte == 100000
    assert info.symmetric_matvec == 1 and 0.0 <= info.asymmetry < 1e-9
    q = cfg4.pre_force()
    assert abs(q.sum()) < 1e-12                       # total electroneutrality, north_star
    assert np.abs(q).max() > 1e-4                     # a real solve, not zeros
    # projected matrix annihilates the constant vector: S.e = 0 (fix_conp.cpp:1011-1020)
    Se = cfg4.ctx.matvec(np.ones(info.n_ele))
    S_scale = np.abs(cfg4.ctx.matvec(np.where(np.arange(info.n_ele) % 2 == 0, 1.0, -1.0))).max()
    assert np.abs(Se).max() < 1e-9 * max(S_scale, 1.0)


def test_charges_are_affine_in_the_voltage(cfg4):
    ctx, s = cfg4.ctx, cfg4.lmp.system
    x = s.x[cfg4.owned]
    q0, _ = ctx.pre_force(x, cfg4.kspace_mode, 0, 0.0)
    q1, _ = ctx.pre_force(x, cfg4.kspace_mode, 0, 1.0)
    q2, sc2 = ctx.pre_force(x, cfg4.kspace_mode, 0, 2.0)
    q0, q1, q2 = q0.copy(), q1.copy(), q2.copy()
    d = np.abs((q2 - q0) - 2.0 * (q1 - q0)).max()
    assert d <= 1e-9 * np.abs(q2).max() + 1e-12
    # conq fed with the right-electrode charge of the dV = 2 run returns dV = 2 (tests/cond/input:56-66)
    side = np.asarray(cfg4.side)
    qr = q2[side == -1].sum()
    _, dv = ctx.pre_force(x, cfg4.kspace_mode, 1, qr)
    assert abs(dv - 2.0) < 1e-8


def test_step_is_bit_reproducible_in_the_matvec(cfg4):
    n = cfg4.ctx.info().n_ele
    v = np.random.default_rng(5).standard_normal(n)
    a = cfg4.ctx.matvec(v)
    b = cfg4.ctx.matvec(v)
    assert np.array_equal(a, b)


def test_symmetric_product_equals_dense_product_of_the_same_matrix(cfg4):
    S = cfg4.ctx.get_matrix()
    n = S.shape[0]
    assert np.array_equal(S, S.T)                     # symmetrised at setup
    v = np.random.default_rng(6).standard_normal(n)
    ref = S @ v
    got = cfg4.ctx.matvec(v)
    assert np.abs(got - ref).max() <= 1e-13 * np.abs(S).sum(axis=1).max() * np.abs(v).max()


def _oracle_step_with_gpu_matrix(fix, name):
    """The CPU oracle's per-step path (b_cal + update_charge) on the `inv`-file setup
    (fix_conp.cpp:442-445) with the matrix this GPU built, on jittered positions."""
    import conp_oracle as O
    S = fix.ctx.get_matrix()
    lmp2, arg2 = synthetic(name, mode="pppm", accuracy=1e-4)
    ref = O.OracleFixConp(lmp2, arg2)
    ref.setup_preinverted(S)
    rng = np.random.default_rng(11)
    x = lmp2.system.x[ref.oth_idx] + rng.normal(0.0, 0.05, (len(ref.oth_idx), 3))
    lmp2.system.x[ref.oth_idx] = x
    qr = ref.pre_force().copy()
    fix.lmp.system.x[fix.owned] = x
    q = fix.pre_force()
    b, _ = fix.ctx.get_b()
    assert np.abs(b - ref.bbb_all).max() <= 5e-11 * np.abs(ref.bbb_all).max()
    assert np.abs(q - qr).max() <= 1e-9 * np.abs(qr).max() + 1e-12       # north_star: 1e-9 relative, 1e-12 e
    assert abs(q.sum()) < 1e-12
    assert abs(fix.scalar_output - ref.scalar_output) <= 1e-9 * abs(ref.scalar_output) + 1e-12
    # electrode density handed to the force pass (pppm_conp.cpp:385-426)
    rho_e = fix.ctx.get_density(1)
    assert np.abs(rho_e - ref.ele_density).max() <= 1e-9 * np.abs(ref.ele_density).max() + 1e-15


def test_cfg4_matches_oracle_with_gpu_built_matrix(cfg4):
    _oracle_step_with_gpu_matrix(cfg4, "cfg4")


def test_density_region_handoff_is_a_slice_of_the_brick_and_fast(cfg4):
    """Force-pass hand-off (PPPMCONP::make_rho, pppm_conp.cpp:428-450): a rank's sub-brick of elyte + electrode
    density comes from conp_get_density_region -- no allocation, no full-mesh transfer.  It must equal the same
    slice of the full brick, and at this size cost well under a step (device side, incl. the copy to the host)."""
    import torch
    ctx = cfg4.ctx
    cfg4.pre_force()
    nx, ny, nz = ctx.mesh
    full = ctx.get_density(2).reshape(nz, ny, nx)
    assert np.abs(full).max() > 0.0
    # the brick a rank of a 2 x 2 x 2 processor grid would own, and one that crosses the planes holding charge
    for lo, hi in (((0, 0, 0), (nx // 2 - 1, ny // 2 - 1, nz // 2 - 1)),
                   ((nx // 2, ny // 2, nz // 4), (nx - 1, ny - 1, nz // 4 + nz // 8))):
        reg = ctx.get_density_region(2, lo, hi)
        want = full[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
        assert reg.shape == want.shape and np.array_equal(reg, want)
    # the brick of one rank of a 2 x 2 x 4 processor grid, into page-locked host memory (3.5 MB)
    lo, hi = (0, 0, nz // 4), (nx // 2 - 1, ny // 2 - 1, nz // 4 + nz // 4 - 1)
    shape = (hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)
    out = torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
    ctx.get_density_region(2, lo, hi, out=out)  # warm: sizes the staging buffer
    ctx.timer_record(0)
    ctx.get_density_region(2, lo, hi, out=out)
    ctx.timer_record(1)
    ctx.sync()
    assert np.array_equal(out, full[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1])
    assert ctx.timer_elapsed_ms(0, 1) < 0.3


def test_cfg5_headline_size_matches_oracle():
    """BASELINE configs[4] (40 000 electrode atoms / 500 000 charges): the configuration bench.py is
    quoted on.  Full setup on the GPU (Gram + inversion, ~90 s), then one jittered update against the oracle."""
    lmp, arg = synthetic("cfg5", mode="pppm", accuracy=1e-4)
    fix = make_fix(lmp, arg)
    try:
        fix.setup()
        info = fix.ctx.info()
        assert info.n_ele == 40000 and info.n_elyte == 500000 and info.symmetric_matvec == 1
        _oracle_step_with_gpu_matrix(fix, "cfg5")
    finally:
        fix.close()
