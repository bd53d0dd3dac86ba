"""GPU parity tests proper: the CUDA path, called through the C ABI
(libconp_b200.so), against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): electrode charges 1e-9 relative and
1e-12 e absolute, total electroneutrality 1e-12 e, energies/forces 1e-8
relative.  The matrix and b vector are checked tighter because q = S.b
amplifies their error by ~cond(A).
"""
import numpy as np
import pytest

import conp_oracle as O
from cases import dilute, il, synthetic
from conp_b200 import abi
from conp_b200.fix_conp import make_fix

pytestmark = pytest.mark.gpu

Q_RTOL, Q_ATOL = 1e-9, 1e-12


def q_close(q, qr):
    scale = np.abs(qr).max()
    err = np.abs(q - qr).max()
    assert err <= Q_RTOL * scale + Q_ATOL, f"max|dq| = {err:.3e} vs |q|max {scale:.3e}"
    assert abs(q.sum() - qr.sum()) < 1e-12


def run_pair(case, *a, mod=None, **kw):
    lmp, arg = case(*a, **kw)
    lmp2, arg2 = case(*a, **kw)
    fix = make_fix(lmp, arg)
    ref = O.OracleFixConp(lmp2, arg2)
    for m in mod or []:
        fix.modify_param(m.split())
        ref.modify_param(m.split())
    fix.setup_post_neighbor()
    fix.setup_pre_force(first_solve=False)
    ref.setup()
    q = fix.pre_force()
    qr = ref.pre_force()
    return fix, ref, q, qr


def check_all(fix, ref, q, qr, btol=5e-12):
    b, bk = fix.ctx.get_b()
    bs = np.abs(ref.bbb_all).max()
    assert np.abs(bk - ref.b_kspace).max() <= btol * max(bs, 1.0), np.abs(bk - ref.b_kspace).max()
    assert np.abs(b - ref.bbb_all).max() <= btol * max(bs, 1.0), np.abs(b - ref.bbb_all).max()
    S = fix.ctx.get_matrix()
    assert np.abs(S - ref.S).max() <= 1e-10 * np.abs(ref.S).max()
    q_close(q, qr)
    assert abs(fix.scalar_output - ref.scalar_output) <= 1e-9 * abs(ref.scalar_output) + 1e-12
    assert abs(fix.totsetq - ref.totsetq) <= 1e-10 * abs(ref.totsetq)


def test_a_matrix_dilute_ffield():
    """FP64 tensor-core Gram + real-space + self vs the oracle's A."""
    lmp, arg = dilute(2)
    fix = make_fix(lmp, arg)
    fix.setup_post_neighbor()
    fix.ctx.build_A()
    A = fix.ctx.get_matrix()
    ref = O.OracleFixConp(*dilute(2))
    ref.setup()
    info = fix.ctx.info()
    assert (info.kcount, info.kcount_flat, info.kcount_expand) == (8095, 147, 3974)
    assert (info.kxmax, info.kymax, info.kzmax) == (8, 7, 58)
    assert np.abs(A - ref.A).max() <= 1e-12 * np.abs(ref.A).max()
    assert np.abs(A - A.T).max() <= 1e-13 * np.abs(A).max()
    fix.close()


def test_golden_persist_log_on_gpu():
    """tests/dilute/persist.log:143 through the CUDA path."""
    fix, ref, q, qr = run_pair(dilute, 2)
    qleft = q[fix.side == 1].sum()
    assert abs(qleft - 0.044057154) < 5e-10
    assert abs(fix.ee - 0.1702472657) < 1e-9 and abs(fix.dd - 0.00731711766) < 1e-10
    check_all(fix, ref, q, qr)
    fix.close()


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5])
def test_dilute_trials_ewald(n):
    """tests/dilute/input trial matrix (slab, etypes, noslab zneutr sym/anti, ffield)."""
    fix, ref, q, qr = run_pair(dilute, n)
    check_all(fix, ref, q, qr)
    fix.close()


@pytest.mark.parametrize("n", [0, 2, 3])
def test_dilute_trials_pppm(n):
    fix, ref, q, qr = run_pair(dilute, n, pppm=True)
    check_all(fix, ref, q, qr)
    u = fix.ctx.get_potential_brick()
    assert np.abs(u - ref.u_brick).max() <= 1e-11 * np.abs(ref.u_brick).max()
    rho = fix.ctx.get_density(0)
    assert np.abs(rho - ref.elyte_density).max() <= 1e-12 * np.abs(ref.elyte_density).max()
    rho_e = fix.ctx.get_density(1)
    assert np.abs(rho_e - ref.ele_density).max() <= 1e-9 * np.abs(ref.ele_density).max() + 1e-15
    tot = fix.ctx.get_density(2)
    assert np.abs(tot - (rho + rho_e)).max() == 0.0
    fix.close()


def test_conq_and_cond_variants():
    """fix conq (fix_conq.cpp:74-80) and fix cond (fix_cond.cpp:57-126) epilogues."""
    for style, value in (("conq", 0.03), ("cond", 0.02)):
        def case():
            lmp, arg = dilute(2)
            arg[2] = style
            arg[6] = repr(value)
            return lmp, arg
        fix, ref, q, qr = run_pair(case)
        check_all(fix, ref, q, qr)
        if style == "conq":
            assert abs(q[fix.side == -1].sum() - value) < 1e-12
        fix.close()


def test_ehgo_tables_and_qinit():
    """EHGO pair mode with non-trivial kappa/u0 (fix_conp.cpp:1517-1573) + qinit."""
    def case():
        lmp, arg = dilute(2)
        return lmp, arg + ["ehgo", "qinit"]
    mod = ["ehgo kappa 0.7", "ehgo coeff 3 1.979 2.1", "ehgo coeff 1*2 1.2 0.9", "ehgo coeff 4 0.8 auto"]
    fix, ref, q, qr = run_pair(case, mod=mod)
    check_all(fix, ref, q, qr)
    fix.close()


def test_one_electrode_and_nonneutral():
    def case1():
        lmp, arg = dilute(2)
        lmp.group_molecule("eleleft", 81, 82)
        arg[4] = "eleleft"
        return lmp, arg
    fix, ref, q, qr = run_pair(case1)
    check_all(fix, ref, q, qr)
    fix.close()

    def case2():
        lmp, arg = dilute(0)
        return lmp, arg + ["nonneutral"]
    fix, ref, q, qr = run_pair(case2)
    b, _ = fix.ctx.get_b()
    assert np.abs(fix.ctx.get_matrix() - ref.S).max() <= 1e-10 * np.abs(ref.S).max()
    assert np.abs(q - qr).max() <= 1e-9 * np.abs(qr).max() + 1e-12
    fix.close()


@pytest.mark.parametrize("n,twolayer", [(1, False), (2, True), (3, True)])
def test_il_configs(n, twolayer):
    """BASELINE configs 1-3: il_onelayer conp Ewald slab (+etypes); il_twolayer
    with pppm/conp (conq); ffield on il_twolayer."""
    fix, ref, q, qr = run_pair(il, n, twolayer=twolayer, value=(0.35 if n == 2 else None))
    assert fix.N == (1664 if twolayer else 832)
    check_all(fix, ref, q, qr)
    fix.close()


def test_post_force_energy_and_forces():
    """force_cal (fix_conp.cpp:1163-1201, 1368-1444) incl. the reference's guard
    eta^2 r^2 < 5.8; a compressed copy of the dilute cell puts atoms in range."""
    def case():
        lmp, arg = dilute(2)
        s = lmp.system
        rng = np.random.default_rng(5)
        ele = np.isin(s.mol, (81, 82))
        idx = np.nonzero(~ele)[0][:40]
        tgt = np.nonzero(ele)[0][:40]
        s.x[idx] = s.x[tgt] + rng.normal(0, 0.5, (40, 3)) + np.array([0, 0, 0.4])
        return lmp, arg
    fix, ref, q, qr = run_pair(case)
    q_close(q, qr)
    f, ecoul, eself, vir = fix.post_force()
    fr, ecr, esr, vr = ref.post_force()
    assert np.abs(fr).max() > 0
    own = fix.owned
    fr_own = np.zeros((fix.lmp.system.natoms, 3))
    fr_own[ref.oth_idx] = fr
    assert np.abs(f - fr_own[own]).max() <= 1e-8 * np.abs(fr).max()
    assert abs(ecoul - ecr) <= 1e-8 * abs(ecr)
    assert abs(eself - esr) <= 1e-8 * abs(esr)
    assert np.abs(vir - vr).max() <= 1e-8 * np.abs(vr).max()
    fix.close()


@pytest.mark.parametrize("name,mode,ff", [("tiny", "pppm", "slab"), ("small", "pppm", "slab"),
                                         ("small", "ewald", "slab"), ("small", "pppm", "ffield"),
                                         ("medium", "pppm", "slab")])
def test_synthetic_capacitor(name, mode, ff):
    """SURVEY 8d recipe at oracle-tractable sizes (config 4/5 geometry family)."""
    fix, ref, q, qr = run_pair(synthetic, name, mode=mode, ff=ff, h=1.25, accuracy=1e-4)
    check_all(fix, ref, q, qr)
    fix.close()


def test_matrix_file_roundtrip(tmp_path, monkeypatch):
    """matout -> org / inv (fix_conp.cpp:721-773, 833-849, 960-977)."""
    monkeypatch.chdir(tmp_path)
    lmp, arg = dilute(2)
    fix = make_fix(lmp, arg + ["matout"])
    fix.setup()
    q0 = fix.pre_force()
    fix.close()
    for kw, fname, tol in (("org", "amatrix", 1e-7), ("inv", "inv_a_matrix", 1e-6)):
        lmp, arg = dilute(2)
        f2 = make_fix(lmp, arg + [kw, fname])
        f2.setup()
        q = f2.pre_force()
        assert np.abs(q - q0).max() <= tol * np.abs(q0).max()  # %20.12f / %20.10f text precision
        f2.close()


def test_one_electrode_inverse_file_is_projected_after_setq(tmp_path, monkeypatch):
    """one_electrode + `inv <file>`: the file holds the UNprojected inverse (fix_conp.cpp:958-977) and the
    projection runs after get_setq (:1115), also for a matrix that was read in."""
    monkeypatch.chdir(tmp_path)

    def case(extra):
        lmp, arg = dilute(2)
        lmp.group_molecule("eleleft", 81, 82)
        arg[4] = "eleleft"
        return lmp, arg + extra
    lmp, arg = case(["matout"])
    f0 = make_fix(lmp, arg)
    f0.setup()
    q0 = f0.pre_force()
    f0.close()
    lmp, arg = case(["inv", "inv_a_matrix"])
    lmp2, arg2 = case(["inv", "inv_a_matrix"])
    fix = make_fix(lmp, arg)
    ref = O.OracleFixConp(lmp2, arg2)
    fix.setup()
    ref.setup()
    q, qr = fix.pre_force(), ref.pre_force()
    assert np.abs(fix.ctx.get_matrix() - ref.S).max() <= 1e-10 * np.abs(ref.S).max()
    q_close(q, qr)
    assert np.abs(q - q0).max() <= 1e-6 * np.abs(q0).max()        # %20.10f text precision
    fix.close()


def test_step_is_repeatable_and_tracks_moving_atoms():
    lmp, arg = synthetic("small", h=1.25, accuracy=1e-4)
    lmp2, arg2 = synthetic("small", h=1.25, accuracy=1e-4)
    fix = make_fix(lmp, arg)
    ref = O.OracleFixConp(lmp2, arg2)
    fix.setup()
    ref.setup()
    rng = np.random.default_rng(3)
    for step in range(3):
        d = rng.normal(0, 0.05, (len(fix.owned), 3))
        lmp.system.x[fix.owned] += d
        lmp2.system.x[fix.owned] += d
        q_close(fix.pre_force(), ref.pre_force())
    qa = fix.pre_force()
    qb = fix.pre_force()
    assert np.abs(qa - qb).max() <= 1e-12 * np.abs(qa).max()  # atomics reorder sums
    fix.close()


def test_error_paths():
    lmp, arg = dilute(2)
    fix = make_fix(lmp, arg)
    with pytest.raises(abi.ConpError):
        fix.ctx.pre_force(np.zeros((1, 3)), 0, 0, 1.0)  # setup incomplete
    fix.close()
    lmp, arg = dilute(2, pppm=True)
    fix = make_fix(lmp, arg)
    fix.setup()
    lmp.system.x[fix.owned[0]] = np.nan
    with pytest.raises(abi.ConpError) as e:
        fix.pre_force()
    assert "Out of range atoms" in str(e.value)
    fix.close()
    # singular matrix -> "Inversion failed!"
    lmp, arg = dilute(2)
    fix = make_fix(lmp, arg)
    fix.setup_post_neighbor()
    fix.ctx.load_matrix(np.zeros((fix.N, fix.N)), False)
    with pytest.raises(abi.ConpError) as e:
        fix.ctx.invert_project()
    assert "Inversion failed" in str(e.value)
    fix.close()


@pytest.mark.parametrize("case", ["dilute0", "dilute2", "dilute4", "il1", "small"])
def test_ewald_gemm_form(case, monkeypatch):
    """Ewald mode with the structure factors and the b extraction computed as matrix products
    (ewald_gemm_sfac / ewald_gemm_bextract; automatic only for large M*K, forced here) against the
    oracle's loops (km_ewald.cpp:668-825): slab, ffield, noslab zneutr, il_onelayer, synthetic."""
    monkeypatch.setenv("CONP_EWALD_GEMM", "1")
    if case.startswith("dilute"):
        fix, ref, q, qr = run_pair(dilute, int(case[-1]))
    elif case == "il1":
        fix, ref, q, qr = run_pair(il, 1, twolayer=False)
    else:
        fix, ref, q, qr = run_pair(synthetic, "small", mode="ewald", ff="slab", h=1.25, accuracy=1e-4)
    check_all(fix, ref, q, qr)
    # second call runs eagerly again or is captured, third replays the CUDA graph (cuBLAS nodes inside)
    fix.pre_force()
    q_close(fix.pre_force(), qr)
    fix.close()


def test_compute_potential_atom_on_the_electrodes():
    """`compute potential/atom` (compute_potential_atom.cpp:120-182, pppm_conp.cpp:452-488) for the electrode
    atoms: GPU vs the oracle's restatement (brute-force image sum + full-mesh Poisson solve), and the
    physics it exists for: after the solve every atom of an electrode sits at the same potential and the
    two electrodes differ by the applied dV (to the accuracy of PPPM vs the Ewald-built A matrix)."""
    lmp, arg = dilute(0, pppm=True)      # slab, conp dV = 1.0 V
    lmp2, arg2 = dilute(0, pppm=True)
    fix = make_fix(lmp, arg)
    ref = O.OracleFixConp(lmp2, arg2)
    fix.setup()
    ref.setup()
    q, qr = fix.pre_force(), ref.pre_force()
    q_close(q, qr)
    eta = fix.args.eta
    for kw in (dict(), dict(pair=False), dict(kspace=False), dict(qsum=False)):
        phi = fix.compute_potential_atom(eta, **kw)
        phir = ref.potential_atom(eta, **kw)
        assert np.abs(phi - phir).max() <= 1e-8 * np.abs(phir).max() + 1e-10, kw
    x = lmp.system.x[fix.ele_idx]
    u = fix.ctx.mesh_potential(x + 0.3)                       # arbitrary positions, not only electrode sites
    assert np.abs(u - ref.mesh_potential(x + 0.3)).max() <= 1e-9 * np.abs(u).max()
    phi = fix.compute_potential_atom(eta)
    left, right = phi[fix.side == 1], phi[fix.side == -1]
    assert np.ptp(left) < 2e-3 and np.ptp(right) < 2e-3       # equipotential electrodes (PPPM-level agreement)
    assert abs((right.mean() - left.mean()) - 1.0) < 2e-3     # ... dV apart (group2 is the positive side)
    fix.close()
    # Ewald mode: the reference's compute refuses without a pppm/conp style
    lmp, arg = dilute(0)
    fix = make_fix(lmp, arg)
    fix.setup()
    fix.pre_force()
    with pytest.raises(abi.ConpError) as e:
        fix.compute_potential_atom(eta)
    assert "compatible KSpace provider" in str(e.value)
    fix.close()
