"""CPU tests that pin the oracle (oracle/) before it is trusted as the parity
checker: the reference's only stored known-answer (tests/dilute/persist.log:143)
plus the physics cross-checks of SURVEY.md 8(c)."""
import numpy as np
import pytest

import conp_oracle as O
from cases import PINS, dilute, il


def run(lmp, arg, **kw):
    fix = O.OracleFixConp(lmp, arg, **kw)
    fix.setup()
    q = fix.pre_force()
    return fix, q


@pytest.fixture(scope="module")
def golden():
    return run(*dilute(2))


def test_persist_log_step0(golden):
    """persist.log:143 -- c_qleft = 0.044057154, c_qright = -0.044057154,
    c_qall ~ 1e-16 (conp/v4 ... etypes 1 3 ffield; g_ewald known to 8 digits,
    so this is a 1e-7-relative pin)."""
    fix, q = golden
    pin = PINS["step0"]
    qleft, qright = q[fix.side == 1].sum(), q[fix.side == -1].sum()
    assert fix.ewald.kcount == 8095 and (fix.ewald.kxmax, fix.ewald.kymax, fix.ewald.kzmax) == (8, 7, 58)
    assert abs(qleft - pin["c_qleft"]) < 5e-10  # all 8 printed digits
    assert abs(qright - pin["c_qright"]) < 5e-10
    assert abs(q.sum()) < 1e-12
    assert abs(fix.scalar_output - qleft) < 1e-14  # f_e = total left charge for conp
    assert abs(fix.ee - 0.1702472657) < 1e-9 and abs(fix.dd - 0.00731711766) < 1e-10


def test_projected_matrix_is_neutral(golden):
    fix, _ = golden
    assert np.abs(fix.S.sum(axis=1)).max() < 1e-13
    assert np.abs(fix.S - fix.S.T).max() < 1e-12


def test_lowmem_himem_identical(golden):
    fix, q = golden
    lmp, arg = dilute(2)
    fix2, q2 = run(lmp, arg + ["himem"])
    assert np.abs(fix2.A - fix.A).max() < 1e-13
    assert np.abs(q2 - q).max() < 1e-14


def test_cell_search_equals_brute_force(golden):
    fix, q = golden
    fix2, q2 = run(*dilute(2), brute_pairs=True)
    assert np.array_equal(fix2.A, fix.A)
    assert np.abs(q2 - q).max() < 1e-15


def test_etypes_does_not_change_charges():
    """etypes only prunes pairs with zero-charge partners (README 'etypes')."""
    _, q0 = run(*dilute(0))
    _, q1 = run(*dilute(1))
    assert np.abs(q0 - q1).max() < 1e-14


def test_slab_vs_ffield_agree_physically(golden):
    """compare.gnu overlays these; SURVEY 8c quotes 0.04405271594 for slab."""
    fix, q = golden
    fs, qs = run(*dilute(0))
    assert fs.ewald.kcount == 22687
    assert abs(qs[fs.side == 1].sum() - 0.04405271594) < 1e-10
    assert abs(qs[fs.side == 1].sum() - q[fix.side == 1].sum()) < 1e-4 * abs(q[fix.side == 1].sum()) * 2


def test_conq_inverts_conp():
    """tests/cond/input:56-66 pattern: conq(QR from a conp run) returns dV."""
    fs, qs = run(*dilute(0))
    lmp, arg = dilute(0)
    arg[2] = "conq"
    arg[6] = "%.17g" % qs[fs.side == -1].sum()
    fq, qq = run(lmp, arg)
    assert abs(fq.scalar_output - 1.0) < 1e-12
    assert np.abs(qq - qs).max() < 1e-14


def test_pppm_mode_converges_to_ewald(golden):
    """mesh 27x24x144 order 5 -> |db_k| ~ 3e-6, qleft 0.044057071."""
    fix, q = golden
    fp, qp = run(*dilute(2, pppm=True))
    assert np.abs(fp.b_kspace - fix.b_kspace).max() < 1e-5
    assert abs(qp[fp.side == 1].sum() - 0.044057071) < 2e-9


def test_noslab_zneutr_doubled_cell():
    """dilute/input n=3,4: each half-cell neutral; sym and anti give the same
    single-cell electrode charge as ffield to ~1e-3 relative."""
    _, qf = run(*dilute(2))
    for n in (3, 4):
        lmp, arg = dilute(n)
        fix, q = run(lmp, arg)
        z = lmp.system.x[fix.ele_idx, 2]
        assert abs(q[z > 0].sum()) < 1e-12 and abs(q[z < 0].sum()) < 1e-12
        qleftneg = q[(fix.side == 1) & (z < 0)].sum()
        assert abs(qleftneg - 0.04405715384) < 2e-3 * 0.044


def test_ehgo_auto_reduces_to_eta():
    """il_onelayer/input:104-106: kappa 0 + coeff <etype> eta auto == ETA mode."""
    lmp, arg = dilute(2)
    f1, q1 = run(lmp, arg)
    lmp2, arg2 = dilute(2)
    f2 = O.OracleFixConp(lmp2, arg2 + ["ehgo"])
    f2.modify_param("ehgo kappa 0".split())
    f2.modify_param("ehgo coeff 3 1.979 auto".split())
    f2.setup()
    q2 = f2.pre_force()
    # eta_ij = eta/sqrt(2) for electrode pairs (A) but eta for electrode-electrolyte (b): auto u0 matches
    assert np.abs(f2.A - f1.A).max() < 1e-12
    assert np.abs(q2 - q1).max() < 1e-12


def test_il_onelayer_runs_and_is_neutral():
    fix, q = run(*il(1))
    assert fix.N == 832
    assert abs(q.sum()) < 1e-12
    assert q[fix.side == 1].sum() > 0


def test_potential_atom_twin_gives_equipotential_electrodes():
    """The oracle's restatement of `compute potential/atom` (compute_potential_atom.cpp:120-182) is pinned
    by the physics it measures: after a conp solve the potential is constant on each electrode and the two
    electrodes differ by the applied dV = 1.0 V (to the PPPM-vs-Ewald consistency level)."""
    import numpy as np
    import conp_oracle as O
    from cases import dilute
    lmp, arg = dilute(0, pppm=True)
    ref = O.OracleFixConp(lmp, arg)
    ref.setup()
    ref.pre_force()
    phi = ref.potential_atom(ref.args.eta)
    left, right = phi[ref.side == 1], phi[ref.side == -1]
    assert np.ptp(left) < 1e-4 and np.ptp(right) < 1e-4
    assert abs((right.mean() - left.mean()) - 1.0) < 1e-4
