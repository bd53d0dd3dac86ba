"""Stand-in for the pieces of the LAMMPS host that the charge solve consumes
but does not own: unit constants, groups, pair cut-offs and the host PPPM's
mesh tables (``rho_coeff``, ``greensfn``).

In a real deployment LAMMPS computes these (PPPM::compute_rho_coeff,
PPPM::compute_gf_ik, KSpace::g_ewald ...) and the shim classes hand the
arrays through the C ABI (include/conp_b200.h).  LAMMPS is not available in
this build environment, so the formulas below are restated from the
published PPPM algorithm (Hockney-Eastwood ``ik`` influence function as used
by LAMMPS 27May2021 ``src/KSPACE/pppm.cpp``); they are *inputs* to both the
CUDA path and the oracle, so parity between the two does not depend on them
("parity unpinned" with respect to LAMMPS itself).
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

from .system import System

# LAMMPS `units real` (update.cpp): force->qqr2e, qqrd2e, qe2f
QQR2E = 332.06371
QQRD2E = 332.06371
QE2F = 23.060549
TWO_CHARGE_FORCE = QQR2E  # qqr2e * qelectron^2 / angstrom^2 in `units real`
EPS_HOC = 1.0e-7
OFFSET = 16384


def good_fft_size(n: int) -> int:
    """Smallest 2^a 3^b 5^c >= n (LAMMPS PPPM::factorable)."""
    n = max(int(n), 2)
    while True:
        m = n
        for f in (2, 3, 5):
            while m % f == 0:
                m //= f
        if m == 1:
            return n
        n += 1


def compute_rho_coeff(order: int) -> np.ndarray:
    """PPPM::compute_rho_coeff: rho_coeff[l][k - nlower], l = 0..order-1,
    k = nlower..nupper, so that w[k] = sum_l rho_coeff[l][k] d^l."""
    a = np.zeros((order, 2 * order + 1))
    off = order  # a[l][k + off], k = -order..order

    a[0][0 + off] = 1.0
    for j in range(1, order):
        for k in range(-j, j + 1, 2):
            s = 0.0
            for l in range(j):
                a[l + 1][k + off] = (a[l][k + 1 + off] - a[l][k - 1 + off]) / (l + 1)
                s += 0.5 ** (l + 1) * (a[l][k - 1 + off] + (-1.0) ** l * a[l][k + 1 + off]) / (l + 1)
            a[0][k + off] = s
    nlower = -((order - 1) // 2)  # C integer division truncates toward zero
    rho = np.zeros((order, order))
    m = nlower
    for k in range(-(order - 1), order, 2):
        for l in range(order):
            rho[l][m - nlower] = a[l][k + off]
        m += 1
    return rho


def compute_gf_denom(order: int) -> np.ndarray:
    gf_b = np.zeros(order)
    gf_b[0] = 1.0
    for m in range(1, order):
        for l in range(m, 0, -1):
            gf_b[l] = 4.0 * (gf_b[l] * (l - m) * (l - m - 0.5) - gf_b[l - 1] * (l - m - 1) * (l - m - 1))
        gf_b[0] = 4.0 * (gf_b[0] * (0 - m) * (0 - m - 0.5))
    ifact = math.factorial(2 * order - 1)
    return gf_b / ifact


def compute_greensfn_ik(mesh, order: int, prd, slab_volfactor: float, g_ewald: float) -> np.ndarray:
    """PPPM::compute_gf_ik on the whole mesh; returns greensfn[nz][ny][nx]
    flattened x-fastest (the FFT-grid order of pppm_conp.cpp:245-249)."""
    nx, ny, nz = (int(v) for v in mesh)
    xprd, yprd = float(prd[0]), float(prd[1])
    zprd_slab = float(prd[2]) * slab_volfactor
    unitk = np.array([2 * math.pi / xprd, 2 * math.pi / yprd, 2 * math.pi / zprd_slab])
    L = (xprd, yprd, zprd_slab)
    n = (nx, ny, nz)
    nb = [int((g_ewald * L[c] / (math.pi * n[c])) * math.pow(-math.log(EPS_HOC), 0.25)) for c in range(3)]
    twoorder = 2 * order
    gf_b = compute_gf_denom(order)

    def per(nn):
        m = np.arange(nn)
        return m - nn * (2 * m // nn)

    kper, lper, mper = per(nx), per(ny), per(nz)

    def axis_terms(c, kp):
        # returns qq[i][alias], s[i][alias]*w[i][alias]
        al = np.arange(-nb[c], nb[c] + 1)
        q = unitk[c] * (kp[:, None] + n[c] * al[None, :])
        s = np.exp(-0.25 * (q / g_ewald) ** 2)
        arg = 0.5 * q * L[c] / n[c]
        with np.errstate(divide="ignore", invalid="ignore"):
            w = np.where(arg == 0.0, 1.0, (np.sin(arg) / arg) ** twoorder)
        return q, s * w

    qx, swx = axis_terms(0, kper)
    qy, swy = axis_terms(1, lper)
    qz, swz = axis_terms(2, mper)
    snx = np.sin(0.5 * unitk[0] * kper * xprd / nx) ** 2
    sny = np.sin(0.5 * unitk[1] * lper * yprd / ny) ** 2
    snz = np.sin(0.5 * unitk[2] * mper * zprd_slab / nz) ** 2

    def horner(x):
        s = np.zeros_like(x)
        for l in range(order - 1, -1, -1):
            s = gf_b[l] + s * x
        return s

    dx, dy, dz = horner(snx), horner(sny), horner(snz)
    g = np.zeros((nz, ny, nx))
    kx = unitk[0] * kper
    ky = unitk[1] * lper
    kz = unitk[2] * mper
    # loop over z planes to bound memory; aliases summed with einsum per plane
    for m in range(nz):
        # sum over aliases: sum1 = sum (k.q / q.q) * sx sy sz wx wy wz
        # shapes: x:(nx,ax) y:(ny,ay) z:(az,)
        qzm, swzm = qz[m], swz[m]
        sum1 = np.zeros((ny, nx))
        for iz in range(qzm.shape[0]):
            dot1 = (kx[None, :, None, None] * qx[None, :, None, :] + ky[:, None, None, None] * qy[:, None, :, None]
                    + kz[m] * qzm[iz])
            dot2 = qx[None, :, None, :] ** 2 + qy[:, None, :, None] ** 2 + qzm[iz] ** 2
            with np.errstate(divide="ignore", invalid="ignore"):
                t = np.where(dot2 == 0.0, 0.0, dot1 / dot2)
            t = t * swx[None, :, None, :] * swy[:, None, :, None] * swzm[iz]
            sum1 += t.sum(axis=(2, 3))
        sqk = kx[None, :] ** 2 + ky[:, None] ** 2 + kz[m] ** 2
        denom = (dx[None, :] * dy[:, None] * dz[m]) ** 2
        with np.errstate(divide="ignore", invalid="ignore"):
            g[m] = np.where(sqk != 0.0, (12.5663706 / sqk) * sum1 / denom, 0.0)
    return np.ascontiguousarray(g.reshape(-1))


@dataclasses.dataclass
class PPPMTables:
    mesh: tuple
    order: int
    rho_coeff: np.ndarray  # (order, order)
    greensfn: np.ndarray  # (nz*ny*nx,)
    shift: float
    shiftone: float


def pppm_tables(mesh, order, prd, slab_volfactor, g_ewald) -> PPPMTables:
    shift = OFFSET + 0.5 if order % 2 else float(OFFSET)
    shiftone = 0.0 if order % 2 else 0.5
    return PPPMTables(tuple(int(v) for v in mesh), int(order), compute_rho_coeff(order),
                      compute_greensfn_ik(mesh, order, prd, slab_volfactor, g_ewald), shift, shiftone)


def mesh_for_spacing(prd, slab_volfactor: float, h: float):
    L = (prd[0], prd[1], prd[2] * slab_volfactor)
    return tuple(good_fft_size(math.ceil(l / h)) for l in L)


class MockLammps:
    """The slice of LAMMPS state that ``fix conp`` reads: atoms, groups, the
    Coulomb pair style's cut-offs, and the KSpace settings (``g_ewald``,
    accuracy, slab factor, PPPM mesh/order).  Mirrors the deck commands of
    tests/*/input: ``boundary``, ``pair_style lj/cut/coul/long``,
    ``kspace_style``, ``kspace_modify slab``, ``group``.
    g_ewald and the mesh must be given explicitly because LAMMPS' auto-tuning
    (KSpace::adjust_gewald / PPPM::set_grid_global) is not available here."""

    def __init__(self, system: System, boundary: str = "p p p"):
        self.system = system
        b = boundary.split()
        self.periodic = [1 if c == "p" else 0 for c in b]
        self.groups = {"all": np.ones(system.natoms, dtype=bool)}
        self.cut_coul = None
        self.cutsq = None
        self.kspace_style = None
        self.accuracy_relative = None
        self.slab_volfactor = 1.0
        self.slabflag = 0
        self.g_ewald = None
        self.mesh = None
        self.order = 5
        self.newton_pair = 0
        self.qqrd2e = QQRD2E
        self.qqr2e = QQR2E
        self.qe2f = QE2F
        self.dielectric = 1.0

    # group commands -------------------------------------------------------
    def group_molecule(self, name, *mols):
        self.groups[name] = np.isin(self.system.mol, mols)

    def group_type(self, name, *types):
        self.groups[name] = np.isin(self.system.type, types)

    def group_union(self, name, *names):
        m = np.zeros(self.system.natoms, dtype=bool)
        for n in names:
            m |= self.groups[n]
        self.groups[name] = m

    # pair_style lj/cut/coul/long <cut> [cut_coul] ---------------------------
    def pair_style_coul_long(self, cut_lj: float, cut_coul: float | None = None):
        cut_coul = cut_lj if cut_coul is None else cut_coul
        self.cut_coul = float(cut_coul)
        nt = self.system.ntypes
        c = max(cut_lj, cut_coul)
        self.cutsq = np.full((nt + 1, nt + 1), c * c)
        self.cutsq[0, :] = 0.0
        self.cutsq[:, 0] = 0.0

    # kspace_style pppm|pppm/conp|ewald acc ; kspace_modify slab/mesh/gewald/order
    def kspace(self, style: str, accuracy_relative: float, g_ewald: float, slab: float | None = None,
               mesh=None, order: int = 5):
        self.kspace_style = style
        self.accuracy_relative = float(accuracy_relative)
        self.g_ewald = float(g_ewald)
        if slab is not None:
            self.slab_volfactor = float(slab)
            self.slabflag = 1
        self.mesh = None if mesh is None else tuple(int(v) for v in mesh)
        self.order = int(order)

    # quantities KSpace exposes -----------------------------------------------
    @property
    def accuracy(self) -> float:
        return self.accuracy_relative * TWO_CHARGE_FORCE

    def q2(self) -> float:
        """qqrd2e * sum(q^2) / dielectric over *all* atoms (km_ewald.cpp:72-79)."""
        return float(np.sum(self.system.q ** 2)) * self.qqrd2e / self.dielectric

    def pppm_tables(self) -> PPPMTables:
        if self.mesh is None:
            raise ValueError("kspace mesh not set")
        return pppm_tables(self.mesh, self.order, self.system.prd, self.slab_volfactor, self.g_ewald)
