"""TEST INFRASTRUCTURE ONLY -- Python driver of the CPU oracle.

Wraps oracle/conp_oracle.c (ctypes) and restates the *driver* logic of the
reference fix (hook order, which vector is computed when) for one rank:

    setup_post_neighbor / setup_pre_force   fix_conp.cpp:382-464
    pre_force  -> b_cal, update_charge       fix_conp.cpp:543-573, 677-695, 1120-1161
    post_force -> force_cal                  fix_conp.cpp:577-580, 1163-1201
    FixConq / FixCond update_charge          fix_conq.cpp:41-90, fix_cond.cpp:46-126
    PPPMCONP::b_cal / elyte_poisson          pppm_conp.cpp:109-124, 230-316

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.  The product package never does.

Parity pin: tests/dilute/persist.log:143 (see tests/test_oracle_golden.py).
All other modes are "parity unpinned" in the reference's own tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

MY_PIS = 1.77245385090551602729  # sqrt(pi), math_const.h

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.normpath(os.path.join(_HERE, "..", "lammps-user-conp2_b200"))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from conp_b200.fixargs import (FF_FFIELD, FF_NORMAL, FF_NOSLAB, PAIR_EHGO, PAIR_ETA,  # noqa: E402
                               VARIANT_COND, VARIANT_CONP, VARIANT_CONQ, FixError, parse_fix_args)

_LIB = None
_SO = os.path.join(_HERE, "_build", "libconp_oracle.so")
_SRC = os.path.join(_HERE, "conp_oracle.c")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """gcc the C restatement into oracle/_build/libconp_oracle.so."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        cmd = ["gcc", "-O3", "-mavx2", "-mfma", "-fopenmp", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"]
        subprocess.check_call(cmd)
    return _SO


def dp(a):
    return a.ctypes.data_as(c_dp)


def ip(a):
    return a.ctypes.data_as(c_ip)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(_SO)
        L.orc_erfcr_sqrt.restype = C.c_double
        L.orc_erfcr_sqrt.argtypes = [C.c_double]
        L.orc_ferfcr_sqrt.restype = C.c_double
        L.orc_ferfcr_sqrt.argtypes = [C.c_double]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_ehgo_setup_tables.restype = C.c_int
        L.orc_ehgo_setup_tables.argtypes = [C.c_int, C.c_double, c_dp, c_dp, c_dp, c_dp]
        L.orc_ewald_create.restype = C.c_void_p
        L.orc_ewald_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_longlong, c_dp, C.c_int,
                                       C.c_double, C.c_int]
        L.orc_ewald_destroy.argtypes = [C.c_void_p]
        L.orc_ewald_info.argtypes = [C.c_void_p, c_ip, c_dp]
        L.orc_ewald_get_kvecs.argtypes = [C.c_void_p, c_ip, c_ip, c_ip, c_dp]
        L.orc_ewald_get_sfac.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_ewald_a_read.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.orc_ewald_aaa.argtypes = [C.c_void_p, c_dp, c_dp]
        L.orc_ewald_sincos_b.argtypes = [C.c_void_p, C.c_int, c_dp, c_dp]
        L.orc_ewald_bbb.argtypes = [C.c_void_p, c_dp]
        L.orc_slabcorr.argtypes = [C.c_double, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp]
        pair_common = [c_dp, c_dp, c_ip, C.c_int, C.c_double, C.c_int, c_dp, c_dp, c_dp, C.c_double]
        L.orc_alist_coul_cal.argtypes = [C.c_int, c_dp, c_ip] + pair_common + [C.c_double, C.c_int, c_ip,
                                                                                C.c_int, c_dp]
        L.orc_blist_coul_cal.argtypes = [C.c_int, c_dp, c_ip, C.c_int, c_dp, c_dp, c_ip] + pair_common + [
            C.c_double, C.c_int, c_ip, C.c_int, c_dp]
        L.orc_force_cal.argtypes = [C.c_int, c_dp, c_ip, c_dp, C.c_int, c_dp, c_dp, c_ip, c_dp, c_dp, c_ip,
                                    C.c_int, C.c_double, C.c_int, c_dp, c_dp, c_dp, c_dp, C.c_double,
                                    C.c_double, C.c_int, c_ip, C.c_int, c_dp, c_dp]
        L.orc_a_self.argtypes = [C.c_int, c_ip, C.c_int, C.c_double, c_dp, c_dp]
        L.orc_a_symmetrize.argtypes = [C.c_int, c_dp]
        L.orc_b_setq_cal.argtypes = [C.c_int, c_dp, c_ip, C.c_int, C.c_double, C.c_double, C.c_double, c_dp]
        L.orc_inv.restype = C.c_int
        L.orc_inv.argtypes = [C.c_int, c_dp]
        L.orc_inv_project.restype = C.c_double
        L.orc_inv_project.argtypes = [C.c_int, c_dp, C.c_int, C.c_int, c_ip]
        L.orc_matvec.argtypes = [C.c_int, c_dp, c_dp, c_dp]
        L.orc_totsetq.restype = C.c_double
        L.orc_totsetq.argtypes = [C.c_int, c_dp, c_ip]
        L.orc_update_charge.restype = C.c_double
        L.orc_update_charge.argtypes = [C.c_int, C.c_int, c_dp, c_dp, c_dp, c_ip, C.c_double, C.c_double,
                                        C.c_int, c_dp, c_dp, c_dp]
        L.orc_cond_vmult.restype = C.c_double
        L.orc_cond_vmult.argtypes = [C.c_int, c_dp, c_dp, C.c_double, C.c_double, C.c_double]
        L.orc_pppm_make_rho.restype = C.c_int
        L.orc_pppm_make_rho.argtypes = [c_ip, C.c_int, c_dp, c_dp, c_dp, C.c_int, c_dp, c_dp, c_dp]
        L.orc_pppm_map_ele.argtypes = [c_ip, C.c_int, c_dp, c_dp, c_dp, C.c_int, c_dp, c_ip, c_dp]
        L.orc_pppm_gather_b.argtypes = [c_ip, C.c_int, C.c_int, c_ip, c_dp, c_dp, c_dp]
        L.orc_pppm_ele_make_rho.argtypes = [c_ip, C.c_int, c_dp, C.c_int, c_ip, c_dp, c_dp, c_dp]
        _LIB = L
    return _LIB


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class OracleEwald:
    """KSpaceModuleEwald (km_ewald.{h,cpp})."""

    def __init__(self, g_ewald, accuracy, q2, natoms, prd, slabflag, slab_volfactor, lowmem=True):
        self.L = lib()
        self._prd = f64(prd)
        self.h = self.L.orc_ewald_create(g_ewald, accuracy, q2, int(natoms), dp(self._prd), int(slabflag),
                                         float(slab_volfactor), int(bool(lowmem)))
        info = np.zeros(16, dtype=np.int32)
        dinfo = np.zeros(8)
        self.L.orc_ewald_info(self.h, ip(info), dp(dinfo))
        (self.kxmax, self.kymax, self.kzmax, self.kcount, self.kcount_flat, self.kcount_expand, self.kcount_a,
         self.kmax3d) = (int(v) for v in info[:8])
        self.kcount_dims = [int(v) for v in info[8:15]]
        self.gsqmx, self.ug_tot, self.volume = float(dinfo[0]), float(dinfo[1]), float(dinfo[2])
        self.unitk = dinfo[3:6].copy()
        self.nele = 0

    def __del__(self):
        try:
            self.L.orc_ewald_destroy(self.h)
        except Exception:
            pass

    def kvecs(self):
        kx = np.zeros(self.kcount, dtype=np.int32)
        ky = np.zeros_like(kx)
        kz = np.zeros_like(kx)
        ug = np.zeros(self.kcount)
        self.L.orc_ewald_get_kvecs(self.h, ip(kx), ip(ky), ip(kz), dp(ug))
        return kx, ky, kz, ug

    def a_read(self, xele):
        xele = f64(xele)
        self.nele = xele.shape[0]
        self.L.orc_ewald_a_read(self.h, self.nele, dp(xele))

    def a_cal(self, xele, aaa):
        """KSpaceModuleEwald::a_cal km_ewald.cpp:147-151 (aaa zero-filled by caller)."""
        xele = f64(xele)
        self.a_read(xele)
        self.L.orc_ewald_aaa(self.h, dp(xele), dp(aaa))

    def b_cal(self, x, q, xele, bbb):
        """KSpaceModuleEwald::b_cal km_ewald.cpp:153-167: x,q = all non-electrode atoms."""
        x, q, xele = f64(x), f64(q), f64(xele)
        self.L.orc_ewald_sincos_b(self.h, x.shape[0], dp(x), dp(q))
        self.L.orc_ewald_bbb(self.h, dp(bbb))

    def sfac(self):
        re = np.zeros(self.kcount)
        im = np.zeros(self.kcount)
        self.L.orc_ewald_get_sfac(self.h, dp(re), dp(im))
        return re, im


def pppm_poisson(brick, greensfn, mesh, workers=None):
    """PPPMCONP::elyte_poisson pppm_conp.cpp:230-267: complex FFT of the real
    density, multiply by greensfn/(nx ny nz), inverse FFT (unnormalised, sign
    -1 in LAMMPS' convention = numpy's forward-unnormalised ifft * N), keep the
    real part."""
    import scipy.fft as sfft
    nx, ny, nz = mesh
    rho = brick.reshape(nz, ny, nx).astype(np.complex128)
    work = sfft.fftn(rho, workers=workers)
    work *= greensfn.reshape(nz, ny, nx) / float(nx * ny * nz)
    # LAMMPS fft2->compute(...,-1) is the unnormalised backward transform
    u = sfft.ifftn(work, workers=workers) * float(nx * ny * nz)
    return np.ascontiguousarray(u.real.reshape(-1))


class OracleFixConp:
    """One-rank restatement of FixConp/FixConq/FixCond driven in the
    reference's hook order.  ``lmp`` is a conp_b200.mockhost.MockLammps."""

    def __init__(self, lmp, arg, brute_pairs=False, fft_workers=None):
        self.L = lib()
        self.lmp = lmp
        s = lmp.system
        self.args = a = parse_fix_args(arg, s.ntypes)
        if a.group not in lmp.groups:
            raise FixError("Could not find fix group ID")
        if a.group2 not in lmp.groups:
            raise FixError("Fix conp group ID does not exist")  # fix_conp.cpp:106-107
        if a.splitflag:
            raise FixError("split is out of scope (experimental KSpaceModuleEwaldSplit)")
        if a.potdiffstr is not None:
            raise FixError("Fix conp potential difference variable does not exist")  # no variables in mock host
        self.brute = int(bool(brute_pairs))
        self.fft_workers = fft_workers
        g1, g2 = lmp.groups[a.group], lmp.groups[a.group2]
        # electrode_check fix_conp.cpp:599-605
        side_all = np.where(g1, 1, np.where(g2, -1, 0)).astype(np.int32)
        self.one_electrode_flag = bool(np.array_equal(g1, g2))  # groupbit == jgroupbit :295
        self.ele_idx = np.nonzero(side_all != 0)[0]
        self.oth_idx = np.nonzero(side_all == 0)[0]
        self.side = i32(side_all[self.ele_idx])
        self.N = int(self.ele_idx.shape[0])
        self.potdiff = a.potdiff
        self.runstage = 0
        self.ehgo = dict(kappa=1.0, eta_i=np.zeros(s.ntypes + 1), u0_i=np.zeros(s.ntypes + 1))
        self.eta_ij = np.zeros((s.ntypes + 1) ** 2)
        self.fo_ij = np.zeros((s.ntypes + 1) ** 2)
        self.pairmode = a.pairmode
        self.is_eletype = np.zeros(s.ntypes + 1, dtype=np.int32)
        for t in a.eletypes:
            self.is_eletype[t] = 1
        self.evscale = lmp.qe2f / lmp.qqr2e  # :412
        self.scalar_output = 0.0
        self.eleinitq = None
        self.pppm = None

    # fix_modify ID ehgo kappa K | coeff <types> eta u0|auto   fix_conp.cpp:1482-1515
    def modify_param(self, arg):
        if self.pairmode == PAIR_ETA:
            raise FixError("Can't fix_modify conp parameters in basic pair mode")
        CON_s2overPIS = np.sqrt(2.0) / 1.77245385090551602729
        if arg[0] == "ehgo":
            if arg[1] == "kappa":
                if len(arg) != 3:
                    raise FixError("Invalid number of inputs for EHGO coeff setting")
                self.ehgo["kappa"] = float(arg[2])
                return 3
            elif arg[1] == "coeff":
                if len(arg) != 5:
                    raise FixError("Invalid number of inputs for EHGO coeff setting")
                nt = self.lmp.system.ntypes
                lohi = str(arg[2])
                if "*" in lohi:
                    lo_s, hi_s = lohi.split("*")
                    ilo = int(lo_s) if lo_s else 1
                    ihi = int(hi_s) if hi_s else nt
                else:
                    ilo = ihi = int(lohi)
                eta_one = float(arg[3])
                u0_one = CON_s2overPIS * eta_one / self.evscale if arg[4] == "auto" else float(arg[4])
                if ilo > ihi:
                    raise FixError("Couldn't set EHGO coeffs with mintype more than maxtype")
                for i in range(ilo, ihi + 1):
                    self.ehgo["eta_i"][i] = eta_one
                    self.ehgo["u0_i"][i] = u0_one * self.evscale
                return 5
            raise FixError("Invalid entry for EHGO coeff setting")
        return 0

    # -- helpers ---------------------------------------------------------
    def _geom(self):
        s = self.lmp.system
        return f64(s.boxlo), f64(s.prd), i32(self.lmp.periodic)

    def _ele(self):
        s = self.lmp.system
        return f64(s.x[self.ele_idx]), i32(s.type[self.ele_idx])

    def _oth(self):
        s = self.lmp.system
        return f64(s.x[self.oth_idx]), f64(s.q[self.oth_idx]), i32(s.type[self.oth_idx])

    def _pair_args(self):
        lmp = self.lmp
        boxlo, prd, per = self._geom()
        self._keep = (boxlo, prd, per, f64(lmp.cutsq).reshape(-1))
        return [dp(boxlo), dp(prd), ip(per), int(self.pairmode), float(self.args.eta), int(lmp.system.ntypes),
                dp(self.eta_ij), dp(self.fo_ij), dp(self._keep[3]), float(lmp.cut_coul)]

    # -- setup -------------------------------------------------------------
    def setup(self):
        """setup_post_neighbor + setup_pre_force without the trailing
        pre_force (fix_conp.cpp:382-464)."""
        lmp, a, L = self.lmp, self.args, self.L
        s = lmp.system
        if lmp.cut_coul is None:
            raise FixError("Fix conp couldn't detect a Coulombic pair style")  # :258
        if a.pppmflag and lmp.kspace_style != "pppm/conp":
            raise FixError("Fix conp couldn't detect a pppm/conp kspace style (which is required with the pppm flag)")
        if self.pairmode == PAIR_EHGO:  # init() :296-299
            ok = L.orc_ehgo_setup_tables(s.ntypes, self.ehgo["kappa"], dp(self.ehgo["eta_i"]),
                                         dp(self.ehgo["u0_i"]), dp(self.eta_ij), dp(self.fo_ij))
            if not ok:
                self.pairmode = PAIR_ETA  # warning "No EHGO settings found" :1553-1558
        self.g_ewald = lmp.g_ewald
        xele, tele = self._ele()
        N = self.N
        # linalg_init :400-410 (the A matrix always comes from an Ewald module,
        # also in pppm mode: PPPMCONP::a_cal pppm_conp.cpp:91-101)
        self.ewald = OracleEwald(lmp.g_ewald, lmp.accuracy, lmp.q2(), s.natoms, s.prd, lmp.slabflag,
                                 lmp.slab_volfactor, a.lowmemflag)
        # a_cal :777-861
        if a.a_matrix_f == 0:
            aaa = np.zeros((N, N))
            self.ewald.a_cal(xele, aaa)
            L.orc_a_self(N, ip(tele), int(self.pairmode), float(a.eta), dp(self.ehgo["u0_i"]), dp(aaa))
            L.orc_a_symmetrize(N, dp(aaa))
            pa = self._pair_args()
            L.orc_alist_coul_cal(N, dp(xele), ip(tele), *pa, float(self.g_ewald), int(a.smartlist),
                                 ip(self.is_eletype), self.brute, dp(aaa))
            self.aaa_all = aaa
            self.A = aaa.copy()
        else:
            tags, mat = read_matrix_file(a.a_matrix_file, N)
            self.aaa_all = mat
            self.A = mat.copy() if a.a_matrix_f == 1 else None
            self.ewald.a_read(xele)
        if self.args.matoutflag:
            write_matrix_file("amatrix", s.id[self.ele_idx], self.aaa_all, "%20.12f")
        if a.pppmflag:
            self._pppm_setup()
        # b_setq_cal :609-637
        d = np.zeros(N)
        L.orc_b_setq_cal(N, dp(xele), ip(self.side), int(a.ff_flag), self.evscale, float(s.boxlo[2]),
                         float(s.prd[2]), dp(d))
        self.d = d
        # cond_setup fix_cond.cpp:46-55
        self.setzvec = d / self.evscale
        # equation_solve -> inv :932-980
        self.ee = None
        if a.a_matrix_f < 2:
            info = L.orc_inv(N, dp(self.aaa_all))
            if info != 0:
                raise FixError("Inversion failed!")
            if not self.one_electrode_flag:
                self._inv_project()
            if a.matoutflag:
                write_matrix_file("inv_a_matrix", s.id[self.ele_idx], self.aaa_all, "%20.10f")
        # get_setq :1071-1116
        self.elesetq = np.zeros(N)
        L.orc_matvec(N, dp(self.aaa_all), dp(d), dp(self.elesetq))
        self.totsetq = L.orc_totsetq(N, dp(self.elesetq), ip(self.side))
        if a.qinitflag:
            self.eleinitq = f64(s.q[self.ele_idx]).copy()
        if self.one_electrode_flag:
            self._inv_project()
        self.dd = -self.totsetq  # "<d,d>" log line :458-461
        self.vmult = None
        self.runstage = 3
        self.S = self.aaa_all

    def setup_preinverted(self, S):
        """Setup with an already inverted+projected matrix (the `inv <file>`
        path, a_matrix_f == 2, fix_conp.cpp:442-445, 935): used by bench.py to
        time the per-step path at sizes where the O(N^2 K) A build is out of
        reach for a CPU."""
        lmp, a, L = self.lmp, self.args, self.L
        s = lmp.system
        N = self.N
        self.g_ewald = lmp.g_ewald
        xele, _ = self._ele()
        if a.pppmflag:
            self._pppm_setup()
            self.ewald = None
        else:
            self.ewald = OracleEwald(lmp.g_ewald, lmp.accuracy, lmp.q2(), s.natoms, s.prd, lmp.slabflag,
                                     lmp.slab_volfactor, a.lowmemflag)
            self.ewald.a_read(xele)
        self.aaa_all = self.S = np.ascontiguousarray(S, dtype=np.float64)
        d = np.zeros(N)
        L.orc_b_setq_cal(N, dp(xele), ip(self.side), int(a.ff_flag), self.evscale, float(s.boxlo[2]),
                         float(s.prd[2]), dp(d))
        self.d = d
        self.setzvec = d / self.evscale
        self.elesetq = np.zeros(N)
        L.orc_matvec(N, dp(self.S), dp(d), dp(self.elesetq))
        self.totsetq = L.orc_totsetq(N, dp(self.elesetq), ip(self.side))
        self.vmult = None
        self.runstage = 3

    def _inv_project(self):
        s = self.lmp.system
        zhalf = 0.5 * s.prd[2] + s.boxlo[2]
        zpos = i32(s.x[self.ele_idx, 2] > zhalf)
        tot = self.L.orc_inv_project(self.N, dp(self.aaa_all), int(self.args.nullneutralflag),
                                     int(self.args.zneutrflag), ip(zpos))
        self.ee = tot * self.evscale  # "<e,e>" :1006-1009

    def _pppm_setup(self):
        lmp = self.lmp
        s = lmp.system
        t = lmp.pppm_tables()
        self.pppm = t
        self.mesh = i32(t.mesh)
        self.prd_slab = f64([s.prd[0], s.prd[1], s.prd[2] * lmp.slab_volfactor])
        xele, _ = self._ele()
        self.part2grid = np.zeros((self.N, 3), dtype=np.int32)
        self.ele2rho = np.zeros((self.N, 3, t.order))
        self._rho = f64(t.rho_coeff)
        self.L.orc_pppm_map_ele(ip(self.mesh), t.order, dp(f64(s.boxlo)), dp(self.prd_slab), dp(self._rho), self.N,
                                dp(xele), ip(self.part2grid), dp(self.ele2rho))
        self.volume = float(np.prod(self.prd_slab))

    # -- per step -----------------------------------------------------------
    def b_cal(self):
        """update_bk fix_conp.cpp:684-695: k-space part (overwrites), then the
        real-space pair part (accumulates)."""
        lmp, a, L = self.lmp, self.args, self.L
        s = lmp.system
        xele, tele = self._ele()
        x, q, typ = self._oth()
        bbb = np.zeros(self.N)
        if a.pppmflag:
            t = self.pppm
            ng = int(np.prod(t.mesh))
            brick = np.zeros(ng)
            bad = L.orc_pppm_make_rho(ip(self.mesh), t.order, dp(f64(s.boxlo)), dp(self.prd_slab), dp(self._rho),
                                      x.shape[0], dp(x), dp(q), dp(brick))
            if bad:
                raise FixError("Out of range atoms - cannot compute PPPM")
            self.elyte_density = brick
            self.u_brick = pppm_poisson(brick, t.greensfn, t.mesh, self.fft_workers)
            L.orc_pppm_gather_b(ip(self.mesh), t.order, self.N, ip(self.part2grid), dp(self.ele2rho),
                                dp(self.u_brick), dp(bbb))
            if lmp.slabflag == 1:
                L.orc_slabcorr(self.volume, x.shape[0], dp(x), dp(q), self.N, dp(xele), dp(bbb))
        else:
            self.ewald.b_cal(x, q, xele, bbb)
            if lmp.slabflag:
                L.orc_slabcorr(self.ewald.volume, x.shape[0], dp(x), dp(q), self.N, dp(xele), dp(bbb))
        self.b_kspace = bbb.copy()
        pa = self._pair_args()
        L.orc_blist_coul_cal(self.N, dp(xele), ip(tele), x.shape[0], dp(x), dp(q), ip(typ), *pa,
                             float(self.g_ewald), int(a.smartlist), ip(self.is_eletype), self.brute, dp(bbb))
        self.bbb_all = bbb
        return bbb

    def update_charge(self):
        lmp, a, L = self.lmp, self.args, self.L
        s = lmp.system
        N = self.N
        eleallq = np.zeros(N)
        L.orc_matvec(N, dp(self.S), dp(self.bbb_all), dp(eleallq))
        self.eleallq = eleallq
        q_out = np.zeros(N)
        init = dp(self.eleinitq) if self.eleinitq is not None else None
        aux = np.zeros(3)
        if a.variant == VARIANT_COND:
            if self.vmult is None:  # cond_setup2 fix_cond.cpp:57-68
                self.vmult = L.orc_cond_vmult(N, dp(self.elesetq), dp(self.setzvec), float(s.prd[2]),
                                              float(s.prd[0] * s.prd[1]), self.evscale)
            x, q, _ = self._oth()
            aux[0] = float(-np.sum(q * x[:, 2]))  # dipole fix_cond.cpp:101-106
            aux[1] = float(s.prd[2])
            aux[2] = self.vmult
        self.scalar_output = L.orc_update_charge(int(a.variant), N, dp(eleallq), dp(self.elesetq), init,
                                                 ip(self.side), self.totsetq, float(self.potdiff),
                                                 int(self.one_electrode_flag), dp(self.setzvec), dp(aux),
                                                 dp(q_out))
        s.q[self.ele_idx] = q_out
        self.q_ele = q_out
        if a.pppmflag:  # kspmod->update_charge() -> ele_make_rho pppm_conp.cpp:385-426
            t = self.pppm
            brick = np.zeros(int(np.prod(t.mesh)))
            L.orc_pppm_ele_make_rho(ip(self.mesh), t.order, dp(self.prd_slab), N, ip(self.part2grid),
                                    dp(self.ele2rho), dp(q_out), dp(brick))
            self.ele_density = brick
        return q_out

    def pre_force(self):
        """FixConp::pre_force fix_conp.cpp:543-573 (every-step case)."""
        self.b_cal()
        return self.update_charge()

    def post_force(self):
        """force_cal fix_conp.cpp:1163-1201: returns (f_on_non_electrode (n,3),
        ecoul_pair, eself, virial[6])."""
        lmp, a, L = self.lmp, self.args, self.L
        s = lmp.system
        xele, tele = self._ele()
        x, q, typ = self._oth()
        qele = f64(s.q[self.ele_idx])
        f = np.zeros((x.shape[0], 3))
        out = np.zeros(8)
        boxlo, prd, per = self._geom()
        cutsq = f64(lmp.cutsq).reshape(-1)
        L.orc_force_cal(self.N, dp(xele), ip(tele), dp(qele), x.shape[0], dp(x), dp(q), ip(typ), dp(boxlo),
                        dp(prd), ip(per), int(self.pairmode), float(a.eta), int(s.ntypes), dp(self.eta_ij),
                        dp(self.fo_ij), dp(self.ehgo["u0_i"]), dp(cutsq), float(lmp.cut_coul), float(lmp.qqrd2e),
                        int(a.smartlist), ip(self.is_eletype), self.brute, dp(f), dp(out))
        return f, float(out[0]), float(out[1]), out[2:8].copy()

    def compute_scalar(self):
        return self.scalar_output

    # -- compute potential/atom for the electrode atoms ---------------------------
    def mesh_potential(self, xyz):
        """Mesh sum of PPPMCONP::compute_particle_potential (pppm_conp.cpp:452-484) at positions xyz with
        u_brick = potential of the total (electrolyte + electrode) density of the last update."""
        t = self.pppm
        s = self.lmp.system
        u_tot = pppm_poisson(self.elyte_density + self.ele_density, t.greensfn, t.mesh, self.fft_workers)
        xyz = f64(xyz).reshape(-1, 3)
        n = xyz.shape[0]
        p2g = np.zeros((n, 3), dtype=np.int32)
        w = np.zeros((n, 3, t.order))
        self.L.orc_pppm_map_ele(ip(self.mesh), t.order, dp(f64(s.boxlo)), dp(self.prd_slab), dp(self._rho), n,
                                dp(xyz), ip(p2g), dp(w))
        out = np.zeros(n)
        self.L.orc_pppm_gather_b(ip(self.mesh), t.order, n, ip(p2g), dp(w), dp(u_tot), dp(out))
        return -out

    def potential_atom(self, eta, pair=True, kspace=True, qsum=True):
        """ComputePotentialAtom::compute_peratom (compute_potential_atom.cpp:120-182) for the electrode
        atoms after update_charge, in volts: compute_pair_potential (:223-318) by brute force over all
        periodic images, the k-space part (:161-171) with the mesh potential of ALL charges, slabcorr
        (:333-358)."""
        lmp = self.lmp
        s = lmp.system
        N = self.N
        x, q, typ = f64(s.x), f64(s.q), s.type
        is_ele = np.zeros(s.natoms, dtype=bool)
        is_ele[self.ele_idx] = True
        phi = np.zeros(N)
        g = self.g_ewald
        EWALD_P, A = 0.3275911, (0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429)

        def erfc_poly(a_r):   # the polynomial both the fix and the compute use (:296-299)
            tt = 1.0 / (1.0 + EWALD_P * a_r)
            return tt * (A[0] + tt * (A[1] + tt * (A[2] + tt * (A[3] + tt * A[4])))) * np.exp(-a_r * a_r)
        if pair:
            cut_coulsq = min(lmp.cut_coul ** 2, 5.8 ** 2 / (g * g))
            smax = [int(np.ceil(np.sqrt(cut_coulsq) / s.prd[a])) if lmp.periodic[a] else 0 for a in range(3)]
            shifts = [(i, j, k) for i in range(-smax[0], smax[0] + 1) for j in range(-smax[1], smax[1] + 1)
                      for k in range(-smax[2], smax[2] + 1)]
            for a, i in enumerate(self.ele_idx):
                acc = 0.0
                for sh in shifts:
                    d = x[i] - (x + np.array(sh) * s.prd)
                    rsq = np.maximum((d * d).sum(axis=1), 1e-10)
                    m = (rsq < lmp.cutsq[typ[i], typ]) & (rsq < cut_coulsq) & ((q[i] != 0) | (q != 0))
                    if sh == (0, 0, 0):
                        m[i] = False
                    r = np.sqrt(rsq[m])
                    dudq = erfc_poly(g * r) / r
                    if eta != 0.0:
                        etar = np.where(is_ele[m], eta * r / np.sqrt(2.0), eta * r)
                        dudq = dudq - np.where(etar < 5.8, erfc_poly(etar) / r, 0.0)
                    acc += float((q[m] * dudq).sum())
                phi[a] += acc
        if kspace:
            if self.pppm is None:
                raise FixError("Compute requires a compatible KSpace provider like pppm/conp")
            qe = q[self.ele_idx]
            phi += self.mesh_potential(x[self.ele_idx]) - 2.0 * g * qe / MY_PIS
            if eta != 0.0:
                phi += eta * qe * np.sqrt(2.0) / MY_PIS
            if lmp.slabflag:
                pi2vol = 2.0 * np.pi / self.volume
                slabcorr = 2.0 * pi2vol * float((q * x[:, 2]).sum())
                z = x[self.ele_idx, 2]
                phi += z * slabcorr
                if qsum:
                    phi -= pi2vol * float(q.sum()) * z * z
        return phi * (lmp.qqr2e / lmp.qe2f)


# -- matout / org / inv file formats (fix_conp.cpp:721-773, 833-849, 960-977) ---

def write_matrix_file(path, tags, mat, fmt):
    n = mat.shape[0]
    with open(path, "w") as fh:
        fh.write(" " + "".join("%20d" % t for t in tags) + "\n")
        for i in range(n):
            fh.write(" " + ("" if fmt == "%20.12f" else "").join(fmt % v for v in mat[i]) + "\n")


def read_matrix_file(path, n):
    with open(path) as fh:
        toks = fh.read().split()
    if len(toks) > n + n * n:
        raise FixError("Too many entries in A matrix file")
    if len(toks) < n + n * n:
        raise FixError("Too few entries in A matrix file")
    tags = np.array([int(t) for t in toks[:n]], dtype=np.int32)
    mat = np.array([float(t) for t in toks[n:]], dtype=np.float64).reshape(n, n)
    return tags, mat
