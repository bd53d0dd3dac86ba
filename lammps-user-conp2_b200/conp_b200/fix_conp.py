"""Host-side mirror of the reference's operator interface for the charge solve:
``fix conp`` / ``fix conq`` / ``fix cond`` (+ the ``pppm/conp`` KSpaceModule
seam), driving libconp_b200.so through the C ABI in the reference's hook order

    setup_post_neighbor -> setup_pre_force -> [post_neighbor] -> pre_force
    -> post_force                       (fix_conp.cpp:233-241, 382-391, 543-588)

Same constructor tokens, keywords and error messages as fix_conp.cpp:79-201.
Everything numerical happens on the GPU; this file is argument parsing, O(N)
index bookkeeping (what FixConp::post_neighbor does with its cross-lists) and
the text I/O of ``matout``/``org``/``inv``.  The C++ shim a LAMMPS build would
compile (shim/) does exactly the same calls.
"""
from __future__ import annotations

import math

import numpy as np

from . import abi
from .fixargs import (FF_FFIELD, FF_NORMAL, FF_NOSLAB, PAIR_EHGO, PAIR_ETA, VARIANT_COND, VARIANT_CONP,
                      VARIANT_CONQ, FixArgs, FixError, parse_fix_args)

MY_PIS = 1.77245385090551602729


def ehgo_setup_tables(ntypes, kappa, eta_i, u0_i):
    """FixConp::ehgo_setup_tables (fix_conp.cpp:1517-1559).  Returns
    (eta_ij, fo_ij) as (ntypes+1)^2 arrays, or None when no coefficient was
    set (the reference then warns and falls back to ETA mode, :1553-1558)."""
    if not (np.any(eta_i[1:] != 0) or np.any(u0_i[1:] != 0)):
        return None
    n1 = ntypes + 1
    eta_ij = np.zeros((n1, n1))
    fo_ij = np.zeros((n1, n1))
    s2 = math.sqrt(2.0) / MY_PIS
    f_i = u0_i - s2 * eta_i
    for i in range(1, n1):
        for j in range(1, i + 1):
            if eta_i[i] and eta_i[j]:
                etaprod = eta_i[i] * eta_i[j]
                e = etaprod / math.sqrt(eta_i[i] ** 2 + eta_i[j] ** 2)
                eta_ij[i, j] = e
                fo_ij[i, j] = 0.5 * kappa * (f_i[i] + f_i[j]) * math.sqrt(8.0) * e * e * e / (etaprod * math.sqrt(etaprod))
            else:
                eta_ij[i, j] = eta_i[i] + eta_i[j]
            eta_ij[j, i] = eta_ij[i, j]
            fo_ij[j, i] = fo_ij[i, j]
    return eta_ij, fo_ij


class FixConp:
    """``fix ID group conp Nevery group2 eta dV logfile [keywords]``."""

    style = "conp"

    def __init__(self, lmp, arg, device=0, rank=0, nranks=1, unique_id=None, owned=None):
        self.lmp = lmp
        s = lmp.system
        self.args = a = parse_fix_args(arg, s.ntypes)
        if a.style != self.style:
            raise FixError(f"fix style {a.style} given to Fix{self.style.capitalize()}")
        if a.group not in lmp.groups:
            raise FixError("Could not find fix group ID")
        if a.group2 not in lmp.groups:
            raise FixError("Fix conp group ID does not exist")  # fix_conp.cpp:106-107
        if a.splitflag:
            raise FixError("Invalid fix conp commmand (unknown option: split)")  # experimental variant not provided
        if a.potdiffstr is not None and a.potdiffstr not in getattr(lmp, "variables", {}):
            raise FixError("Fix conp potential difference variable does not exist")  # :267-269
        self.rank, self.nranks = rank, nranks
        g1, g2 = lmp.groups[a.group], lmp.groups[a.group2]
        side_all = np.where(g1, 1, np.where(g2, -1, 0)).astype(np.int32)  # electrode_check :599-605
        self.one_electrode_flag = bool(np.array_equal(g1, g2))  # :295
        self.side_all = side_all
        self.ele_idx = np.nonzero(side_all != 0)[0]  # eleall order
        self.side = side_all[self.ele_idx]
        self.N = int(self.ele_idx.shape[0])
        oth = np.nonzero(side_all == 0)[0]
        if owned is None:  # contiguous chunk of the non-electrode atoms per rank
            lo = (len(oth) * rank) // nranks
            hi = (len(oth) * (rank + 1)) // nranks
            owned = oth[lo:hi]
        self.owned = np.asarray(owned)
        self.everynum = a.everynum
        self.pairmode = a.pairmode
        self.kappa = 1.0
        self.eta_i = np.zeros(s.ntypes + 1)
        self.u0_i = np.zeros(s.ntypes + 1)
        self.evscale = lmp.qe2f / lmp.qqr2e  # :412
        self.runstage = 0
        self.scalar_output = 0.0
        self.ctx = abi.Context(device, rank, nranks, unique_id)
        self.kspace_mode = abi.KSPACE_PPPM if a.pppmflag else abi.KSPACE_EWALD
        self.log_lines = []

    # -- fix_modify ID ehgo ... (fix_conp.cpp:1482-1515) ------------------------
    def modify_param(self, arg):
        if self.pairmode == PAIR_ETA:
            raise FixError("Can't fix_modify conp parameters in basic pair mode")
        if arg[0] == "ehgo":
            if arg[1] == "kappa":
                if len(arg) != 3:
                    raise FixError("Invalid number of inputs for EHGO coeff setting")
                self.kappa = float(arg[2])
                return 3
            if arg[1] == "coeff":
                if len(arg) != 5:
                    raise FixError("Invalid number of inputs for EHGO coeff setting")
                nt = self.lmp.system.ntypes
                spec = str(arg[2])
                if "*" in spec:
                    a, b = spec.split("*")
                    ilo, ihi = (int(a) if a else 1), (int(b) if b else nt)
                else:
                    ilo = ihi = int(spec)
                eta_one = float(arg[3])
                u0_one = math.sqrt(2.0) / MY_PIS * eta_one / self.evscale if arg[4] == "auto" else float(arg[4])
                if ilo > ihi:
                    raise FixError("Couldn't set EHGO coeffs with mintype more than maxtype")
                self.eta_i[ilo:ihi + 1] = eta_one
                self.u0_i[ilo:ihi + 1] = u0_one * self.evscale
                return 5
            raise FixError("Invalid entry for EHGO coeff setting")
        return 0

    def _potdiff(self):
        a = self.args
        if a.potdiffstr is not None:  # equal-style variable, fix_conp.cpp:1143
            return float(self.lmp.variables[a.potdiffstr]())
        return a.potdiff

    # -- setup hooks -------------------------------------------------------------
    def setup_post_neighbor(self):
        """linalg_init + post_neighbor (fix_conp.cpp:382-385, 393-424)."""
        lmp, a, ctx = self.lmp, self.args, self.ctx
        s = lmp.system
        if self.runstage == 0:
            if lmp.cut_coul is None:
                raise FixError("Fix conp couldn't detect a Coulombic pair style")  # :258
            if a.pppmflag and lmp.kspace_style != "pppm/conp":
                raise FixError("Fix conp couldn't detect a pppm/conp kspace style "
                               "(which is required with the pppm flag)")  # :402-404
            ctx.set_cell(s.boxlo, s.prd, lmp.periodic, lmp.slabflag, lmp.slab_volfactor, a.ff_flag)
            ctx.set_ewald(lmp.g_ewald, lmp.accuracy, lmp.q2(), s.natoms, a.lowmemflag)
            eta_ij = fo_ij = u0 = None
            if self.pairmode == PAIR_EHGO:  # init() :296-299
                t = ehgo_setup_tables(s.ntypes, self.kappa, self.eta_i, self.u0_i)
                if t is None:
                    self.pairmode = PAIR_ETA
                    self.log_lines.append("WARNING: No EHGO settings found, switching back to ETA mode")
                else:
                    eta_ij, fo_ij = t
                    u0 = self.u0_i
            is_eletype = np.zeros(s.ntypes + 1, dtype=np.int32)
            for t_ in a.eletypes:
                is_eletype[t_] = 1
            ctx.set_pair(self.pairmode, a.eta, lmp.cut_coul, s.ntypes, lmp.cutsq, eta_ij, fo_ij, u0, a.smartlist,
                         is_eletype if a.smartlist else None)
            ctx.set_electrodes(s.id[self.ele_idx], s.type[self.ele_idx], self.side, s.x[self.ele_idx])
        self.post_neighbor()

    def post_neighbor(self):
        """FixConp::post_neighbor (fix_conp.cpp:468-539): hand the locally owned
        non-electrode atoms' static data to the device."""
        s = self.lmp.system
        own = self.owned
        self.ctx.post_neighbor(s.q[own], s.type[own])

    def setup_pre_force(self, first_solve=True):
        """kspace->setup + linalg_setup + pre_force (fix_conp.cpp:387-391, 426-464)."""
        a, ctx = self.args, self.ctx
        s = self.lmp.system
        if self.runstage == 0:
            if a.pppmflag:
                # force->kspace->setup() (fix_conp.cpp:388): `pppm/conp` hands its mesh tables over here, i.e.
                # AFTER the first post_neighbor -- the order the LAMMPS shim produces (shim/fix_conp.cpp)
                t = self.lmp.pppm_tables()
                ctx.pppm_setup(t.mesh, t.order, t.rho_coeff, t.greensfn, t.shift, t.shiftone)
            if a.a_matrix_f == 0:
                ctx.build_A()
            else:
                tags, mat = read_matrix_file(a.a_matrix_file, self.N)
                mat = self._permute_to_eleall(tags, mat)
                ctx.load_matrix(mat, is_inverse=(a.a_matrix_f == 2))
            if a.matoutflag and self.rank == 0 and a.a_matrix_f == 0:
                write_amatrix("amatrix", s.id[self.ele_idx], self._full_matrix())
            self.runstage = 1
            ee = ctx.invert_project(a.nullneutralflag, a.zneutrflag, self.one_electrode_flag)
            if not self.one_electrode_flag and a.a_matrix_f < 2:
                self.log_lines.append("conp output: <e,e> = %.8g" % (ee * self.evscale))  # :1006-1009
            if a.matoutflag and self.rank == 0 and a.a_matrix_f < 2:
                write_inv_a_matrix("inv_a_matrix", s.id[self.ele_idx], self._full_matrix())
            qinit = s.q[self.ele_idx].copy() if a.qinitflag else None
            self.totsetq = ctx.set_unit_voltage(self.evscale, qinit, self.one_electrode_flag, a.nullneutralflag,
                                                a.zneutrflag)
            self.log_lines.append("conp output: <d,d> = %.8g" % (-self.totsetq))  # :458-461
            self.ee = ee * self.evscale
            self.dd = -self.totsetq
            self.runstage = 3
            self.post_neighbor()  # the A build reuses the binning buffers
        if first_solve:
            return self.pre_force()
        return None

    def setup(self):
        """setup_post_neighbor + setup_pre_force without the first solve."""
        self.setup_post_neighbor()
        self.setup_pre_force(first_solve=False)

    def _full_matrix(self):
        if self.nranks != 1:
            raise FixError("matout is written by rank 0 from the full matrix; run it on one GPU")
        return self.ctx.get_matrix()

    def _permute_to_eleall(self, tags, mat):
        """a_read adopts the file's tag order as the eleall order
        (fix_conp.cpp:750-759); results per tag are identical if instead the
        matrix is permuted into the current eleall order, which is what we do."""
        cur = self.lmp.system.id[self.ele_idx]
        if np.array_equal(tags, cur):
            return mat
        pos = {int(t): i for i, t in enumerate(tags)}
        try:
            perm = np.array([pos[int(t)] for t in cur])
        except KeyError:
            raise FixError("A matrix file does not list the electrode atoms of this run") from None
        return np.ascontiguousarray(mat[np.ix_(perm, perm)])

    # -- per-step hooks ------------------------------------------------------------
    def _value(self):
        return self._potdiff()

    def pre_force(self):
        """FixConp::pre_force (fix_conp.cpp:543-573): b_cal + update_charge on the
        device, then scatter the new charges to the host's atoms (:1153-1158)."""
        s = self.lmp.system
        q, scalar = self.ctx.pre_force(s.x[self.owned], self.kspace_mode, self.args.variant, self._value())
        s.q[self.ele_idx] = q
        self.q_ele = q
        self.scalar_output = scalar
        return q

    def post_force(self, want_forces=True):
        """force_cal (fix_conp.cpp:1163-1201).  Returns (f[n_owned,3], ecoul,
        eself, virial[6])."""
        f, en = self.ctx.post_force(self.lmp.qqrd2e, want_forces)
        return f, float(en[0]), float(en[1]), en[2:8].copy()

    def compute_scalar(self):
        return self.scalar_output

    def compute_potential_atom(self, eta=None, pair=True, kspace=True, qsum=True):
        """``compute ID <electrode groups> potential/atom [pair] [kspace] eta <eta> <molL> <molR> [noqsum]``
        (compute_potential_atom.cpp:47-182) for the electrode atoms, in volts: the potential the solve is
        supposed to have made constant on each electrode."""
        eta = self.args.eta if eta is None else eta
        phi = self.ctx.electrode_potential(pair, kspace, eta, qsum)
        return phi * (self.lmp.qqr2e / self.lmp.qe2f)   # evscale of the compute, :106

    def close(self):
        self.ctx.close()


class FixConq(FixConp):
    """``fix conq``: value = charge of the right electrode (fix_conq.cpp:41-90)."""
    style = "conq"


class FixCond(FixConp):
    """``fix cond``: constant-D finite-field variant (fix_cond.cpp)."""
    style = "cond"


class FixZmirror:
    """``fix ID group zmirror Nevery group2`` (fix_zmirror.cpp:30-64, 124-220): the atoms of group2 become
    the mirror image, in the plane z = (zlo + zhi)/2, of the atoms of ``group``, matched by tag offset.
    Host-side position plumbing of the doubled-cell decks; mirrors shim/fix_zmirror.cpp."""

    def __init__(self, lmp, arg):
        if len(arg) != 5:
            raise FixError("Illegal fix zmirror command (incorrect no. of parameters)")
        if arg[1] not in lmp.groups:
            raise FixError("Could not find fix group ID")
        if arg[4] not in lmp.groups:
            raise FixError("Fix zmirror group ID does not exist")
        self.lmp, self.everynum = lmp, int(arg[3])
        self.g1, self.g2 = lmp.groups[arg[1]], lmp.groups[arg[4]]

    def setup(self):
        tag = self.lmp.system.id
        t1, t2 = tag[self.g1], tag[self.g2]
        self.send_mintag, self.recv_mintag = int(t1.min()), int(t2.min())
        if int(t1.max()) - self.send_mintag != int(t2.max()) - self.recv_mintag:
            raise FixError("Groups do not have same number of tags")

    def post_integrate(self, ntimestep=0):
        if ntimestep % self.everynum:
            return
        s = self.lmp.system
        zoffset = 2.0 * s.boxlo[2] + s.prd[2]
        src = np.nonzero(self.g1)[0]
        dst = np.nonzero(self.g2)[0]
        table = {int(s.id[i]) - self.send_mintag: s.x[i].copy() for i in src}
        if len(table) != len(dst):
            raise FixError("Incorrect number of atoms communicated")
        for i in dst:
            x = table[int(s.id[i]) - self.recv_mintag]
            s.x[i] = (x[0], x[1], zoffset - x[2])


def make_fix(lmp, arg, **kw):
    """``fix`` command dispatch on the style token (FixStyle(conp|conq|cond))."""
    style = str(arg[2])
    if style == "zmirror":
        return FixZmirror(lmp, arg)
    cls = {"conp": FixConp, "conq": FixConq, "cond": FixCond}.get(style)
    if cls is None:
        raise FixError(f"Unknown fix style {style}")
    return cls(lmp, arg, **kw)


# -- matout / org / inv text formats (fix_conp.cpp:721-773, 833-849, 960-977) --------

def write_amatrix(path, tags, mat):
    with open(path, "w") as fh:
        fh.write(" " + "".join("%20d" % int(t) for t in tags) + "\n")
        for row in mat:
            fh.write(" " + "".join("%20.12f" % v for v in row) + "\n")


def write_inv_a_matrix(path, tags, mat):
    n = mat.shape[0]
    with open(path, "w") as fh:
        fh.write(" " + "".join("%20d" % int(t) for t in tags) + "\n")
        for row in mat:
            fh.write(" ".join("%20.10f" % v for v in row) + "\n")
    return n


def read_matrix_file(path, n):
    """FixConp::a_read tokenizer (fix_conp.cpp:729-748)."""
    try:
        with open(path) as fh:
            toks = fh.read().split()
    except OSError:
        raise FixError("Invalid fix conp command (Cannot open A matrix file)") from None
    if len(toks) > n + n * n:
        raise FixError("Too many entries in A matrix file")
    if len(toks) < n + n * n:
        raise FixError("Too few entries in A matrix file")
    tags = np.array([int(t) for t in toks[:n]], dtype=np.int32)
    mat = np.array([float(t) for t in toks[n:]], dtype=np.float64).reshape(n, n)
    return tags, mat
