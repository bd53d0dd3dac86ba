"""ctypes binding of libconp_b200.so (include/conp_b200.h).

This is the only way the Python host mirror reaches the compute path; if the
shared library (or a CUDA device) is missing the calls fail loudly -- there
is no eager/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "libconp_b200.so"))

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)

CONP_UNIQUE_ID_BYTES = 128
KSPACE_EWALD, KSPACE_PPPM = 0, 1

# every symbol include/conp_b200.h declares (checked by tests/test_abi_symbols.py)
SYMBOLS = [
    "conp_abi_version", "conp_device_count", "conp_get_unique_id", "conp_create", "conp_destroy", "conp_last_error", "conp_get_info",
    "conp_set_cell", "conp_set_ewald", "conp_set_pair", "conp_set_electrodes", "conp_pppm_setup", "conp_build_A",
    "conp_load_matrix", "conp_get_matrix", "conp_invert_project", "conp_set_unit_voltage", "conp_post_neighbor",
    "conp_pre_force", "conp_solve_device", "conp_get_charges", "conp_get_b", "conp_get_density",
    "conp_get_density_region", "conp_mesh_potential", "conp_electrode_potential",
    "conp_get_potential_brick", "conp_post_force", "conp_stream", "conp_sync", "conp_timer_record",
    "conp_timer_elapsed_ms", "conp_stage_times", "conp_bench_gemv", "conp_bench_dgemm_tflops",
    "conp_matvec", "conp_plan_symv", "conp_plan_spread", "conp_row_block", "conp_plan_pair_runs", "conp_plan_zconv",
    "conp_plan_sweep",
]


class ConpInfo(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int), ("device", C.c_int), ("rank", C.c_int), ("nranks", C.c_int),
        ("n_ele", C.c_int), ("row_begin", C.c_int), ("row_end", C.c_int), ("n_elyte", C.c_int),
        ("kxmax", C.c_int), ("kymax", C.c_int), ("kzmax", C.c_int),
        ("kcount", C.c_int), ("kcount_flat", C.c_int), ("kcount_expand", C.c_int),
        ("mesh", C.c_int * 3), ("order", C.c_int),
        ("matrix_pitch", C.c_longlong), ("launches", C.c_longlong),
        ("setup_build_ms", C.c_double), ("setup_invert_ms", C.c_double),
        ("ee", C.c_double), ("dd", C.c_double), ("totsetq", C.c_double),
        ("symmetric_matvec", C.c_int), ("reserved0", C.c_int), ("asymmetry", C.c_double),
    ]


def plan_symv(n, row0, nrows, num_sms=148):
    """Strip decomposition of the symmetric matvec (host-only entry point): (strips[nstrips][2], L) or None."""
    L = load_library()
    ns, sl = C.c_int(0), C.c_int(0)
    cap = max(num_sms, nrows // 256 + 2, 1)
    out = np.zeros((cap, 2), dtype=np.int32)
    rc = L.conp_plan_symv(int(n), int(row0), int(nrows), int(num_sms), cap, out.ctypes.data_as(c_ip), C.byref(ns),
                          C.byref(sl))
    if rc != 0:
        raise RuntimeError(f"conp_plan_symv: status {rc}")
    if ns.value < 0:
        return None
    return out[:ns.value].copy(), sl.value


def plan_spread(mesh, order, shift, boxlo, prd, periodic, slab_volfactor, rc, zin_lo, nzi, zs_lo, zs_n, num_sms=148):
    """Tile decomposition of the owner-computes PPPM spread (host-only entry point).  Returns a dict with the
    tile geometry, the sort-cell grid and, per tile, the list of [c0, c1) candidate cell ranges."""
    L = load_library()
    m, lo, pr, pe = i32(mesh), f64(boxlo), f64(prd), i32(periodic)
    geom = np.zeros(12, dtype=np.int32)
    nt, nr = C.c_int(0), C.c_int(0)
    args = (_ip(m), int(order), float(shift), _dp(lo), _dp(pr), _ip(pe), float(slab_volfactor), float(rc), int(zin_lo),
            int(nzi), int(zs_lo), int(zs_n), int(num_sms))
    rc_ = L.conp_plan_spread(*args, _ip(geom), None, 0, None, 0, C.byref(nt), C.byref(nr))
    if rc_:
        raise RuntimeError(f"conp_plan_spread: status {rc_}")
    rs = np.zeros(nt.value + 1, dtype=np.int32)
    rr = np.zeros((max(nr.value, 1), 2), dtype=np.int32)
    L.conp_plan_spread(*args, _ip(geom), _ip(rs), nt.value, _ip(rr), nr.value, C.byref(nt), C.byref(nr))
    keys = ("tz", "ty", "tx", "ntz", "nty", "ntx", "halo_z", "halo_y", "halo_x", "ncx", "ncy", "ncz")
    out = {k: int(v) for k, v in zip(keys, geom)}
    out.update(ntiles=nt.value, run_start=rs, runs=rr[:nr.value])
    return out


def plan_sweep(mesh, order, nzi, zs_lo, zs_n, num_sms=148):
    """Work plan of the z-sweep spread (host-only entry point)."""
    L = load_library()
    m = i32(mesh)
    geom = np.zeros(8, dtype=np.int32)
    n = C.c_int(0)
    args = (_ip(m), int(order), int(nzi), int(zs_lo), int(zs_n), int(num_sms))
    st = L.conp_plan_sweep(*args, _ip(geom), None, 0, C.byref(n))
    if st:
        raise RuntimeError(f"conp_plan_sweep: status {st}")
    items = np.zeros((max(n.value, 1), 3), dtype=np.int32)
    L.conp_plan_sweep(*args, _ip(geom), _ip(items), n.value, C.byref(n))
    keys = ("usable", "ncolx", "ncoly", "pz_lo", "npz", "wrap_z", "nbins", "grid")
    out = {k: int(v) for k, v in zip(keys, geom)}
    out["items"] = items[:n.value]
    return out


def plan_zconv(ncol, nz, nzi, zs_lo, nzl, zin_lo, krad, zout, real_kernel=True):
    """Work plan of the windowed z-convolution (host-only entry point): dict with the narrow groups (c0, rblock,
    intervals [(lo, hi, base)], np), the wide columns, the output planes in compact coordinates, rcap and npcap."""
    L = load_library()
    kr, zo = i32(krad), i32(zout)
    ng, nw = C.c_int(0), C.c_int(0)
    caps = np.zeros(2, dtype=np.int32)
    args = (int(ncol), int(nz), int(nzi), int(zs_lo), int(nzl), int(zin_lo), _ip(kr), int(zo.size), _ip(zo),
            int(bool(real_kernel)))
    st = L.conp_plan_zconv(*args, None, 0, None, 0, None, _ip(caps), C.byref(ng), C.byref(nw))
    if st:
        raise RuntimeError(f"conp_plan_zconv: status {st}")
    groups = np.zeros((max(ng.value, 1), 32), dtype=np.int32)
    wide = np.zeros(max(nw.value, 1), dtype=np.int32)
    aout = np.zeros(max(zo.size, 1), dtype=np.int32)
    L.conp_plan_zconv(*args, _ip(groups), ng.value, _ip(wide), nw.value, _ip(aout), _ip(caps), C.byref(ng),
                      C.byref(nw))
    out = []
    for g in groups[:ng.value]:
        nint = int(g[2])
        out.append(dict(c0=int(g[0]), rblock=int(g[1]), np=int(g[3]),
                        intervals=[(int(g[4 + i]), int(g[12 + i]), int(g[20 + i])) for i in range(nint)]))
    return dict(narrow=out, wide=wide[:nw.value].copy(), aout=aout[:zo.size].copy(), rcap=int(caps[0]),
                npcap=int(caps[1]))


def plan_pair_runs(boxlo, prd, periodic, rc, xyz):
    """Static candidate list of the pair kernels (host-only entry point): (nc[3], run_start[n+1], runs[nruns,5])
    with runs = (c0, c1, sx, sy, sz)."""
    L = load_library()
    lo, pr, pe = f64(boxlo), f64(prd), i32(periodic)
    x = f64(np.ascontiguousarray(xyz).reshape(-1))
    n = x.size // 3
    nc = np.zeros(3, dtype=np.int32)
    nr = C.c_int(0)
    st = L.conp_plan_pair_runs(_dp(lo), _dp(pr), _ip(pe), float(rc), n, _dp(x), _ip(nc), None, None, 0, C.byref(nr))
    if st:
        raise RuntimeError(f"conp_plan_pair_runs: status {st}")
    rs = np.zeros(n + 1, dtype=np.int32)
    rr = np.zeros((max(nr.value, 1), 5), dtype=np.int32)
    L.conp_plan_pair_runs(_dp(lo), _dp(pr), _ip(pe), float(rc), n, _dp(x), _ip(nc), _ip(rs), _ip(rr), nr.value,
                          C.byref(nr))
    return nc, rs, rr[:nr.value]


def row_block(n_ele, nranks, rank):
    """(row_begin, row_end, rows_per_rank) of the library's row partition (host-only entry point)."""
    L = load_library()
    a, b, r = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = L.conp_row_block(int(n_ele), int(nranks), int(rank), C.byref(a), C.byref(b), C.byref(r))
    if rc:
        raise RuntimeError(f"conp_row_block: status {rc}")
    return a.value, b.value, r.value


class ConpError(RuntimeError):
    """A non-zero status from the C ABI (the LAMMPS shim raises error->all)."""

    def __init__(self, code, msg):
        super().__init__(f"[conp_b200 status {code}] {msg}")
        self.code = code
        self.msg = msg


_lib = None


def load_library(path: str | None = None):
    """dlopen libconp_b200.so; raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ConpError(-1, f"{p} not found: build it with `make -C lammps-user-conp2_b200/csrc` "
                            "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(p, mode=C.RTLD_GLOBAL)
    vp = C.c_void_p
    L.conp_abi_version.restype = C.c_int
    L.conp_get_unique_id.argtypes = [vp]
    L.conp_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp]
    L.conp_destroy.argtypes = [vp]
    L.conp_destroy.restype = None
    L.conp_last_error.argtypes = [vp]
    L.conp_last_error.restype = C.c_char_p
    L.conp_get_info.argtypes = [vp, C.POINTER(ConpInfo)]
    L.conp_set_cell.argtypes = [vp, c_dp, c_dp, c_ip, C.c_int, C.c_double, C.c_int]
    L.conp_set_ewald.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_longlong, C.c_int]
    L.conp_set_pair.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_int, c_dp, c_dp, c_dp, c_dp, C.c_int, c_ip]
    L.conp_set_electrodes.argtypes = [vp, C.c_int, c_ip, c_ip, c_ip, c_dp]
    L.conp_pppm_setup.argtypes = [vp, c_ip, C.c_int, c_dp, c_dp, C.c_double, C.c_double]
    L.conp_build_A.argtypes = [vp]
    L.conp_load_matrix.argtypes = [vp, c_dp, C.c_int]
    L.conp_get_matrix.argtypes = [vp, c_dp]
    L.conp_invert_project.argtypes = [vp, C.c_int, C.c_int, C.c_int, c_dp]
    L.conp_set_unit_voltage.argtypes = [vp, C.c_double, c_dp, C.c_int, C.c_int, C.c_int, c_dp]
    L.conp_post_neighbor.argtypes = [vp, C.c_int, c_dp, c_ip, c_ip, C.c_int]
    L.conp_device_count.argtypes = [c_ip]
    L.conp_pre_force.argtypes = [vp, c_dp, C.c_int, C.c_int, C.c_double, c_dp, c_dp]
    L.conp_solve_device.argtypes = [vp, vp, C.c_int, C.c_int, C.c_double]
    L.conp_get_charges.argtypes = [vp, c_dp, c_dp]
    L.conp_get_b.argtypes = [vp, c_dp, c_dp]
    L.conp_get_density.argtypes = [vp, C.c_int, c_dp]
    L.conp_get_potential_brick.argtypes = [vp, c_dp]
    L.conp_get_density_region.argtypes = [vp, C.c_int, c_ip, c_ip, c_dp]
    L.conp_mesh_potential.argtypes = [vp, C.c_int, c_dp, c_dp]
    L.conp_electrode_potential.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int, c_dp]
    L.conp_post_force.argtypes = [vp, C.c_double, c_dp, c_dp]
    L.conp_stream.argtypes = [vp]
    L.conp_stream.restype = vp
    L.conp_sync.argtypes = [vp]
    L.conp_timer_record.argtypes = [vp, C.c_int]
    L.conp_timer_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.conp_stage_times.argtypes = [vp, C.c_int, c_dp]
    L.conp_bench_gemv.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    L.conp_bench_dgemm_tflops.argtypes = [vp, C.c_int, c_dp]
    L.conp_matvec.argtypes = [vp, c_dp, c_dp]
    L.conp_plan_symv.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, C.POINTER(C.c_int),
                                 C.POINTER(C.c_int)]
    L.conp_plan_spread.argtypes = [c_ip, C.c_int, C.c_double, c_dp, c_dp, c_ip, C.c_double, C.c_double, C.c_int,
                                   C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, C.c_int, c_ip, C.c_int,
                                   C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.conp_row_block.argtypes = [C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip]
    L.conp_plan_sweep.argtypes = [c_ip, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, C.c_int,
                                  C.POINTER(C.c_int)]
    L.conp_plan_zconv.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, C.c_int, c_ip, C.c_int,
                                  c_ip, C.c_int, c_ip, C.c_int, c_ip, c_ip, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.conp_plan_pair_runs.argtypes = [c_dp, c_dp, c_ip, C.c_double, C.c_int, c_dp, c_ip, c_ip, c_ip, C.c_int,
                                      C.POINTER(C.c_int)]
    if path is None:
        _lib = L
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def get_unique_id() -> bytes:
    L = load_library()
    buf = C.create_string_buffer(CONP_UNIQUE_ID_BYTES)
    rc = L.conp_get_unique_id(buf)
    if rc:
        raise ConpError(rc, L.conp_last_error(None).decode())
    return buf.raw


class Context:
    """Thin OO wrapper over conp_ctx; one per GPU."""

    def __init__(self, device=0, rank=0, nranks=1, unique_id: bytes | None = None):
        self.L = load_library()
        self.h = C.c_void_p()
        uid = C.create_string_buffer(unique_id, CONP_UNIQUE_ID_BYTES) if unique_id else None
        rc = self.L.conp_create(C.byref(self.h), device, rank, nranks, uid)
        if rc:
            self.h = None
            raise ConpError(rc, self.L.conp_last_error(None).decode())
        self.rank, self.nranks, self.device = rank, nranks, device
        self.n_ele = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.conp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise ConpError(rc, self.L.conp_last_error(self.h).decode())

    def info(self) -> ConpInfo:
        o = ConpInfo()
        self._ck(self.L.conp_get_info(self.h, C.byref(o)))
        return o

    # setup ---------------------------------------------------------------
    def set_cell(self, boxlo, prd, periodic, slabflag, slab_volfactor, ff_flag):
        a, b, p = f64(boxlo), f64(prd), i32(periodic)
        self._ck(self.L.conp_set_cell(self.h, _dp(a), _dp(b), _ip(p), int(slabflag), float(slab_volfactor),
                                      int(ff_flag)))

    def set_ewald(self, g_ewald, accuracy_abs, q2, natoms, lowmem=True):
        self._ck(self.L.conp_set_ewald(self.h, float(g_ewald), float(accuracy_abs), float(q2), int(natoms),
                                       int(bool(lowmem))))

    def set_pair(self, pairmode, eta, cut_coul, ntypes, cutsq, eta_ij=None, fo_ij=None, u0_i=None, smartlist=False,
                 is_eletype=None):
        cs, e, f, u, t = f64(cutsq), f64(eta_ij), f64(fo_ij), f64(u0_i), i32(is_eletype)
        self._ck(self.L.conp_set_pair(self.h, int(pairmode), float(eta), float(cut_coul), int(ntypes),
                                      _dp(cs.reshape(-1)), _dp(e), _dp(f), _dp(u), int(bool(smartlist)), _ip(t)))

    def set_electrodes(self, tag, typ, side, xyz):
        t, ty, s, x = i32(tag), i32(typ), i32(side), f64(xyz)
        self.n_ele = int(ty.shape[0])
        self._ck(self.L.conp_set_electrodes(self.h, self.n_ele, _ip(t), _ip(ty), _ip(s), _dp(x)))

    def pppm_setup(self, mesh, order, rho_coeff, greensfn, shift, shiftone):
        m, r, g = i32(mesh), f64(rho_coeff), f64(greensfn)
        self._ck(self.L.conp_pppm_setup(self.h, _ip(m), int(order), _dp(r), _dp(g), float(shift), float(shiftone)))
        self.ngrid = int(np.prod(m))
        self.mesh = tuple(int(v) for v in m)

    def build_A(self):
        self._ck(self.L.conp_build_A(self.h))

    def load_matrix(self, full, is_inverse):
        m = f64(full)
        if m.shape != (self.n_ele, self.n_ele):
            raise ConpError(1, "Too few entries in A matrix file" if m.size < self.n_ele ** 2
                            else "Too many entries in A matrix file")
        self._ck(self.L.conp_load_matrix(self.h, _dp(m), int(bool(is_inverse))))

    def get_matrix(self):
        i = self.info()
        out = np.zeros((i.row_end - i.row_begin, self.n_ele))
        self._ck(self.L.conp_get_matrix(self.h, _dp(out)))
        return out

    def invert_project(self, nullneutral=True, zneutr=False, one_electrode=False):
        ee = C.c_double(0)
        self._ck(self.L.conp_invert_project(self.h, int(bool(nullneutral)), int(bool(zneutr)),
                                            int(bool(one_electrode)), C.byref(ee)))
        return ee.value

    def set_unit_voltage(self, evscale, q_init=None, one_electrode=False, nullneutral=True, zneutr=False):
        t = C.c_double(0)
        qi = f64(q_init)
        self._ck(self.L.conp_set_unit_voltage(self.h, float(evscale), _dp(qi), int(bool(one_electrode)),
                                              int(bool(nullneutral)), int(bool(zneutr)), C.byref(t)))
        return t.value

    # per step ------------------------------------------------------------------
    def post_neighbor(self, q, typ, mask=None, ele_bits=0):
        qq, tt, mm = f64(q), i32(typ), i32(mask)
        self._nlocal = int(qq.shape[0])
        self._ck(self.L.conp_post_neighbor(self.h, self._nlocal, _dp(qq), _ip(tt), _ip(mm), int(ele_bits)))

    def pre_force(self, x, kspace_mode, variant, value):
        xx = f64(x)
        q = np.zeros(self.n_ele)
        sc = C.c_double(0)
        self._ck(self.L.conp_pre_force(self.h, _dp(xx), int(kspace_mode), int(variant), float(value), _dp(q),
                                       C.byref(sc)))
        return q, sc.value

    def pre_force_into(self, x_ptr, kspace_mode, variant, value, q_out_ptr):
        """Raw-pointer variant for pinned torch buffers (bench e2e leg)."""
        sc = C.c_double(0)
        self._ck(self.L.conp_pre_force(self.h, C.cast(x_ptr, c_dp), int(kspace_mode), int(variant), float(value),
                                       C.cast(q_out_ptr, c_dp), C.byref(sc)))
        return sc.value

    def solve_device(self, x_device_ptr, kspace_mode, variant, value):
        self._ck(self.L.conp_solve_device(self.h, C.c_void_p(x_device_ptr), int(kspace_mode), int(variant),
                                          float(value)))

    def get_charges(self):
        q = np.zeros(self.n_ele)
        sc = C.c_double(0)
        self._ck(self.L.conp_get_charges(self.h, _dp(q), C.byref(sc)))
        return q, sc.value

    def get_b(self):
        b, bk = np.zeros(self.n_ele), np.zeros(self.n_ele)
        self._ck(self.L.conp_get_b(self.h, _dp(b), _dp(bk)))
        return b, bk

    def get_density(self, which):
        out = np.zeros(self.ngrid)
        self._ck(self.L.conp_get_density(self.h, int(which), _dp(out)))
        return out

    def get_density_region(self, which, lo, hi, out=None):
        """Density on the sub-brick lo..hi (inclusive mesh indices x, y, z); returns [nz][ny][nx].  `out`: a
        caller-owned C-contiguous float64 array of that shape (e.g. pinned memory, like a registered LAMMPS brick)."""
        lo_, hi_ = i32(lo), i32(hi)
        shape = tuple(int(hi_[a] - lo_[a] + 1) for a in (2, 1, 0))
        if out is None:
            out = np.zeros(shape)
        elif out.shape != shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("get_density_region: out must be a C-contiguous float64 array of the region's shape")
        self._ck(self.L.conp_get_density_region(self.h, int(which), _ip(lo_), _ip(hi_), _dp(out)))
        return out

    def mesh_potential(self, xyz):
        x = f64(xyz).reshape(-1, 3)
        out = np.zeros(x.shape[0])
        self._ck(self.L.conp_mesh_potential(self.h, int(x.shape[0]), _dp(x), _dp(out)))
        return out

    def electrode_potential(self, pair=True, kspace=True, eta=0.0, qsum=True):
        out = np.zeros(self.n_ele)
        self._ck(self.L.conp_electrode_potential(self.h, int(bool(pair)), int(bool(kspace)), float(eta),
                                                 int(bool(qsum)), _dp(out)))
        return out

    def get_potential_brick(self):
        out = np.zeros(self.ngrid)
        self._ck(self.L.conp_get_potential_brick(self.h, _dp(out)))
        return out

    def post_force(self, qqrd2e, want_forces=True):
        f = np.zeros((self._nlocal, 3)) if want_forces else None
        en = np.zeros(8)
        self._ck(self.L.conp_post_force(self.h, float(qqrd2e), _dp(f), _dp(en)))
        return f, en

    # instrumentation -------------------------------------------------------------
    def stream(self) -> int:
        return int(self.L.conp_stream(self.h) or 0)

    def sync(self):
        self._ck(self.L.conp_sync(self.h))

    def timer_record(self, slot):
        self._ck(self.L.conp_timer_record(self.h, int(slot)))

    def timer_elapsed_ms(self, a, b) -> float:
        ms = C.c_float(0)
        self._ck(self.L.conp_timer_elapsed_ms(self.h, int(a), int(b), C.byref(ms)))
        return ms.value

    def stage_times(self, enable=True):
        out = np.zeros(8)
        n = self.L.conp_stage_times(self.h, int(bool(enable)), _dp(out))
        return n, out

    def bench_gemv(self, reps=20) -> float:
        ms = C.c_float(0)
        self._ck(self.L.conp_bench_gemv(self.h, int(reps), C.byref(ms)))
        return ms.value

    def matvec(self, v) -> np.ndarray:
        """S.v through the kernel the step uses (symmetric or general)."""
        v = np.ascontiguousarray(v, dtype=np.float64)
        out = np.zeros_like(v)
        self._ck(self.L.conp_matvec(self.h, _dp(v), _dp(out)))
        return out

    def bench_dgemm_tflops(self, n=8192) -> float:
        t = C.c_double(0)
        self._ck(self.L.conp_bench_dgemm_tflops(self.h, int(n), C.byref(t)))
        return t.value
