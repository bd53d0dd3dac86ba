"""The LAMMPS-side shim classes (lammps-user-conp2_b200/shim: FixConpB200 / FixConqB200 / PPPMCONPB200) compiled
against the behaving single-rank mock of the LAMMPS classes (tests/lammps_stubs) and RUN on the GPU through the
hook order of a real run: init -> setup_post_neighbor -> setup_pre_force (which calls kspace->setup(), i.e.
conp_pppm_setup AFTER the first conp_post_neighbor) -> post_force -> PPPM make_rho -> one more step with moved
atoms.  What a LAMMPS user would see (atom->q, the fix scalar, forces, energies, the density brick the force
PPPM transforms, the fix's log file) is compared with the CPU oracle.  No LAMMPS tree exists on this machine,
so this is as close to the reference's own decks (tests/dilute/input, tests/il_twolayer/input) as the shim
gets executed."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import conp_oracle as O
from cases import dilute, synthetic

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "lammps-user-conp2_b200", "shim")
STUBS = os.path.join(ROOT, "tests", "lammps_stubs")


@pytest.fixture(scope="module")
def mock_lammps(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    libdir = os.path.join(ROOT, "lammps-user-conp2_b200")
    exe = str(tmp_path_factory.mktemp("mock") / "mock_lammps")
    cmd = [gxx, "-std=c++17", "-O1", "-Wall", "-Werror", "-I", STUBS, "-I", os.path.join(ROOT, "include"), "-I", SHIM,
           os.path.join(STUBS, "mock_lammps_main.cpp"), os.path.join(SHIM, "fix_conp.cpp"),
           os.path.join(SHIM, "pppm_conp.cpp"), "-L", libdir, "-lconp_b200", f"-Wl,-rpath,{libdir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return exe


def write_blob(path, **arrays):
    with open(path, "wb") as fh:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            kind = 0 if a.dtype.kind in "iu" else 1
            a = a.astype(np.int32 if kind == 0 else np.float64).reshape(-1)
            nb = name.encode()
            fh.write(struct.pack("<i", len(nb)) + nb + struct.pack("<iq", kind, a.size) + a.tobytes())


def read_blob(path):
    out, data, o = {}, open(path, "rb").read(), 0
    while o < len(data):
        (nl,) = struct.unpack_from("<i", data, o); o += 4
        name = data[o:o + nl].decode(); o += nl
        kind, cnt = struct.unpack_from("<iq", data, o); o += 12
        dt = np.int32 if kind == 0 else np.float64
        out[name] = np.frombuffer(data, dtype=dt, count=cnt, offset=o).copy(); o += cnt * dt().itemsize
    return out


def system_blob(lmp, x2=None, **extra):
    s = lmp.system
    g1, g2 = lmp.groups["eleleft"], lmp.groups["eleright"]
    d = dict(natoms=[s.natoms], ntypes=[s.ntypes], tag=s.id, type=s.type, mask=1 + 2 * g1.astype(int) + 4 * g2.astype(int),
             q=s.q, x=s.x, boxlo=s.boxlo, prd=s.prd, periodic=lmp.periodic, cut_coul=[lmp.cut_coul], cutsq=lmp.cutsq,
             kspace_is_conp=[int(lmp.kspace_style == "pppm/conp")], g_ewald=[lmp.g_ewald], accuracy=[lmp.accuracy],
             slabflag=[lmp.slabflag], slab_volfactor=[lmp.slab_volfactor])
    if lmp.kspace_style == "pppm/conp":
        t = lmp.pppm_tables()
        d.update(mesh=t.mesh, order=[t.order], rho_coeff=t.rho_coeff, greensfn=t.greensfn, shift=[t.shift],
                 shiftone=[t.shiftone])
    if x2 is not None:
        d["x2"] = x2
    d.update(extra)
    return d


def run_case(exe, tmp_path, case, kw=None, **extra):
    kw = kw or {}
    lmp, arg = case(**kw)
    lmp2, arg2 = case(**kw)
    ref = O.OracleFixConp(lmp2, arg2)
    ref.setup()
    rng = np.random.default_rng(17)
    x2 = lmp.system.x.copy()
    x2[ref.oth_idx] += rng.normal(0.0, 0.05, (len(ref.oth_idx), 3))
    sysf, outf = str(tmp_path / "system.bin"), str(tmp_path / "out.bin")
    write_blob(sysf, **system_blob(lmp, x2=x2, **extra))
    r = subprocess.run([exe, sysf, outf] + [str(a) for a in arg], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = read_blob(outf)
    steps = []
    for k, x in enumerate((lmp.system.x, x2)):
        lmp2.system.x[:] = x
        q = ref.pre_force().copy()
        f, ecoul, eself, vir = ref.post_force()
        steps.append(dict(q=q, scalar=ref.scalar_output, f=f, ecoul=ecoul, eself=eself,
                          rho=(ref.elyte_density + ref.ele_density) if ref.pppm is not None else None))
    return lmp, ref, out, steps, r.stdout, tmp_path / arg[7]


def check(ref, out, steps):
    for k, st in enumerate(steps):
        q = out[f"q{k}"][ref.ele_idx]
        assert np.abs(q - st["q"]).max() <= 1e-9 * np.abs(st["q"]).max() + 1e-12      # charges scattered to atom->q
        assert abs(out[f"scalar{k}"][0] - st["scalar"]) <= 1e-9 * abs(st["scalar"]) + 1e-12
        f = out[f"f{k}"].reshape(-1, 3)
        assert np.abs(f[ref.oth_idx] - st["f"]).max() <= 1e-8 * max(np.abs(st["f"]).max(), 1e-300)
        assert np.abs(f[ref.ele_idx]).max() == 0.0
        assert abs((out[f"kspace_energy{k}"][0] - 1.0) - st["eself"]) <= 1e-8 * abs(st["eself"])
        assert abs(out[f"eng_coul{k}"][0] - st["ecoul"]) <= 1e-8 * abs(st["ecoul"]) + 1e-12
        if st["rho"] is not None:   # make_rho override: the density the force PPPM transforms
            assert np.abs(out[f"density{k}"] - st["rho"]).max() <= 1e-9 * np.abs(st["rho"]).max()
            assert out[f"ghost{k}"][0] == 0.0


def test_dilute_ffield_ewald_through_the_shim(mock_lammps, tmp_path):
    lmp, ref, out, steps, stdout, log = run_case(mock_lammps, tmp_path, dilute, dict(n=2))
    check(ref, out, steps)
    assert "conp output: <e,e> = %.8g" % ref.ee in stdout and "conp output: <d,d> = %.8g" % ref.dd in stdout
    lines = open(log).read()
    for key in ("A matrix calculating ...", "A matrix calculation time  =", "B vector calculation time =",
                "Coulomb calculation time =", "Kspace calculation time ="):      # fix_conp.cpp:553-568, 787, 857
        assert key in lines


def test_pppm_conp_kspace_through_the_shim(mock_lammps, tmp_path):
    """`pppm` keyword: FixConpB200 finds the pppm/conp style, attaches, and kspace->setup() (called from
    setup_pre_force, AFTER the first post_neighbor) hands the mesh tables to the library."""
    lmp, ref, out, steps, stdout, log = run_case(mock_lammps, tmp_path, synthetic, dict(name="small", h=1.25, accuracy=1e-4))
    check(ref, out, steps)
    assert out["kspace_setups"][0] >= 1


def test_conq_and_variable_potential(mock_lammps, tmp_path):
    def case():
        lmp, arg = dilute(1)
        arg[2], arg[6] = "conq", "0.03"
        return lmp, arg
    lmp, ref, out, steps, _, _ = run_case(mock_lammps, tmp_path, case)
    check(ref, out, steps)

    def case_v():
        return dilute(0)
    lmp, arg = case_v()
    sysf, outf = str(tmp_path / "s2.bin"), str(tmp_path / "o2.bin")
    write_blob(sysf, **system_blob(lmp, variable_value=[1.0]))
    arg_v = list(arg)
    arg_v[6] = "v_dv"                                   # equal-style variable, fix_conp.cpp:1143
    r = subprocess.run([mock_lammps, sysf, outf] + arg_v, capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    lmp2, arg2 = case_v()
    ref2 = O.OracleFixConp(lmp2, arg2)
    ref2.setup()
    q = ref2.pre_force()
    assert np.abs(read_blob(outf)["q0"][ref2.ele_idx] - q).max() <= 1e-9 * np.abs(q).max() + 1e-12


def test_pppm_keyword_without_the_kspace_style_is_the_reference_error(mock_lammps, tmp_path):
    lmp, arg = dilute(2)                                # plain pppm kspace style ...
    sysf, outf = str(tmp_path / "s.bin"), str(tmp_path / "o.bin")
    write_blob(sysf, **system_blob(lmp))
    r = subprocess.run([mock_lammps, sysf, outf] + list(arg) + ["pppm"], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1                            # ... with the pppm keyword: fix_conp.cpp:402-404
    assert "couldn't detect a pppm/conp kspace style" in r.stderr
