"""The LAMMPS-side shim (lammps-user-conp2_b200/shim) is shipped as source because no LAMMPS tree is
available here.  This test keeps it honest against the C ABI: both translation units must pass
`g++ -fsyntax-only -Wall -Werror` with include/conp_b200.h and minimal stand-ins for the LAMMPS
headers they include (tests/lammps_stubs: signatures only), so a change of an entry point's
signature or of conp_info that the shim does not follow fails on the CPU suite."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "lammps-user-conp2_b200", "shim")


@pytest.mark.parametrize("unit", ["fix_conp.cpp", "pppm_conp.cpp", "fix_zmirror.cpp"])
def test_shim_unit_compiles_against_the_abi(unit):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    cmd = [gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "lammps_stubs"),
           "-I", os.path.join(ROOT, "include"), "-I", SHIM, os.path.join(SHIM, unit)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_shim_registers_the_reference_style_names():
    """FixStyle(conp|conq|cond, ...) and KSpaceStyle(pppm/conp, ...) as in fix_conp.h:19-21, fix_conq.h:21,
    fix_cond.h:21, pppm_conp.h:19-21 of the reference."""
    fix_h = open(os.path.join(SHIM, "fix_conp.h")).read()
    for name in ("conp", "conq", "cond"):
        assert f"FixStyle({name}," in fix_h
    assert "KSpaceStyle(pppm/conp," in open(os.path.join(SHIM, "pppm_conp.h")).read()
    assert "FixStyle(zmirror," in open(os.path.join(SHIM, "fix_zmirror.h")).read()   # fix_zmirror.h:16
