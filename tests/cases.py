"""The reference's own test matrix (tests/dilute/input, tests/il_onelayer/input,
tests/il_twolayer/input, tests/cond/input) expressed against MockLammps.
g_ewald / mesh are explicit because LAMMPS' auto-tuning is unavailable
(dilute: persist.log:112-113; il: SURVEY.md 8d)."""
import json
import os

from conp_b200 import MockLammps, load_reference_case, make_workload

PINS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_pins.json")))


def dilute(n, pppm=False):
    """tests/dilute/input trial n: 0 slab, 1 slab etypes, 2 ffield etypes,
    3 noslab zneutr sym, 4 noslab zneutr anti, 5 ffield (no etypes)."""
    s = load_reference_case("dilute")
    if n in (3, 4):
        s = s.doubled_cell(sym=(n == 3), molleft=81, molright=82, molmax=82)
    lmp = MockLammps(s, "p p f" if n <= 1 else "p p p")
    lmp.pair_style_coul_long(4.0)
    mesh = None
    if pppm:
        mesh = (27, 24, 432) if n <= 1 else ((27, 24, 288) if n in (3, 4) else tuple(PINS["mesh"]))
    lmp.kspace("pppm/conp" if pppm else "pppm", 1e-6, PINS["g_ewald"], slab=3.0 if n <= 1 else None, mesh=mesh)
    lmp.group_molecule("eleleft", 81)
    lmp.group_molecule("eleright", 82)
    tail = {0: "", 1: " etypes 1 3", 2: " etypes 1 3 ffield", 3: " etypes 1 3 noslab zneutr",
            4: " etypes 1 3 noslab zneutr", 5: " ffield"}[n]
    arg = ("e eleleft conp 1 eleright 1.979 1.0 log_conp" + tail + (" pppm" if pppm else "")).split()
    return lmp, arg


def il(n, twolayer=False, style=None, value=None, g_ewald=0.21, accuracy=1e-7):
    """tests/il_onelayer/input (and il_twolayer with merged molecules,
    tests/il_twolayer/input:41-42) trial n: 0 conp slab, 1 +etypes,
    2 conq etypes pppm, 3 ffield etypes, 4 pppm ffield ehgo, 5/6 noslab zneutr."""
    s = load_reference_case("il")
    if twolayer:
        s.mol[s.mol == 643] = 641
        s.mol[s.mol == 644] = 642
    if n in (5, 6):
        s = s.doubled_cell(sym=(n == 5), molleft=641, molright=642, molmax=646)
    lmp = MockLammps(s, "p p f" if n <= 2 else "p p p")
    lmp.pair_style_coul_long(16.0)
    pppm = n in (2, 4)
    mesh = None
    if pppm:
        mesh = (36, 36, 432) if n <= 2 else (36, 36, 144)
    lmp.kspace("pppm/conp" if pppm else "pppm", accuracy, g_ewald, slab=3.0 if n <= 2 else None, mesh=mesh)
    lmp.group_molecule("eleleft", 641)
    lmp.group_molecule("eleright", 642)
    v = "2.0" if value is None else repr(float(value))
    style = style or {2: "conq"}.get(n, "conp")
    tail = {0: "", 1: " etypes 1 5", 2: " etypes 1 5 pppm", 3: " etypes 1 5 ffield",
            4: " etypes 1 5 pppm ffield ehgo", 5: " etypes 1 5 noslab zneutr", 6: " etypes 1 5 noslab zneutr"}[n]
    arg = (f"e eleleft {style} 1 eleright 1.979 {v} iter" + tail).split()
    return lmp, arg


def synthetic(name, mode="pppm", ff="slab", g_ewald=0.26, cut=12.0, h=1.0, accuracy=1e-6, style="conp", value=2.0):
    """SURVEY.md 8d synthetic capacitor recipe at any size."""
    from conp_b200.mockhost import mesh_for_spacing
    s = make_workload(name)
    slab = 3.0 if ff == "slab" else None
    lmp = MockLammps(s, "p p f" if ff == "slab" else "p p p")
    lmp.pair_style_coul_long(cut)
    mesh = mesh_for_spacing(s.prd, slab or 1.0, h) if mode == "pppm" else None
    lmp.kspace("pppm/conp" if mode == "pppm" else "pppm", accuracy, g_ewald, slab=slab, mesh=mesh)
    lmp.group_molecule("eleleft", 1)
    lmp.group_molecule("eleright", 2)
    tail = " etypes 1 3" + (" pppm" if mode == "pppm" else "") + ("" if ff == "slab" else f" {ff}")
    arg = (f"e eleleft {style} 1 eleright 1.979 {value!r} log_conp" + tail).split()
    return lmp, arg
