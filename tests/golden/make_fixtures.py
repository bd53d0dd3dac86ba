"""Regenerates the committed fixtures from the reference's own test inputs.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):   python tests/golden/make_fixtures.py

* dilute_atoms.npz / il_atoms.npz : atoms of tests/dilute/data and
  tests/il_onelayer/data (byte-identical to il_twolayer/cond/zmirror data),
  columns id mol type q x y z + box, nothing else.
* reference_pins.json : the only numbers the reference's tests store for this
  path -- tests/dilute/persist.log:112-114,143 (g_ewald, mesh, step-0 charges).
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "lammps-user-conp2_b200"))
from conp_b200.system import read_lammps_data, save_fixture  # noqa: E402

REF = "/root/reference/tests"


def main():
    save_fixture(os.path.join(HERE, "dilute_atoms.npz"), read_lammps_data(f"{REF}/dilute/data"))
    save_fixture(os.path.join(HERE, "il_atoms.npz"), read_lammps_data(f"{REF}/il_onelayer/data"))
    log = open(f"{REF}/dilute/persist.log").read().splitlines()
    pins = {"source": "tests/dilute/persist.log"}
    for i, ln in enumerate(log, 1):
        m = re.match(r"\s*G vector \(1/distance\) = (\S+)", ln)
        if m:
            pins["g_ewald"] = float(m.group(1)); pins["g_ewald_line"] = i
        m = re.match(r"\s*grid = (\d+) (\d+) (\d+)", ln)
        if m:
            pins["mesh"] = [int(m.group(k)) for k in (1, 2, 3)]; pins["mesh_line"] = i
        m = re.match(r"\s*stencil order = (\d+)", ln)
        if m:
            pins["order"] = int(m.group(1))
        m = re.match(r"\s*0\s+0\s+0\s+(\S+)\s+(\S+)\s+(\S+)\s*$", ln)
        if m:
            pins["step0"] = {"c_qleft": float(m.group(1)), "c_qright": float(m.group(2)),
                             "c_qall": float(m.group(3)), "line": i}
        m = re.match(r"fix e all conp/v4 (.*)$", ln)
        if m:
            pins["fix_line_v4_syntax"] = ln.strip()
        m = re.match(r"kspace_style\s+pppm\s+(\S+)", ln)
        if m:
            pins["kspace_accuracy"] = float(m.group(1))
        m = re.match(r"pair_style\s+lj/cut/coul/long\s+(\S+)", ln)
        if m:
            pins["pair_cut"] = float(m.group(1))
    json.dump(pins, open(os.path.join(HERE, "reference_pins.json"), "w"), indent=1)
    print(pins)


if __name__ == "__main__":
    main()
