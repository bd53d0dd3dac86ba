"""Atom systems for the electrode charge solve: LAMMPS data-file reader, the
reference test geometries and the synthetic graphite capacitors of
BASELINE.json configs 4/5.

This is *host-side input plumbing* (what LAMMPS' ``read_data``/``group``/
``replicate`` commands do in the reference decks, e.g.
tests/il_onelayer/input:29-61); nothing here is on the hot path.
"""
from __future__ import annotations

import dataclasses
import os
import re

import numpy as np


@dataclasses.dataclass
class System:
    """Atoms of one LAMMPS ``atom_style full`` configuration (tags are id)."""

    boxlo: np.ndarray  # (3,)
    boxhi: np.ndarray  # (3,)
    id: np.ndarray  # (n,) int32, 1-based tags
    mol: np.ndarray  # (n,) int32
    type: np.ndarray  # (n,) int32, 1-based
    q: np.ndarray  # (n,) float64
    x: np.ndarray  # (n,3) float64
    ntypes: int

    @property
    def natoms(self) -> int:
        return int(self.id.shape[0])

    @property
    def prd(self) -> np.ndarray:
        return self.boxhi - self.boxlo

    def copy(self) -> "System":
        return System(self.boxlo.copy(), self.boxhi.copy(), self.id.copy(), self.mol.copy(),
                      self.type.copy(), self.q.copy(), self.x.copy(), self.ntypes)

    # -- the geometry edits the reference decks perform -------------------
    def doubled_cell(self, sym: bool, molleft: int, molright: int, molmax: int) -> "System":
        """``replicate 1 1 2`` + ``change_box z final -lz/2 lz/2 remap`` and the
        group/molecule edits of tests/il_onelayer/input:37-51 (n == 5: mirror
        image, "sym"; n == 6: plain translate with swapped electrode roles,
        "anti")."""
        n = self.natoms
        lz = self.prd[2]
        lo = self.boxlo.copy()
        hi = self.boxhi.copy()
        x2 = self.x.copy()
        x2[:, 2] += lz
        x = np.concatenate([self.x, x2])
        mol = np.concatenate([self.mol, np.where(self.mol > 0, self.mol + molmax, 0)])
        # replicated box is [zlo, zlo+2lz]; remap to [-lz, lz]
        newlo = -lz
        x[:, 2] = (x[:, 2] - lo[2]) + newlo
        lo[2], hi[2] = -lz, lz
        pos = x[:, 2] >= 0.0  # region pos block ... 0 EDGE
        if sym:
            x[pos, 2] = lz - x[pos, 2]  # variable newz atom lz/2-z  (lz = new box length/2 ... )
            mol = np.where(mol == molmax + molright, molright, mol)
            mol = np.where(mol == molmax + molleft, molleft, mol)
        else:
            mol = np.where(mol == molmax + molright, molleft, mol)
            mol = np.where(mol == molmax + molleft, molright, mol)
        return System(lo, hi, np.arange(1, 2 * n + 1, dtype=np.int32), mol.astype(np.int32),
                      np.concatenate([self.type, self.type]), np.concatenate([self.q, self.q]), x,
                      self.ntypes)


def read_lammps_data(path: str) -> System:
    """Minimal ``read_data`` for ``atom_style full`` files: header counts, box
    bounds and the ``Atoms`` section (rows ``id mol type q x y z [ix iy iz]``,
    optional trailing ``# comment``), cf. tests/il_onelayer/data:39-41 and
    tests/dilute/data:33-35.  Image flags are ignored (coordinates stay as
    written, inside the box)."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    natoms = ntypes = None
    lo = np.zeros(3)
    hi = np.zeros(3)
    i = 0
    atoms_at = None
    for i, ln in enumerate(lines):
        s = ln.split("#")[0].strip()
        if not s:
            continue
        m = re.match(r"^(\d+)\s+atoms$", s)
        if m:
            natoms = int(m.group(1))
        m = re.match(r"^(\d+)\s+atom types$", s)
        if m:
            ntypes = int(m.group(1))
        for ax, tag in enumerate(("xlo xhi", "ylo yhi", "zlo zhi")):
            if s.endswith(tag):
                a, b = s.split()[:2]
                lo[ax], hi[ax] = float(a), float(b)
        if s.split()[0] == "Atoms":
            atoms_at = i
            break
    if natoms is None or ntypes is None or atoms_at is None:
        raise ValueError(f"{path}: not a LAMMPS data file with an Atoms section")
    rows = []
    for ln in lines[atoms_at + 1:]:
        s = ln.split("#")[0].split()
        if not s:
            if rows:
                break
            continue
        if not s[0].lstrip("-").isdigit():
            break
        rows.append(s[:7])
        if len(rows) == natoms:
            break
    if len(rows) != natoms:
        raise ValueError(f"{path}: expected {natoms} atoms, found {len(rows)}")
    arr = np.array(rows, dtype=np.float64)
    order = np.argsort(arr[:, 0], kind="stable")
    arr = arr[order]
    return System(lo, hi, arr[:, 0].astype(np.int32), arr[:, 1].astype(np.int32),
                  arr[:, 2].astype(np.int32), arr[:, 3].copy(), arr[:, 4:7].copy(), ntypes)


def save_fixture(path: str, s: System, **extra) -> None:
    np.savez_compressed(path, boxlo=s.boxlo, boxhi=s.boxhi, id=s.id, mol=s.mol, type=s.type, q=s.q,
                        x=s.x, ntypes=np.int32(s.ntypes), **extra)


def load_fixture(path: str) -> System:
    z = np.load(path)
    return System(z["boxlo"].copy(), z["boxhi"].copy(), z["id"].copy(), z["mol"].copy(),
                  z["type"].copy(), z["q"].copy(), z["x"].copy(), int(z["ntypes"]))


GOLDEN_DIR = os.path.normpath(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden"))


def load_reference_case(name: str) -> System:
    """Committed copies of the reference's own test inputs (atoms only), made by
    tests/golden/make_fixtures.py from /root/reference/tests/*/data."""
    return load_fixture(os.path.join(GOLDEN_DIR, f"{name}_atoms.npz"))


# ---------------------------------------------------------------------------
# synthetic capacitors
# ---------------------------------------------------------------------------

GRAPHENE_A = 2.46
GRAPHENE_B = 4.26
LAYER_DZ = 3.35


def graphene_layer(ncx: int, ncy: int, z: float) -> np.ndarray:
    """Rectangular 4-atom graphene cell (a = 2.46 A, b = 4.26 A), ncx x ncy cells."""
    a, b = GRAPHENE_A, GRAPHENE_B
    basis = np.array([[0.0, 0.0], [a / 2, b / 6], [a / 2, b / 2], [0.0, 2 * b / 3]])
    ix, iy = np.meshgrid(np.arange(ncx), np.arange(ncy), indexing="ij")
    cell = np.stack([ix.ravel() * a, iy.ravel() * b], axis=1)
    xy = (cell[:, None, :] + basis[None, :, :]).reshape(-1, 2)
    # small offset keeps atoms off the periodic seam
    xy += np.array([0.25 * a, 0.05 * b])
    return np.concatenate([xy, np.full((xy.shape[0], 1), z)], axis=1)


def make_capacitor(ncx: int, ncy: int, nlayers: int, n_elyte: int, seed: int,
                   density: float = 0.05, exclusion: float = 3.0, qmag: float = 0.8,
                   zmargin: float = 5.0) -> System:
    """Two graphene electrodes (``nlayers`` each) facing each other across z with
    ``n_elyte`` point charges (+-qmag alternating, net zero) uniformly placed in
    the gap.  Types: 1 cation, 2 anion, 3 electrode carbon.  Molecule ids: 0 for
    the electrolyte, 1 = left (low z) electrode, 2 = right electrode."""
    rng = np.random.default_rng(seed)
    lx, ly = ncx * GRAPHENE_A, ncy * GRAPHENE_B
    gap = n_elyte / (density * lx * ly)
    zin = 0.5 * gap + exclusion  # inner electrode planes at +-zin
    layers_l, layers_r = [], []
    for k in range(nlayers):
        layers_l.append(graphene_layer(ncx, ncy, -(zin + k * LAYER_DZ)))
        layers_r.append(graphene_layer(ncx, ncy, +(zin + k * LAYER_DZ)))
    xl = np.concatenate(layers_l)
    xr = np.concatenate(layers_r)
    xe = np.empty((n_elyte, 3))
    xe[:, 0] = rng.uniform(0.0, lx, n_elyte)
    xe[:, 1] = rng.uniform(0.0, ly, n_elyte)
    xe[:, 2] = rng.uniform(-0.5 * gap, 0.5 * gap, n_elyte)
    qe = np.where(np.arange(n_elyte) % 2 == 0, qmag, -qmag)
    te = np.where(np.arange(n_elyte) % 2 == 0, 1, 2)
    zmax = zin + (nlayers - 1) * LAYER_DZ + zmargin
    x = np.concatenate([xe, xl, xr])
    n = x.shape[0]
    q = np.concatenate([qe, np.zeros(xl.shape[0] + xr.shape[0])])
    typ = np.concatenate([te, np.full(xl.shape[0] + xr.shape[0], 3)]).astype(np.int32)
    mol = np.concatenate([np.zeros(n_elyte), np.full(xl.shape[0], 1), np.full(xr.shape[0], 2)]).astype(np.int32)
    return System(np.array([0.0, 0.0, -zmax]), np.array([lx, ly, zmax]),
                  np.arange(1, n + 1, dtype=np.int32), mol, typ, q, x, 3)


# named workloads -------------------------------------------------------------
WORKLOADS = {
    # BASELINE.json configs[3]: 10k electrode / 100k electrolyte
    "cfg4": dict(ncx=25, ncy=25, nlayers=2, n_elyte=100_000, seed=20261018),
    # BASELINE.json configs[4]: 40k electrode / 500k electrolyte
    "cfg5": dict(ncx=50, ncy=50, nlayers=2, n_elyte=500_000, seed=20261019),
    # small stand-ins used by the tests (same recipe)
    "tiny": dict(ncx=4, ncy=3, nlayers=2, n_elyte=400, seed=7),
    "small": dict(ncx=8, ncy=5, nlayers=2, n_elyte=3000, seed=11),
    "medium": dict(ncx=13, ncy=8, nlayers=2, n_elyte=12000, seed=13),
}


def make_workload(name: str) -> System:
    return make_capacitor(**WORKLOADS[name])
