"""Stand-alone roofline probe of the TMA GEMV: random S of a given size is
loaded as the (already inverted) matrix and streamed `reps` times."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lammps-user-conp2_b200"))
from conp_b200 import abi  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 10000, 20000]
    for n in sizes:
        ctx = abi.Context()
        ctx.set_cell([0, 0, -50], [60, 100, 100], [1, 1, 0], 1, 3.0, 0)
        ctx.set_ewald(0.26, 1e-2, 1000.0, 1000)
        ctx.set_pair(0, 1.979, 12.0, 1, np.full((2, 2), 144.0))
        rng = np.random.default_rng(0)
        xyz = rng.uniform(0, 50, (n, 3))
        side = np.where(np.arange(n) % 2 == 0, 1, -1)
        ctx.set_electrodes(np.arange(1, n + 1), np.ones(n), side, xyz)
        S = rng.standard_normal((n, n))
        t = time.time()
        ctx.load_matrix(S, True)
        t_load = time.time() - t
        tot = ctx.set_unit_voltage(0.0694)
        d = -0.5 * 0.0694 * side
        ref = (S @ d)[side == 1].sum()
        ms = ctx.bench_gemv(50)
        gb = (8.0 * n * n + 16.0 * n) / 1e9
        print(f"N={n}: gemv {ms*1e3:.1f} us  {gb/ (ms*1e-3):.0f} GB/s  (load {t_load:.1f}s)  "
              f"totsetq err {abs(tot-ref)/abs(ref):.2e}", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
