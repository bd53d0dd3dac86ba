"""Row-sharded multi-GPU path: one process per GPU, NCCL allgather of
the packed charges, b and S.b; every rank must reproduce the oracle's charges.
Runs only where >= 2 CUDA devices are visible (gpurun --gpus 2)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, uid_file, case_name, out_file):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "lammps-user-conp2_b200"), os.path.join(root, "tests")):
        sys.path.insert(0, p)
    import time
    from cases import dilute, synthetic
    from conp_b200 import abi
    from conp_b200.fix_conp import make_fix
    if rank == 0:
        uid = abi.get_unique_id()
        with open(uid_file + ".tmp", "wb") as fh:
            fh.write(uid)
        os.rename(uid_file + ".tmp", uid_file)
    else:
        while not os.path.exists(uid_file):
            time.sleep(0.05)
        uid = open(uid_file, "rb").read()
    if case_name == "dilute_ewald":
        lmp, arg = dilute(0)
    elif case_name == "dilute_cond":
        lmp, arg = dilute(2)
        arg[2], arg[6] = "cond", "0.02"
    elif case_name == "small_pppm_sweep":   # the z-sweep tensor-core spread on each rank's slab of planes
        os.environ["CONP_SPREAD"] = "sweep"
        lmp, arg = synthetic("small", h=0.5, accuracy=1e-4)
    else:
        lmp, arg = synthetic("small", h=1.25, accuracy=1e-4)
    fix = make_fix(lmp, arg, device=rank, rank=rank, nranks=world, unique_id=uid)
    fix.setup()
    q = fix.pre_force()
    rng = np.random.default_rng(3)
    others = np.nonzero(fix.side_all == 0)[0]
    lmp.system.x[others] += rng.normal(0, 0.05, (len(others), 3))  # same move on every rank
    q2 = fix.pre_force()
    f, ecoul, eself, vir = fix.post_force()
    b, bk = fix.ctx.get_b()
    info = fix.ctx.info()
    extra = {}
    if case_name.startswith("small_pppm"):  # bricks are sharded (electrolyte by z-slab, electrode by rows): gathered on demand
        extra = dict(rho=fix.ctx.get_density(0), rho_e=fix.ctx.get_density(1), u=fix.ctx.get_potential_brick())
    np.savez(out_file % rank, q=q, q2=q2, scalar=fix.scalar_output, b=b, rows=[info.row_begin, info.row_end],
             eself=eself, ecoul=ecoul, **extra)
    fix.close()


@pytest.mark.parametrize("case_name", ["dilute_ewald", "dilute_cond", "small_pppm", "small_pppm_sweep"])
def test_two_ranks_match_oracle(tmp_path, case_name):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import conp_oracle as O
    from cases import dilute, synthetic
    world = 2
    uid_file = str(tmp_path / "uid.bin")
    out = str(tmp_path / "rank%d.npz")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, uid_file, case_name, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    if case_name == "dilute_ewald":
        lmp, arg = dilute(0)
    elif case_name == "dilute_cond":
        lmp, arg = dilute(2)
        arg[2], arg[6] = "cond", "0.02"
    elif case_name == "small_pppm_sweep":
        lmp, arg = synthetic("small", h=0.5, accuracy=1e-4)
    else:
        lmp, arg = synthetic("small", h=1.25, accuracy=1e-4)
    ref = O.OracleFixConp(lmp, arg)
    ref.setup()
    qr = ref.pre_force()
    rng = np.random.default_rng(3)
    lmp.system.x[ref.oth_idx] += rng.normal(0, 0.05, (len(ref.oth_idx), 3))
    qr2 = ref.pre_force()
    res = [np.load(out % r) for r in range(world)]
    rows = sorted(tuple(r["rows"]) for r in res)
    assert rows[0][0] == 0 and rows[0][1] == rows[1][0] and rows[1][1] == ref.N  # contiguous row blocks
    for r in res:
        for q, qq in ((r["q"], qr), (r["q2"], qr2)):
            assert np.abs(q - qq).max() <= 1e-9 * np.abs(qq).max() + 1e-12
            assert abs(q.sum()) < 1e-12 or case_name == "dilute_cond" and abs(q.sum()) < 1e-12
        assert np.abs(r["b"] - ref.bbb_all).max() <= 5e-12 * max(np.abs(ref.bbb_all).max(), 1.0)
        assert abs(float(r["scalar"]) - ref.scalar_output) <= 1e-9 * abs(ref.scalar_output) + 1e-12
    assert np.array_equal(res[0]["q2"], res[1]["q2"])  # replicated epilogue is bitwise identical
    if case_name.startswith("small_pppm"):
        for r in res:
            assert np.abs(r["rho"] - ref.elyte_density).max() <= 1e-12 * np.abs(ref.elyte_density).max()
            assert np.abs(r["rho_e"] - ref.ele_density).max() <= 1e-9 * np.abs(ref.ele_density).max() + 1e-15
            assert np.abs(r["u"] - ref.u_brick).max() <= 1e-11 * np.abs(ref.u_brick).max()
