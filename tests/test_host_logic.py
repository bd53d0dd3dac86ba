"""CPU tests of the host-side logic: fix argument parsing (reference syntax and
error messages), the data-file reader, ownership partition across ranks, the
matrix-file formats, EHGO tables, and a world_size-2 gloo run of the rank
plumbing bench.py uses (unique-id broadcast, max-over-ranks reduction)."""
import os
import sys

import numpy as np
import pytest

from cases import dilute
from conp_b200 import MockLammps, load_reference_case
from conp_b200.fixargs import FF_FFIELD, FF_NOSLAB, PAIR_EHGO, FixError, parse_fix_args
from conp_b200.fix_conp import ehgo_setup_tables, read_matrix_file, write_amatrix, write_inv_a_matrix
from conp_b200.mockhost import compute_rho_coeff, good_fft_size, pppm_tables


def test_parse_reference_decks():
    a = parse_fix_args("e eleleft conp 1 eleright 1.979 2.0 iter etypes 1 5 pppm ffield ehgo".split(), 5)
    assert (a.style, a.everynum, a.group, a.group2, a.eta, a.potdiff, a.logfile) == \
        ("conp", 1, "eleleft", "eleright", 1.979, 2.0, "iter")
    assert a.smartlist and a.eletypes == (5,) and a.pppmflag and a.ff_flag == FF_FFIELD and a.pairmode == PAIR_EHGO
    a = parse_fix_args("e eleleft conq 1 eleright 1.979 v_q log noslab zneutr himem nonneutral qinit matout".split(), 5)
    assert a.potdiffstr == "q" and a.potdiff is None and a.ff_flag == FF_NOSLAB and a.zneutrflag
    assert not a.lowmemflag and not a.nullneutralflag and a.qinitflag and a.matoutflag and a.variant == 1


@pytest.mark.parametrize("tokens,msg", [
    ("e g conp 1 g2 1.9 1.0", "too few input parameters"),
    ("e g conp 1 g2 1.9 1.0 log ffield noslab", "ffield and noslab cannot both be chosen"),
    ("e g conp 1 g2 1.9 1.0 log org", "No A matrix filename given"),
    ("e g conp 1 g2 1.9 1.0 log org a inv b", "A matrix file specified more than once"),
    ("e g conp 1 g2 1.9 1.0 log etypes 1", "Insufficient input entries for etypes"),
    ("e g conp 1 g2 1.9 1.0 log etypes 1 9 x", "Invalid atom type in etypes"),
    ("e g conp 1 g2 1.9 1.0 log bogus", "unknown option: bogus"),
])
def test_parse_errors_match_reference_messages(tokens, msg):
    with pytest.raises(FixError) as e:
        parse_fix_args(tokens.split(), 5)
    assert msg in str(e.value)


def test_data_reader_fixture_shapes():
    s = load_reference_case("dilute")
    assert s.natoms == 432 and s.ntypes == 4
    assert (s.mol == 81).sum() == 96 and (s.mol == 82).sum() == 96
    assert np.allclose(s.prd, [9.838, 8.52, 88.4])
    il = load_reference_case("il")
    assert il.natoms == 3776 and (il.type == 5).sum() == 2496 and (il.mol == 641).sum() == 416
    d = s.doubled_cell(sym=True, molleft=81, molright=82, molmax=82)
    assert d.natoms == 864 and np.allclose(d.prd[2], 176.8) and (d.mol == 81).sum() == 192
    assert np.allclose(np.sort(d.x[d.mol == 81, 2]), np.sort(-d.x[d.mol == 81, 2]))  # mirror symmetric


def test_rank_ownership_partition_is_disjoint_and_complete():
    lmp, arg = dilute(2)
    g1, g2 = lmp.groups["eleleft"], lmp.groups["eleright"]
    others = np.nonzero(~(g1 | g2))[0]
    for world in (1, 2, 3, 8):
        chunks = []
        for rank in range(world):
            lo = (len(others) * rank) // world
            hi = (len(others) * (rank + 1)) // world
            chunks.append(others[lo:hi])
        allc = np.concatenate(chunks)
        assert np.array_equal(np.sort(allc), others) and len(set(allc.tolist())) == len(allc)


def test_row_blocks_of_the_library():
    """conp_row_block (the rule conp_set_electrodes applies): contiguous blocks that tile [0, n), equal and a
    multiple of 16 rows except at the end, also when n < 16 * world."""
    from conp_b200 import abi
    import __graft_entry__
    import os
    if not os.path.exists(abi.LIB_PATH):
        __graft_entry__.build()
    for n, world in ((192, 2), (10000, 8), (40000, 8), (833, 4), (5, 2), (40000, 1)):
        blocks = [abi.row_block(n, world, r) for r in range(world)]
        rpr = blocks[0][2]
        assert rpr % 16 == 0 and rpr * world >= n and (rpr - 16) * world < n
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert all(b - a == rpr or b == n for a, b, _ in blocks)


def test_matrix_file_formats(tmp_path):
    rng = np.random.default_rng(0)
    n = 7
    tags = np.arange(11, 11 + n)
    m = rng.standard_normal((n, n))
    pa, pi = str(tmp_path / "amatrix"), str(tmp_path / "inv_a_matrix")
    write_amatrix(pa, tags, m)
    write_inv_a_matrix(pi, tags, m)
    first = open(pa).readline()
    assert first == " " + "".join("%20d" % t for t in tags) + "\n"  # fix_conp.cpp:836-838
    t2, m2 = read_matrix_file(pa, n)
    assert np.array_equal(t2, tags) and np.abs(m2 - m).max() < 1e-12
    t3, m3 = read_matrix_file(pi, n)
    assert np.abs(m3 - m).max() < 1e-10
    with pytest.raises(FixError) as e:
        read_matrix_file(pa, n + 1)
    assert "Too few entries" in str(e.value)
    with pytest.raises(FixError) as e:
        read_matrix_file(pa, n - 1)
    assert "Too many entries" in str(e.value)


def test_ehgo_tables_agree_with_oracle():
    import conp_oracle as O
    nt = 4
    eta_i = np.array([0.0, 1.2, 0.0, 1.979, 0.8])
    u0_i = np.array([0.0, 0.9, 0.0, 2.1, 0.5]) * 0.0694
    t = ehgo_setup_tables(nt, 0.7, eta_i, u0_i)
    e2, f2 = np.zeros((nt + 1) ** 2), np.zeros((nt + 1) ** 2)
    assert O.lib().orc_ehgo_setup_tables(nt, 0.7, O.dp(eta_i), O.dp(u0_i), O.dp(e2), O.dp(f2)) == 1
    assert np.abs(t[0].reshape(-1) - e2).max() < 1e-15 and np.abs(t[1].reshape(-1) - f2).max() < 1e-15
    assert ehgo_setup_tables(nt, 1.0, np.zeros(nt + 1), np.zeros(nt + 1)) is None


def test_pppm_host_tables():
    for order in (3, 4, 5, 6, 7):
        rc = compute_rho_coeff(order)
        for d in (-0.5, -0.2, 0.0, 0.31, 0.5):
            w = sum(rc[l] * d ** l for l in range(order))
            assert abs(w.sum() - 1.0) < 1e-13 and (w > -1e-13).all()  # B-spline weights
    assert [good_fft_size(n) for n in (7, 11, 97, 1000, 1025)] == [8, 12, 100, 1000, 1080]
    t = pppm_tables((8, 9, 10), 5, [9.0, 8.0, 20.0], 1.0, 0.8)
    g = t.greensfn.reshape(10, 9, 8)
    assert g[0, 0, 0] == 0.0 and (g >= 0).all()
    assert np.allclose(g[1:, 1:, 1:], g[1:, 1:, 1:][::-1, ::-1, ::-1])  # inversion symmetric


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # unique-id broadcast as bench.py does it (the id itself needs NCCL; any 128 bytes do here)
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf.copy_(torch.arange(128, dtype=torch.uint8))
    dist.broadcast(buf, 0)
    # max-over-ranks timing reduction
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    q.put((rank, bytes(buf.numpy().tobytes()), float(t.item())))
    dist.destroy_process_group()


def _gather_worker(rank, world, port, q):
    """bench.py's parity leg at N > 1: the row blocks of S (conp_row_block's partition) are assembled on rank 0
    and every rank slices the same global jittered position sets."""
    import sys
    import types
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    import bench
    from conp_b200 import abi
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 200
    S = np.arange(n * n, dtype=np.float64).reshape(n, n)
    a, b, _ = abi.row_block(n, world, rank)
    ctx = types.SimpleNamespace(get_matrix=lambda: S[a:b].copy())
    info = types.SimpleNamespace(row_begin=a, row_end=b)
    full = bench.gather_matrix_to_rank0(ctx, info, n, rank, world)
    x = np.linspace(0.0, 1.0, 30).reshape(10, 3)
    sets = bench.jitter_sets(x, 3, seed=1234)
    ok = (full is None) if rank else bool(np.array_equal(full, S))
    q.put((rank, ok, (a, b), [s_.tobytes() for s_ in sets]))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_matrix_gather_and_position_sets():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)                                   # rank 0 holds the whole S, the others nothing
    assert res[0][2][0] == 0 and res[0][2][1] == res[1][2][0] and res[1][2][1] == 200  # contiguous row blocks
    assert res[0][3] == res[1][3]                                   # every rank solves the same systems


def test_gloo_world2_rank_plumbing():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] == bytes(range(128)) for r in res)
    assert all(r[2] == 2.0 for r in res)


def test_fix_zmirror_mirrors_group2_by_tag_offset():
    """fix zmirror (fix_zmirror.cpp:124-220): group2 <- mirror image of group in z = (zlo+zhi)/2, by tag offset."""
    import numpy as np
    from conp_b200.fix_conp import make_fix, FixError
    from conp_b200 import MockLammps, load_reference_case
    s = load_reference_case("dilute").doubled_cell(sym=False, molleft=81, molright=82, molmax=82)
    lmp = MockLammps(s, "p p p")
    n = s.natoms // 2
    lmp.groups["lower"] = np.arange(s.natoms) < n
    lmp.groups["upper"] = np.arange(s.natoms) >= n
    rng = np.random.default_rng(2)
    s.x[:n] += rng.normal(0, 0.1, (n, 3))
    fix = make_fix(lmp, "zm lower zmirror 1 upper".split())
    fix.setup()
    fix.post_integrate(0)
    zoff = 2 * s.boxlo[2] + s.prd[2]
    assert np.array_equal(s.x[n:, :2], s.x[:n, :2])
    assert np.allclose(s.x[n:, 2], zoff - s.x[:n, 2], rtol=0, atol=0)
    lmp.groups["short"] = np.arange(s.natoms) >= n + 1
    bad = make_fix(lmp, "zm lower zmirror 1 short".split())
    import pytest
    with pytest.raises(FixError, match="same number of tags"):
        bad.setup()
