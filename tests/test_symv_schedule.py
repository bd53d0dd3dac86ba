"""Host-side check of the symmetric matvec's work decomposition (no GPU needed).

`conp_plan_symv` (host code of the library, the same function the GPU path calls) gives the strips;
the loops below restate, in numpy, what symv_tma_kernel / symv_reduce_kernel do with them
(gemv.cu: sy_chunk, sy_in_band, the masked and the fast path, the column-slice layout).  The test
asserts that every element of a symmetric matrix is used exactly once per unordered pair (plus the
diagonal), that the unmasked fast path is only taken where no mask is needed, that no strip's column
slice is overrun or read before it is written, and that row blocks of several ranks add up to S.b."""
import numpy as np
import pytest

from conp_b200 import abi

R, C = 8, 512  # rows per stage, columns per chunk (gemv.cu)


@pytest.fixture(scope="module", autouse=True)
def _library():
    import os
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def chunks(a, bnd, N, H):
    CS, CE = a & ~1, bnd + H
    end1 = min(CE, N)
    len1 = end1 - CS
    n1 = -(-len1 // C)
    len2 = max(CE - N, 0)
    n2 = -(-len2 // C)
    for k in range(n1 + n2):
        if k < n1:
            cb = CS + k * C
            cev, seg2, cact, joff = min(cb + C, end1), False, cb, k * C
        else:
            kk = k - n1
            cb = N + kk * C
            cev, seg2, cact, joff = min(cb + C, CE), True, kk * C, ((len1 + 1) & ~1) + kk * C
        yield cb, cev, (cev - cb + 1) & ~1, cact, joff, seg2


def symv_model(S_rows, b, row0, nrows, N, ncols_pad, num_sms):
    plan = abi.plan_symv(N, row0, nrows, num_sms)
    assert plan is not None
    strips, L = plan
    H, tie = N // 2, N % 2 == 0
    rowpart = np.zeros(N)
    colpart = np.full((len(strips), L), np.nan)      # NaN: a slot read before it was written shows up
    Sp = np.zeros((S_rows.shape[0], ncols_pad))
    Sp[:, :N] = S_rows
    bp = np.zeros(ncols_pad)
    bp[:N] = b
    count = np.zeros((N, N), dtype=np.int32)
    for s, (a, bnd) in enumerate(strips):
        if a >= bnd:
            continue
        assert bnd - a <= 512 and (bnd - a) + H + 2 <= N
        for cb, cev, w, cact, joff, seg2 in chunks(a, bnd, N, H):
            assert cact + w <= ncols_pad and joff + w <= L
            col = np.zeros(w)
            cu = cb + np.arange(w)
            for rg in range(a, bnd, R):
                nr = min(R, bnd - rg)
                if not (rg <= cev - 1 and rg + nr - 1 + H >= cb):
                    continue
                fast = cb - (rg + nr - 1) >= 1 and (cb + w - 1) - rg <= H - 1 and (seg2 or cb + w <= N)
                for r in range(rg, rg + nr):
                    d = cu - r
                    ok = (d >= 0) & (d <= H) & ~((d == H) & tie & (r >= H)) & (seg2 | (cu < N))
                    if fast:
                        assert ok.all() and (d >= 1).all()
                    x = np.where(ok, Sp[r - row0, cact:cact + w], 0.0)
                    rowpart[r] += x @ bp[cact:cact + w]
                    col += np.where(d >= 1, x, 0.0) * b[r]
                    cols = (cact + np.arange(w))[ok]
                    count[r, cols] += 1
                    count[cols[d[ok] >= 1], r] += 1
            colpart[s, joff:joff + w] = col
    out = np.zeros(N)
    for c in range(N):
        v = rowpart[c] if row0 <= c < row0 + nrows else 0.0
        for s, (a, bnd) in enumerate(strips):
            if a >= bnd:
                continue
            CS, CE = a & ~1, bnd + H
            end1 = min(CE, N)
            if CS <= c < end1:
                v += colpart[s, c - CS]
            elif c < CE - N:
                v += colpart[s, ((end1 - CS + 1) & ~1) + c]
        out[c] = v
    return out, count


@pytest.mark.parametrize("N,num_sms,nranks", [(64, 148, 1), (65, 148, 2), (191, 148, 3), (192, 148, 1), (600, 148, 2),
                                              (300, 4, 1), (1200, 7, 2), (1500, 3, 1), (1031, 148, 8)])
def test_every_pair_once_and_rank_blocks_add_up(N, num_sms, nranks):
    rng = np.random.default_rng(N)
    A = rng.standard_normal((N, N))
    S = A + A.T
    b = rng.standard_normal(N)
    pad = -(-N // 16) * 16
    rpr = -(-(-(-N // nranks)) // 16) * 16            # the library's row blocks (conp_set_electrodes)
    tot = np.zeros(N)
    cnt = np.zeros((N, N), dtype=np.int32)
    for rk in range(nranks):
        r0 = min(N, rk * rpr)
        r1 = min(N, r0 + rpr)
        if r1 <= r0:                                  # rank without rows: zero partial, empty plan
            strips, _ = abi.plan_symv(N, r0, 0, num_sms)
            assert len(strips) == 0
            continue
        o, c = symv_model(S[r0:r1], b, r0, r1 - r0, N, pad, num_sms)
        tot += o
        cnt += c
    assert (cnt == 1).all()
    assert np.abs(tot - S @ b).max() <= 1e-12 * np.abs(S).sum(axis=1).max() * np.abs(b).max()


def test_small_matrices_keep_the_general_kernel():
    assert abi.plan_symv(63, 0, 63) is None
    assert abi.plan_symv(10, 0, 10) is None


def test_plan_at_bench_sizes():
    for N, nranks in ((10000, 1), (40000, 1), (40000, 8)):
        rpr = -(-(-(-N // nranks)) // 16) * 16
        strips, L = abi.plan_symv(N, 0, min(rpr, N), 148)
        h = (strips[:, 1] - strips[:, 0])
        assert len(strips) == 148 and h.max() - h.min() <= 1 and h.max() <= 512
        assert strips[0, 0] == 0 and strips[-1, 1] == min(rpr, N) and (strips[1:, 0] == strips[:-1, 1]).all()
        assert L >= h.max() + N // 2 + 2
