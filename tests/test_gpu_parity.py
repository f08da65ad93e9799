"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through
the C ABI (hgaprec_b200.Engine -> libhpf_b200.so) and is compared with the fp64
oracle / the reference's own golden states.  Tolerances: util.TOL_* (SURVEY.md 8c)."""
import numpy as np
import pytest

import util
import hgaprec_b200 as H
from hgaprec_b200 import synth
from oracle import hpf_oracle as O

pytestmark = pytest.mark.gpu


def make_engine(state, n_users=None, **kw):
    return H.Engine(state.n if n_users is None else n_users, state.m, state.k, flags=util.engine_flags(state), **kw)


def run_engine(state, rp, ci, y, iters):
    with make_engine(state) as e:
        e.set_ratings_csr(rp, ci, y)
        util.push_state(e, state)
        e.iterate(iters)
        return util.pull_state(e, state), e.stats()


# ------------------------------------------------------------------ goldens
@pytest.mark.parametrize("mode", util.MODES)
def test_one_iteration_matches_reference_golden(mode):
    g = util.load_golden(mode)
    s0 = util.golden_state(g, 0)
    got, st = run_engine(s0, g["csr.row_ptr"], g["csr.col_idx"], g["csr.y"], 1)
    bad = util.compare_states(got, util.golden_state(g, 1))
    assert not bad, bad
    assert st["slow_path_nnz"] == 0 and st["kernel_launches"] > 0


@pytest.mark.parametrize("mode", util.MODES)
def test_three_iterations_and_heldout_match_reference_golden(mode):
    g = util.load_golden(mode)
    s0 = util.golden_state(g, 0)
    with make_engine(s0) as e:
        e.set_ratings_csr(g["csr.row_ptr"], g["csr.col_idx"], g["csr.y"])
        util.push_state(e, s0)
        # the reference's initial expectations are not a function of its rates:
        # held-out ll of the uploaded state must reproduce the reference's T=0 number
        for split in ("validation", "test"):
            ll0 = e.heldout_loglik(g[split + ".u"], g[split + ".i"], g[split + ".y"])
            ref0 = float(g["T0/%s.ll_sum" % split][0])
            assert abs(ll0 - ref0) <= 2e-5 * max(1.0, abs(ref0))
        e.iterate(3)
        got = util.pull_state(e, s0)
        bad = util.compare_states(got, util.golden_state(g, 3), rel=6e-5, elog_abs=6e-5)
        assert not bad, bad
        for split in ("validation", "test"):
            npairs = len(g[split + ".u"])
            ll = e.heldout_loglik(g[split + ".u"], g[split + ".i"], g[split + ".y"]) / npairs
            ref = float(g["T3/%s.ll_sum" % split][0]) / npairs
            assert abs(ll - ref) <= util.TOL_LL_20IT, (split, ll, ref)


# --------------------------------------------------- oracle at config shapes
def _oracle_case(n, m, nnz, k, flags, seed, binary=False):
    d = synth.make_ratings(n, m, nnz, binary=binary, seed=seed, heldout=0.05)
    s = O.OracleState(d["n"], d["m"], k, flags).init(seed + 1)
    return d, s


@pytest.mark.parametrize("name,n,m,nnz,k,flags,binary", [
    ("C1-shape hier K=100", 6040, 3681, 792166, 100, H.HIER, False),
    ("C3-shape hier binary K=200", 5000, 1900, 240000, 200, H.HIER | H.BINARY, True),
    ("C4-shape bpf bias K=100", 8000, 800, 400000, 100, H.BIAS, False),
    ("hier bias K=100", 3000, 1500, 200000, 100, H.HIER | H.BIAS, False),
])
def test_single_iteration_matches_oracle(name, n, m, nnz, k, flags, binary):
    d, s = _oracle_case(n, m, nnz, k, flags, seed=11, binary=binary)
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], 1, nthreads=8)
    got, st = run_engine(s, d["row_ptr"], d["col_idx"], d["y"], 1)
    bad = util.compare_states(got, want)
    assert not bad, (name, bad)
    assert st["slow_path_nnz"] == 0


def test_twenty_iterations_movielens_shape_trajectory():
    """C1-sized problem, K=100, -hier: 20 iterations against the fp64 oracle."""
    d, s = _oracle_case(6040, 3681, 792166, 100, H.HIER, seed=111)
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], 20, nthreads=8)
    hu, hi, hy = d["heldout"]
    with make_engine(s) as e:
        e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
        util.push_state(e, s)
        e.iterate(20)
        got = util.pull_state(e, s)
        ll = e.heldout_loglik(hu, hi, hy) / len(hu)
    ref = want.heldout(hu, hi, hy) / len(hu)
    assert abs(ll - ref) <= util.TOL_LL_20IT, (ll, ref)
    for gname in ("theta", "beta"):
        rf = util.rel_fro(got.p[gname]["Ev"], want.p[gname]["Ev"])
        assert rf <= util.TOL_RELFRO_20IT, (gname, rf)
        rel = np.abs(got.p[gname]["Ev"] - want.p[gname]["Ev"]) / np.maximum(np.abs(want.p[gname]["Ev"]), 1e-3)
        assert np.percentile(rel, 99) <= 5e-4


# ---------------------------------------------------------------- edge cases
@pytest.mark.parametrize("k", [1, 3, 5, 33, 100, 130, 257])
def test_ragged_k(k):
    d, s = _oracle_case(300, 120, 6000, k, H.HIER, seed=5 + k)
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], 2)
    got, _ = run_engine(s, d["row_ptr"], d["col_idx"], d["y"], 2)
    bad = util.compare_states(got, want, rel=4e-5, elog_abs=4e-5)
    assert not bad, bad


def test_empty_rows_long_rows_duplicates_and_big_ratings():
    rng = np.random.default_rng(3)
    n, m, k = 400, 700, 20
    rows = []
    for u in range(n):
        if u in (0, 17, 399):
            rows.append(np.zeros(0, dtype=np.uint32))            # users without ratings
        elif u == 5:
            rows.append(rng.permutation(600).astype(np.uint32))  # > seg_len: split user row
        else:
            others = 1 + rng.choice(599, size=rng.integers(1, 30), replace=False)
            rows.append(np.concatenate([[0], others]).astype(np.uint32))  # item 0: > seg_len users -> split item row
    rows[7] = np.array([3, 9, 3, 3], dtype=np.uint32)            # duplicate (u,i): processed once per entry
    ci = np.concatenate(rows)                                    # items 600..699 have no ratings
    rp = np.zeros(n + 1, dtype=np.uint64)
    rp[1:] = np.cumsum([len(r) for r in rows])
    y = rng.integers(1, 6, size=len(ci)).astype(np.uint8)
    y[::7] = 255                                                 # yval_t is uint8 (env.hh:20)
    s = O.OracleState(n, m, k, H.HIER | H.BIAS).init(9)
    want = s.copy().iterate(rp, ci, y, 2)
    got, _ = run_engine(s, rp, ci, y, 2)
    bad = util.compare_states(got, want, rel=4e-5, elog_abs=4e-5)
    assert not bad, bad


@pytest.mark.parametrize("flags,tile_kb", [(H.HIER, 64), (H.HIER | H.BIAS, 200), (H.BIAS, 37)])
def test_l2_tiled_orderings_match_oracle(monkeypatch, flags, tile_kb):
    """Both sweeps regroup the nonzeros by (tile of the gathered side, row) once
    the gathered factor rows exceed the L2 budget; force many tiles on a small
    problem (also: split rows inside tiles) and compare with the oracle."""
    monkeypatch.setenv("HPF_L2_TILE_KB", str(tile_kb))
    monkeypatch.setenv("HPF_SEG_LEN", "32")
    d, s = _oracle_case(3000, 1500, 200000, 100, flags, seed=31)
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], 2, nthreads=8)
    got, st = run_engine(s, d["row_ptr"], d["col_idx"], d["y"], 2)
    bad = util.compare_states(got, want, rel=4e-5, elog_abs=4e-5)
    assert not bad, bad
    assert st["slow_path_nnz"] == 0


@pytest.mark.parametrize("dense", ["1", "0"])
@pytest.mark.parametrize("name,n,m,nnz,k,flags,binary", [
    ("hier K=100", 3000, 1500, 200000, 100, H.HIER, False),
    ("bpf K=33 fewer items than head slots", 700, 90, 20000, 33, 0, False),
    ("hier binary K=128", 1000, 400, 60000, 128, H.HIER | H.BINARY, True),
    ("bpf jacobi-less K=5, users not a multiple of 128", 517, 300, 15000, 5, 0, False),
    ("hier bias K=100", 3000, 1500, 200000, 100, H.HIER | H.BIAS, False),
    ("bpf bias -novb K=64 (bias columns open a new 32-column chunk)", 900, 500, 40000, 64, H.BIAS | H.JACOBI, False),
    ("bpf bias K=124: Kp + 2 = 126 operand columns, the widest that fits", 600, 300, 20000, 124, H.BIAS, False),
])
def test_dense_head_matches_oracle_and_gather_plan(monkeypatch, dense, name, n, m, nnz, k, flags, binary):
    """The tcgen05 dense head (forced on / off) against the oracle: three iterations, duplicates included."""
    monkeypatch.setenv("HPF_DENSE_HEAD", dense)
    monkeypatch.setenv("HPF_DENSE_BLOCK_SHARE", "0")  # as many 128-item head blocks as there are items (up to 4)
    d, s = _oracle_case(n, m, nnz, k, flags, seed=53, binary=binary)
    rp, ci, y = d["row_ptr"].astype(np.int64), d["col_idx"].copy(), d["y"]
    ci[rp[3] + 1] = ci[rp[3]]  # a repeated (user, item) entry: walked twice by the reference
    want = s.copy().iterate(rp, ci, y, 3, nthreads=8)
    got, st = run_engine(s, rp, ci, y, 3)
    assert (st["head_nnz"] > 0) == (dense == "1") and st["slow_path_nnz"] == 0
    bad = util.compare_states(got, want, rel=6e-5, elog_abs=6e-5)
    assert not bad, (name, bad)


@pytest.mark.parametrize("flags", [H.HIER, H.HIER | H.BIAS])
def test_dense_head_cell_overflow_plans_without_the_head(monkeypatch, flags):
    """The dense ratings block keeps one byte per (user, head item).  Repeated lines for one pair add up (the
    reference walks every line, hgaprec.cc:1340-1366); a sum past 255 does not fit, the set-up has to notice and
    plan without the dense head.  A sum of exactly 255 still fits."""
    monkeypatch.setenv("HPF_DENSE_HEAD", "1")
    monkeypatch.setenv("HPF_DENSE_BLOCK_SHARE", "0")  # m = 300: every item is a head item
    d, s = _oracle_case(900, 300, 40000, 32, flags, seed=71)
    rp, ci, y = d["row_ptr"].astype(np.int64), d["col_idx"].copy(), d["y"].copy()
    u = int(np.argmax(np.diff(rp) >= 3))
    b = int(rp[u])
    ci[b + 1] = ci[b + 2] = ci[b]  # three lines for one (user, item) pair
    for last, head_on in ((6, False), (5, True)):
        y[b], y[b + 1], y[b + 2] = 200, 50, last
        want = s.copy().iterate(rp, ci, y, 2, nthreads=8)
        got, st = run_engine(s, rp, ci, y, 2)
        assert (st["head_nnz"] > 0) == head_on
        bad = util.compare_states(got, want, rel=6e-5, elog_abs=6e-5)
        assert not bad, (last, bad)


def test_no_ratings_at_all_and_y_null():
    n, m, k = 10, 12, 4
    s = O.OracleState(n, m, k, 0).init(2)
    rp = np.zeros(n + 1, dtype=np.uint64)
    want = s.copy().iterate(rp, np.zeros(0, np.uint32), None, 1)
    got, _ = run_engine(s, rp, np.zeros(0, np.uint32), None, 1)
    assert not util.compare_states(got, want)
    d, s = _oracle_case(200, 90, 3000, 10, H.HIER | H.BINARY, seed=4, binary=True)
    assert d["y"] is None
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], None, 1)
    got, _ = run_engine(s, d["row_ptr"], d["col_idx"], None, 1)
    assert not util.compare_states(got, want)


def test_exact_fallback_when_products_underflow():
    """Rows whose log-expectations peak in different factors by > 100 nats make
    the fp32 product form underflow; the engine must take the log-domain path
    and still match the oracle."""
    d, s = _oracle_case(64, 48, 900, 8, H.HIER, seed=21)
    el_t, el_b = s.p["theta"]["Elogv"], s.p["beta"]["Elogv"]
    el_t[:, :] = -120.0
    el_t[:, 0] = 0.0          # theta rows peak at k=0
    el_b[:, :] = -130.0
    el_b[:, 5] = 1.0          # beta rows peak at k=5
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], 1)
    got, st = run_engine(s, d["row_ptr"], d["col_idx"], d["y"], 1)
    assert st["slow_path_nnz"] == 2 * len(d["col_idx"])  # both passes
    bad = util.compare_states(got, want, rel=4e-5, elog_abs=4e-5)
    assert not bad, bad
    # and the next iteration (ordinary state again) runs on the fast path
    want.iterate(d["row_ptr"], d["col_idx"], d["y"], 1)
    with make_engine(s) as e:
        e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
        util.push_state(e, s)
        e.iterate(2)
        got2 = util.pull_state(e, s)
        assert e.stats()["slow_path_nnz"] == 2 * len(d["col_idx"])
    assert not util.compare_states(got2, want, rel=6e-5, elog_abs=6e-5)


@pytest.mark.parametrize("flags,dense", [(H.HIER | H.BIAS, True), (H.HIER, True), (0, True)])
def test_default_plan_is_bitwise_deterministic(flags, dense):
    """The default plan (gather kernel for the tail, dense tcgen05 head where it applies: K (+2 with bias) <= 128)
    sums in a fixed order."""
    d, s = _oracle_case(2000, 700, 90000, 100, flags, seed=8)
    a, st = run_engine(s, d["row_ptr"], d["col_idx"], d["y"], 3)
    b, _ = run_engine(s, d["row_ptr"], d["col_idx"], d["y"], 3)
    assert (st["head_nnz"] > 0) == dense
    for gname in util.groups(a):
        for f in O.FIELDS:
            np.testing.assert_array_equal(a.p[gname][f], b.p[gname][f])


@pytest.mark.parametrize("seg_len,tile_kb,dense,pack", [("64", None, "0", "1"), ("8", None, "0", "1"), ("64", 300, "0", "1"),
                                                        ("64", None, "1", "1"), ("16", 150, "1", "1"), ("64", None, "0", "0"),
                                                        ("16", 150, "1", "0")])
@pytest.mark.parametrize("flags", [H.HIER, H.BIAS])
def test_every_sweep_plan_matches_oracle(monkeypatch, flags, seg_len, tile_kb, dense, pack):
    """The device-built work lists in every shape they take: short segments (rows split over many partial slots and
    combined in a fixed order), L2-tiled orderings of both passes, with and without the dense tcgen05 head (whose
    items have no rows in the item pass and no nonzeros in the user pass), with index and rating packed into one word
    per nonzero (the default whenever the gathered side has < 2^24 rows) and in separate streams."""
    monkeypatch.setenv("HPF_DENSE_HEAD", dense)
    monkeypatch.setenv("HPF_PACK", pack)
    monkeypatch.setenv("HPF_SEG_LEN", seg_len)
    if tile_kb:
        monkeypatch.setenv("HPF_L2_TILE_KB", str(tile_kb))
    d, s = _oracle_case(3000, 1500, 200000, 100, flags, seed=41)
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], 2, nthreads=8)
    got, st = run_engine(s, d["row_ptr"], d["col_idx"], d["y"], 2)
    assert (st["head_nnz"] > 0) == (dense == "1")
    assert (st["user_l2_tiles"] > 1 and st["item_l2_tiles"] > 1) == bool(tile_kb)
    bad = util.compare_states(got, want, rel=4e-5, elog_abs=4e-5)
    assert not bad, bad
    assert st["slow_path_nnz"] == 0
    again, _ = run_engine(s, d["row_ptr"], d["col_idx"], d["y"], 2)   # and every plan sums in a fixed order
    for gname in util.groups(got):
        for f in O.FIELDS:
            np.testing.assert_array_equal(got.p[gname][f], again.p[gname][f])


def test_expectations_are_materialised_on_demand():
    """Per iteration the engine stores A, E[log v] and the shape only; the rate matrix and E[v] are derived from
    the shape and the two rate terms when a consumer asks (hpf_get_state, held-out ll, top-N).  Reading the state
    between iterations must not change the trajectory, and partial reads must agree with full ones."""
    d, s = _oracle_case(1500, 600, 60000, 100, H.HIER, seed=77)
    with make_engine(s) as e:
        e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
        util.push_state(e, s)
        e.iterate(3)
        a = util.pull_state(e, s)
    with make_engine(s) as e:
        e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
        util.push_state(e, s)
        hu, hi, hy = d["heldout"]
        for _ in range(3):
            e.iterate(1)
            e.heldout_loglik(hu, hi, hy)
            only_ev = e.get_state(H.THETA, fields=("Ev",))["Ev"]
        b = util.pull_state(e, s)
    for gname in util.groups(a):
        for f in O.FIELDS:
            np.testing.assert_array_equal(a.p[gname][f], b.p[gname][f], err_msg="%s.%s" % (gname, f))
    np.testing.assert_array_equal(only_ev, b.p["theta"]["Ev"])
    # E[v] = shape / rate and E[log v] = psi(shape) - log(rate) hold between the exported arrays (fp32 rounding)
    for gname in ("theta", "beta"):
        p = b.p[gname]
        np.testing.assert_allclose(p["Ev"], p["shape"] / p["rate"], rtol=3e-7)


def test_errors_are_reported_not_fatal():
    with H.Engine(4, 4, 3) as e:
        with pytest.raises(H.HpfError):
            e.iterate(1)  # nothing set yet
        rp = np.array([0, 1, 2, 2, 3], dtype=np.uint64)
        with pytest.raises(H.HpfError):
            e.set_ratings_csr(rp, np.array([0, 9, 1], dtype=np.uint32))  # item 9 >= m
    with pytest.raises(H.HpfError):
        H.Engine(4, 4, 0)


# ----------------------------------- size-independent property at full size
def test_full_size_netflix_shape_mass_conservation():
    """BASELINE config C2 (480,189 x 17,770, 1e8 nnz, K=100, -hier): phi sums to
    one per nonzero, so for every user sum_k (shape_uk - prior) == sum_i y_ui
    and likewise per item (hgaprec.cc:1355-1359).  Checked on every row."""
    c = synth.CONFIGS["netflix"]
    d = synth.make_ratings(c["n"], c["m"], c["nnz"], seed=c["seed"])
    n, m, k = d["n"], d["m"], 100
    rng = np.random.default_rng(0)
    with H.Engine(n, m, k, flags=H.HIER) as e:
        e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
        for which, rows in ((H.THETA, n), (H.BETA, m)):
            shp = 0.3 + 0.01 * rng.random((rows, k))
            rate = 0.3 + 0.1 * rng.random((rows, k))
            e.set_state(which, shp, rate, shp / rate, np.log(shp / rate) - 0.5 / shp)
        for which, rows in ((H.THETARATE, n), (H.BETARATE, m)):
            e.set_state(which, np.full(rows, 0.3), np.full(rows, 100.3), np.full(rows, 0.3 / 100.3))
        e.iterate(2)
        th = e.get_state(H.THETA, fields=("shape",))["shape"]
        be = e.get_state(H.BETA, fields=("shape",))["shape"]
        assert e.stats()["slow_path_nnz"] == 0
    rp = d["row_ptr"].astype(np.int64)
    ysum_u = np.add.reduceat(np.concatenate([d["y"].astype(np.float64), [0.0]]), rp[:-1])[:n]
    ysum_u[rp[1:] == rp[:-1]] = 0.0
    ysum_i = np.bincount(d["col_idx"], weights=d["y"].astype(np.float64), minlength=m)
    np.testing.assert_allclose((th - 0.3).sum(1), ysum_u, rtol=2e-5, atol=1e-3)
    np.testing.assert_allclose((be - 0.3).sum(1), ysum_i, rtol=2e-5, atol=1e-3)
    assert np.isfinite(th).all() and np.isfinite(be).all()


def test_full_size_netflix_one_iteration_matches_oracle():
    """BASELINE config C2 at FULL size against the fp64 restatement of the reference loop (all host cores, rows are
    independent; ~1 minute): one iteration from the reference's own initialize() law, every element of every parameter
    set under the single-iteration gate of SURVEY.md 8c, and the held-out log-likelihood."""
    import os
    c = synth.CONFIGS["netflix"]
    d = synth.make_ratings(c["n"], c["m"], c["nnz"], seed=c["seed"], heldout=0.002)
    n, m, k = d["n"], d["m"], 100
    s = O.OracleState(n, m, k, H.HIER).init(20131104)
    with make_engine(s) as e:
        e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
        util.push_state(e, s)
        e.iterate(1)
        got = util.pull_state(e, s)
        hu, hi, hy = d["heldout"]
        ll = e.heldout_loglik(hu, hi, hy)
        st = e.stats()
    want = s.iterate(d["row_ptr"], d["col_idx"], d["y"], 1, nthreads=os.cpu_count() or 1)   # in place: 3 GB of fp64 state
    assert st["slow_path_nnz"] == 0 and st["head_nnz"] > 0                  # the default plan: dense head + gather tail
    bad = util.compare_states(got, want)
    assert not bad, bad
    assert abs(ll - want.heldout(hu, hi, hy)) / len(hu) <= util.TOL_LL_20IT


# ------------------------------------------------ -gen-ranking: scoring + top-N
def _check_topn(items, scores, o_items, o_scores, m, topn, rel=1e-4):
    """Engine (split-bf16 tensor-core scores, fp32) against the fp64 oracle:
    the sorted score lists agree to `rel`, and the item at a position agrees
    wherever the oracle's score there is separated from its neighbours by more
    than the tolerance (near-ties may legitimately swap)."""
    kk = min(topn, m)
    np.testing.assert_allclose(scores[:, :kk], o_scores[:, :kk], rtol=rel, atol=1e-12)
    if topn > m:
        assert (items[:, m:] == 0xFFFFFFFF).all() and (scores[:, m:] == 0).all()
    gap_hi = np.abs(np.diff(o_scores[:, :kk], axis=1, prepend=np.inf))
    # the list's last entry competes with an item we do not see: never "clear"
    gap_lo = np.abs(np.diff(o_scores[:, :kk], axis=1, append=o_scores[:, kk - 1:kk]))
    clear = (np.minimum(gap_hi, gap_lo) > 4 * rel * np.abs(o_scores[:, :kk])) & (o_scores[:, :kk] > 0)
    assert clear.mean() > 0.2
    assert (items[:, :kk][clear] == o_items[:, :kk][clear]).all()
    zero = o_scores[:, :kk] == 0  # excluded items: exact zeros, ties broken by ascending item
    if kk == m:
        assert (items[:, :kk][zero] == o_items[:, :kk][zero]).all() and (scores[:, :kk][zero] == 0).all()
    overlap = np.mean([len(set(a[:kk]) & set(b[:kk])) / kk for a, b in zip(items, o_items)])
    assert overlap >= 0.99, overlap


@pytest.mark.parametrize("name,n,m,nnz,k,flags,topn,iters", [
    ("hier K=100 top-100", 700, 1000, 40000, 100, H.HIER, 100, 2),
    ("bpf bias K=20 top-10", 300, 530, 9000, 20, H.BIAS, 10, 2),
    ("hier bias K=64 top-256, m < topn", 150, 200, 3000, 64, H.HIER | H.BIAS, 256, 1),
    ("hier K=130 (3 K-blocks) top-50", 260, 777, 12000, 130, H.HIER, 50, 1),
])
def test_topn_matches_oracle(name, n, m, nnz, k, flags, topn, iters):
    d, s = _oracle_case(n, m, nnz, k, flags, seed=17)
    s.iterate(d["row_ptr"], d["col_idx"], d["y"], iters, nthreads=8)  # a fitted-looking state
    rng = np.random.default_rng(5)
    users = rng.permutation(n)[: max(1, (2 * n) // 3)].astype(np.uint32)
    users[3] = users[0]  # a user may be listed twice
    rp = d["row_ptr"].astype(np.int64)
    excl = [d["col_idx"][rp[u]:rp[u + 1]] for u in users]  # training items, file order (unsorted)
    excl[1] = np.zeros(0, np.uint32)                          # a user without exclusions
    ep = np.zeros(len(users) + 1, np.uint64)
    ep[1:] = np.cumsum([len(x) for x in excl])
    ei = np.concatenate(excl).astype(np.uint32)
    o_items, o_scores = s.topn(users, ep, ei, topn)
    with make_engine(s) as e:
        util.push_state(e, s)
        items, scores = e.topn(users, ep, ei, topn)
        assert e.stats()["kernel_launches"] > 0
    _check_topn(items, scores.astype(np.float64), o_items, o_scores, m, topn)
    # excluded items never appear with a positive score
    for a in range(len(users)):
        hit = np.isin(items[a], excl[a]) & (items[a] != 0xFFFFFFFF)
        assert (scores[a][hit] == 0).all()


@pytest.mark.parametrize("flags,k", [(H.HIER, 100), (H.BIAS, 20)])
def test_item_ranks_match_full_sorted_list(flags, k):
    """hpf_item_ranks against the oracle's complete sorted list (topn with topn = m)."""
    n, m = 400, 900
    d, s = _oracle_case(n, m, 30000, k, flags, seed=29)
    s.iterate(d["row_ptr"], d["col_idx"], d["y"], 2, nthreads=8)
    rng = np.random.default_rng(7)
    users = rng.permutation(n)[:150].astype(np.uint32)
    rp = d["row_ptr"].astype(np.int64)
    excl = [d["col_idx"][rp[u]:rp[u + 1]] for u in users]
    ep = np.zeros(len(users) + 1, np.uint64)
    ep[1:] = np.cumsum([len(x) for x in excl])
    ei = np.concatenate(excl).astype(np.uint32)
    qs = [rng.choice(m, size=rng.integers(0, 25), replace=False).astype(np.uint32) for _ in users]
    qs[2] = np.concatenate([qs[2], excl[2][:3]]).astype(np.uint32)  # queries that are excluded items: rank among the zeros
    qs[5] = rng.choice(m, size=70, replace=False).astype(np.uint32)   # more than one batch of 32: several counting passes
    qs[140] = np.arange(m, dtype=np.uint32)                           # every item of one user: ranks are a permutation
    qp = np.zeros(len(users) + 1, np.uint64)
    qp[1:] = np.cumsum([len(x) for x in qs])
    qi = np.concatenate(qs).astype(np.uint32)
    o_items, o_scores = s.topn(users, ep, ei, m)
    with make_engine(s) as e:
        util.push_state(e, s)
        ranks, scores = e.item_ranks(users, ep, ei, qp, qi)
    exact = 0
    for a in range(len(users)):
        pos = np.empty(m, np.int64)
        pos[o_items[a]] = np.arange(m)
        for q in range(int(qp[a]), int(qp[a + 1])):
            it, want = qi[q], pos[qi[q]]
            sc = o_scores[a][want]
            assert abs(scores[q] - sc) <= 2e-5 * max(sc, 1e-12), (a, it)
            # fp32 vs fp64 may swap near-ties: allow the position to move within the run of almost-equal scores
            near = np.sum(np.abs(o_scores[a] - sc) <= 4e-6 * max(sc, 1e-30)) if sc > 0 else 1
            assert abs(int(ranks[q]) - want) < near, (a, it, ranks[q], want)
            exact += int(ranks[q]) == want
    assert exact >= 0.98 * len(qi)
    full = ranks[int(qp[140]):int(qp[141])]
    # all items queried: every position once -- up to pairs of scores closer than the fp32 / tensor-core rounding gap,
    # where each of the two may see the other one ahead
    assert int(full.max()) <= m - 1 and len(set(full.tolist())) >= m - 6


def test_topn_rejects_bad_arguments():
    d, s = _oracle_case(50, 40, 500, 8, H.HIER, seed=3)
    with make_engine(s) as e:
        util.push_state(e, s)
        ep = np.zeros(2, np.uint64)
        with pytest.raises(H.HpfError):
            e.topn(np.array([50], np.uint32), ep, np.zeros(0, np.uint32), 10)   # user out of range
        with pytest.raises(H.HpfError):
            e.topn(np.array([1], np.uint32), ep, np.zeros(0, np.uint32), 1000)  # topn too large
        ep[1] = 1
        with pytest.raises(H.HpfError):
            e.topn(np.array([1], np.uint32), ep, np.array([99], np.uint32), 10)  # excluded item >= m
