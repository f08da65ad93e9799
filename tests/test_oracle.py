"""CPU tests: the oracle (oracle/hpf_oracle.c) against the reference.

Pinned three ways: (1) committed golden states produced by the unmodified
reference binary (tests/golden/make_golden.py); (2) where /root/reference and the
oracle/_ref build exist, a fresh run of the reference harness; (3) digamma
against scipy (the GSL stand-in both the reference build and the oracle use)."""
import os
import tempfile

import numpy as np
import pytest

import util
from oracle import hpf_oracle as O
from hgaprec_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_digamma_matches_scipy():
    from scipy.special import digamma
    xs = np.concatenate([np.logspace(-30, 6, 300), np.linspace(0.3, 30, 400)])
    err = max(abs(O.digamma(x) - digamma(x)) / max(1.0, abs(digamma(x))) for x in xs)
    assert err < 1e-13


@pytest.mark.parametrize("mode", util.MODES)
def test_oracle_matches_reference_golden(mode):
    g = util.load_golden(mode)
    s = util.golden_state(g, 0)
    rp, ci, y = g["csr.row_ptr"], g["csr.col_idx"], g["csr.y"]
    done = 0
    for t in (1, 3):
        s.iterate(rp, ci, y, t - done, nthreads=1)
        done = t
        want = util.golden_state(g, t)
        for gname in util.groups(want):
            for f in O.FIELDS:
                np.testing.assert_allclose(s.p[gname][f], want.p[gname][f], rtol=1e-12, atol=1e-13,
                                           err_msg="%s %s.%s T=%d" % (mode, gname, f, t))
        for split in ("validation", "test"):
            ll = s.heldout(g[split + ".u"], g[split + ".i"], g[split + ".y"])
            ref = float(g["T%d/%s.ll_sum" % (t, split)][0])
            assert abs(ll - ref) <= 1e-9 * max(1.0, abs(ref)), (mode, split, t)


@pytest.mark.parametrize("mode", util.MODES)
@pytest.mark.parametrize("t", (1, 3))
def test_oracle_elbo_matches_reference_logl(mode, t):
    """hpf_oracle_elbo against the value the reference's own HGAPRec::logl() wrote to logl.txt on the same state
    (oracle/ref_harness.cc calls it; "%.5f"), both from the reference's state and from the oracle's own iterations."""
    g = util.load_golden(mode)
    rp, ci, y = g["csr.row_ptr"], g["csr.col_idx"], g["csr.y"]
    ref = float(g["T%d/elbo" % t][0])
    on_ref_state = util.golden_state(g, t).elbo(rp, ci, y, rate_prior=util.golden_rate_prior(g, t))
    assert abs(on_ref_state - ref) <= util.TOL_ELBO_REF_PRINT
    iterated = util.golden_state(g, 0).iterate(rp, ci, y, t)
    assert abs(iterated.elbo(rp, ci, y) - ref) <= util.TOL_ELBO_REF_PRINT
    # one call of t iterations and t calls of one iteration keep the same rate priors
    step = util.golden_state(g, 0)
    for _ in range(t):
        step.iterate(rp, ci, y, 1)
    assert step.elbo(rp, ci, y) == iterated.elbo(rp, ci, y)


def _device_formulas_elbo(s, rp, ci, y, pri):
    """numpy mirror of hgaprec_b200/csrc/hpf_elbo.cuh (same algebra, fp32-stored state, fp64 sums): per nonzero
    y^2 (logsumexp(x) - log y) - E[theta].E[beta] (- bias expectations); Gamma terms element by element with the
    hier rate prior rebuilt from the previous xi / eta (shape, rate)."""
    from scipy.special import digamma, gammaln
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    P = {g: {f: f32(s.p[g][f]) for f in O.FIELDS} for g in s.p}
    u = np.repeat(np.arange(s.n), np.diff(rp.astype(np.int64)))
    i = ci.astype(np.int64)
    yy = np.ones(len(i)) if y is None else y.astype(np.float64)
    x = P["theta"]["Elogv"][u] + P["beta"]["Elogv"][i]
    sub = (P["theta"]["Ev"][u] * P["beta"]["Ev"][i]).sum(1)
    if s.bias:
        x = np.concatenate([x, P["thetabias"]["Elogv"][u, None], P["betabias"]["Elogv"][i, None]], axis=1)
        sub = sub + P["thetabias"]["Ev"][u] + P["betabias"]["Ev"][i]
    mx = x.max(1)
    lse = mx + np.log(np.exp(x - mx[:, None]).sum(1))
    tot = float((yy * yy * (lse - np.log(yy)) - sub).sum())

    def gamma_terms(g, rate, rp_, lrp_):  # gamma_matrix_kernel
        a, b = np.maximum(g["shape"], 1e-30), np.maximum(rate, 1e-30)
        t = 0.3 * lrp_ + (0.3 - 1) * g["Elogv"] - (rp_ * g["Ev"] + gammaln(0.3))
        t = t - (a * np.log(b) + (a - 1) * g["Elogv"]) + b * g["Ev"] + gammaln(a)
        return float(t.sum())

    for gname, bname, k in (("theta", "thetarate", 0), ("beta", "betarate", 2)):
        g = P[gname]
        if s.hier:
            pa, pb = np.maximum(f32(pri[k]), 1e-30), np.maximum(f32(pri[k + 1]), 1e-30)  # previous (shape, rate)
            tot += gamma_terms(g, g["rate"], (pa / pb)[:, None], (digamma(pa) - np.log(pb))[:, None])
            a, b = np.maximum(P[bname]["shape"], 1e-30), np.maximum(P[bname]["rate"], 1e-30)  # gamma_array_kernel
            ev, el = a / b, digamma(a) - np.log(b)
            tot += float((0.3 * np.log(0.3) + (0.3 - 1) * el - (0.3 * ev + gammaln(0.3))
                          - (a * np.log(b) + (a - 1) * el) + b * ev + gammaln(a)).sum())
        else:
            tot += gamma_terms(g, g["rate"][None, :], 0.3, np.log(0.3))
    if s.bias:
        for gname in ("thetabias", "betabias"):
            g = P[gname]
            tot += gamma_terms(g, g["rate"], 0.3, np.log(0.3))
    return tot


@pytest.mark.parametrize("mode", util.MODES)
def test_device_elbo_algebra_matches_reference_logl(mode):
    """hpf_elbo.cuh does not walk the reference's literal sum: it uses y^2 (logsumexp - log y) per nonzero and
    rebuilds the hier rate priors from the previous xi / eta (shape, rate).  That algebra, on fp32-rounded state,
    must reproduce the reference's own logl() value (CPU check of the formulas; the kernels are checked on the GPU)."""
    g = util.load_golden(mode)
    rp, ci, y = g["csr.row_ptr"], g["csr.col_idx"], g["csr.y"]
    prev = util.golden_state(g, 0)
    done = 0
    for t in (1, 3):
        if t - done > 1:
            prev.iterate(rp, ci, y, t - done - 1)
        pri = (prev.p["thetarate"]["shape"], prev.p["thetarate"]["rate"], prev.p["betarate"]["shape"], prev.p["betarate"]["rate"])
        pri = tuple(a.copy() for a in pri)
        prev.iterate(rp, ci, y, 1)
        done = t
        ref = float(g["T%d/elbo" % t][0])
        got = _device_formulas_elbo(prev, rp, ci, y, pri)
        assert abs(got - ref) <= 2e-7 * abs(ref), (mode, t, got, ref)


@pytest.mark.parametrize("mode", ("hier", "bpf_bias"))
def test_oracle_init_matches_reference_golden(mode):
    g = util.load_golden(mode)
    want = util.golden_state(g, 0)
    s = O.OracleState(want.n, want.m, want.k, want.flags).init(777)
    for gname in util.groups(want):
        for f in O.FIELDS:
            np.testing.assert_array_equal(s.p[gname][f], want.p[gname][f], err_msg="%s.%s" % (gname, f))


@pytest.mark.parametrize("mode", ("hier", "bpf_bias"))
def test_threaded_oracle_equals_sequential(mode):
    g = util.load_golden(mode)
    a, b = util.golden_state(g, 0), util.golden_state(g, 0)
    rp, ci, y = g["csr.row_ptr"], g["csr.col_idx"], g["csr.y"]
    a.iterate(rp, ci, y, 3, nthreads=1)
    b.iterate(rp, ci, y, 3, nthreads=4)
    for gname in util.groups(a):
        np.testing.assert_allclose(b.p[gname]["Ev"], a.p[gname]["Ev"], rtol=1e-11)


def test_oracle_topn_orders_and_masks():
    g = util.load_golden("hier")
    s = util.golden_state(g, 3)
    users = np.array([0, 5, 7], dtype=np.uint32)
    rp, ci = g["csr.row_ptr"].astype(np.int64), g["csr.col_idx"]
    excl = [ci[rp[u]:rp[u + 1]] for u in users]
    ep = np.zeros(len(users) + 1, dtype=np.uint64)
    ep[1:] = np.cumsum([len(e) for e in excl])
    items, scores = s.topn(users, ep, np.concatenate(excl), 10)
    full = s.p["theta"]["Ev"][users] @ s.p["beta"]["Ev"].T
    for a, u in enumerate(users):
        full[a, excl[a]] = 0.0
        order = np.lexsort((np.arange(s.m), -full[a]))[:10]
        np.testing.assert_array_equal(items[a], order)
        np.testing.assert_allclose(scores[a], full[a, order], rtol=1e-12)
        assert not set(items[a]) & set(excl[a].tolist())


@pytest.mark.skipif(not (os.path.isdir("/root/reference/src") and os.path.exists(O.REF_HARNESS)),
                    reason="needs /root/reference and the oracle/_ref build")
def test_oracle_matches_live_reference_run():
    """Fresh reference run (different seed/shape than the goldens), K=12."""
    from hgaprec_b200 import synth
    d = synth.make_ratings(90, 60, 1500, seed=99, heldout=0.1, device="cpu")
    deg = np.diff(d["row_ptr"].astype(np.int64))
    tu = np.repeat(np.arange(90), deg)
    with tempfile.TemporaryDirectory() as tmp:
        data = os.path.join(tmp, "data")
        os.makedirs(data)
        with open(os.path.join(data, "train.tsv"), "w") as f:
            for a, b, c in zip(tu, d["col_idx"], d["y"]):
                f.write("%d\t%d\t%d\n" % (a + 1, b + 1, c))
        hu, hi, hy = d["heldout"]
        for name in ("validation.tsv", "test.tsv"):
            with open(os.path.join(data, name), "w") as f:
                for a, b, c in zip(hu, hi, hy):
                    f.write("%d\t%d\t%d\n" % (a + 1, b + 1, c))
        O.run_ref_harness(data, 90, 60, 12, [0], os.path.join(tmp, "d"), tmp, hier=True, bias=True, seed=5)
        O.run_ref_harness(data, 90, 60, 12, [2], os.path.join(tmp, "d"), tmp, hier=True, bias=True, seed=5)
        d0, d2 = O.read_dump(os.path.join(tmp, "d_0.bin")), O.read_dump(os.path.join(tmp, "d_2.bin"))
    s = O.state_from_dump(d0)
    s.iterate(d0["csr.row_ptr"], d0["csr.col_idx"], d0["csr.y"], 2)
    want = O.state_from_dump(d2)
    for gname in util.groups(want):
        for f in O.FIELDS:
            np.testing.assert_allclose(s.p[gname][f], want.p[gname][f], rtol=1e-12, atol=1e-13)
    # and the reference's own logl() on that state (the harness calls it; printed with "%.5f")
    assert abs(s.elbo(d0["csr.row_ptr"], d0["csr.col_idx"], d0["csr.y"]) - float(d2["elbo"][0])) <= util.TOL_ELBO_REF_PRINT


# ------------------------------------------------------------------ C1: the real MovieLens-1M fixture, K=100, 20 iterations
def test_oracle_and_host_reader_pinned_on_movielens_k100_t21(tmp_path):
    """BASELINE configs[0]: example/HPF-KDD-movielens.tgz with scripts/run.pl:109-111's flags, -hier, K=100, seed 111.
    The C++ host reader + start state (hgaprec_hostcheck) on the real files, then the plain-C restatement for 11 and
    21 iterations (all host cores; the states the CLI reports as "iteration 10" and "iteration 20": it counts from 0
    and reports after the update), against what the UNMODIFIED reference produced (tests/golden/movielens/ref_T21.npz:
    fp64 row sums, column sums and 4096 sampled entries of every matrix, its held-out sums).  Generation-time figures
    over the full states: bit-identical single-threaded, 1.6e-8 max rel threaded (oracle_pin.json)."""
    import json
    import subprocess
    host = os.path.join(ROOT, "hgaprec_b200", "host")
    subprocess.check_call(["make", "-C", host, "../bin/hgaprec_hostcheck"], stdout=subprocess.DEVNULL)
    pin = json.load(open(os.path.join(util.MOVIELENS, "oracle_pin.json")))
    assert pin["T21_max_rel_oracle_vs_reference_1thread"] == 0.0 and pin["T21_max_rel_oracle_vs_reference_threaded"] < 1e-6
    data = util.write_movielens(str(tmp_path / "movielens"))
    out = str(tmp_path / "dump.bin")
    subprocess.check_call([os.path.join(ROOT, "hgaprec_b200", "bin", "hgaprec_hostcheck"), "-dir", data, "-out", out] + util.MOVIELENS_FLAGS)
    d = O.read_dump(out)
    z = np.load(os.path.join(util.MOVIELENS, "ref_T21.npz"))
    # the reader on the real files: same seq numbering, same CSR (checksums), same held-out counts as the reference
    np.testing.assert_array_equal(d["seq2user"], z["seq2user"])
    np.testing.assert_array_equal(d["seq2movie"], z["seq2movie"])
    chk = [int(d["csr.row_ptr"].sum()), int(d["csr.col_idx"].astype(np.uint64).sum()), int(d["csr.y"].astype(np.uint64).sum()),
           len(d["csr.col_idx"]), len(d["validation.u"]), len(d["test.u"])]
    assert chk == [int(v) for v in z["csr.checksum"]]
    assert len(d["csr.col_idx"]) == pin["nnz_train"] == 792166
    s = O.state_from_dump(dict(d, meta=np.array([6040, 3681, 100, 0, 1, 0, 0, 1], dtype=np.float64)))
    csr = (d["csr.row_ptr"], d["csr.col_idx"], d["csr.y"])
    for T in (11, 21):
        s.iterate(*csr, 11 if T == 11 else 10, nthreads=os.cpu_count() or 1)
        fp = util.fingerprint(s)
        for key, val in fp.items():
            ref = z["T%d/fp/%s" % (T, key)]
            if key.endswith("sample_idx"):
                np.testing.assert_array_equal(val, ref)
            else:
                np.testing.assert_allclose(val, ref, rtol=2e-6, atol=1e-12, err_msg="T=%d %s" % (T, key))
        for split in ("validation", "test"):
            ll = s.heldout(d[split + ".u"], d[split + ".i"], d[split + ".y"])
            assert abs(ll - float(z["T%d/%s.ll_sum" % (T, split)][0])) <= 1e-6 * abs(ll), (T, split)
    # and the reference CLI's own report agrees with its harness dump (validation.txt row of iteration 20 = 21 sweeps)
    row = open(os.path.join(util.MOVIELENS, "cli", "validation.txt")).read().splitlines()[-1].split("\t")
    assert int(row[0]) == 20 and int(row[3]) == 8001
    assert abs(float(row[2]) - float(z["T21/validation.ll_sum"][0]) / 8001) <= 1e-8


def test_numpy_ranking_checkers_agree_with_the_oracle():
    """tests/util.py's check_topn_rows / check_rank_rows (used by the 60K-user GPU ranking test, where the oracle's
    full sort per user would take minutes) accept the oracle's own lists and reject a damaged one."""
    n, m, k = 300, 500, 20
    d = synth.make_ratings(n, m, 9000, seed=3)
    s = O.OracleState(n, m, k, 1).init(4)   # 1 = HPF_HIER
    s.iterate(d["row_ptr"], d["col_idx"], d["y"], 2)
    users = np.array([0, 7, 299, 150, 33], np.uint32)
    rp = d["row_ptr"].astype(np.int64)
    ep = np.zeros(len(users) + 1, np.uint64)
    ep[1:] = np.cumsum([rp[u + 1] - rp[u] for u in users])
    ei = np.concatenate([d["col_idx"][rp[u]:rp[u + 1]] for u in users]).astype(np.uint32)
    Et, Eb = s.p["theta"]["Ev"], s.p["beta"]["Ev"]
    items, scores = s.topn(users, ep, ei, 50)
    assert util.check_topn_rows(Et, Eb, users, ep, ei, items, scores, range(len(users))) == len(users)
    full, _ = s.topn(users, ep, ei, m)
    qp = np.arange(0, 4 * len(users) + 1, 4, dtype=np.uint64)
    qi = np.tile(np.array([3, 499, 250, 0], np.uint32), len(users))
    qi[0] = ei[0]                                              # an excluded item: ranks among the zeros
    ranks = np.array([int(np.where(full[a] == qi[q])[0][0]) for a in range(len(users)) for q in range(4 * a, 4 * a + 4)])
    assert util.check_rank_rows(Et, Eb, users, ep, ei, qp, qi, ranks, range(len(users))) == len(qi)
    bad = scores.copy(); bad[1, 10] *= 1.01
    with pytest.raises(AssertionError):
        util.check_topn_rows(Et, Eb, users, ep, ei, items, bad, range(len(users)))
    swapped = items.copy(); swapped[2, [0, 40]] = swapped[2, [40, 0]]
    with pytest.raises(AssertionError):
        util.check_topn_rows(Et, Eb, users, ep, ei, swapped, scores, range(len(users)))
    off = ranks.copy(); off[5] += 3
    with pytest.raises(AssertionError):
        util.check_rank_rows(Et, Eb, users, ep, ei, qp, qi, off, range(len(users)))
