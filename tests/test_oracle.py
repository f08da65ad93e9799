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
