"""-gen-ranking (BASELINE config 5) at a production-like shape, on the GPU, against fp64 numpy.
Runs last in the suite (file name): it is the largest ranking case, the small ones against the oracle's own full sort
are in test_gpu_parity.py."""
import numpy as np
import pytest

import hgaprec_b200 as H
from hgaprec_b200 import synth
import util


def _ranking_case(n, m, nnz, make_engine, n_random_rows):
    """hpf_topn top-100 and hpf_item_ranks with 20 queries per user for ALL n users in one call each, every user's
    training items excluded.  Whole result: order, range, no duplicates; first / last / CTA-tile-boundary / random users
    against fp64 numpy."""
    k, topn, nq = 100, 100, 20
    d = synth.make_ratings(n, m, nnz, seed=31)
    rng = np.random.default_rng(5)
    users = np.arange(n, dtype=np.uint32)
    qp = np.arange(0, (n + 1) * nq, nq, dtype=np.uint64)
    qi = rng.integers(0, m, n * nq).astype(np.uint32)
    with make_engine(n, m, k) as e:
        for which, rows in ((H.THETA, n), (H.BETA, m)):
            shp = rng.gamma(0.3, 1.0, size=(rows, k)) + 0.3
            rate = 0.3 + rng.random((rows, k)) * 10
            e.set_state(which, shp, rate, shp / rate, np.log(shp / rate))
        items, scores = e.topn(users, d["row_ptr"], d["col_idx"], topn)
        ranks, rscores = e.item_ranks(users, d["row_ptr"], d["col_idx"], qp, qi)
        Et, Eb = e.get_state(H.THETA, ("Ev",))["Ev"], e.get_state(H.BETA, ("Ev",))["Ev"]
    assert items.shape == (n, topn) and (items < m).all()
    assert (np.diff(scores, axis=1) <= 0).all() and (scores[:, -1] > 0).all()
    assert (np.diff(np.sort(items, axis=1).astype(np.int64), axis=1) > 0).all()      # no item twice in a row's list
    assert (ranks < m).all() and (rscores >= 0).all()
    rows = sorted(set([0, 1, 127, 128, 129, 255, 256, n // 2, n - 129, n - 128, n - 2, n - 1]) |
                  set(rng.choice(n, n_random_rows, replace=False).tolist()))
    assert util.check_topn_rows(Et, Eb, users, d["row_ptr"], d["col_idx"], items, scores, rows) == len(rows)
    assert util.check_rank_rows(Et, Eb, users, d["row_ptr"], d["col_idx"], qp, qi, ranks, rows) == len(rows) * nq
    # a query's score is the score hpf_topn reports for the same (user, item)
    checked = 0
    for a in rows:
        hit = {int(i): float(s) for i, s in zip(items[a], scores[a])}
        for q in range(a * nq, (a + 1) * nq):
            if int(qi[q]) in hit:
                assert abs(rscores[q] - hit[int(qi[q])]) <= 1e-5 * hit[int(qi[q])], (a, q)
                checked += 1
    return checked


@pytest.mark.gpu
def test_gen_ranking_at_60k_users_against_numpy():
    """60,000 users x the 17,770 items of config 5, K=100: 200 users checked against numpy."""
    _ranking_case(60000, 17770, 12_600_000, lambda n, m, k: H.Engine(n, m, k, flags=H.HIER), 188)


class _NumpyStandIn:
    """The engine's ranking calls restated in fp32 numpy -- ONLY so that the checking code above can be exercised
    where there is no GPU (the driver's CPU run); nothing in the product knows it."""
    def __init__(self, n, m, k):
        self.Ev = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def set_state(self, which, shape, rate, Ev, Elog):
        self.Ev[which] = np.asarray(Ev, np.float32)

    def get_state(self, which, fields):
        return {"Ev": self.Ev[which].astype(np.float64)}

    def _scores(self, u, ep, ei, a):
        sc = self.Ev[H.THETA][u] @ self.Ev[H.BETA].T
        sc[ei[int(ep[a]):int(ep[a + 1])]] = 0
        return sc

    def topn(self, users, ep, ei, topn):
        items = np.empty((len(users), topn), np.uint32)
        scores = np.empty((len(users), topn), np.float32)
        for a, u in enumerate(users):
            sc = self._scores(u, ep, ei, a)
            order = np.lexsort((np.arange(len(sc)), -sc))[:topn]
            items[a], scores[a] = order, sc[order]
        return items, scores

    def item_ranks(self, users, ep, ei, qp, qi):
        ranks = np.zeros(len(qi), np.uint32)
        scores = np.zeros(len(qi), np.float32)
        for a, u in enumerate(users):
            sc = self._scores(u, ep, ei, a)
            idx = np.arange(len(sc))
            for q in range(int(qp[a]), int(qp[a + 1])):
                it = int(qi[q])
                ranks[q] = np.sum(sc > sc[it]) + np.sum((sc == sc[it]) & (idx < it))
                scores[q] = sc[it]
        return ranks, scores


def test_the_large_ranking_check_itself_on_a_numpy_stand_in():
    assert _ranking_case(700, 1300, 60000, _NumpyStandIn, 40) > 0
