"""Shared helpers for the parity tests: golden fixtures, oracle <-> engine state
transfer and the comparison metrics the tolerances are stated in."""
import os

import numpy as np

from oracle import hpf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODES = ("hier", "hier_bias", "hier_binary", "bpf", "bpf_bias", "bpf_bias_novb")

# fp32 gates, SURVEY.md 8c "stated tolerance" (derived from the reference rebuilt
# in float): single iteration from identical state -- max rel 2e-5 on shapes /
# rates / E, abs 2e-5 on Elog; 20 iterations -- relFro 1e-4, |d mean ll| 2e-4.
TOL_REL_1IT = 2e-5
TOL_ELOG_ABS_1IT = 2e-5
TOL_RELFRO_20IT = 1e-4
TOL_LL_20IT = 2e-4


def load_golden(mode):
    z = np.load(os.path.join(GOLDEN, "ref_%s.npz" % mode))
    return {k: z[k] for k in z.files}


def golden_state(g, t):
    """OracleState holding the reference's own numbers after t iterations."""
    d = {k[len("T%d/" % t):]: v for k, v in g.items() if k.startswith("T%d/" % t)}
    return O.state_from_dump(d)


def golden_rate_prior(g, t):
    """The rate priors HGAPRec::logl() saw after t iterations of a -hier golden run (_hier_rprior, _hier_log_rprior
    of htheta and hbeta, gpbase.hh:163-173), or None for the non-hier modes."""
    keys = ["T%d/%s" % (t, k) for k in ("htheta.hier_rprior", "htheta.hier_log_rprior", "hbeta.hier_rprior", "hbeta.hier_log_rprior")]
    return tuple(g[k] for k in keys) if keys[0] in g else None


TOL_ELBO_REF_PRINT = 6e-6  # the reference prints logl with "%.5f" (hgaprec.cc:2253)


def groups(state):
    gs = ["theta", "beta"]
    if state.hier:
        gs += ["thetarate", "betarate"]
    if state.bias:
        gs += ["thetabias", "betabias"]
    return gs


_IDS = {"theta": 0, "beta": 1, "thetarate": 2, "betarate": 3, "thetabias": 4, "betabias": 5}


def engine_flags(state):
    return state.flags  # HIER/BIAS/BINARY/JACOBI share values with hpf_cuda.h


def push_state(engine, state, users=None):
    """hpf_set_state for every parameter set; `users` selects a shard's rows."""
    for gname in groups(state):
        p = state.p[gname]
        sl = (lambda a: a) if users is None or gname.startswith("beta") else (lambda a: a[users])
        rate = p["rate"]
        if gname in ("theta", "beta") and not state.hier:
            rate_arg = rate  # k-vector, shared
        else:
            rate_arg = sl(rate)
        engine.set_state(_IDS[gname], sl(p["shape"]), rate_arg, sl(p["Ev"]), sl(p["Elogv"]))


def pull_state(engine, like):
    out = O.OracleState(engine.n, engine.m, engine.k, like.flags)
    for gname in groups(like):
        got = engine.get_state(_IDS[gname])
        for f in O.FIELDS:
            out.p[gname][f][...] = got[f].reshape(out.p[gname][f].shape)
    return out


def max_rel(a, b, floor=1e-300):
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max()) if a.size else 0.0


def rel_fro(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def compare_states(got, want, rel=TOL_REL_1IT, elog_abs=TOL_ELOG_ABS_1IT):
    """Return list of violations of the single-iteration gate."""
    bad = []
    for gname in groups(want):
        for f in ("shape", "rate", "Ev"):
            r = max_rel(got.p[gname][f], want.p[gname][f])
            if not r <= rel:
                bad.append("%s.%s max rel %.3g > %.3g" % (gname, f, r, rel))
        d = float(np.abs(got.p[gname]["Elogv"] - want.p[gname]["Elogv"]).max())
        if not d <= elog_abs:
            bad.append("%s.Elogv max abs %.3g > %.3g" % (gname, d, elog_abs))
    return bad


def write_dataset(g, path, extra_lines=(), lift=0):
    """TSVs that reproduce the golden's CSR: users in seq order, rows in walk order.
    `lift` raises every rating to at least that value (the -binary-data goldens hold
    the 1s the reference stored, the file must pass its -rating-threshold)."""
    os.makedirs(path, exist_ok=True)
    uid, iid = g["seq2user"], g["seq2movie"]
    rp = g["csr.row_ptr"].astype(np.int64)
    with open(os.path.join(path, "train.tsv"), "w") as f:
        for u in range(len(rp) - 1):
            for j in range(rp[u], rp[u + 1]):
                f.write("%d\t%d\t%d\n" % (uid[u], iid[g["csr.col_idx"][j]], max(int(g["csr.y"][j]), lift)))
        for line in extra_lines:
            f.write(line)
    for split in ("validation", "test"):
        with open(os.path.join(path, split + ".tsv"), "w") as f:
            for u, i, y in zip(g[split + ".u"], g[split + ".i"], g[split + ".y"]):
                f.write("%d\t%d\t%d\n" % (uid[u], iid[i], max(int(y), lift)))
            f.write("%d\t%d\t%d\n" % (uid[0], 987654, 3))   # item never seen in training: dropped (ratings.cc:79-81)
    with open(os.path.join(path, "test_users.tsv"), "w") as f:
        for u in sorted(set(g["test.u"].tolist())):
            f.write("%d\n" % uid[u])


# ---- the real MovieLens-1M fixture (BASELINE configs[0]); tests/golden/make_movielens_golden.py
MOVIELENS = os.path.join(GOLDEN, "movielens")
MOVIELENS_FLAGS = ["-n", "6040", "-m", "3681", "-k", "100", "-rating-threshold", "4", "-hier", "-seed", "111"]


def write_movielens(path):
    """train.tsv / validation.tsv / test.tsv / test_users.tsv exactly as the reference's example archive holds them
    (same lines in the same order), from the compact copy under tests/golden/movielens/data.npz."""
    os.makedirs(path, exist_ok=True)
    z = np.load(os.path.join(MOVIELENS, "data.npz"))
    for split in ("train", "validation", "test"):
        a = np.stack([z[split + "_u"].astype(np.int64), z[split + "_i"].astype(np.int64), z[split + "_y"].astype(np.int64)], 1)
        np.savetxt(os.path.join(path, split + ".tsv"), a, fmt="%d", delimiter="\t")
    np.savetxt(os.path.join(path, "test_users.tsv"), z["test_users"].astype(np.int64), fmt="%d")
    return path


def fingerprint(state, rng_seed=20131103, sample=4096):
    """Same reduction as make_movielens_golden.fingerprint: fp64 row sums, column sums and fixed sampled entries."""
    rng = np.random.default_rng(rng_seed)
    out = {}
    for g in ("theta", "beta", "thetarate", "betarate"):
        for f in O.FIELDS:
            a = state.p[g][f]
            key = "%s.%s" % (g, f)
            if a.ndim == 2:
                out[key + ".rowsum"] = a.sum(axis=1)
                out[key + ".colsum"] = a.sum(axis=0)
                idx = rng.integers(0, a.size, sample)
                out[key + ".sample_idx"] = idx.astype(np.uint32)
                out[key + ".sample"] = a.reshape(-1)[idx]
            else:
                out[key] = a.copy()
    return out


# ---------------------------------------------------------------- ranking checks against plain numpy (any size)
def _scores_fp64(Et, Eb, u, excl):
    sc = np.asarray(Et[u], np.float64) @ np.asarray(Eb, np.float64).T
    sc[excl] = 0.0          # excluded items stay candidates with score 0 (src/hgaprec.cc:1736-1741)
    return sc


def check_topn_rows(Et, Eb, users, excl_ptr, excl_idx, items, scores, rows, rel=1e-4):
    """For the listed rows of a topn result: the scores are the best topn of E[theta_u].E[beta]^T (excluded -> 0) within
    `rel`, every returned item really has the score returned with it, and no item appears twice.  Returns the number
    of rows checked; raises AssertionError with the row otherwise."""
    topn = items.shape[1]
    for a in rows:
        u = int(users[a])
        sc = _scores_fp64(Et, Eb, u, excl_idx[int(excl_ptr[a]):int(excl_ptr[a + 1])])
        want = np.sort(sc)[::-1][:topn]
        got = np.asarray(scores[a], np.float64)
        assert np.all(np.abs(got - want) <= rel * np.maximum(want, 1e-30)), ("scores", a, u)
        assert np.all(np.abs(sc[items[a]] - got) <= rel * np.maximum(got, 1e-30)), ("item/score pairing", a, u)
        assert len(set(items[a].tolist())) == topn, ("duplicate item", a, u)
        assert np.all(np.diff(got) <= 0), ("order", a, u)
    return len(rows)


def check_rank_rows(Et, Eb, users, excl_ptr, excl_idx, query_ptr, query_idx, ranks, rows, rel=4e-6):
    """For the listed rows of an item_ranks result: rank = number of items ahead in the descending list (ties by
    ascending item), allowed to move within the run of scores closer than `rel` to the query's."""
    m = Eb.shape[0]
    idx = np.arange(m)
    n = 0
    for a in rows:
        u = int(users[a])
        sc = _scores_fp64(Et, Eb, u, excl_idx[int(excl_ptr[a]):int(excl_ptr[a + 1])])
        for q in range(int(query_ptr[a]), int(query_ptr[a + 1])):
            it = int(query_idx[q])
            want = int(np.sum(sc > sc[it]) + np.sum((sc == sc[it]) & (idx < it)))
            near = int(np.sum(np.abs(sc - sc[it]) <= rel * max(sc[it], 1e-30))) if sc[it] > 0 else 1
            assert abs(int(ranks[q]) - want) < near + 1, ("rank", a, u, it, int(ranks[q]), want, near)
            n += 1
    return n
