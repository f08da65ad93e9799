#!/usr/bin/env python
"""Timings that are not part of the bench line (measurement aid, one GPU):
  elbo   hpf_elbo at the Netflix-scale workload (-hier, K=100), wall clock around the call (it synchronises)
  c4     one rank's shard of the BPF -bias 1e9-nnz workload (1.25M of 10M users x 1M items, K=100): iteration
         time with the dense tcgen05 head off / on (the head takes -bias since round 1f)
usage: python tools/bench_extras.py [elbo] [c4]      -> one JSON line per measurement on stdout"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hgaprec_b200 as H
from hgaprec_b200 import synth


def start_state(e, n, m, k, hier, bias, rng):
    def rs(rows):
        shp = 0.3 + 0.01 * rng.random((rows, k))
        rate = 0.3 + 0.1 * rng.random((rows, k))
        return shp, (rate if hier else rate[0]), shp / rate, np.log(shp / rate) - 0.5 / shp
    e.set_state(H.THETA, *rs(n))
    e.set_state(H.BETA, *rs(m))
    if hier:
        e.set_state(H.THETARATE, np.full(n, 0.3), np.full(n, 0.3 + k), np.full(n, 0.3 / (0.3 + k)))
        e.set_state(H.BETARATE, np.full(m, 0.3), np.full(m, 0.3 + k), np.full(m, 0.3 / (0.3 + k)))
    if bias:
        for which, rows, other in ((H.THETABIAS, n, m), (H.BETABIAS, m, n)):
            shp, rate = 0.3 + 0.01 * rng.random(rows), np.full(rows, 0.3 + other)
            e.set_state(which, shp, rate, shp / rate, np.log(shp / rate) - 0.5 / shp)


def elbo():
    cfg = synth.CONFIGS["netflix"]
    d = synth.make_ratings(cfg["n"], cfg["m"], cfg["nnz"], seed=cfg["seed"])
    n, m, k = d["n"], d["m"], cfg["k"]
    with H.Engine(n, m, k, flags=H.HIER | H.LOGL) as e:
        e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
        start_state(e, n, m, k, True, False, np.random.default_rng(1))
        e.iterate(3)
        e.elbo()
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            v = e.elbo()
            t.append((time.perf_counter() - t0) * 1e3)
        e.iterate(10)
        it_ms = e.stats()["last_iterate_ms"] / 10
    print(json.dumps({"what": "hpf_elbo, netflix-scale -hier K=100", "nnz": int(len(d["col_idx"])), "elbo": v,
                      "elbo_ms": min(t), "iteration_ms_with_HPF_LOGL": it_ms}), flush=True)


def c4():
    cfg = synth.CONFIGS["bpf-1b"]
    ranks = 8
    d = synth.make_ratings(cfg["n"], cfg["m"], cfg["nnz"], seed=cfg["seed"], users_lo=0, users_hi=cfg["n"] // ranks)
    n, m, k = d["n"], d["m"], cfg["k"]
    for mode in ("0", "-1"):
        os.environ["HPF_DENSE_HEAD"] = mode
        with H.Engine(n, m, k, flags=H.BIAS, n_users_global=cfg["n"]) as e:
            e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
            start_state(e, n, m, k, False, True, np.random.default_rng(2))
            e.iterate(3)
            p = e.iterate_profiled(5)
            e.iterate(10)
            st = e.stats()
        print(json.dumps({"what": "one of 8 shards of bpf-1b (-bias K=100)", "HPF_DENSE_HEAD": mode, "users": n, "items": m,
                          "nnz": int(len(d["col_idx"])), "head_nnz": int(st["head_nnz"]), "iteration_ms": st["last_iterate_ms"] / 10,
                          "profile_ms": {a: round(b, 4) for a, b in p.items()}}), flush=True)
    os.environ.pop("HPF_DENSE_HEAD", None)


if __name__ == "__main__":
    todo = sys.argv[1:] or ["elbo", "c4"]
    for name in todo:
        {"elbo": elbo, "c4": c4}[name]()
