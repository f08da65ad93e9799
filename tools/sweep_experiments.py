#!/usr/bin/env python
"""Timing sweep over the engine's tuning knobs (env vars read at hpf_create) on
one workload; prints one line per setting.  Measurement aid, not a test."""
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hgaprec_b200 as H
from hgaprec_b200 import synth


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "netflix"
    knobs = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {"HPF_SWEEP_G": ["4", "8", "16", "32"],
                                                              "HPF_SEG_LEN": ["128", "256", "512"]}
    cfg = dict(synth.CONFIGS[workload])
    k = cfg["k"]
    d = synth.make_ratings(cfg["n"], cfg["m"], cfg["nnz"], binary=cfg["binary"], seed=cfg["seed"])
    n, m = d["n"], d["m"]
    rng = np.random.default_rng(1)

    def rs(rows):
        shp = 0.3 + 0.01 * rng.random((rows, k))
        rate = 0.3 + 0.1 * rng.random((rows, k))
        return shp, rate, shp / rate, np.log(shp / rate) - 0.5 / shp
    st_t, st_b = rs(n), rs(m)
    names = sorted(knobs)
    for combo in itertools.product(*[knobs[nm] for nm in names]):
        for nm, v in zip(names, combo):
            os.environ[nm] = v
        try:
            with H.Engine(n, m, k, flags=H.HIER | (H.BINARY if cfg["binary"] else 0)) as e:
                e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
                e.set_state(H.THETA, *st_t)
                e.set_state(H.BETA, *st_b)
                e.set_state(H.THETARATE, np.full(n, 0.3), np.full(n, 0.3 + k), np.full(n, 0.3 / (0.3 + k)))
                e.set_state(H.BETARATE, np.full(m, 0.3), np.full(m, 0.3 + k), np.full(m, 0.3 / (0.3 + k)))
                e.iterate(3)
                p = e.iterate_profiled(5)
                e.iterate(10)
                ms = e.stats()["last_iterate_ms"] / 10
            print(dict(zip(names, combo)), "iter %.3f ms |" % ms, " ".join("%s=%.3f" % (a[:-3], b) for a, b in p.items()),
                  flush=True)
        except Exception as ex:
            print(dict(zip(names, combo)), "FAILED", ex, flush=True)


if __name__ == "__main__":
    main()
