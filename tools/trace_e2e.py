#!/usr/bin/env python
"""Where the end-to-end step goes: HPF_TRACE=1 phases of hpf_set_ratings_csr at Netflix scale."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["HPF_TRACE"] = "1"
import numpy as np, torch
import hgaprec_b200 as H
from hgaprec_b200 import synth
c = synth.CONFIGS["netflix"]
d = synth.make_ratings(c["n"], c["m"], c["nnz"], seed=c["seed"], heldout=0.002)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
rp, ci, y = pin(d["row_ptr"]), pin(d["col_idx"]), pin(d["y"])
hu, hi, hy = d["heldout"]
e = H.Engine(d["n"], d["m"], 100, flags=H.HIER)
for rep in range(3):
    t0 = time.time(); e.set_ratings_csr(rp, ci, y); print("set_ratings_csr total %.1f ms" % ((time.time() - t0) * 1e3), file=sys.stderr)
rng = np.random.default_rng(0)
for which, rows in ((H.THETA, d["n"]), (H.BETA, d["m"])):
    shp = 0.3 + 0.01 * rng.random((rows, 100)); rate = 0.3 + 0.1 * rng.random((rows, 100))
    t0 = time.time(); e.set_state(which, shp, rate, shp / rate, np.log(shp / rate)); print("set_state %d %.1f ms" % (which, (time.time() - t0) * 1e3), file=sys.stderr)
for which, rows in ((H.THETARATE, d["n"]), (H.BETARATE, d["m"])):
    e.set_state(which, np.full(rows, 0.3), np.full(rows, 100.3), np.full(rows, 0.3 / 100.3))
for rep in range(3):
    t0 = time.time(); e.iterate(1); t1 = time.time(); ll = e.heldout_loglik(hu, hi, hy); t2 = time.time()
    print("iterate(1) %.1f ms, heldout(%d pairs) %.1f ms" % ((t1 - t0) * 1e3, len(hu), (t2 - t1) * 1e3), file=sys.stderr)
t0 = time.time(); st = e.get_state(H.THETA); print("get_state THETA %.1f ms" % ((time.time() - t0) * 1e3), file=sys.stderr)
