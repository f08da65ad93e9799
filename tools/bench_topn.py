#!/usr/bin/env python
"""-gen-ranking at BASELINE config C5: theta.beta^T top-100 for every user of the Netflix-scale
problem (480,189 x 17,770, K=100), exclusion lists = each user's training items.
Prints one JSON line: wall time through the C ABI (host lists in, host top-100 out) and the
algorithmic tensor rate 2*nu*m*K / time (the kernel itself issues 3x that in split-bf16 MMAs)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hgaprec_b200 as H
from hgaprec_b200 import synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
c = synth.CONFIGS["netflix"]
d = synth.make_config("netflix", scale=scale)
n, m, k = d["n"], d["m"], 100
rng = np.random.default_rng(0)
with H.Engine(n, m, k, flags=H.HIER) as e:
    for which, rows in ((H.THETA, n), (H.BETA, m)):
        shp = rng.gamma(0.3, 1.0, size=(rows, k)) + 0.3
        rate = 0.3 + rng.random((rows, k)) * 10
        e.set_state(which, shp, rate, shp / rate, np.log(shp / rate))
    users = np.arange(n, dtype=np.uint32)
    e.topn(users[:4096], d["row_ptr"][:4097], d["col_idx"][: int(d["row_ptr"][4096])], 100)  # warm
    t0 = time.time()
    items, scores = e.topn(users, d["row_ptr"], d["col_idx"], 100)
    dt = time.time() - t0
    kernel_ms = e.stats()["last_topn_ms"]
    # spot check against numpy on a few users
    Et, Eb = e.get_state(H.THETA, ("Ev",))["Ev"], e.get_state(H.BETA, ("Ev",))["Ev"]
    ok = True
    for u in (0, n // 3, n - 1):
        sc = Et[u] @ Eb.T
        sc[d["col_idx"][int(d["row_ptr"][u]):int(d["row_ptr"][u + 1])]] = 0
        want = np.sort(sc)[::-1][:100]
        ok &= bool(np.allclose(scores[u], want, rtol=1e-4))
    # compute_itemrank at the same scale: every user, 20 query items each (the reference asks for the user's test hits)
    nq = 20
    qp = np.arange(0, (n + 1) * nq, nq, dtype=np.uint64)
    qi = rng.integers(0, m, n * nq).astype(np.uint32)
    e.item_ranks(users[:4096], d["row_ptr"][:4097], d["col_idx"][: int(d["row_ptr"][4096])], qp[:4097], qi[:4096 * nq])  # warm
    t0 = time.time()
    ranks, rscores = e.item_ranks(users, d["row_ptr"], d["col_idx"], qp, qi)
    dt_rank = time.time() - t0
    rank_kernel_ms = e.stats()["last_topn_ms"]
    for u in (0, n // 3, n - 1):
        sc = (Et[u] @ Eb.T).astype(np.float32)
        sc[d["col_idx"][int(d["row_ptr"][u]):int(d["row_ptr"][u + 1])]] = 0
        for j in range(nq):
            it = int(qi[u * nq + j])
            want_rank = int(np.sum(sc > sc[it]) + np.sum((sc == sc[it]) & (np.arange(m) < it)))
            ok &= abs(int(ranks[u * nq + j]) - want_rank) <= 2      # fp32 summation order near ties
flop = 2.0 * n * m * k
print(json.dumps({"item_ranks": {"what": "hpf_item_ranks, all users, %d queries each, host in/out" % nq, "seconds": dt_rank,
                                 "kernel_ms": rank_kernel_ms, "queries": int(n * nq)}, "what": "hpf_topn, all users, top-100, host in/out", "users": n, "items": m, "k": k, "excluded": int(len(d["col_idx"])),
                  "seconds": dt, "kernel_ms": kernel_ms, "kernel_algorithmic_tflops": flop / (kernel_ms * 1e-3) / 1e12,
                  "kernel_issued_tflops_bf16": 3 * 2.0 * n * (-(-m // 256) * 256) * 128 / (kernel_ms * 1e-3) / 1e12,
                  "algorithmic_tflops": flop / dt / 1e12, "scores_per_s": n * m / dt, "spot_check_ok": ok}))
