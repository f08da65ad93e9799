#!/usr/bin/env python
"""Generate tests/golden/ref_*.npz from the UNMODIFIED reference.

Runs oracle/_ref/ref_harness (the reference's own sources behind a dump main,
built by oracle/Makefile from /root/reference/src) on a small seeded synthetic
data set in every mode of the hot path and stores, per mode: the training CSR
exactly as the reference walks it, the state right after initialize() (T=0),
the states after 1 and 3 iterations, and the reference's own held-out
log-likelihood sums.  Only runnable where /root/reference exists; the .npz
files are committed so tests can run anywhere.

    python tests/golden/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hpf_oracle as O  # noqa: E402
from hgaprec_b200 import synth  # noqa: E402

N, M, K, NNZ, SEED = 120, 80, 6, 1800, 424242
ITERS = (1, 3)
MODES = {
    # name: harness switches
    "hier": dict(hier=True),
    "hier_bias": dict(hier=True, bias=True),
    "hier_binary": dict(hier=True, binary=True, rating_threshold=3),
    "bpf": dict(),
    "bpf_bias": dict(bias=True),
    "bpf_bias_novb": dict(bias=True, novb=True),
}


def write_tsv(path, u, i, y, uid, iid):
    with open(path, "w") as f:
        for a, b, c in zip(u, i, y):
            f.write("%d\t%d\t%d\n" % (uid[a], iid[b], c))


def main():
    O.build()
    assert os.path.exists(O.REF_HARNESS), "needs oracle/_ref (make -C oracle ref)"
    d = synth.make_ratings(N, M, NNZ, seed=SEED, heldout=0.15, device="cpu")
    rng = np.random.default_rng(SEED)
    uid = rng.permutation(np.arange(1000, 1000 + N))   # arbitrary external ids
    iid = rng.permutation(np.arange(5000, 5000 + M))
    deg = np.diff(d["row_ptr"].astype(np.int64))
    tu = np.repeat(np.arange(N), deg)
    hu, hi, hy = d["heldout"]
    half = len(hu) // 3
    with tempfile.TemporaryDirectory() as tmp:
        data = os.path.join(tmp, "data")
        os.makedirs(data)
        write_tsv(os.path.join(data, "train.tsv"), tu, d["col_idx"], d["y"], uid, iid)
        write_tsv(os.path.join(data, "validation.tsv"), hu[:half], hi[:half], hy[:half], uid, iid)
        write_tsv(os.path.join(data, "test.tsv"), hu[half:], hi[half:], hy[half:], uid, iid)
        with open(os.path.join(data, "test_users.tsv"), "w") as f:
            for a in sorted(set(hu[half:].tolist())):
                f.write("%d\n" % uid[a])
        for name, sw in MODES.items():
            run = os.path.join(tmp, name)
            os.makedirs(run)
            kw = dict(hier=sw.get("hier", False), bias=sw.get("bias", False), binary=sw.get("binary", False),
                      novb=sw.get("novb", False), rating_threshold=sw.get("rating_threshold", 1), seed=777)
            O.run_ref_harness(data, N, M, K, [0], os.path.join(run, "d"), run, **kw)
            O.run_ref_harness(data, N, M, K, list(ITERS), os.path.join(run, "d"), run, **kw)
            out = {}
            for t in (0,) + ITERS:
                dump = O.read_dump(os.path.join(run, "d_%d.bin" % t))
                for key, val in dump.items():
                    if key.startswith(("csr.", "seq2", "validation.u", "validation.i", "validation.y",
                                       "test.u", "test.i", "test.y")):
                        if t == 0:
                            out[key] = val
                    else:
                        out["T%d/%s" % (t, key)] = val
            path = os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name)
            np.savez_compressed(path, **out)
            print("wrote", path, os.path.getsize(path), "bytes; n,m =", out["T0/meta"][:2],
                  "nnz =", len(out["csr.col_idx"]))


if __name__ == "__main__":
    main()
