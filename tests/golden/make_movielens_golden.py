#!/usr/bin/env python
"""Generate tests/golden/movielens/ -- BASELINE.json configs[0] (C1), the only data set the reference ships.

Runs the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile from /root/reference/src) on
/root/reference/example/HPF-KDD-movielens.tgz with the flag set of scripts/run.pl:109-111 plus -hier, K=100
(SURVEY.md Appendix D):

    hgaprec -dir movielens -n 6040 -m 3681 -k 100 -rfreq 10 -rating-threshold 4 -hier -seed 111 -max-iterations 20

and keeps
  data.npz            the four input files as compact integer arrays in FILE ORDER (seq ids are first-appearance
                      ordinals, src/ratings.hh:117-151, so the order is part of the input); tests write them back
                      out as TSVs -- the GPU box has no /root/reference
  cli/                validation.txt, test.txt, precision.txt, max.txt, param.txt as the reference CLI wrote them
  ref_T21.npz         E[theta], E[beta] after 21 iterations from oracle/_ref/ref_harness -- the state the CLI saves in its
                      report window of iteration 20: vb_hier numbers its iterations from 0 and reports AFTER the update
                      (src/hgaprec.cc:1337-1425), so "iteration 10 / 20" in validation.txt is the state after 11 / 21
                      sweeps -- as fp32 copies (the 1e-4 relative-Frobenius gate is three orders above fp32 rounding),
                      the reference's own held-out sums, plus an fp64 FINGERPRINT of the full T=11 / T=21 states
                      (row sums, column sums and 4096 sampled entries of every matrix) that pins oracle/hpf_oracle.c
                      at K=100 over 20 iterations without committing 30 MB of doubles
  oracle_pin.json     max relative difference oracle-vs-reference over the FULL fp64 states at T=11 / T=21, measured
                      here at generation time (the CPU test re-checks it through the fingerprint anywhere, and in
                      full where /root/reference exists)

Only runnable where /root/reference exists (about 5 minutes on one core); the outputs are committed.

    python tests/golden/make_movielens_golden.py
"""
import json
import os
import shutil
import subprocess
import sys
import tarfile
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hpf_oracle as O  # noqa: E402

TGZ = "/root/reference/example/HPF-KDD-movielens.tgz"
N, M, K, SEED, RFREQ, MAXIT, THRESH = 6040, 3681, 100, 111, 10, 20, 4
DST = os.path.join(ROOT, "tests", "golden", "movielens")
KEEP = ("validation.txt", "test.txt", "precision.txt", "max.txt", "param.txt")
SAMPLE = 4096


def read_triples(path):
    a = np.loadtxt(path, dtype=np.int64, delimiter="\t")
    assert a.shape[1] == 3 and a[:, :2].max() < 65536 and a[:, 2].max() < 256
    return a[:, 0].astype(np.uint16), a[:, 1].astype(np.uint16), a[:, 2].astype(np.uint8)


def fingerprint(state, rng_seed=20131103):
    """fp64 row sums, column sums and SAMPLE fixed entries of every array of an OracleState."""
    rng = np.random.default_rng(rng_seed)
    out = {}
    for g in ("theta", "beta", "thetarate", "betarate"):
        for f in O.FIELDS:
            a = state.p[g][f]
            key = "%s.%s" % (g, f)
            if a.ndim == 2:
                out[key + ".rowsum"] = a.sum(axis=1)
                out[key + ".colsum"] = a.sum(axis=0)
                idx = rng.integers(0, a.size, SAMPLE)
                out[key + ".sample_idx"] = idx.astype(np.uint32)
                out[key + ".sample"] = a.reshape(-1)[idx]
            else:
                out[key] = a.copy()
    return out


def main():
    O.build()
    assert os.path.exists(O.REF_BINARY) and os.path.exists(O.REF_HARNESS), "needs oracle/_ref (make -C oracle ref)"
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(os.path.join(DST, "cli"))
    with tempfile.TemporaryDirectory() as tmp:
        with tarfile.open(TGZ) as t:
            t.extractall(tmp)
        data = os.path.join(tmp, "movielens")
        tr, va, te = (read_triples(os.path.join(data, f + ".tsv")) for f in ("train", "validation", "test"))
        test_users = np.loadtxt(os.path.join(data, "test_users.tsv"), dtype=np.int64).astype(np.uint16)
        np.savez_compressed(os.path.join(DST, "data.npz"), train_u=tr[0], train_i=tr[1], train_y=tr[2],
                            validation_u=va[0], validation_i=va[1], validation_y=va[2],
                            test_u=te[0], test_i=te[1], test_y=te[2], test_users=test_users)
        # ---- the reference CLI
        base = ["-dir", data, "-n", str(N), "-m", str(M), "-k", str(K), "-rating-threshold", str(THRESH), "-hier", "-seed", str(SEED)]
        cli = subprocess.Popen([O.REF_BINARY] + base + ["-rfreq", str(RFREQ), "-max-iterations", str(MAXIT), "-label", "ml"],
                               cwd=tmp, stdout=subprocess.DEVNULL)
        # ---- the same run behind the dump harness (T=0 needs its own process)
        run = os.path.join(tmp, "h")
        os.makedirs(run)
        h0 = subprocess.Popen([O.REF_HARNESS] + base + ["-iters", "0", "-dump", os.path.join(run, "d"), "-label", "h0"],
                              cwd=run, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        h1 = subprocess.Popen([O.REF_HARNESS] + base + ["-iters", "11,21", "-dump", os.path.join(run, "d"), "-label", "h1"],
                              cwd=run, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert h0.wait() == 0
        d0 = O.read_dump(os.path.join(run, "d_0.bin"))
        s = O.state_from_dump(d0)
        csr = (d0["csr.row_ptr"], d0["csr.col_idx"], d0["csr.y"])
        # ---- the plain-C restatement from the reference's own T=0 state, all cores (rows are independent)
        nt = os.cpu_count() or 1
        s10 = s.copy().iterate(*csr, 11, nthreads=nt)
        s20 = s10.copy().iterate(*csr, 10, nthreads=nt)
        assert h1.wait() == 0 and cli.wait() == 0
        # ... and single-threaded (the reference's own summation order)
        q10 = s.copy().iterate(*csr, 11, nthreads=1)
        q20 = q10.copy().iterate(*csr, 10, nthreads=1)
        pin, out = {"threads": nt}, {}

        def worst_rel(mine, ref):
            w = 0.0
            for g in ("theta", "beta", "thetarate", "betarate"):
                for f in O.FIELDS:
                    a, b = mine.p[g][f], ref.p[g][f]
                    w = max(w, float((np.abs(a - b) / np.maximum(np.abs(b), 1e-300)).max()))
            return w
        for T, mine, single in ((11, s10, q10), (21, s20, q20)):
            d = O.read_dump(os.path.join(run, "d_%d.bin" % T))
            ref = O.state_from_dump(d)
            pin["T%d_max_rel_oracle_vs_reference_threaded" % T] = worst_rel(mine, ref)
            pin["T%d_max_rel_oracle_vs_reference_1thread" % T] = worst_rel(single, ref)
            for kk, v in fingerprint(ref).items():
                out["T%d/fp/%s" % (T, kk)] = v
            for split in ("validation", "test"):
                out["T%d/%s.ll_sum" % (T, split)] = d[split + ".ll_sum"]
            if T == 21:
                out["T%d/theta.Ev" % T] = ref.p["theta"]["Ev"].astype(np.float32)
                out["T%d/beta.Ev" % T] = ref.p["beta"]["Ev"].astype(np.float32)
        for kk in ("seq2user", "seq2movie"):   # the CSR itself is re-derived from data.npz by the host reader under test
            out[kk] = d0[kk]
        out["csr.checksum"] = np.array([int(d0["csr.row_ptr"].sum()), int(d0["csr.col_idx"].astype(np.uint64).sum()),
                                        int(d0["csr.y"].astype(np.uint64).sum()), len(d0["csr.col_idx"]),
                                        len(d0["validation.u"]), len(d0["test.u"])], dtype=np.uint64)
        np.savez_compressed(os.path.join(DST, "ref_T21.npz"), **out)
        fit = [x for x in os.listdir(tmp) if x.startswith("n%d-" % N)]
        assert len(fit) == 1, fit
        for f in KEEP:
            shutil.copy(os.path.join(tmp, fit[0], f), os.path.join(DST, "cli", f))
        open(os.path.join(DST, "cli", "dirname.txt"), "w").write(fit[0] + "\n")
        pin["flags"] = "-n %d -m %d -k %d -rfreq %d -rating-threshold %d -hier -seed %d -max-iterations %d" % (N, M, K, RFREQ, THRESH, SEED, MAXIT)
        pin["nnz_train"] = int(len(d0["csr.col_idx"]))
        json.dump(pin, open(os.path.join(DST, "oracle_pin.json"), "w"), indent=1, sort_keys=True)
        print(json.dumps(pin))
    for f in sorted(os.listdir(DST)):
        p = os.path.join(DST, f)
        print(f, os.path.getsize(p) if os.path.isfile(p) else sorted(os.listdir(p)))


if __name__ == "__main__":
    main()
