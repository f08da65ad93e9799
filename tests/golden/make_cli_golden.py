#!/usr/bin/env python
"""Generate tests/golden/cli_<mode>/ from the UNMODIFIED reference command line.

Runs oracle/_ref/hgaprec_ref (src/main.cc ... built by oracle/Makefile) on the
data set of tests/golden/ref_<mode>.npz, first a fit (-rfreq 2 -max-iterations 4 -logl:
logl.txt gets one ELBO line per report window, nothing else changes),
then -gen-ranking from inside the fit's output directory,
and keeps the report files the new driver has to reproduce: validation.txt,
test.txt, max.txt, precision.txt, the model TSVs and ranking.tsv.  Only runnable
where /root/reference exists; the outputs are committed.

    python tests/golden/make_cli_golden.py
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from oracle import hpf_oracle as O  # noqa: E402

MODES = {"hier": ["-hier"], "hier_bias": ["-hier", "-bias"]}
KEEP = ("validation.txt", "test.txt", "max.txt", "precision.txt", "param.txt", "byusers.tsv", "byitems.tsv", "logl.txt")


def main():
    O.build()
    assert os.path.exists(O.REF_BINARY)
    for mode, sw in MODES.items():
        g = util.load_golden("hier" if mode == "hier" else mode)
        n, m, k = (int(v) for v in g["T0/meta"][:3])
        dst = os.path.join(ROOT, "tests", "golden", "cli_" + mode)
        shutil.rmtree(dst, ignore_errors=True)
        os.makedirs(os.path.join(dst, "fit"))
        os.makedirs(os.path.join(dst, "rank"))
        with tempfile.TemporaryDirectory() as tmp:
            data = os.path.join(tmp, "data")
            util.write_dataset(g, data)
            base = [O.REF_BINARY, "-dir", data, "-n", str(n), "-m", str(m), "-k", str(k), "-seed", "777", "-label", "g"] + sw
            subprocess.check_call(base + ["-rfreq", "2", "-max-iterations", "4", "-logl"], cwd=tmp, stdout=subprocess.DEVNULL)
            fit = [d for d in os.listdir(tmp) if d.startswith("n%d-" % n)]
            assert len(fit) == 1, fit
            fitdir = os.path.join(tmp, fit[0])
            open(os.path.join(dst, "dirname.txt"), "w").write(fit[0] + "\n")
            for f in os.listdir(fitdir):
                if f in KEEP or (f.endswith(".tsv") and f not in ("ranking.tsv", "itemrank.tsv")):
                    shutil.copy(os.path.join(fitdir, f), os.path.join(dst, "fit", f))
            subprocess.check_call(base + ["-gen-ranking"], cwd=fitdir, stdout=subprocess.DEVNULL)
            rank = os.path.join(fitdir, fit[0])
            for f in ("precision.txt", "meanrank.txt"):
                shutil.copy(os.path.join(rank, f), os.path.join(dst, "rank", f))
            for f in ("ranking.tsv", "itemrank.tsv"):
                with open(os.path.join(rank, f), "rb") as fi, open(os.path.join(dst, "rank", f + ".gz"), "wb") as raw, \
                        gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as fo:  # mtime=0: reruns are byte-identical
                    fo.write(fi.read())
        print("wrote", dst, sorted(os.listdir(os.path.join(dst, "fit"))), sorted(os.listdir(os.path.join(dst, "rank"))))


if __name__ == "__main__":
    main()
