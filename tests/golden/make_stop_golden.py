#!/usr/bin/env python
"""Generate tests/golden/cli_stop_<mode>/ : the reference CLI run TO ITS STOPPING RULE.

vb() and vb_bias() have no iteration cap (src/hgaprec.cc:919-980, 1219-1319): they end only when
compute_likelihood(validation) decides to stop (src/hgaprec.cc:1476-1500: after iteration 30, either the
validation log-likelihood improved by less than 1e-6 relative -- why 0 -- or it fell in more than two
report windows in a row -- why 1), which calls do_on_stop() (save_model + gen_ranking_for_users) and exit(0).
vb_hier() has the same rule under its -max-iterations cap.  This script runs oracle/_ref/hgaprec_ref (the
unmodified sources) with -rfreq 5 and no usable cap on the data set of tests/golden/ref_<mode>.npz and keeps
what the run leaves behind: validation.txt / test.txt (one row per report window), max.txt (iteration,
seconds, mean ll, why), precision.txt, the final model files and the ranking.tsv that do_on_stop() wrote.

Only runnable where /root/reference exists; the outputs are committed.

    python tests/golden/make_stop_golden.py
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

# mode -> (golden data set, CLI switches)
MODES = {
    "bpf": ("bpf", []),
    "bpf_bias": ("bpf_bias", ["-bias"]),
    "bpf_bias_novb": ("bpf_bias_novb", ["-bias", "-novb"]),
    "hier": ("hier", ["-hier", "-max-iterations", "2000"]),
}
RFREQ = 5
KEEP = ("validation.txt", "test.txt", "max.txt", "precision.txt", "meanrank.txt")


def main():
    O.build()
    assert os.path.exists(O.REF_BINARY)
    for mode, (gname, sw) in MODES.items():
        g = util.load_golden(gname)
        n, m, k = (int(v) for v in g["T0/meta"][:3])
        dst = os.path.join(ROOT, "tests", "golden", "cli_stop_" + mode)
        shutil.rmtree(dst, ignore_errors=True)
        os.makedirs(dst)
        with tempfile.TemporaryDirectory() as tmp:
            data = os.path.join(tmp, "data")
            util.write_dataset(g, data)
            cmd = [O.REF_BINARY, "-dir", data, "-n", str(n), "-m", str(m), "-k", str(k), "-seed", "777", "-label", "s",
                   "-rfreq", str(RFREQ)] + sw
            subprocess.run(cmd, cwd=tmp, stdout=subprocess.DEVNULL, timeout=600, check=True)  # exit(0) from the rule
            fit = [d for d in os.listdir(tmp) if d.startswith("n%d-" % n)]
            assert len(fit) == 1, fit
            fitdir = os.path.join(tmp, fit[0])
            open(os.path.join(dst, "dirname.txt"), "w").write(fit[0] + "\n")
            for f in sorted(os.listdir(fitdir)):
                if f in KEEP or (f.endswith(".tsv") and f not in ("ranking.tsv", "itemrank.tsv", "byusers.tsv", "byitems.tsv")):
                    shutil.copy(os.path.join(fitdir, f), os.path.join(dst, f))
            for f in ("ranking.tsv", "itemrank.tsv"):
                with open(os.path.join(fitdir, f), "rb") as fi, open(os.path.join(dst, f + ".gz"), "wb") as raw, \
                        gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as fo:
                    fo.write(fi.read())
        print(mode, open(os.path.join(dst, "max.txt")).read().strip(), sorted(os.listdir(dst)))


if __name__ == "__main__":
    main()
