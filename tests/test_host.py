"""Host-side logic of the C++ driver (hgaprec_b200/host) against the reference's own
dumps -- no GPU involved.  tests/golden/ref_*.npz were written by the unmodified
reference behind oracle/ref_harness.cc: the CSR as the reference walks it, its id
maps, its held-out maps and the state right after HGAPRec::initialize().  The
driver's reader + CSR builder + start state must reproduce them from the same TSVs."""
import os
import subprocess

import numpy as np
import pytest

import util
from oracle import hpf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "hgaprec_b200", "host")
CHECK = os.path.join(ROOT, "hgaprec_b200", "bin", "hgaprec_hostcheck")

SWITCHES = {
    "hier": ["-hier"], "hier_bias": ["-hier", "-bias"], "hier_binary": ["-hier", "-binary-data", "-rating-threshold", "3"],
    "bpf": [], "bpf_bias": ["-bias"], "bpf_bias_novb": ["-bias", "-novb"],
}


@pytest.fixture(scope="module")
def hostcheck():
    subprocess.check_call(["make", "-C", HOST, "../bin/hgaprec_hostcheck"], stdout=subprocess.DEVNULL)
    return CHECK


@pytest.mark.parametrize("mode", util.MODES)
def test_reader_csr_and_start_state_match_reference(hostcheck, tmp_path, mode):
    g = util.load_golden(mode)
    n, m, k = (int(v) for v in g["T0/meta"][:3])
    data = str(tmp_path / "data")
    util.write_dataset(g, data, lift=3 if "binary" in mode else 0)
    out = str(tmp_path / "dump.bin")
    subprocess.check_call([hostcheck, "-dir", data, "-n", str(n), "-m", str(m), "-k", str(k), "-seed", "777",
                           "-out", out] + SWITCHES[mode])
    d = O.read_dump(out)
    # integer work: bit exact
    for key in ("csr.row_ptr", "csr.col_idx", "seq2user", "seq2movie", "validation.u", "validation.i", "validation.y",
                "test.u", "test.i", "test.y"):
        np.testing.assert_array_equal(d[key], g[key], err_msg=key)
    if "binary" not in mode:
        np.testing.assert_array_equal(d["csr.y"], g["csr.y"])
    else:
        assert (d["csr.y"] == 1).all()
    # start state: same mt19937 stream, fp64; psi implementations differ by rounding only
    want = util.golden_state(g, 0)
    got = O.state_from_dump(dict(d, meta=g["T0/meta"]))
    for gname in util.groups(want):
        for f in ("shape", "rate", "Ev"):
            np.testing.assert_allclose(got.p[gname][f], want.p[gname][f], rtol=1e-14, atol=0, err_msg="%s.%s" % (gname, f))
        np.testing.assert_allclose(got.p[gname]["Elogv"], want.p[gname]["Elogv"], rtol=0, atol=2e-13, err_msg=gname)


def test_reader_edge_cases(hostcheck, tmp_path):
    """Capacity limits, duplicate lines, zero ratings and uint8 wrap (src/ratings.cc:63-119,
    src/env.hh:20), checked against hand-derived expectations of the reference's rules."""
    data = str(tmp_path / "d")
    os.makedirs(data)
    lines = ["10\t100\t5", "10\t101\t0",      # rating 0: dropped, item 101 gets no seq here
             "10\t102\t3", "11\t100\t4",
             "10\t102\t1",                     # duplicate (10,102): walked twice, carries the LAST value
             "12\t103\t2",                     # third user: over -n 2 -> dropped
             "11\t104\t2",                     # third item: over -m 2 -> dropped
             "11\t102\t260"]                   # 260 wraps to 4 in a uint8
    open(os.path.join(data, "train.tsv"), "w").write("\n".join(lines) + "\n")
    open(os.path.join(data, "validation.tsv"), "w").write("10\t100\t2\n12\t100\t1\n11\t102\t0\n")
    open(os.path.join(data, "test.tsv"), "w").write("11\t100\t5\n11\t100\t1\n")   # same pair twice: last wins
    out = str(tmp_path / "o.bin")
    subprocess.check_call([hostcheck, "-dir", data, "-n", "2", "-m", "2", "-k", "3", "-hier", "-out", out])
    d = O.read_dump(out)
    np.testing.assert_array_equal(d["seq2user"], [10, 11])
    np.testing.assert_array_equal(d["seq2movie"], [100, 102])
    np.testing.assert_array_equal(d["csr.row_ptr"], [0, 3, 5])
    np.testing.assert_array_equal(d["csr.col_idx"], [0, 1, 1, 0, 1])
    np.testing.assert_array_equal(d["csr.y"], [5, 1, 1, 4, 4])
    np.testing.assert_array_equal(d["validation.u"], [0])
    np.testing.assert_array_equal(d["test.y"], [1])


@pytest.mark.parametrize("bad,where", [("user\titem\trating\n1\t2\t3\n", "train"),       # header line
                                       ("1\t2\t3\n4\t5\t2.5\n6\t7\t1\n", "train"),          # float rating
                                       ("1\t2\t3.5\n4\t5\t2.5\n6\t7\t1.5\n", "train"),      # ... on every line: 12 tokens, 4 "triples"
                                       ("1\t2\t3\n4x\t5\t2\n", "train"),                      # digits followed by a letter
                                       ("1\t2\t3\n4\t5\n", "train"),                          # line cut short by the end
                                       ("1\t2\t3\n4\tx\t1\n", "validation")])
def test_reader_stops_on_malformed_input_instead_of_truncating(hostcheck, tmp_path, bad, where):
    """A token that is not an unsigned decimal is not the end of the input: the reader names the file and the byte and
    the run stops -- it must not train on (or write a -csr-cache of) the lines before it."""
    data = str(tmp_path / "d")
    os.makedirs(data)
    good = "1\t2\t3\n4\t5\t1\n"
    for name in ("train", "validation", "test"):
        open(os.path.join(data, name + ".tsv"), "w").write(bad if name == where else good)
    p = subprocess.run([hostcheck, "-dir", data, "-n", "4", "-m", "8", "-k", "3", "-hier", "-csr-cache", "-out", str(tmp_path / "o.bin")],
                       capture_output=True, text=True)
    assert p.returncode != 0
    assert "malformed input" in p.stderr and where + ".tsv" in p.stderr and "byte" in p.stderr, p.stderr
    if where == "train":
        assert not os.path.exists(os.path.join(data, "train.tsv.hpfcsr"))
    assert not os.path.exists(str(tmp_path / "o.bin"))


def _run_check(hostcheck, data, out, *extra):
    p = subprocess.run([hostcheck, "-dir", data, "-out", out] + list(extra), check=True, capture_output=True, text=True)
    return p.stdout


@pytest.mark.parametrize("mode", ("hier", "hier_binary"))
def test_csr_cache_reload_is_bit_identical_and_invalidates(hostcheck, tmp_path, mode):
    """-csr-cache: the second run loads <dir>/train.tsv.hpfcsr instead of parsing the text and
    hands over exactly the same CSR, id maps, held-out maps, item marginals and start state; a cache made for another
    train.tsv, other -n/-m caps or another rating class is not used; a damaged cache is ignored."""
    g = util.load_golden(mode)
    n, m, k = (int(v) for v in g["T0/meta"][:3])
    data = str(tmp_path / "data")
    util.write_dataset(g, data, lift=3 if "binary" in mode else 0)
    # a repeated (user, item) line: the cache has to carry the fixed-up values
    lines = open(os.path.join(data, "train.tsv")).read().splitlines()
    a, b, _ = lines[0].split("\t")
    lines.insert(5, "%s\t%s\t4" % (a, b))
    open(os.path.join(data, "train.tsv"), "w").write("\n".join(lines) + "\n")
    base = ["-n", str(n), "-m", str(m), "-k", str(k), "-seed", "777"] + SWITCHES[mode]
    cache = os.path.join(data, "train.tsv.hpfcsr")
    plain, first, second = (str(tmp_path / f) for f in ("plain.bin", "first.bin", "second.bin"))

    _run_check(hostcheck, data, plain, *base)
    assert not os.path.exists(cache)                      # nothing is written without the flag
    assert "csr cache: written" in _run_check(hostcheck, data, first, *(base + ["-csr-cache"]))
    assert "csr cache: loaded" in _run_check(hostcheck, data, second, *(base + ["-csr-cache"]))
    assert open(plain, "rb").read() == open(first, "rb").read() == open(second, "rb").read()
    assert [f for f in os.listdir(data) if ".tmp." in f] == []

    other = str(tmp_path / "other.bin")
    # other caps / other rating class: parsed again (and the cache replaced)
    assert "csr cache: written" in _run_check(hostcheck, data, other, "-n", str(n - 1), "-m", str(m), "-k", str(k), "-csr-cache")
    assert "csr cache: written" in _run_check(hostcheck, data, other, *(base + ["-csr-cache"]))
    if "binary" in mode:
        assert "csr cache: written" in _run_check(hostcheck, data, other, "-n", str(n), "-m", str(m), "-k", str(k), "-hier",
                                                  "-binary-data", "-rating-threshold", "4", "-csr-cache")
        assert "csr cache: written" in _run_check(hostcheck, data, other, *(base + ["-csr-cache"]))
    # damaged payload (checksum), truncated file, garbage: ignored, parsed again, result unchanged
    blob = bytearray(open(cache, "rb").read())
    for damage in ("flip", "truncate", "garbage"):
        bad = bytearray(blob)
        if damage == "flip":
            bad[len(bad) // 2] ^= 0x40
        elif damage == "truncate":
            bad = bad[:-7]
        else:
            bad = bytearray(b"not a cache")
        open(cache, "wb").write(bytes(bad))
        st = os.stat(os.path.join(data, "train.tsv"))
        assert "csr cache: written" in _run_check(hostcheck, data, other, *(base + ["-csr-cache"])), damage
        assert open(other, "rb").read() == open(plain, "rb").read(), damage
        assert os.stat(os.path.join(data, "train.tsv")).st_mtime_ns == st.st_mtime_ns
    # train.tsv changed (one more line): the old cache is not used, the new result differs
    open(os.path.join(data, "train.tsv"), "a").write(lines[1] + "\n")
    assert "csr cache: written" in _run_check(hostcheck, data, other, *(base + ["-csr-cache"]))
    fresh = str(tmp_path / "fresh.bin")
    _run_check(hostcheck, data, fresh, *base)
    assert open(other, "rb").read() == open(fresh, "rb").read() != open(plain, "rb").read()


def test_csr_cache_in_a_read_only_directory_is_not_an_error(hostcheck, tmp_path):
    g = util.load_golden("bpf")
    n, m, k = (int(v) for v in g["T0/meta"][:3])
    data = str(tmp_path / "data")
    util.write_dataset(g, data)
    os.chmod(data, 0o555)
    try:
        out = _run_check(hostcheck, data, str(tmp_path / "o.bin"), "-n", str(n), "-m", str(m), "-k", str(k), "-csr-cache")
        if os.geteuid() != 0:  # root writes anywhere
            assert "csr cache: not written" in out
    finally:
        os.chmod(data, 0o755)


@pytest.mark.skipif(not (os.path.isdir("/root/reference/src") and os.path.exists(O.REF_HARNESS)),
                    reason="needs /root/reference and the oracle/_ref build")
def test_reader_matches_the_live_reference_on_random_files(hostcheck, tmp_path):
    """Differential test against the unmodified reference reader (Ratings::read_generic, src/ratings.cc:63-119, behind
    oracle/_ref/ref_harness): random files with repeated lines, zero ratings, values past uint8, -n / -m below and above
    what the file holds (the reference shrinks both to what it read, so held-out lines never add users or items),
    held-out lines naming unseen users / items, -binary-data with random thresholds."""
    for seed in range(30):
        rng = np.random.default_rng(seed)
        nu, ni = int(rng.integers(4, 25)), int(rng.integers(4, 25))
        uids = rng.choice(np.arange(100, 400), nu, replace=False)
        iids = rng.choice(np.arange(1000, 1400), ni, replace=False)

        def lines(cnt, strangers):
            out = []
            for _ in range(cnt):
                u, i = int(rng.choice(uids)), int(rng.choice(iids))
                if strangers and rng.random() < 0.2:
                    i = int(rng.integers(2000, 2010))
                if strangers and rng.random() < 0.1:
                    u = int(rng.integers(500, 505))
                out.append("%d\t%d\t%d" % (u, i, int(rng.choice([0, 1, 2, 3, 4, 5, 5, 255, 256, 260, 300]))))
            return out

        train = lines(int(rng.integers(30, 200)), False)
        if rng.random() < 0.5:
            train.sort(key=lambda s: int(s.split("\t")[0]))  # grouped by user, as real files are
        binary, thr = bool(rng.random() < 0.4), int(rng.integers(1, 5))
        seen_u, seen_i = len({l.split("\t")[0] for l in train}), len({l.split("\t")[1] for l in train})
        n_cap = max(1, seen_u - int(rng.integers(0, 3)))
        m_cap = max(1, seen_i - int(rng.integers(0, 3))) + (3 if seed % 2 else 0)
        run = tmp_path / ("c%d" % seed)
        data = str(run / "data")
        os.makedirs(data)
        open(os.path.join(data, "train.tsv"), "w").write("\n".join(train) + "\n")
        open(os.path.join(data, "validation.tsv"), "w").write("\n".join(lines(20, True)) + "\n")
        open(os.path.join(data, "test.tsv"), "w").write("\n".join(lines(40, True)) + "\n")
        O.run_ref_harness(data, n_cap, m_cap, 3, [0], str(run / "r"), str(run), hier=True, binary=binary,
                          rating_threshold=thr, seed=3)
        ref = O.read_dump(str(run / "r_0.bin"))
        args = [hostcheck, "-dir", data, "-n", str(n_cap), "-m", str(m_cap), "-k", "3", "-hier", "-seed", "3",
                "-rating-threshold", str(thr), "-out", str(run / "h.bin")] + (["-binary-data"] if binary else [])
        subprocess.check_call(args)
        got = O.read_dump(str(run / "h.bin"))
        for key in ("csr.row_ptr", "csr.col_idx", "seq2user", "seq2movie", "validation.u", "validation.i", "validation.y",
                    "test.u", "test.i", "test.y"):
            np.testing.assert_array_equal(got[key], ref[key], err_msg="seed %d %s" % (seed, key))
        if not binary:  # a rating that wrapped to 0 in the uint8 is walked as 1 (the loop only scales y > 1)
            np.testing.assert_array_equal(got["csr.y"], np.where(ref["csr.y"] == 0, 1, ref["csr.y"]), err_msg="seed %d y" % seed)
        # same model dimensions and the same start state (same number of generator draws)
        np.testing.assert_allclose(got["htheta.shape"], ref["htheta.shape"], rtol=1e-14, err_msg="seed %d" % seed)
        np.testing.assert_allclose(got["hbeta.Ev"], ref["hbeta.Ev"], rtol=1e-13, err_msg="seed %d" % seed)
