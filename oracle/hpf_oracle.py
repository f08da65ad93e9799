"""TEST INFRASTRUCTURE ONLY -- ctypes front-end for oracle/libhpf_oracle.so (the
plain-C fp64 restatement of the reference CAVI path, oracle/hpf_oracle.c) and a
reader for the state dumps written by oracle/_ref/ref_harness.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference leg import this module.  hgaprec_b200/ never does.
"""
import ctypes
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhpf_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_HARNESS = os.path.join(REF_DIR, "ref_harness")
REF_BINARY = os.path.join(REF_DIR, "hgaprec_ref")

HIER, BIAS, BINARY, JACOBI = 1, 2, 4, 8
SHAPE, RATE, EV, ELOGV = 0, 1, 2, 3
FIELDS = ("shape", "rate", "Ev", "Elogv")

_P4 = ctypes.POINTER(ctypes.c_double) * 4


class _CState(ctypes.Structure):
    _fields_ = [("n", ctypes.c_uint32), ("m", ctypes.c_uint32), ("k", ctypes.c_uint32),
                ("flags", ctypes.c_uint32),
                ("theta", _P4), ("beta", _P4), ("thetarate", _P4), ("betarate", _P4),
                ("thetabias", _P4), ("betabias", _P4)]


_lib = None


def build(force=False):
    """Compile the C restatement (and, when /root/reference is present, the
    reference binaries under oracle/_ref) with oracle/Makefile."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "hpf_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src") and not (os.path.exists(REF_HARNESS) and os.path.exists(REF_BINARY)):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.hpf_oracle_digamma.restype = ctypes.c_double
        L.hpf_oracle_digamma.argtypes = [ctypes.c_double]
        L.hpf_oracle_init.argtypes = [ctypes.POINTER(_CState), ctypes.c_ulong]
        L.hpf_oracle_iterate.argtypes = [ctypes.POINTER(_CState), ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int]
        L.hpf_oracle_heldout.restype = ctypes.c_double
        L.hpf_oracle_heldout.argtypes = [ctypes.POINTER(_CState), ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_uint64]
        L.hpf_oracle_elbo.restype = ctypes.c_double
        L.hpf_oracle_elbo.argtypes = [ctypes.POINTER(_CState)] + [ctypes.c_void_p] * 7
        L.hpf_oracle_topn.argtypes = [ctypes.POINTER(_CState), ctypes.c_void_p, ctypes.c_uint32,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                      ctypes.c_void_p, ctypes.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleState:
    """fp64 variational state laid out like the reference's GPMatrix / GPMatrixGR /
    GPArray members (src/gpbase.hh): dict name -> {shape, rate, Ev, Elogv}."""

    def __init__(self, n, m, k, flags):
        self.n, self.m, self.k, self.flags = int(n), int(m), int(k), int(flags)
        hier = bool(flags & HIER)
        z = lambda *s: np.zeros(s, dtype=np.float64)
        self.p = {
            "theta": {"shape": z(n, k), "rate": z(n, k) if hier else z(k), "Ev": z(n, k), "Elogv": z(n, k)},
            "beta": {"shape": z(m, k), "rate": z(m, k) if hier else z(k), "Ev": z(m, k), "Elogv": z(m, k)},
            "thetarate": {f: z(n) for f in FIELDS},
            "betarate": {f: z(m) for f in FIELDS},
            "thetabias": {f: z(n) for f in FIELDS},
            "betabias": {f: z(m) for f in FIELDS},
        }

    @property
    def hier(self):
        return bool(self.flags & HIER)

    @property
    def bias(self):
        return bool(self.flags & BIAS)

    def copy(self):
        o = OracleState(self.n, self.m, self.k, self.flags)
        for g in self.p:
            for f in FIELDS:
                o.p[g][f][...] = self.p[g][f]
        return o

    def _c(self):
        c = _CState(self.n, self.m, self.k, self.flags)
        for g in self.p:
            arr = getattr(c, g)
            for j, f in enumerate(FIELDS):
                a = self.p[g][f]
                assert a.flags["C_CONTIGUOUS"] and a.dtype == np.float64
                arr[j] = a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        return c

    # ---- reference operations -------------------------------------------
    def init(self, seed):
        c = self._c()
        lib().hpf_oracle_init(ctypes.byref(c), int(seed))
        return self

    def iterate(self, row_ptr, col_idx, y, niters=1, nthreads=1):
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
        yy = None if y is None else np.ascontiguousarray(y, dtype=np.uint8)
        c = self._c()
        # the last iteration runs on its own so that the rate priors it saw (set_prior_rate, gpbase.hh:163-173:
        # E[xi], E[log xi], E[eta], E[log eta] BEFORE that iteration's xi / eta update) can be kept for elbo()
        for part in (int(niters) - 1, 1) if niters >= 1 else ():
            if part == 1:
                self.rate_prior = tuple(self.p[g][f].copy() for g in ("thetarate", "betarate") for f in ("Ev", "Elogv"))
            if part > 0:
                lib().hpf_oracle_iterate(ctypes.byref(c), _ptr(row_ptr), _ptr(col_idx),
                                         None if yy is None else _ptr(yy), part, int(nthreads))
        return self

    def elbo(self, row_ptr, col_idx, y, rate_prior=None):
        """HGAPRec::logl (hgaprec.cc:2160-2255).  rate_prior = (E[xi], E[log xi], E[eta], E[log eta]) as the last
        iteration's set_prior_rate saw them; default: what the last iterate() call kept (hier only)."""
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
        yy = None if y is None else np.ascontiguousarray(y, dtype=np.uint8)
        pri = [None] * 4
        if self.hier:
            rp = rate_prior if rate_prior is not None else getattr(self, "rate_prior", None)
            assert rp is not None, "hier ELBO needs the rate priors of the last iteration (iterate() first)"
            pri = [np.ascontiguousarray(a, dtype=np.float64) for a in rp]
        c = self._c()
        return lib().hpf_oracle_elbo(ctypes.byref(c), _ptr(row_ptr), _ptr(col_idx), None if yy is None else _ptr(yy),
                                     *[None if a is None else _ptr(a) for a in pri])

    def heldout(self, u, i, y):
        u = np.ascontiguousarray(u, dtype=np.uint32)
        i = np.ascontiguousarray(i, dtype=np.uint32)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        c = self._c()
        return lib().hpf_oracle_heldout(ctypes.byref(c), _ptr(u), _ptr(i), _ptr(y), len(u))

    def topn(self, users, excl_ptr, excl_idx, topn):
        users = np.ascontiguousarray(users, dtype=np.uint32)
        excl_ptr = np.ascontiguousarray(excl_ptr, dtype=np.uint64)
        excl_idx = np.ascontiguousarray(excl_idx, dtype=np.uint32)
        items = np.zeros((len(users), topn), dtype=np.uint32)
        scores = np.zeros((len(users), topn), dtype=np.float64)
        c = self._c()
        lib().hpf_oracle_topn(ctypes.byref(c), _ptr(users), len(users), _ptr(excl_ptr), _ptr(excl_idx),
                              int(topn), _ptr(items), _ptr(scores))
        return items, scores


def digamma(x):
    return lib().hpf_oracle_digamma(float(x))


# ---------------------------------------------------------------- ref dumps
_DT = {0: np.float64, 1: np.uint32, 2: np.uint8, 3: np.uint64}


def read_dump(path):
    """Read a HPFDUMP1 file written by oracle/ref_harness.cc -> dict of arrays."""
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"HPFDUMP1", path
        while True:
            h = f.read(4)
            if not h:
                break
            (nl,) = struct.unpack("<I", h)
            name = f.read(nl).decode()
            dtype, ndim = struct.unpack("<II", f.read(8))
            dims = struct.unpack("<%dQ" % ndim, f.read(8 * ndim))
            cnt = int(np.prod(dims)) if ndim else 1
            dt = np.dtype(_DT[dtype])
            out[name] = np.frombuffer(f.read(cnt * dt.itemsize), dtype=dt).reshape(dims).copy()
    return out


def state_from_dump(d):
    """OracleState carrying the reference's own numbers from a dump."""
    n, m, k, _t, hier, bias, binary, vb = (int(v) for v in d["meta"])
    flags = (HIER if hier else 0) | (BIAS if bias else 0) | (BINARY if binary else 0) | (0 if vb else JACOBI)
    s = OracleState(n, m, k, flags)
    tn, bn = ("htheta", "hbeta") if hier else ("theta", "beta")
    for f in FIELDS:
        s.p["theta"][f][...] = d["%s.%s" % (tn, f)]
        s.p["beta"][f][...] = d["%s.%s" % (bn, f)]
    if hier:
        for g in ("thetarate", "betarate"):
            for f in FIELDS:
                s.p[g][f][...] = d["%s.%s" % (g, f)]
    if bias:
        for g in ("thetabias", "betabias"):
            for f in FIELDS:
                s.p[g][f][...] = d["%s.%s" % (g, f)].reshape(-1)
    return s


def run_ref_harness(data_dir, n, m, k, iters, dump_prefix, cwd, hier=False, bias=False, binary=False,
                    novb=False, rating_threshold=1, seed=0):
    """Run the unmodified reference behind oracle/_ref/ref_harness; returns the
    dump paths.  Needs the _ref build (only available where /root/reference is)."""
    args = [REF_HARNESS, "-dir", data_dir, "-n", str(n), "-m", str(m), "-k", str(k),
            "-rating-threshold", str(rating_threshold), "-seed", str(seed),
            "-iters", ",".join(str(t) for t in iters), "-dump", dump_prefix]
    if hier:
        args.append("-hier")
    if bias:
        args.append("-bias")
    if binary:
        args.append("-binary-data")
    if novb:
        args.append("-novb")
    subprocess.check_call(args, cwd=cwd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return ["%s_%d.bin" % (dump_prefix, t) for t in iters]
