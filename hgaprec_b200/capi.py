"""ctypes binding of include/hpf_cuda.h (libhpf_b200.so).

Mirrors the reference's host-side use of the path: HGAPRec keeps GPMatrix /
GPMatrixGR / GPArray objects (src/gpbase.hh) and runs vb_hier()/vb()/vb_bias()
(src/hgaprec.cc:1321-1436, 919-980, 1219-1319); here the same parameter sets are
uploaded with set_state(), advanced with iterate() and read back with
get_state().  If the CUDA library is missing this module raises -- there is no
CPU path.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HPF_LIB") or os.path.join(HERE, "libhpf_b200.so")  # HPF_LIB: tuning variants
CSRC = os.path.join(HERE, "csrc")

ABI_VERSION = 2
MAX_DEVICES = 16
HIER, BIAS, BINARY, JACOBI, LOGL = 1, 2, 4, 8, 16
THETA, BETA, THETARATE, BETARATE, THETABIAS, BETABIAS = range(6)
COMM_ID_BYTES = 128


class HpfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("hpf error %d: %s" % (code, msg))
        self.code = code


class _Config(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_uint32), ("n_users", ctypes.c_uint32), ("n_items", ctypes.c_uint32),
                ("k", ctypes.c_uint32), ("flags", ctypes.c_uint32), ("device", ctypes.c_int32),
                ("n_users_global", ctypes.c_uint64),
                ("theta_shape", ctypes.c_double), ("theta_rate", ctypes.c_double),
                ("beta_shape", ctypes.c_double), ("beta_rate", ctypes.c_double),
                ("thetarate_shape", ctypes.c_double), ("thetarate_rate", ctypes.c_double),
                ("betarate_shape", ctypes.c_double), ("betarate_rate", ctypes.c_double),
                ("thetabias_shape", ctypes.c_double), ("thetabias_rate", ctypes.c_double),
                ("betabias_shape", ctypes.c_double), ("betabias_rate", ctypes.c_double),
                ("n_devices", ctypes.c_uint32), ("devices", ctypes.c_int32 * MAX_DEVICES)]


class Stats(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_uint64), ("iterations", ctypes.c_uint64),
                ("slow_path_nnz", ctypes.c_uint64), ("nnz", ctypes.c_uint64), ("device_bytes", ctypes.c_uint64),
                ("last_iterate_ms", ctypes.c_float), ("sweep_group", ctypes.c_uint32), ("sweep_vec", ctypes.c_uint32),
                ("user_l2_tiles", ctypes.c_uint32), ("item_l2_tiles", ctypes.c_uint32), ("head_nnz", ctypes.c_uint64),
                ("item_chunks", ctypes.c_uint32), ("mg_exact", ctypes.c_uint32), ("last_topn_ms", ctypes.c_float),
                ("n_devices", ctypes.c_uint32), ("beta_sharded", ctypes.c_uint32)]


class IterProfile(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in ("sweep_user_ms", "sweep_item_ms", "combine_ms", "update_theta_ms",
                                             "allreduce_ms", "update_beta_ms", "total_ms", "sweep_user_head_ms")]


_lib = None


def build_library():
    """nvcc-compile libhpf_b200.so for sm_100a (cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC], stdout=subprocess.DEVNULL)


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make -C %s` (or __graft_entry__.build()); "
                          "hgaprec_b200 has no CPU fallback" % (LIB_PATH, CSRC))
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, u64, cint = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
    L.hpf_config_default.argtypes = [ctypes.POINTER(_Config)]
    L.hpf_config_default.restype = None
    L.hpf_create.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(vp)]
    L.hpf_destroy.argtypes = [vp]
    L.hpf_destroy.restype = None
    L.hpf_last_error.argtypes = [vp]
    L.hpf_last_error.restype = ctypes.c_char_p
    L.hpf_set_ratings_csr.argtypes = [vp, vp, vp, vp]
    L.hpf_set_state.argtypes = [vp, cint, vp, vp, vp, vp]
    L.hpf_get_state.argtypes = [vp, cint, vp, vp, vp, vp]
    L.hpf_iterate.argtypes = [vp, u32]
    L.hpf_heldout_loglik.argtypes = [vp, vp, vp, vp, u64, ctypes.POINTER(ctypes.c_double)]
    L.hpf_elbo.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    L.hpf_topn.argtypes = [vp, vp, u32, vp, vp, u32, vp, vp]
    L.hpf_item_ranks.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp, vp]
    L.hpf_partition_users.argtypes = [vp, u32, u32, vp]
    L.hpf_comm_unique_id.argtypes = [vp, ctypes.c_size_t]
    L.hpf_comm_init.argtypes = [vp, cint, cint, vp, ctypes.c_size_t]
    L.hpf_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.hpf_iterate_profiled.argtypes = [vp, u32, ctypes.POINTER(IterProfile)]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def partition_users(row_ptr, nranks):
    """hpf_partition_users: nranks + 1 user bounds, contiguous ranges balanced by nonzeros."""
    L = load_library()
    rp = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    out = np.zeros(nranks + 1, dtype=np.uint32)
    rc = L.hpf_partition_users(_p(rp), len(rp) - 1, int(nranks), _p(out))
    if rc != 0:
        raise HpfError(rc, L.hpf_last_error(None).decode())
    return out


def comm_unique_id():
    """NCCL unique id (bytes) for Engine.comm_init; call on rank 0 and broadcast."""
    L = load_library()
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    rc = L.hpf_comm_unique_id(buf, COMM_ID_BYTES)
    if rc != 0:
        raise HpfError(rc, L.hpf_last_error(None).decode())
    return buf.raw


class Engine:
    """One hpf_ctx: the variational state of one user shard on one GPU."""

    def __init__(self, n_users, n_items, k, flags=HIER, device=0, n_users_global=0, devices=None, **priors):
        """devices: a list of CUDA ordinals -> ONE ctx that shards its users over those GPUs (hpf_config.n_devices)."""
        self._L = load_library()
        cfg = _Config()
        self._L.hpf_config_default(ctypes.byref(cfg))
        cfg.n_users, cfg.n_items, cfg.k, cfg.flags, cfg.device = int(n_users), int(n_items), int(k), int(flags), int(device)
        cfg.n_users_global = int(n_users_global)
        if devices is not None:
            cfg.n_devices = len(devices)
            for j, dv in enumerate(devices):
                cfg.devices[j] = int(dv)
        for name, val in priors.items():
            setattr(cfg, name, float(val))
        self.n, self.m, self.k, self.flags = int(n_users), int(n_items), int(k), int(flags)
        self._ctx = ctypes.c_void_p()
        rc = self._L.hpf_create(ctypes.byref(cfg), ctypes.byref(self._ctx))
        if rc != 0:
            self._ctx = None
            raise HpfError(rc, self._L.hpf_last_error(None).decode())

    # -- plumbing
    def _check(self, rc):
        if rc != 0:
            raise HpfError(rc, self._L.hpf_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.hpf_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def hier(self):
        return bool(self.flags & HIER)

    @property
    def bias(self):
        return bool(self.flags & BIAS)

    def _rows(self, which):
        return self.n if which in (THETA, THETARATE, THETABIAS) else self.m

    def _shapes(self, which):
        rows = self._rows(which)
        if which in (THETA, BETA):
            full = (rows, self.k)
            return full, (full if self.hier else (self.k,))
        return (rows,), (rows,)

    # -- the ABI
    def set_ratings_csr(self, row_ptr, col_idx, y=None):
        rp = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        ci = np.ascontiguousarray(col_idx, dtype=np.uint32)
        yy = None if y is None else np.ascontiguousarray(y, dtype=np.uint8)
        assert rp.shape == (self.n + 1,)
        self._check(self._L.hpf_set_ratings_csr(self._ctx, _p(rp), _p(ci), _p(yy)))

    def set_state(self, which, shape, rate, Ev, Elogv=None):
        full, rshape = self._shapes(which)
        s, r, e, l = _f64(shape), _f64(rate), _f64(Ev), _f64(Elogv)
        assert s.reshape(full).shape == full and r.reshape(rshape).shape == rshape
        self._check(self._L.hpf_set_state(self._ctx, which, _p(s), _p(r), _p(e), _p(l)))

    def get_state(self, which, fields=("shape", "rate", "Ev", "Elogv")):
        full, rshape = self._shapes(which)
        out = {f: np.empty(rshape if f == "rate" else full, dtype=np.float64) for f in fields}
        self._check(self._L.hpf_get_state(self._ctx, which, _p(out.get("shape")), _p(out.get("rate")),
                                          _p(out.get("Ev")), _p(out.get("Elogv"))))
        return out

    def iterate(self, n_iters=1):
        self._check(self._L.hpf_iterate(self._ctx, int(n_iters)))

    def iterate_profiled(self, n_iters=1):
        pr = IterProfile()
        self._check(self._L.hpf_iterate_profiled(self._ctx, int(n_iters), ctypes.byref(pr)))
        return {f[0]: getattr(pr, f[0]) for f in IterProfile._fields_}

    def heldout_loglik(self, u, i, y):
        u = np.ascontiguousarray(u, dtype=np.uint32)
        i = np.ascontiguousarray(i, dtype=np.uint32)
        y = np.ascontiguousarray(y, dtype=np.uint8)
        out = ctypes.c_double(0.0)
        self._check(self._L.hpf_heldout_loglik(self._ctx, _p(u), _p(i), _p(y), len(u), ctypes.byref(out)))
        return out.value

    def elbo(self):
        """HGAPRec::logl() (hgaprec.cc:2160-2255); the engine must have been created with the LOGL flag."""
        out = ctypes.c_double(0.0)
        self._check(self._L.hpf_elbo(self._ctx, ctypes.byref(out)))
        return out.value

    def topn(self, users, excl_ptr, excl_idx, topn):
        users = np.ascontiguousarray(users, dtype=np.uint32)
        ep = np.ascontiguousarray(excl_ptr, dtype=np.uint64)
        ei = np.ascontiguousarray(excl_idx, dtype=np.uint32)
        items = np.empty((len(users), topn), dtype=np.uint32)
        scores = np.empty((len(users), topn), dtype=np.float32)
        self._check(self._L.hpf_topn(self._ctx, _p(users), len(users), _p(ep), _p(ei), int(topn), _p(items), _p(scores)))
        return items, scores

    def item_ranks(self, users, excl_ptr, excl_idx, query_ptr, query_idx):
        users = np.ascontiguousarray(users, dtype=np.uint32)
        ep = np.ascontiguousarray(excl_ptr, dtype=np.uint64)
        ei = np.ascontiguousarray(excl_idx, dtype=np.uint32)
        qp = np.ascontiguousarray(query_ptr, dtype=np.uint64)
        qi = np.ascontiguousarray(query_idx, dtype=np.uint32)
        ranks = np.zeros(len(qi), dtype=np.uint32)
        scores = np.zeros(len(qi), dtype=np.float32)
        self._check(self._L.hpf_item_ranks(self._ctx, _p(users), len(users), _p(ep), _p(ei), _p(qp), _p(qi), _p(ranks), _p(scores)))
        return ranks, scores

    def comm_init(self, rank, nranks, unique_id):
        buf = ctypes.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._check(self._L.hpf_comm_init(self._ctx, int(rank), int(nranks), buf, COMM_ID_BYTES))

    def stats(self):
        st = Stats()
        self._L.hpf_get_stats(self._ctx, ctypes.byref(st))
        return {f[0]: getattr(st, f[0]) for f in Stats._fields_}
