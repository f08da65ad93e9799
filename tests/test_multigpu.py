"""N > 1: users are sharded over ranks (contiguous ranges balanced by nonzeros),
the item side is replicated and one all-reduce per iteration sums the item-side
block.  CPU part: the partition arithmetic and the world_size-2
plumbing over gloo.  GPU part (needs two devices; run with `gpurun --gpus 2`): two
engines, one per GPU, joined by NCCL, against the single-GPU engine and the oracle."""
import os
import socket
import threading

import numpy as np
import pytest

import util
import hgaprec_b200 as H
from hgaprec_b200 import synth
from oracle import hpf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ CPU
def test_partition_is_contiguous_and_balanced_by_nonzeros():
    rng = np.random.default_rng(0)
    deg = np.concatenate([rng.integers(0, 5, 1000), [5000], rng.integers(0, 50, 3000), np.zeros(40, np.int64)])
    rp = np.zeros(len(deg) + 1, np.uint64)
    rp[1:] = np.cumsum(deg)
    for nr in (1, 2, 3, 4, 8):
        b = H.partition_users(rp, nr)
        assert b[0] == 0 and b[-1] == len(deg) and (np.diff(b.astype(np.int64)) >= 0).all()
        per = np.diff(rp[b].astype(np.int64))
        # no shard exceeds its fair share by more than the heaviest single user
        assert per.max() <= rp[-1] / nr + deg.max()
        assert per.sum() == rp[-1]
    # degenerate inputs: more ranks than users, no ratings at all
    b = H.partition_users(np.array([0, 3, 4], np.uint64), 5)
    assert b[0] == 0 and b[-1] == 2 and (np.diff(b.astype(np.int64)) >= 0).all()
    b = H.partition_users(np.zeros(4, np.uint64), 2)
    assert b[0] == 0 and b[-1] == 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank cuts ITS shard out of the same seeded data set with the library's partition
        n, m, nnz = 4000, 300, 60000
        full = synth.make_ratings(n, m, nnz, seed=99, device="cpu")
        bounds = H.partition_users(full["row_ptr"], world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        frp = full["row_ptr"].astype(np.int64)
        mine = dict(row_ptr=(frp[lo:hi + 1] - frp[lo]).astype(np.uint64), col_idx=full["col_idx"][frp[lo]:frp[hi]],
                    y=full["y"][frp[lo]:frp[hi]])
        # what bench.py does: the engine id travels by broadcast_object_list, timings by all_reduce(MAX)
        obj = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        import torch
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cnt = torch.tensor([float(len(mine["col_idx"]))], dtype=torch.float64)
        dist.all_reduce(cnt)
        # the ranks' shards tile the global CSR: gather the pieces on every rank and compare
        parts = [None] * world
        dist.all_gather_object(parts, (lo, hi, mine["col_idx"], mine["y"]))
        parts.sort(key=lambda p: p[0])
        ok = (parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:])) and
              np.array_equal(np.concatenate([p[2] for p in parts]), full["col_idx"]) and
              np.array_equal(np.concatenate([p[3] for p in parts]), full["y"]) and int(mine["row_ptr"][-1]) == len(mine["col_idx"]))
        q.put((rank, ok, obj[0] == bytes(range(128)), float(t.item()), float(cnt.item()), len(full["col_idx"])))
    finally:
        dist.destroy_process_group()


def test_world_size_2_sharding_plumbing_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, shard_ok, id_ok, tmax, total, nnz in res:
        assert shard_ok and id_ok, res        # shards tile the global CSR exactly; the id arrived intact
        assert tmax == 2.0 and total == nnz   # MAX over ranks; every nonzero owned by exactly one rank


# ------------------------------------------------------------------ GPU (2 devices)
def _two_gpus():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.device_count() >= 2
    except Exception:
        return False


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [H.HIER, H.HIER | H.BIAS, H.BIAS, H.BIAS | H.JACOBI])
def test_two_gpu_shards_match_single_gpu_and_oracle(flags):
    if not _two_gpus():
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    n, m, nnz, k, iters = 5000, 1200, 200000, 100, 3
    d = synth.make_ratings(n, m, nnz, seed=23, heldout=0.05)
    s = O.OracleState(n, m, k, flags).init(24)
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], iters, nthreads=8)
    bounds = H.partition_users(d["row_ptr"], 2)
    hu, hi_, hy = d["heldout"]

    def one_gpu(niter):
        with H.Engine(n, m, k, flags=flags, device=0) as e:
            e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
            util.push_state(e, s)
            e.iterate(niter)
            return util.pull_state(e, s)

    # SURVEY.md 8e gate: G GPUs vs 1 GPU identical up to fp32 summation order, max rel 1e-5 after ONE iteration
    got1 = _stitch(_run_two_ranks(d, s, flags, 1, k), s, n, m, k, flags)
    bad = util.compare_states(got1, one_gpu(1), rel=1e-5, elog_abs=1e-5)
    assert not bad, bad
    # three iterations: against the oracle (single-iteration gate x 3), held-out ll, and the single GPU again
    out = _run_two_ranks(d, s, flags, iters, k, heldout=(hu, hi_, hy, bounds))
    got = _stitch(out, s, n, m, k, flags)
    bad = util.compare_states(got, want, rel=6e-5, elog_abs=6e-5)
    assert not bad, bad
    assert abs((out[0][2] + out[1][2]) - want.heldout(hu, hi_, hy)) / len(hu) <= 2e-4
    bad = util.compare_states(got, one_gpu(iters), rel=3e-5, elog_abs=3e-5)
    assert not bad, bad


def _run_two_ranks(d, s, flags, iters, k, heldout=None, windows=None):
    """Two engines (GPU 0 and 1, one thread each) joined by NCCL over the nnz-balanced user partition.
    Returns per rank (state dict, stats[, held-out ll sum of the rank's users]).  windows: the iterations as
    several hpf_iterate calls (default: one)."""
    n, m = d["n"], d["m"]
    bounds = H.partition_users(d["row_ptr"], 2)
    rp = d["row_ptr"].astype(np.int64)
    uid = H.comm_unique_id()
    out, errs = [None, None], []

    def worker(r):
        try:
            lo, hi = int(bounds[r]), int(bounds[r + 1])
            with H.Engine(hi - lo, m, k, flags=flags, device=r, n_users_global=n) as e:
                e.comm_init(r, 2, uid)
                e.set_ratings_csr(rp[lo:hi + 1] - rp[lo], d["col_idx"][rp[lo]:rp[hi]], None if d["y"] is None else d["y"][rp[lo]:rp[hi]])
                util.push_state(e, s, users=np.arange(lo, hi))
                for w in (windows or [iters]):
                    e.iterate(w)
                res = ({g: e.get_state(util._IDS[g]) for g in util.groups(s)}, e.stats())
                if heldout is not None:
                    hu, hi_, hy, _ = heldout
                    sel = (hu >= lo) & (hu < hi)
                    res = res + (e.heldout_loglik(hu[sel] - lo, hi_[sel], hy[sel]),)
                out[r] = res
        except Exception as ex:  # surfaced in the main thread
            errs.append(ex)

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=600)
    assert not errs, errs
    return out


def _stitch(out, s, n, m, k, flags):
    got = O.OracleState(n, m, k, flags)
    for g in util.groups(s):
        for f in O.FIELDS:
            if g.startswith("beta") or (g == "theta" and f == "rate" and not (flags & H.HIER)):
                np.testing.assert_array_equal(out[0][0][g][f], out[1][0][g][f])  # replicas stay bitwise identical
                got.p[g][f][...] = out[0][0][g][f].reshape(got.p[g][f].shape)
            else:
                got.p[g][f][...] = np.concatenate([out[0][0][g][f], out[1][0][g][f]]).reshape(got.p[g][f].shape)
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("flags,dense", [(H.HIER | H.BIAS, "0"), (0, "0"), (H.HIER, "1")])
def test_two_gpu_chunked_allreduce_is_bitwise_the_single_allreduce(monkeypatch, flags, dense):
    """The item pass runs in chunks of items and the all-reduce of a finished chunk's T_beta rows runs on the
    communication stream under the sweeps that follow (HPF_AR_CHUNKS; default from the payload size).  Same operands,
    two ranks: every sum is a + b whatever the chunking, so the state must not change by a bit -- a missing
    stream dependency would show up here."""
    if not _two_gpus():
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    monkeypatch.setenv("HPF_DENSE_HEAD", dense)
    n, m, nnz, k, iters = 4000, 1000, 150000, 64, 3
    d = synth.make_ratings(n, m, nnz, seed=41, heldout=0.05)
    s = O.OracleState(n, m, k, flags).init(42)
    runs = {}
    for chunks in ("1", "5"):
        monkeypatch.setenv("HPF_AR_CHUNKS", chunks)
        runs[chunks] = _run_two_ranks(d, s, flags, iters, k)
        assert runs[chunks][0][1]["item_chunks"] == int(chunks)
        assert (runs[chunks][0][1]["head_nnz"] > 0) == (dense == "1")
    for r in range(2):
        for g in util.groups(s):
            for f in O.FIELDS:
                np.testing.assert_array_equal(runs["1"][r][0][g][f], runs["5"][r][0][g][f], err_msg="rank %d %s.%s" % (r, g, f))
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], iters, nthreads=8)
    assert not util.compare_states(_stitch(runs["5"], s, n, m, k, flags), want, rel=6e-5, elog_abs=6e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [H.HIER, H.BIAS])
def test_two_gpu_exact_fallback_is_reduced_over_the_ranks(flags):
    """A nonzero whose product form leaves the fp32 range adds y*phi to rank-local fallback buffers; on the item side
    those have to be summed over the ranks like T_beta.  The engine runs optimistically (no extra traffic), the ranks
    agree on a 'fallback fired' flag that rides with the column sums, and hpf_iterate re-runs its window from a
    snapshot with the buffers inside the all-reduce (stats.mg_exact).  State as in the single-GPU test: user rows
    peaked in one factor, item rows in another, 200 nats apart -- EVERY nonzero takes the fallback."""
    if not _two_gpus():
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    n, m, nnz, k, iters = 900, 400, 30000, 12, 2
    d = synth.make_ratings(n, m, nnz, seed=5)
    s = O.OracleState(n, m, k, flags).init(6)
    s.p["theta"]["Elogv"][:, 1:] -= 200.0
    s.p["beta"]["Elogv"][:, :-1] -= 200.0
    if flags & H.BIAS:   # the two bias slots of phi must not rescue the normaliser
        s.p["thetabias"]["Elogv"] -= 300.0
        s.p["betabias"]["Elogv"] -= 300.0
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], iters)
    out = _run_two_ranks(d, s, flags, iters, k)
    assert out[0][1]["mg_exact"] == 1 and out[1][1]["mg_exact"] == 1
    assert out[0][1]["slow_path_nnz"] > 0
    bad = util.compare_states(_stitch(out, s, n, m, k, flags), want, rel=6e-5, elog_abs=6e-5)
    assert not bad, bad


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [H.HIER, H.BIAS])
def test_one_ctx_driving_two_gpus_is_bitwise_the_two_rank_run(flags):
    """hpf_config.n_devices = 2: ONE ctx shards its users over both GPUs (the form the `hgaprec -gpus N` command line
    uses; SURVEY.md 8b: one process).  Same partition, same kernels, same NCCL sums as two one-device ctxs joined with
    hpf_comm_init, so the state must be bit-identical; held-out ll, top-N and item ranks take GLOBAL user numbers."""
    if not _two_gpus():
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    n, m, nnz, k, iters = 4000, 1000, 150000, 64, 3
    d = synth.make_ratings(n, m, nnz, seed=41, heldout=0.05)
    s = O.OracleState(n, m, k, flags).init(42)
    two = _stitch(_run_two_ranks(d, s, flags, iters, k), s, n, m, k, flags)
    hu, hi, hy = d["heldout"]
    users = np.array([3, 3999, 1700, 0, 2500, 2100], dtype=np.uint32)          # both shards, arbitrary order
    rp = d["row_ptr"].astype(np.int64)
    ep = np.zeros(len(users) + 1, np.uint64)
    ep[1:] = np.cumsum([rp[u + 1] - rp[u] for u in users])
    ei = np.concatenate([d["col_idx"][rp[u]:rp[u + 1]] for u in users])
    qp = np.arange(0, 3 * len(users) + 1, 3, dtype=np.uint64)
    qi = np.tile(np.array([5, 77, 900], dtype=np.uint32), len(users))
    with H.Engine(n, m, k, flags=flags, devices=[0, 1]) as e:
        e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
        util.push_state(e, s)
        e.iterate(iters)
        got = util.pull_state(e, s)
        st = e.stats()
        ll = e.heldout_loglik(hu, hi, hy)
        items, scores = e.topn(users, ep, ei, 20)
        ranks, rscores = e.item_ranks(users, ep, ei, qp, qi)
    assert st["n_devices"] == 2 and st["nnz"] == len(d["col_idx"])
    for g in util.groups(s):
        for f in O.FIELDS:
            np.testing.assert_array_equal(got.p[g][f], two.p[g][f], err_msg="%s.%s" % (g, f))
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], iters, nthreads=8)
    assert abs(ll - want.heldout(hu, hi, hy)) / len(hu) <= 2e-4
    # ranking calls against a single-GPU ctx holding the same state
    with H.Engine(n, m, k, flags=flags, device=0) as e1:
        util.push_state(e1, got)
        items1, scores1 = e1.topn(users, ep, ei, 20)
        ranks1, rscores1 = e1.item_ranks(users, ep, ei, qp, qi)
    np.testing.assert_array_equal(items, items1)
    np.testing.assert_array_equal(scores, scores1)
    np.testing.assert_array_equal(ranks, ranks1)
    np.testing.assert_array_equal(rscores, rscores1)


@pytest.mark.gpu
@pytest.mark.parametrize("flags,dense", [(H.HIER, "0"), (0, "0"), (H.HIER, "1")])
def test_two_gpu_sharded_beta_update_matches_the_replicated_one(monkeypatch, flags, dense):
    """Many items (MSD): T_beta is reduce-scattered, each rank updates its slice of ceil(m / N) items, A_beta is
    all-gathered every iteration and the rest of beta's state when the hpf_iterate window ends (HPF_SHARD_BETA; auto
    from the payload).  Against the replicated update only the order of the sums behind sum_i E[beta] changes; m is odd
    so the slices are unequal and the last one runs into the spare rows; two windows, so a window starts from gathered
    state; every rank must hold the SAME beta afterwards (checked bitwise by _stitch)."""
    if not _two_gpus():
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    monkeypatch.setenv("HPF_DENSE_HEAD", dense)
    n, m, nnz, k, iters = 4000, 1001, 150000, 64, 3
    d = synth.make_ratings(n, m, nnz, seed=43, heldout=0.05)
    s = O.OracleState(n, m, k, flags).init(44)
    hu, hi_, hy = d["heldout"]
    bounds = H.partition_users(d["row_ptr"], 2)
    runs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("HPF_SHARD_BETA", mode)
        one = _stitch(_run_two_ranks(d, s, flags, 1, k), s, n, m, k, flags)
        out = _run_two_ranks(d, s, flags, iters, k, heldout=(hu, hi_, hy, bounds), windows=[2, 1])
        assert out[0][1]["beta_sharded"] == int(mode) and out[1][1]["beta_sharded"] == int(mode)
        assert out[0][1]["mg_exact"] == 0 and (out[0][1]["head_nnz"] > 0) == (dense == "1")
        runs[mode] = (one, _stitch(out, s, n, m, k, flags), out[0][2] + out[1][2])
    bad = util.compare_states(runs["1"][0], runs["0"][0], rel=1e-5, elog_abs=1e-5)
    assert not bad, bad
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], iters, nthreads=8)
    bad = util.compare_states(runs["1"][1], want, rel=6e-5, elog_abs=6e-5)
    assert not bad, bad
    assert abs(runs["1"][2] - want.heldout(hu, hi_, hy)) / len(hu) <= 2e-4


@pytest.mark.gpu
def test_two_gpu_sharded_beta_update_falls_back_to_the_exact_replicated_run(monkeypatch):
    """Inside a sharded window the exact fallback would read rows of beta's shape another rank owns, so ANY fallback
    raises the flag: the window re-runs from its snapshot unsharded, fallback buffers inside the all-reduce."""
    if not _two_gpus():
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    monkeypatch.setenv("HPF_SHARD_BETA", "1")
    flags = H.HIER
    n, m, nnz, k, iters = 900, 401, 30000, 12, 2
    d = synth.make_ratings(n, m, nnz, seed=5)
    s = O.OracleState(n, m, k, flags).init(6)
    s.p["theta"]["Elogv"][:, 1:] -= 200.0
    s.p["beta"]["Elogv"][:, :-1] -= 200.0
    want = s.copy().iterate(d["row_ptr"], d["col_idx"], d["y"], iters)
    out = _run_two_ranks(d, s, flags, iters, k)
    assert out[0][1]["mg_exact"] == 1 and out[1][1]["mg_exact"] == 1
    assert out[0][1]["beta_sharded"] == 0 and out[0][1]["slow_path_nnz"] > 0
    bad = util.compare_states(_stitch(out, s, n, m, k, flags), want, rel=6e-5, elog_abs=6e-5)
    assert not bad, bad


@pytest.mark.gpu
def test_one_ctx_driving_two_gpus_shards_the_beta_update_too(monkeypatch):
    if not _two_gpus():
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    monkeypatch.setenv("HPF_SHARD_BETA", "1")
    flags = H.HIER
    n, m, nnz, k, iters = 4000, 1001, 150000, 64, 3
    d = synth.make_ratings(n, m, nnz, seed=43)
    s = O.OracleState(n, m, k, flags).init(44)
    two = _stitch(_run_two_ranks(d, s, flags, iters, k), s, n, m, k, flags)
    with H.Engine(n, m, k, flags=flags, devices=[0, 1]) as e:
        e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
        util.push_state(e, s)
        e.iterate(iters)
        got = util.pull_state(e, s)
        assert e.stats()["beta_sharded"] == 1
    for g in util.groups(s):
        for f in O.FIELDS:
            np.testing.assert_array_equal(got.p[g][f], two.p[g][f], err_msg="%s.%s" % (g, f))


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [H.HIER, H.BIAS])
def test_two_gpus_against_one_after_20_iterations(flags):
    """SURVEY.md 8e, second half of the gate: after 20 iterations the G-GPU fit and the 1-GPU fit differ by fp32
    summation order only -- |delta mean held-out ll| <= 1e-5 (and the states stay within the 20-iteration band)."""
    if not _two_gpus():
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    n, m, nnz, k, iters = 5000, 1200, 200000, 100, 20
    d = synth.make_ratings(n, m, nnz, seed=23, heldout=0.05)
    s = O.OracleState(n, m, k, flags).init(24)
    hu, hi_, hy = d["heldout"]
    res = {}
    for name, kw in (("one", dict(device=0)), ("two", dict(devices=[0, 1]))):
        with H.Engine(n, m, k, flags=flags, **kw) as e:
            e.set_ratings_csr(d["row_ptr"], d["col_idx"], d["y"])
            util.push_state(e, s)
            e.iterate(iters)
            res[name] = (util.pull_state(e, s), e.heldout_loglik(hu, hi_, hy) / len(hu), e.stats())
    assert res["two"][2]["n_devices"] == 2 and res["two"][2]["slow_path_nnz"] == 0
    assert abs(res["one"][1] - res["two"][1]) <= 1e-5, (res["one"][1], res["two"][1])
    bad = util.compare_states(res["two"][0], res["one"][0], rel=2e-3, elog_abs=2e-3)
    assert not bad, bad
