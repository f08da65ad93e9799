#!/usr/bin/env python
"""bench.py -- HPF CAVI throughput (training nonzeros / second) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload netflix|msd|bpf-1b] [--scaling strong|weak]

A "step" is one full CAVI iteration (hpf_iterate(1): user sweep, item sweep,
theta / beta / xi / eta updates).  The workload is ONE fixed problem -- by default
BASELINE.json configs[1]: synthetic Netflix-scale 480,189 x 17,770, 1e8 nnz,
K=100, -hier -- cut into 8 fixed user blocks.  With N ranks (torchrun, one per
GPU) rank r holds blocks [8r/N, 8(r+1)/N): the same data at every N, users
sharded, items replicated, the item-side sums all-reduced over NCCL once per
iteration (chunk by chunk, under the sweeps): STRONG scaling.  --scaling weak is
round 1's measurement (every rank a Netflix-scale shard of an N-times larger
problem); at N > 1 the default run appends it as the secondary key "weak".
--workload msd / bpf-1b are BASELINE configs[2] / [3] (MSD-scale K=200 -hier
-binary-data; BPF -bias 1e9 nnz), same sharding.  One JSON line from rank 0.

At N=1 the default run also reports, as secondary keys, the HBM-bound MSD-scale
workload on the same GPU ("hbm_bound_workload": its roofline is the HBM copy peak)
and a steady-state leg ("steady_state": planted-factor data, timed after 100
iterations, with the number of nonzeros that took the exact fallback).

--impl reference times the reference's own single-threaded CPU loop
(oracle/_ref/hgaprec_ref, the unmodified sources) on a bounded user sample of
the same workload; it is the baseline, not the target.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hpf_cavi_nonzeros_per_sec"
UNIT = "nnz/s"
NBLOCKS = 8
EXTRA_WARMUP = 10  # untimed iterations on top of --warmup before the timed region (reported in config)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def l2_gather_peak():
    """L2 -> SM gather ceiling of this GPU model: random 512-byte row gathers from an L2-resident table, measured by
    tools/gather_bench.cu on the pool's B200 (profiles/l2_gather_peak.json); else the microarchitecture notes' figure."""
    p = os.path.join(ROOT, "profiles", "l2_gather_peak.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["peak_gbs"]), "measured (profiles/l2_gather_peak.json: %s)" % d.get("what", "gather micro-benchmark")
        except Exception:
            pass
    return 12400.0, "fallback (B300_MICROARCH.md: ~6,300 B/clk LTS->SM at 1.965 GHz)"


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        """One nvidia-smi process for all listed GPUs (rank 0 only: N pollers perturb the run)."""
        self.gpu, self.proc, self.path = ",".join(str(g) for g in gpu_indices), None, None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", "-i", self.gpu, "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits", "-lms", "100"],
                                     stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------- reference (CPU) arm
def ref_switches(cfg):
    sw = []
    if cfg.get("hier", True):
        sw.append("-hier")
    if cfg.get("bias"):
        sw.append("-bias")
    if cfg["binary"]:
        sw += ["-binary-data"]
    return sw


def time_reference(sample, cfg, iters, warm, replicas=0):
    """Per-iteration wall time of the unmodified reference binary on `sample`.

    -hier honours -max-iterations (T runs T+1 iterations, src/hgaprec.cc:1337-1339): two runs are differenced,
    (wall(warm+iters) - wall(warm)) / iters.  vb()/vb_bias() have no cap (src/hgaprec.cc:919-980, 1219-1319): the
    harness build (oracle/_ref/ref_harness, same objects, exits after a given iteration) is differenced instead.

    replicas > 1 adds a fourth return value: the per-iteration wall time when that many INDEPENDENT copies of the
    same job run side by side (the reference has no threads, so this is not something it can do for one job: it is
    the upper bound "every host core busy"), measured the same way over min(iters, 2) iterations."""
    from oracle import hpf_oracle as O
    if not os.path.exists(O.REF_BINARY):
        O.build()
    use_ref = os.path.exists(O.REF_BINARY) and os.path.exists(O.REF_HARNESS)
    n, m, k = sample["n"], sample["m"], cfg["k"]
    nnz = len(sample["col_idx"])
    hier = cfg.get("hier", True)
    if not use_ref:
        # the port: same loop, plain C, one thread
        flags = (O.HIER if hier else 0) | (O.BIAS if cfg.get("bias") else 0) | (O.BINARY if cfg["binary"] else 0)
        s = O.OracleState(n, m, k, flags).init(1)
        s.iterate(sample["row_ptr"], sample["col_idx"], sample["y"], max(1, warm), nthreads=1)
        t0 = time.time()
        s.iterate(sample["row_ptr"], sample["col_idx"], sample["y"], iters, nthreads=1)
        return ((time.time() - t0) / iters, "port", nnz) + ((None,) if replicas > 1 else ())
    tmp = tempfile.mkdtemp(prefix="hpf_ref_")
    try:
        data = os.path.join(tmp, "data")
        os.makedirs(data)
        deg = np.diff(sample["row_ptr"].astype(np.int64))
        uu = np.repeat(np.arange(n), deg)
        yy = sample["y"] if sample["y"] is not None else np.ones(nnz, np.uint8)
        np.savetxt(os.path.join(data, "train.tsv"), np.stack([uu + 1, sample["col_idx"].astype(np.int64) + 1, yy], 1),
                   fmt="%d", delimiter="\t")
        hu, hi, hy = sample["heldout"]
        held = np.stack([hu.astype(np.int64) + 1, hi.astype(np.int64) + 1, hy], 1)[:2000]
        for name in ("validation.tsv", "test.tsv"):
            np.savetxt(os.path.join(data, name), held, fmt="%d", delimiter="\t")
        common = ["-dir", data, "-n", str(n), "-m", str(m), "-k", str(k), "-seed", "1"] + ref_switches(cfg)

        def run(T, copies=1):
            t0 = time.time()
            procs = []
            for c in range(copies):
                if hier:
                    cmd = [O.REF_BINARY] + common + ["-rfreq", "100000", "-max-iterations", str(T - 1), "-label", "b%d_%d_%d" % (T, copies, c)]
                else:
                    cmd = [O.REF_HARNESS] + common + ["-iters", str(T), "-dump", os.path.join(tmp, "d%d_%d_%d" % (T, copies, c)),
                                                     "-label", "b%d_%d_%d" % (T, copies, c)]
                procs.append(subprocess.Popen(cmd, cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
            rcs = [p.wait() for p in procs]
            if any(rcs):
                raise subprocess.CalledProcessError(max(rcs), cmd[0])
            return time.time() - t0
        ta = run(max(1, warm))
        tb = run(max(1, warm) + iters)
        if replicas > 1:
            it2 = max(1, min(iters, 2))
            ra = run(1, replicas)
            rb = run(1 + it2, replicas)
            return (tb - ta) / iters, "reference", nnz, (rb - ra) / it2
        return (tb - ta) / iters, "reference", nnz
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def reference_sample(cfg, budget_iters, seconds=20.0, assumed_rate=1.9e5):
    """A contiguous user range of the workload sized for ~`seconds` of CPU work."""
    from hgaprec_b200 import synth
    want_nnz = max(20000, int(seconds * assumed_rate / max(1, budget_iters)))
    frac = min(1.0, want_nnz / cfg["nnz"])
    users = max(50, int(cfg["n"] * frac))
    d = synth.make_ratings(cfg["n"], cfg["m"], cfg["nnz"], binary=cfg["binary"], seed=cfg["seed"], heldout=0.02,
                           users_lo=0, users_hi=users)
    return d, "first %d of %d users of the workload (%d nnz, all %d items), K=%d" % (
        users, cfg["n"], len(d["col_idx"]), cfg["m"], cfg["k"])


# ----------------------------------------------------------------- main arm
def algorithmic_bytes(n, m, nnz, k, has_y, bias):
    """B_nnz = 4 + 1 + 4K + 4K (+8 bias) + (20K n + 12K m)/nnz."""
    per = 4 + (1 if has_y else 0) + 8 * k + (8 if bias else 0)
    return per * nnz + 20 * k * n + 12 * k * m


def sweep_launch_bytes(rows, nnz, k, has_y):
    """One sweep launch: per nonzero the index, the rating and the gathered K-row;
    per row the row-side factor row in and the accumulator row out."""
    return nnz * (4 + (1 if has_y else 0) + 4 * k) + rows * 8 * k


def workload_name(name, cfg):
    return "synthetic %s-scale %dx%d, %d nnz, K=%d%s%s%s" % (
        name, cfg["n"], cfg["m"], cfg["nnz"], cfg["k"], ", -hier" if cfg.get("hier", True) else " (BPF)",
        " -bias" if cfg.get("bias") else "", " -binary-data" if cfg["binary"] else "")


def engine_flags(H, cfg):
    return (H.HIER if cfg.get("hier", True) else 0) | (H.BIAS if cfg.get("bias") else 0) | (H.BINARY if cfg["binary"] else 0)


def start_state(H, eng, n, m, cfg, rng_theta, rng_beta):
    """Random Gamma start in the law of the reference's initialize() (shape 0.3 + 0.01 U, rate 0.3 + 0.1 U)."""
    k, hier, bias = cfg["k"], cfg.get("hier", True), cfg.get("bias", False)

    def rs(rows, rng):
        shp = 0.3 + 0.01 * rng.random((rows, k))
        rate = 0.3 + 0.1 * rng.random((rows, k))
        return shp, (rate if hier else rate[0]), shp / rate, np.log(shp / rate) - 0.5 / shp
    eng.set_state(H.THETA, *rs(n, rng_theta))
    eng.set_state(H.BETA, *rs(m, rng_beta))  # identical on every rank
    if hier:
        eng.set_state(H.THETARATE, np.full(n, 0.3), np.full(n, 0.3 + k), np.full(n, 0.3 / (0.3 + k)))
        eng.set_state(H.BETARATE, np.full(m, 0.3), np.full(m, 0.3 + k), np.full(m, 0.3 / (0.3 + k)))
    if bias:
        for which, rows, other, rng in ((H.THETABIAS, n, m, rng_theta), (H.BETABIAS, m, n, rng_beta)):
            shp, rate = 0.3 + 0.01 * rng.random(rows), np.full(rows, 0.3 + other)
            eng.set_state(which, shp, rate, shp / rate, np.log(shp / rate) - 0.5 / shp)


def roofline_of(prof, stats, n, m, nnz, cfg, has_y, ms_per_step, l2_resident):
    """Roofline block of the dominant kernel, hpf::sweep_kernel (user pass + item pass; with the dense tcgen05 head on,
    the head items' nonzeros run in head_kernel and are not in these launches)."""
    k = cfg["k"]
    head_nnz = int(stats["head_nnz"])
    gather_nnz = nnz - head_nnz
    sweep_bytes = sweep_launch_bytes(n, gather_nnz, k, has_y) + sweep_launch_bytes(m, gather_nnz, k, has_y)
    sweep_ms = prof["sweep_user_ms"] - prof["sweep_user_head_ms"] + prof["sweep_item_ms"]
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    hbm, hbm_src = measured_peaks()
    l2, l2_src = l2_gather_peak()
    launches = 1 + int(stats["item_chunks"])
    r = {"kernel": "hpf::sweep_kernel (user pass + item pass%s)" % (" in %d chunks" % stats["item_chunks"] if stats["item_chunks"] > 1 else ""),
         "achieved": achieved, "unit": "GB/s",
         "algorithmic_bytes_per_launch": sweep_bytes / launches, "avg_launch_ms": sweep_ms / launches, "launches_per_step": launches,
         "share_of_step": sweep_ms / prof["total_ms"],
         "gather_nnz_per_pass": gather_nnz, "dense_head_nnz": head_nnz, "dense_head_ms": prof["sweep_user_head_ms"],
         "per_kernel_ms": prof,
         "iteration_model": {"B_nnz": algorithmic_bytes(n, m, nnz, k, has_y, cfg.get("bias", False)) / max(1, nnz),
                             "whole_iteration_GBps": algorithmic_bytes(n, m, nnz, k, has_y, cfg.get("bias", False)) / (ms_per_step * 1e-3) / 1e9}}
    if l2_resident:
        # the gathered factor rows live in the 126 MB L2 at this size: DRAM sees ~5 % of the algorithmic bytes, the roof
        # that binds is the L2 -> SM gather path
        r.update(bound="l2", peak=l2, frac=achieved / l2, peak_source=l2_src,
                 hbm_frac_of_algorithmic_bytes=achieved / hbm, hbm_peak=hbm,
                 note="algorithmic bytes count a gathered factor row once per nonzero (SURVEY 8d); the rows are L2-resident "
                      "at this size, so the peak is the measured L2->SM gather ceiling, not HBM")
    else:
        r.update(bound="hbm", peak=hbm, frac=achieved / hbm, peak_source=hbm_src)
    tfile = os.path.join(ROOT, "profiles", "sweep_dram_bytes.json")
    r["traffic"] = None
    if l2_resident and os.path.exists(tfile):
        try:
            t = json.load(open(tfile))
            r["traffic"] = t.get("dram_bytes_per_launch")
            r["traffic_source"] = t.get("source", "profiles/sweep_dram_bytes.json (ncu --set full capture of this kernel on this workload; not this run)")
        except Exception:
            pass
    return r


def run_workload(H, synth, torch, dist, name, cfg, rank, world, dev, steps, warmup, scaling, profile_iters=5,
                 e2e_steps=0, sampler=None, pre_iters=0, planted=False):
    """Generate this rank's shard, run W + K iterations, return the measurements (dict) and the engine's data."""
    k = cfg["k"]
    if scaling == "weak":
        n_glob, nnz_glob = cfg["n"] * world, cfg["nnz"] * world
        d = synth.make_ratings(n_glob, cfg["m"], nnz_glob, binary=cfg["binary"], seed=cfg["seed"], heldout=0.002,
                               device="cuda:%d" % dev, users_lo=rank * cfg["n"], users_hi=(rank + 1) * cfg["n"])
    elif planted:
        n_glob = cfg["n"]
        lo, hi = (rank * cfg["n"]) // world, ((rank + 1) * cfg["n"]) // world
        d = synth.make_planted(cfg["n"], cfg["m"], cfg["nnz"], seed=cfg["seed"], device="cuda:%d" % dev, users_lo=lo, users_hi=hi, heldout=0.002)
    else:
        n_glob = cfg["n"]
        d = synth.make_blocks(cfg["n"], cfg["m"], cfg["nnz"], rank * NBLOCKS // world, (rank + 1) * NBLOCKS // world, NBLOCKS,
                              binary=cfg["binary"], seed=cfg["seed"], heldout=0.002, device="cuda:%d" % dev)
    n, m = d["n"], d["m"]
    nnz = len(d["col_idx"])
    has_y = d["y"] is not None
    hu, hi_, hy = d["heldout"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    row_ptr, col_idx = pin(d["row_ptr"]), pin(d["col_idx"])
    yv = pin(d["y"]) if has_y else None
    del d
    torch.cuda.empty_cache()

    eng = H.Engine(n, m, k, flags=engine_flags(H, cfg), device=dev, n_users_global=n_glob)
    if world > 1:
        uid = [H.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(rank, world, uid[0])
    eng.set_ratings_csr(row_ptr, col_idx, yv)
    start_state(H, eng, n, m, cfg, np.random.default_rng(1234 + rank), np.random.default_rng(99))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if pre_iters:
        eng.iterate(pre_iters)
    slow0 = eng.stats()["slow_path_nnz"]
    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps
    if sampler is not None:
        sampler.start()
        time.sleep(1.0)
    barrier()
    eng.iterate(warmup + EXTRA_WARMUP)  # W untimed steps, and a few more: the sampler's start-up second lets the clocks drop
    l0 = eng.stats()["kernel_launches"]
    barrier()
    t_wall = time.time()
    eng.iterate(steps)  # CUDA events on the library's stream bracket exactly these K steps
    ms = eng.stats()["last_iterate_ms"]
    barrier()
    t_wall = time.time() - t_wall
    clocks = sampler.stop() if sampler is not None else None
    st = eng.stats()
    launches = st["kernel_launches"] - l0
    slow = st["slow_path_nnz"] - slow0
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tn = torch.tensor([float(nnz), float(n), float(slow)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tn)
        nnz_total, n_total, slow = float(tn[0].item()), float(tn[1].item()), float(tn[2].item())
    else:
        nnz_total, n_total = float(nnz), float(n)
    ms_per_step = ms / steps
    out = {"value": nnz_total / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "nnz_total": nnz_total, "users_total": n_total,
           "launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
           "slow_path_nnz_per_iteration": slow / (warmup + EXTRA_WARMUP + steps), "n": n, "m": m, "nnz": nnz, "has_y": has_y}
    # ---- per-kernel device times (live, CUDA events on the launching stream)
    prof = eng.iterate_profiled(profile_iters) if profile_iters else None
    out["prof"] = prof
    out["stats"] = eng.stats()

    # ---- end to end through the C ABI with HOST buffers: every step uploads the ratings (pinned host CSR -> device,
    # CSC and work lists built on the device), runs one iteration and reads the held-out log-likelihood back.
    if e2e_steps:
        eng.set_ratings_csr(row_ptr, col_idx, yv); eng.iterate(1); eng.heldout_loglik(hu, hi_, hy)  # warm
        barrier()
        t0 = time.time()
        for _ in range(e2e_steps):
            eng.set_ratings_csr(row_ptr, col_idx, yv)
            eng.iterate(1)
            ll = eng.heldout_loglik(hu, hi_, hy)
        barrier()
        e2e_s = (time.time() - t0) / e2e_steps
        if dist is not None:
            t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        h2d = row_ptr.nbytes + col_idx.nbytes + (yv.nbytes if has_y else 0) + hu.nbytes + hi_.nbytes + hy.nbytes
        out["e2e"] = {"value": nnz_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8 + 4 * min(m, 512) + 256,
                      "ms_per_step": e2e_s * 1e3, "heldout_mean_ll": ll / max(1, len(hu)),
                      "what": "hpf_set_ratings_csr(host CSR) + hpf_iterate(1) + hpf_heldout_loglik(host pairs) per step, per rank; "
                              "bytes are per rank"}
    eng.close()
    del eng
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="netflix", choices=["netflix", "msd", "bpf-1b"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; reported in config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary legs (weak line at N>1; MSD and steady state at N=1)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from hgaprec_b200 import synth

    def config(name):
        cfg = dict(synth.CONFIGS[name])
        if args.scale != 1.0:
            cfg["n"] = max(64, int(cfg["n"] * args.scale)); cfg["m"] = max(64, int(cfg["m"] * args.scale))
            cfg["nnz"] = max(1000, int(cfg["nnz"] * args.scale))
        return cfg
    cfg = config(args.workload)
    k = cfg["k"]
    workload = workload_name(args.workload, cfg)

    # ---------------- reference arm: rank 0 only, CPU only
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample, desc = reference_sample(cfg, 2 * args.warmup + args.steps, seconds=90.0)
        cores = os.cpu_count() or 1
        sec, kind, nnz_s, sec_rep = time_reference(sample, cfg, args.steps, args.warmup, replicas=cores)
        val = nnz_s / sec
        replicas = None if not sec_rep or sec_rep <= 0 else {
            "processes": cores, "value": cores * nnz_s / sec_rep, "unit": UNIT,
            "what": "aggregate of that many independent copies of the same job side by side: an upper bound with every "
                    "host core busy, not something the single-threaded reference can do for one job"}
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "sample": desc},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "sample": desc,
                                 "host_cores": os.cpu_count(), "replicas": replicas},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------- our arm
    import torch
    import hgaprec_b200 as H
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    torch.cuda.set_device(dev)

    # the clock sampler (one nvidia-smi for all GPUs, on rank 0) starts BEFORE the warm-up: its start-up enumerates the
    # devices and would otherwise land inside the ~100 ms timed region
    sampler = ClockSampler(range(world)) if rank == 0 else None
    r = run_workload(H, synth, torch, dist, args.workload, cfg, rank, world, dev, args.steps, args.warmup, args.scaling,
                     e2e_steps=max(1, args.e2e_steps), sampler=sampler)
    l2_resident = (r["n"] + r["m"]) * k * 4 * 2 < 400e6 and args.workload == "netflix"  # both sides' operand rows vs the 126 MB L2
    roofline = roofline_of(r["prof"], r["stats"], r["n"], r["m"], r["nnz"], cfg, r["has_y"], r["ms_per_step"], l2_resident)
    stats = r["stats"]

    extras = {}
    if not args.no_extras and args.scale == 1.0:
        try:
            if world > 1 and args.workload == "netflix" and args.scaling == "strong":
                w = run_workload(H, synth, torch, dist, "netflix", cfg, rank, world, dev, args.steps, args.warmup, "weak", profile_iters=0)
                extras["weak"] = {"value": w["value"], "unit": UNIT, "ms_per_step": w["ms_per_step"], "nnz_total": w["nnz_total"],
                                  "what": "round 1's measurement: every rank a Netflix-scale user shard of an N-times larger problem"}
            if world == 1 and args.workload == "netflix":
                mcfg = config("msd")
                mr = run_workload(H, synth, torch, dist, "msd", mcfg, rank, world, dev, 10, 3, "strong")
                extras["hbm_bound_workload"] = {
                    "workload": workload_name("msd", mcfg) + " (BASELINE configs[2], whole problem on one GPU)",
                    "value": mr["value"], "unit": UNIT, "ms_per_step": mr["ms_per_step"],
                    "roofline": roofline_of(mr["prof"], mr["stats"], mr["n"], mr["m"], mr["nnz"], mcfg, mr["has_y"], mr["ms_per_step"], False)}
                pcfg = dict(cfg)
                ss = run_workload(H, synth, torch, dist, "netflix", pcfg, rank, world, dev, args.steps, args.warmup, "strong",
                                  pre_iters=100, planted=True)
                extras["steady_state"] = {
                    "what": "planted-factor data of the same shape (theta, beta ~ Gamma(0.3), y ~ Poisson(theta.beta), 20 factors), "
                            "timed after 100 iterations of the fit", "value": ss["value"], "unit": UNIT, "ms_per_step": ss["ms_per_step"],
                    "nnz": ss["nnz_total"], "slow_path_nnz_per_iteration": ss["slow_path_nnz_per_iteration"],
                    "dense_head_nnz": int(ss["stats"]["head_nnz"]), "per_kernel_ms": ss["prof"]}
        except Exception as ex:  # a secondary leg must not take the headline down with it
            extras["extras_error"] = repr(ex)
            print("bench.py: a secondary leg FAILED (%r); see extras_error in the line" % (ex,), file=sys.stderr, flush=True)

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sample, desc = reference_sample(cfg, 4, seconds=20.0)
            cores = os.cpu_count() or 1
            sec, kind, nnz_s, sec_rep = time_reference(sample, cfg, 3, 1, replicas=cores)
            cpu = {"value": nnz_s / sec, "unit": UNIT, "cores": 1, "kind": kind, "sample": desc,
                   "host_cores": os.cpu_count(), "s_per_iteration_on_sample": sec,
                   "replicas": None if not sec_rep or sec_rep <= 0 else {
                       "processes": cores, "value": cores * nnz_s / sec_rep, "unit": UNIT,
                       "what": "that many independent copies of the same job side by side (upper bound with every host "
                               "core busy; the reference itself has no threads)"}}
        except Exception as ex:  # the baseline must not take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(ex)}
            print("bench.py: the CPU baseline leg FAILED (%r); the line is printed with cpu_baseline.kind = "
                  "'unavailable'" % (ex,), file=sys.stderr, flush=True)

    if rank == 0:
        shard = " (users sharded over %d GPUs, items replicated)" % world if world > 1 else ""
        if args.scaling == "weak" and world > 1:
            shard = " per GPU (every rank a shard of an %d-times larger problem)" % world
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload + shard,
                           "users_total": int(r["users_total"]), "items": r["m"], "nnz_total": int(r["nnz_total"]),
                           "users_rank0": r["n"], "nnz_rank0": r["nnz"], "k": k,
                           "l2_policy": "inputs larger than L2 (ratings %.0f MB + factor rows %.0f MB per GPU vs 126 MB L2)"
                                        % ((2 * r["nnz"] * 5) / 1e6, (r["n"] + r["m"]) * k * 4 * 2 / 1e6),
                           "sweep_group": stats["sweep_group"], "sweep_vec": stats["sweep_vec"],
                           "item_chunks": stats["item_chunks"], "beta_sharded": stats["beta_sharded"], "extra_untimed_warmup_steps": EXTRA_WARMUP,
                           "sweep_plan": ("gather kernel on both passes" if not stats["head_nnz"] else
                                          "gather kernel for the tail + dense tcgen05 head (%d nonzeros of the most popular items)"
                                          % stats["head_nnz"]),
                           "state_init": "random Gamma(0.3+U,0.3+U) start (reference initialize() law), synthetic"},
                "e2e": r["e2e"], "gpu_launches": r["launches"], "roofline": roofline, "cpu_baseline": cpu,
                "clocks": r["clocks"], "wall_s_timed_region": r["wall_s_timed_region"],
                "slow_path_nnz": int(stats["slow_path_nnz"])}
        line.update(extras)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
