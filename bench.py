#!/usr/bin/env python
"""bench.py -- HPF CAVI throughput (training nonzeros / second) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one full CAVI iteration (hpf_iterate(1): user sweep, item sweep,
theta / beta / xi / eta updates) over the rank's ratings shard.  Workload at
N=1 is BASELINE.json configs[1]: synthetic Netflix-scale 480,189 x 17,770,
1e8 nnz, K=100, -hier.  With N>1 (torchrun, one rank per GPU) every rank holds
a Netflix-scale USER shard of an N-times larger problem (items fixed, beta
replicated, one NCCL all-reduce of the item-side block per iteration):
weak scaling.  One JSON line is printed by rank 0.

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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
def time_reference(sample, k, iters, warm, flags_hier=True, replicas=0):
    """Per-iteration wall time of the unmodified reference binary on `sample`,
    by differencing two runs (-max-iterations T runs T+1 iterations,
    src/hgaprec.cc:1337-1339): (wall(warm+iters) - wall(warm)) / iters.

    replicas > 1 adds a fourth return value: the per-iteration wall time when that many INDEPENDENT copies of the
    same job run side by side (the reference has no threads, so this is not something it can do for one job: it is
    the upper bound "every host core busy"), measured the same way over min(iters, 2) iterations."""
    from oracle import hpf_oracle as O
    if not os.path.exists(O.REF_BINARY):
        O.build()
    use_ref = os.path.exists(O.REF_BINARY)
    n, m = sample["n"], sample["m"]
    nnz = len(sample["col_idx"])
    if not use_ref:
        # the port: same loop, plain C, one thread
        s = O.OracleState(n, m, k, O.HIER).init(1)
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

        def run(T, copies=1):
            t0 = time.time()
            procs = [subprocess.Popen([O.REF_BINARY, "-dir", data, "-n", str(n), "-m", str(m), "-k", str(k), "-hier",
                                       "-rfreq", "100000", "-max-iterations", str(T - 1), "-seed", "1",
                                       "-label", "b%d_%d_%d" % (T, copies, c)],
                                      cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in range(copies)]
            rcs = [p.wait() for p in procs]
            if any(rcs):
                raise subprocess.CalledProcessError(max(rcs), O.REF_BINARY)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="netflix")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (debug only; reported in config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from hgaprec_b200 import synth
    cfg = dict(synth.CONFIGS[args.workload])
    if args.scale != 1.0:
        cfg["n"] = max(64, int(cfg["n"] * args.scale)); cfg["m"] = max(64, int(cfg["m"] * args.scale))
        cfg["nnz"] = max(1000, int(cfg["nnz"] * args.scale))
    k = cfg["k"]
    workload = "synthetic %s-scale %dx%d, %d nnz, K=%d, -hier%s" % (
        args.workload, cfg["n"], cfg["m"], cfg["nnz"], k, " -binary-data" if cfg["binary"] else "")

    # ---------------- reference arm: rank 0 only, CPU only
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample, desc = reference_sample(cfg, 2 * args.warmup + args.steps, seconds=90.0)
        cores = os.cpu_count() or 1
        sec, kind, nnz_s, sec_rep = time_reference(sample, k, args.steps, args.warmup, replicas=cores)
        val = nnz_s / sec
        replicas = None if not sec_rep or sec_rep <= 0 else {
            "processes": cores, "value": cores * nnz_s / sec_rep, "unit": UNIT,
            "what": "aggregate of that many independent copies of the same job side by side: an upper bound with every "
                    "host core busy, not something the single-threaded reference can do for one job"}
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
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

    # this rank's user shard of the (world x larger) problem
    n_glob, nnz_glob = cfg["n"] * world, cfg["nnz"] * world
    d = synth.make_ratings(n_glob, cfg["m"], nnz_glob, binary=cfg["binary"], seed=cfg["seed"], heldout=0.002,
                           device="cuda:%d" % dev, users_lo=rank * cfg["n"], users_hi=(rank + 1) * cfg["n"])
    n, m = d["n"], d["m"]
    nnz = len(d["col_idx"])
    has_y = d["y"] is not None
    hu, hi, hy = d["heldout"]
    # pinned host copies (the e2e leg copies from these every step)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    row_ptr, col_idx = pin(d["row_ptr"]), pin(d["col_idx"])
    yv = pin(d["y"]) if has_y else None
    torch.cuda.empty_cache()

    flags = H.HIER | (H.BINARY if cfg["binary"] else 0)
    eng = H.Engine(n, m, k, flags=flags, device=dev, n_users_global=n_glob)
    if world > 1:
        uid = [H.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(rank, world, uid[0])
    eng.set_ratings_csr(row_ptr, col_idx, yv)
    rng = np.random.default_rng(1234 + rank)

    def random_state(rows, seed_rng):
        shp = 0.3 + 0.01 * seed_rng.random((rows, k))
        rate = 0.3 + 0.1 * seed_rng.random((rows, k))
        return shp, rate, shp / rate, np.log(shp / rate) - 0.5 / shp
    eng.set_state(H.THETA, *random_state(n, rng))
    eng.set_state(H.BETA, *random_state(m, np.random.default_rng(99)))  # identical on every rank
    eng.set_state(H.THETARATE, np.full(n, 0.3), np.full(n, 0.3 + k), np.full(n, 0.3 / (0.3 + k)))
    eng.set_state(H.BETARATE, np.full(m, 0.3), np.full(m, 0.3 + k), np.full(m, 0.3 / (0.3 + k)))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps.  The clock sampler
    # (one nvidia-smi for all GPUs, on rank 0) starts BEFORE the warm-up: its start-up enumerates the
    # devices and would otherwise land inside the ~100 ms timed region.
    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler is not None:
        sampler.start()
        time.sleep(1.0)
    barrier()
    eng.iterate(args.warmup)
    l0 = eng.stats()["kernel_launches"]
    barrier()
    t_wall = time.time()
    eng.iterate(args.steps)  # CUDA events on the library's stream bracket exactly these K steps
    ms = eng.stats()["last_iterate_ms"]
    barrier()
    t_wall = time.time() - t_wall
    clocks = sampler.stop() if sampler is not None else None
    launches = eng.stats()["kernel_launches"] - l0
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tn = torch.tensor([float(nnz)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tn)
        nnz_total = float(tn.item())
    else:
        nnz_total = float(nnz)
    ms_per_step = ms / args.steps
    value = nnz_total / (ms_per_step * 1e-3)

    # ---- per-kernel device times (live, CUDA events on the launching stream)
    prof = eng.iterate_profiled(5)
    peak, peak_src = measured_peaks()
    # dominant kernel: hpf::sweep_kernel, two launches per iteration (user pass, item pass).  When the dense
    # tcgen05 head is on, the nonzeros of the head items are not in these launches (they run in head_kernel).
    stats0 = eng.stats()
    head_nnz = int(stats0["head_nnz"]) if not stats0["item_tiles"] else 0
    gather_nnz = nnz - head_nnz
    sweep_bytes = sweep_launch_bytes(n, gather_nnz, k, has_y) + sweep_launch_bytes(m, gather_nnz, k, has_y)
    sweep_ms = prof["sweep_user_ms"] - prof["sweep_user_head_ms"] + prof["sweep_item_ms"]
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "sweep_dram_bytes.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "hpf::sweep_kernel (2 launches / iteration: user pass + item pass)",
                "note": "algorithmic bytes count a gathered factor row once per nonzero (SURVEY 8d); the rows are L2-resident at "
                        "this size, so frac > 1 is expected and the binding limit is the L2->SM gather rate (see DESIGN.md 5)",
                "l2_gather_TBps": (2 * gather_nnz * (4 * ((k + 3) // 4 * 4) + 5)) / (sweep_ms * 1e-3) / 1e12,
                "gather_nnz_per_launch": gather_nnz, "dense_head_nnz": head_nnz,
                "dense_head_ms": prof["sweep_user_head_ms"],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sweep_bytes / 2,
                "avg_launch_ms": sweep_ms / 2,
                "share_of_step": sweep_ms / prof["total_ms"],
                "per_kernel_ms": prof,
                "iteration_model": {"B_nnz": algorithmic_bytes(n, m, nnz, k, has_y, False) / nnz,
                                    "whole_iteration_GBps": algorithmic_bytes(n, m, nnz, k, has_y, False) / (ms_per_step * 1e-3) / 1e9}}

    # ---- end to end through the C ABI with HOST buffers: every step uploads the
    # ratings (pinned host CSR -> device, CSC built on device), runs one iteration
    # and reads the held-out log-likelihood back.
    e2e_steps = max(1, args.e2e_steps)
    eng.set_ratings_csr(row_ptr, col_idx, yv); eng.iterate(1); eng.heldout_loglik(hu, hi, hy)  # warm
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        eng.set_ratings_csr(row_ptr, col_idx, yv)
        eng.iterate(1)
        ll = eng.heldout_loglik(hu, hi, hy)
    barrier()
    e2e_s = (time.time() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = row_ptr.nbytes + col_idx.nbytes + (yv.nbytes if has_y else 0) + hu.nbytes + hi.nbytes + hy.nbytes
    d2h = 8 + (m + 1) * 8
    e2e = {"value": nnz_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": e2e_s * 1e3, "heldout_mean_ll": ll / max(1, len(hu)),
           "what": "hpf_set_ratings_csr(host CSR) + hpf_iterate(1) + hpf_heldout_loglik(host pairs) per step"}
    stats = eng.stats()
    eng.close()

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            sample, desc = reference_sample(cfg, 4, seconds=20.0)
            cores = os.cpu_count() or 1
            sec, kind, nnz_s, sec_rep = time_reference(sample, k, 3, 1, replicas=cores)
            cpu = {"value": nnz_s / sec, "unit": UNIT, "cores": 1, "kind": kind, "sample": desc,
                   "host_cores": os.cpu_count(), "s_per_iteration_on_sample": sec,
                   "replicas": None if not sec_rep or sec_rep <= 0 else {
                       "processes": cores, "value": cores * nnz_s / sec_rep, "unit": UNIT,
                       "what": "that many independent copies of the same job side by side (upper bound with every host "
                               "core busy; the reference itself has no threads)"}}
        except Exception as ex:  # the baseline must not take the GPU number down with it
            cpu = {"value": None, "unit": UNIT, "cores": 1, "kind": "unavailable", "sample": repr(ex)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload + (" per GPU (users sharded, items replicated)" if world > 1 else ""),
                           "users_per_gpu": n, "items": m, "nnz_per_gpu": nnz, "k": k,
                           "l2_policy": "inputs larger than L2 (ratings %.0f MB + factor rows %.0f MB per GPU vs 126 MB L2)"
                                        % ((2 * nnz * 5) / 1e6, (n + m) * k * 4 * 2 / 1e6),
                           "sweep_group": stats["sweep_group"], "sweep_vec": stats["sweep_vec"],
                           "sweep_plan": ("gather kernel on both passes" if not (stats["item_tiles"] or stats["head_nnz"]) else
                                          "gather kernel for the tail + dense tcgen05 head (%d nonzeros of the most popular items)"
                                          % stats["head_nnz"] if not stats["item_tiles"] else
                                          "tile sweeps: item_tiles=%d head_nnz=%d" % (stats["item_tiles"], stats["head_nnz"])),
                           "state_init": "random Gamma(0.3+U,0.3+U) start (reference initialize() law), synthetic"},
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
                "clocks": clocks, "wall_s_timed_region": t_wall,
                "slow_path_nnz": int(stats["slow_path_nnz"])}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
