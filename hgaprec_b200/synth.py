"""Seeded synthetic ratings in the shape of the reference's data sets.

User activity is log-normal, item popularity is Zipf(s), a user never rates an
item twice, rows keep an arbitrary (shuffled) item order like the reference's
file order (src/ratings.cc:63-119), ratings follow a Netflix-like 1..5 pmf or are
all ones (-binary-data).  Runs on torch (CUDA when available, else CPU) so the
100M-nonzero configurations are generated in seconds on the GPU box; the result
is returned as host numpy arrays in the layout hpf_set_ratings_csr takes.

This is input plumbing for tests and bench.py, not part of the hot path.
"""
import numpy as np
import torch

RATING_PMF = (0.05, 0.10, 0.29, 0.33, 0.23)  # y = 1..5

CONFIGS = {
    # name: n, m, nnz, k, flags-as-names, seed     (BASELINE.json configs / SURVEY.md 8)
    "movielens-shape": dict(n=6040, m=3681, nnz=792_166, k=100, binary=False, seed=20131103),
    "netflix": dict(n=480_189, m=17_770, nnz=100_000_000, k=100, binary=False, seed=20131104),
    "msd": dict(n=1_019_318, m=384_546, nnz=48_000_000, k=200, binary=True, seed=20131105),
    "bpf-1b": dict(n=10_000_000, m=1_000_000, nnz=1_000_000_000, k=100, binary=False, seed=20131106),
}


def _expected_unique(lam, w):
    # E[#distinct items] when a user makes Poisson(lam) Zipf draws
    return (1.0 - torch.exp(-lam[:, None] * w[None, :])).sum(dim=1)


def make_ratings(n, m, nnz, binary=False, seed=0, zipf_s=1.0, sigma=1.0, heldout=0.0,
                 device=None, users_lo=0, users_hi=None):
    """Return dict(n, m, row_ptr, col_idx, y, heldout=(u, i, y) or None).

    users_lo/users_hi generate only a contiguous user range of an n-user problem
    (one rank's shard, without materialising the others): same item popularity,
    same activity law and per-user rates as the global problem, but the draws of a
    range are its own -- the ranges of different calls are statistically alike,
    they do not tile one fixed data set.  Tests that need exact shards slice a
    full CSR with hpf_partition_users instead."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    users_hi = n if users_hi is None else users_hi

    # item popularity (rank -> weight), randomly assigned to item ids
    rank = torch.arange(1, m + 1, dtype=torch.float64, device=dev)
    w = rank.pow(-zipf_s)
    w = (w / w.sum()).float()
    item_of_rank = torch.randperm(m, generator=g, device=dev)
    cdf = torch.cumsum(w.double(), 0).float()
    cdf[-1] = 1.0

    # user activity: log-normal, then a global multiplier c calibrated on a user
    # sample so that the expected number of DISTINCT items sums to nnz
    act = torch.exp(sigma * torch.randn(n, generator=g, device=dev, dtype=torch.float32))
    sample = act[torch.randperm(n, generator=g, device=dev)[: min(n, 2048)]]
    lo, hi = 1e-3, float(m) * 50.0
    target = float(nnz) / n
    for _ in range(60):
        c = (lo * hi) ** 0.5
        lam = torch.clamp(c * sample, max=8.0 * m)
        mean_deg = torch.clamp(_expected_unique(lam, w), min=1.0).clamp(max=m / 2).mean().item()
        if mean_deg < target:
            lo = c
        else:
            hi = c
    lam = torch.clamp(c * act, max=8.0 * m)

    # draws for the selected user range, chunked to bound memory
    sel = torch.arange(users_lo, users_hi, device=dev)
    ndraw = torch.poisson(lam[sel], generator=g).clamp(min=1).long()
    keys = []
    chunk_users = max(1, int(4e7 // max(1.0, ndraw.float().mean().item())))
    cap = m // 2 if m >= 4 else m
    for s0 in range(0, len(sel), chunk_users):
        nd = ndraw[s0:s0 + chunk_users]
        uu = torch.repeat_interleave(sel[s0:s0 + chunk_users], nd)
        r = torch.rand(len(uu), generator=g, device=dev)
        it = torch.searchsorted(cdf, r).clamp(max=m - 1)
        key = torch.unique(uu * m + item_of_rank[it])  # sorted, duplicates dropped
        # cap a user's degree at m/2: drop a random excess
        ku = torch.div(key, m, rounding_mode="floor")
        _, cnt = torch.unique_consecutive(ku, return_counts=True)
        if cnt.max().item() > cap:
            start = torch.cumsum(cnt, 0) - cnt
            rk = torch.rand(len(key), generator=g, device=dev)
            order = torch.argsort(ku.double() + rk.double() * 0.999)
            pos = torch.arange(len(key), device=dev) - torch.repeat_interleave(start, cnt)
            key = torch.sort(key[order][pos < cap])[0]
        keys.append(key)
    key = torch.cat(keys)
    u = torch.div(key, m, rounding_mode="floor")
    i = key - u * m
    # shuffle inside each row: sort by (user, random)
    rnd = torch.randint(0, 2 ** 31 - 1, (len(key),), generator=g, device=dev, dtype=torch.int64)
    order = torch.argsort(u * (2 ** 31) + rnd)
    u, i = u[order], i[order]
    if binary:
        y = None
    else:
        pm = torch.tensor(RATING_PMF, device=dev).cumsum(0)
        pm[-1] = 1.0
        y = (torch.searchsorted(pm, torch.rand(len(key), generator=g, device=dev)).clamp(max=4) + 1).to(torch.uint8)

    held = None
    if heldout > 0:
        hm = torch.rand(len(key), generator=g, device=dev) < heldout
        # keep at least the first rating of every user in training
        first = torch.ones_like(hm)
        first[1:] = u[1:] != u[:-1]
        hm &= ~first
        hy = torch.ones(int(hm.sum()), dtype=torch.uint8, device=dev) if y is None else y[hm]
        held = ((u[hm] - users_lo).cpu().numpy().astype(np.uint32), i[hm].cpu().numpy().astype(np.uint32),
                hy.cpu().numpy())
        u, i = u[~hm], i[~hm]
        if y is not None:
            y = y[~hm]

    counts = torch.bincount(u - users_lo, minlength=users_hi - users_lo)
    row_ptr = torch.zeros(users_hi - users_lo + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(counts, 0)
    return dict(n=users_hi - users_lo, n_global=n, m=m,
                row_ptr=row_ptr.cpu().numpy().astype(np.uint64),
                col_idx=i.cpu().numpy().astype(np.uint32),
                y=None if y is None else y.cpu().numpy(),
                heldout=held)


def make_config(name, scale=1.0, **kw):
    """One of CONFIGS, optionally scaled down (users, items and nnz by `scale`)."""
    c = dict(CONFIGS[name])
    k = c.pop("k")
    n = max(8, int(round(c["n"] * scale)))
    m = max(8, int(round(c["m"] * scale)))
    nnz = max(n, int(round(c["nnz"] * scale)))
    d = make_ratings(n, m, nnz, binary=c["binary"], seed=c["seed"], **kw)
    d["k"] = k
    return d
