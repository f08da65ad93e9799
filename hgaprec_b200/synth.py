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
    "bpf-1b": dict(n=10_000_000, m=1_000_000, nnz=1_000_000_000, k=100, binary=False, seed=20131106, bias=True, hier=False),
}


def _expected_unique(lam, w):
    # E[#distinct items] when a user makes Poisson(lam) Zipf draws
    return (1.0 - torch.exp(-lam[:, None] * w[None, :])).sum(dim=1)


def make_ratings(n, m, nnz, binary=False, seed=0, zipf_s=1.0, sigma=1.0, heldout=0.0,
                 device=None, users_lo=0, users_hi=None):
    """Return dict(n, m, row_ptr, col_idx, y, heldout=(u, i, y) or None).

    users_lo/users_hi generate only a contiguous user range of an n-user problem
    (one rank's shard, without materialising the others): same item popularity,
    same activity law and per-user rates as the global problem; the draws of a
    range depend only on (seed, users_lo), so a fixed split into ranges is one fixed
    data set whoever generates the pieces (make_blocks).  A range is NOT a slice of
    the full-range call's data; tests that need exact shards of one CSR slice it
    with hpf_partition_users instead."""
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

    # draws for the selected user range, chunked to bound memory.  A sub-range gets its own stream (keyed on where
    # it starts), so that the ranges of one problem are independent draws and a fixed set of ranges is a fixed data set
    # however the ranges are spread over ranks (bench.py: strong scaling).
    if users_lo != 0 or users_hi != n:
        g.manual_seed((int(seed) * 1000003 + int(users_lo) * 7919 + 12345) % (2 ** 63 - 1))
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


def concat_blocks(blocks):
    """Stack user ranges returned by make_ratings (consecutive ranges of one problem) into one CSR."""
    if len(blocks) == 1:
        return blocks[0]
    rp = [blocks[0]["row_ptr"]]
    off = int(blocks[0]["row_ptr"][-1])
    for b in blocks[1:]:
        rp.append(b["row_ptr"][1:] + np.uint64(off))
        off += int(b["row_ptr"][-1])
    held, nu = None, 0
    if blocks[0]["heldout"] is not None:
        hu, hi, hy = [], [], []
        for b in blocks:
            hu.append(b["heldout"][0] + np.uint32(nu)); hi.append(b["heldout"][1]); hy.append(b["heldout"][2])
            nu += b["n"]
        held = (np.concatenate(hu), np.concatenate(hi), np.concatenate(hy))
    return dict(n=sum(b["n"] for b in blocks), n_global=blocks[0]["n_global"], m=blocks[0]["m"],
                row_ptr=np.concatenate(rp), col_idx=np.concatenate([b["col_idx"] for b in blocks]),
                y=None if blocks[0]["y"] is None else np.concatenate([b["y"] for b in blocks]), heldout=held)


def make_blocks(n, m, nnz, first, last, nblocks=8, **kw):
    """Blocks first..last-1 of the fixed nblocks-block split of an n-user problem: the same data whoever generates
    them (one rank all eight, or eight ranks one each)."""
    bounds = [(b * n) // nblocks for b in range(nblocks + 1)]
    return concat_blocks([make_ratings(n, m, nnz, users_lo=bounds[b], users_hi=bounds[b + 1], **kw) for b in range(first, last)])


def make_planted(n, m, nnz, k_true=20, seed=0, device=None, users_lo=0, users_hi=None, heldout=0.0):
    """Ratings with PLANTED low-rank structure: theta_u ~ Gamma(0.3, .) (k_true factors, a few active per user),
    beta_i ~ Gamma(0.3, .) scaled by a Zipf popularity, y_ui ~ Poisson(theta_u . beta_i), zeros dropped, counts
    clipped to 255 -- the generative model HPF fits (SURVEY.md 8d asks for structure-free data; a fit of that
    never develops the peaked rows a converged real fit has, so the steady-state measurement uses this one).
    The global scale is calibrated so that the expected number of nonzeros is nnz."""
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    users_hi = n if users_hi is None else users_hi

    def gamma(shape, size):
        # Gamma(shape < 1) by Ahrens-Dieter via the boost trick: G(a) = G(a + 1) * U^(1/a); G(a+1) by Marsaglia-Tsang
        a1 = shape + 1.0
        dd = a1 - 1.0 / 3.0
        cc = 1.0 / (9.0 * dd) ** 0.5
        out = torch.empty(size, device=dev)
        todo = torch.ones(size, dtype=torch.bool, device=dev)
        while bool(todo.any()):
            cnt = int(todo.sum())
            x = torch.randn(cnt, generator=g, device=dev)
            v = (1.0 + cc * x) ** 3
            u = torch.rand(cnt, generator=g, device=dev)
            ok = (v > 0) & (torch.log(u) < 0.5 * x * x + dd - dd * v + dd * torch.log(v.clamp(min=1e-30)))
            idx = todo.nonzero(as_tuple=True)
            sel = tuple(i[ok] for i in idx)
            out[sel] = (dd * v[ok])
            todo[sel] = False
        return out * torch.rand(size, generator=g, device=dev).pow(1.0 / shape)
    rank = torch.arange(1, m + 1, dtype=torch.float32, device=dev)
    pop = rank.pow(-0.8)[torch.randperm(m, generator=g, device=dev)]
    beta = gamma(0.3, (m, k_true)) * pop[:, None]
    act = torch.exp(0.8 * torch.randn(n, generator=g, device=dev))
    # calibrate on a user sample: E[#nonzeros] = sum (1 - exp(-rate))
    ts = gamma(0.3, (min(n, 2048), k_true)) * act[: min(n, 2048), None]
    lo, hi = 1e-6, 1e6
    for _ in range(60):
        c = (lo * hi) ** 0.5
        dens = (1.0 - torch.exp(-(c * ts) @ beta.T)).sum(dim=1).mean().item()
        if dens < float(nnz) / n:
            lo = c
        else:
            hi = c
    g.manual_seed((int(seed) * 1000003 + int(users_lo) * 7919 + 777) % (2 ** 63 - 1))
    us, its, ys = [], [], []
    chunk = max(1, int(6e7 // m))
    for u0 in range(users_lo, users_hi, chunk):
        u1 = min(users_hi, u0 + chunk)
        th = gamma(0.3, (u1 - u0, k_true)) * act[u0:u1, None] * c
        cnt = torch.poisson(th @ beta.T, generator=g)
        nz = cnt.nonzero(as_tuple=False)
        # shuffled inside each row, like a file in arbitrary order
        key = nz[:, 0] * (2 ** 31) + torch.randint(0, 2 ** 31 - 1, (len(nz),), generator=g, device=dev)
        nz = nz[torch.argsort(key)]
        us.append(nz[:, 0] + (u0 - users_lo)); its.append(nz[:, 1]); ys.append(cnt[nz[:, 0], nz[:, 1]].clamp(max=255).to(torch.uint8))
    u, i, y = torch.cat(us), torch.cat(its), torch.cat(ys)
    held = None
    if heldout > 0:
        hm = torch.rand(len(u), generator=g, device=dev) < heldout
        first = torch.ones_like(hm)
        first[1:] = u[1:] != u[:-1]
        hm &= ~first
        held = (u[hm].cpu().numpy().astype(np.uint32), i[hm].cpu().numpy().astype(np.uint32), y[hm].cpu().numpy())
        u, i, y = u[~hm], i[~hm], y[~hm]
    counts = torch.bincount(u, minlength=users_hi - users_lo)
    row_ptr = torch.zeros(users_hi - users_lo + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(counts, 0)
    return dict(n=users_hi - users_lo, n_global=n, m=m, row_ptr=row_ptr.cpu().numpy().astype(np.uint64),
                col_idx=i.cpu().numpy().astype(np.uint32), y=y.cpu().numpy(), heldout=held)


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
