"""Vectorised on-device generator of synthetic matching LPs with the distributions of the reference's generator
(benchmark/generate_synthetic_data.py:27-164): destination breadth Z_j ~ LogN(0,1) -> p_j, scale s_j ~ LogN(0,1),
value v_j ~ LogN(-4,0.75), source affinity u_i ~ LogN(0,0.5), c_ij = min(v_j u_i LogN(0,0.5), 0.5), a_ij = s_j c_ij,
stored c = -c_ij, b_j = U(0.5,1) * (greedy load_j + 1e-8).

The reference draws K_j ~ Poisson(p_j n) distinct sources per destination in a Python loop (8 s per million sources);
here the same bipartite graph law is sampled from the source side: degree_i ~ Poisson(sum_j p_j), destinations drawn
with probability proportional to p_j, duplicates within a column removed, rows sorted within each column.  It is not
stream-identical to the reference generator (different RNG and sampling order), so parity tests use committed
fixtures and this generator feeds both arms of the benchmark.

Shard-direct: a rank generates only its own contiguous column range [col_start, col_end) on its GPU; per-source
randomness is seeded per chunk of `chunk_cols` global columns, so the global problem does not depend on the sharding.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass
class SyntheticShard:
    ccol: torch.Tensor  # int64 (n_local+1)
    row: torch.Tensor  # int64 (E_local)
    a: torch.Tensor  # float32
    c: torch.Tensor  # float32 (negated values)
    greedy_load: torch.Tensor  # float64 (m): this shard's contribution to the greedy destination loads
    n_rows: int
    col_start: int
    col_end: int


def destination_params(num_destinations: int, target_sparsity: float, seed: int, device):
    g = torch.Generator(device="cpu").manual_seed(seed)
    z = torch.exp(torch.randn(num_destinations, generator=g, dtype=torch.float64))
    s = torch.exp(torch.randn(num_destinations, generator=g, dtype=torch.float64))
    v = torch.exp(-4.0 + 0.75 * torch.randn(num_destinations, generator=g, dtype=torch.float64))
    rho = 0.5 + 0.5 * torch.rand(num_destinations, generator=g, dtype=torch.float64)
    avg_degree = target_sparsity * num_destinations
    p = z / z.sum() * avg_degree
    cdf = torch.cumsum(p / p.sum(), 0)
    cdf[-1] = 1.0
    return dict(p=p.to(device), s=s.to(device), v=v.to(device), rho=rho.to(device), cdf=cdf.to(device), avg_degree=avg_degree)


def generate_shard(num_sources: int, num_destinations: int, target_sparsity: float, seed: int, device,
                   col_start: int = 0, col_end: int | None = None, chunk_cols: int = 4_000_000) -> SyntheticShard:
    col_end = num_sources if col_end is None else col_end
    device = torch.device(device)
    dp = destination_params(num_destinations, target_sparsity, seed, device)
    m = num_destinations
    cdf32 = dp["cdf"]
    deg_parts, row_parts, a_parts, c_parts = [], [], [], []
    load = torch.zeros(m, dtype=torch.float64, device=device)
    first_chunk = col_start // chunk_cols
    last_chunk = (max(col_end, col_start + 1) - 1) // chunk_cols
    for ch in range(first_chunk, last_chunk + 1):
        g0, g1 = ch * chunk_cols, min((ch + 1) * chunk_cols, num_sources)
        gen = torch.Generator(device=device).manual_seed(seed * 1_000_003 + 17 * ch + 1)
        nc = g1 - g0
        deg = torch.poisson(torch.full((nc,), float(dp["avg_degree"]), device=device), generator=gen).to(torch.int64)
        deg.clamp_(max=m)
        u = torch.exp(0.5 * torch.randn(nc, device=device, generator=gen))
        total = int(deg.sum().item())
        col = torch.repeat_interleave(torch.arange(nc, device=device), deg, output_size=total)
        dest = torch.searchsorted(cdf32, torch.rand(total, device=device, dtype=torch.float64, generator=gen)).clamp_(max=m - 1)
        eps = torch.exp(0.5 * torch.randn(total, device=device, generator=gen))
        key = col * m + dest
        key, order = torch.sort(key)
        keep = torch.ones(total, dtype=torch.bool, device=device)
        keep[1:] = key[1:] != key[:-1]
        key = key[keep]
        eps = eps[order][keep]
        del order, keep, dest
        col = torch.div(key, m, rounding_mode="floor")
        dest = key - col * m
        del key
        # restrict to this shard's columns inside the chunk
        lo, hi = max(col_start, g0) - g0, min(col_end, g1) - g0
        if lo > 0 or hi < nc:
            sel = (col >= lo) & (col < hi)
            col, dest, eps = col[sel] - lo, dest[sel], eps[sel]
            u = u[lo:hi]
        ncl = hi - lo
        cval = torch.minimum(dp["v"][dest].float() * u[col] * eps, torch.tensor(0.5, device=device))
        aval = dp["s"][dest].float() * cval
        deg = torch.bincount(col, minlength=ncl)
        # greedy load: every source sends its largest a_ij to that destination (generate_synthetic_data.py:145-157)
        colmax = torch.zeros(ncl, dtype=torch.float32, device=device).scatter_reduce_(0, col, aval, reduce="amax", include_self=True)
        is_max = aval == colmax[col]
        load.index_add_(0, dest[is_max], aval[is_max].double())
        deg_parts.append(deg)
        row_parts.append(dest)
        a_parts.append(aval)
        c_parts.append(-cval)
        del col, eps, u, colmax, is_max
    n_local = col_end - col_start
    ccol = torch.zeros(n_local + 1, dtype=torch.int64, device=device)
    if deg_parts:
        torch.cumsum(torch.cat(deg_parts), 0, out=ccol[1:])
    row = torch.cat(row_parts) if row_parts else torch.zeros(0, dtype=torch.int64, device=device)
    a = torch.cat(a_parts) if a_parts else torch.zeros(0, dtype=torch.float32, device=device)
    c = torch.cat(c_parts) if c_parts else torch.zeros(0, dtype=torch.float32, device=device)
    return SyntheticShard(ccol, row, a, c, load, m, col_start, col_end)


def capacity_vector(total_greedy_load: torch.Tensor, num_destinations: int, target_sparsity: float, seed: int, device):
    """b_j = rho_j * (load_j + 1e-8) with rho ~ U(0.5, 1) (generate_synthetic_data.py:159-162)."""
    dp = destination_params(num_destinations, target_sparsity, seed, device)
    return (dp["rho"] * (total_greedy_load.to(device) + 1e-8)).float()


def generate_movielens_shaped(n_users: int, n_movies: int, seed: int, device, mean_log_deg: float = 4.3, sigma_log_deg: float = 1.0):
    """configs[0]-shaped data (reference examples/movielens_matching/movies_lens_matching.py:49-116): one column per user,
    one row per movie, a == 1, c = -rating in {0.5, ..., 5}; column lengths heavy-tailed like ml-20m's (at least 20 ratings
    per user, mean ~140, a few users with thousands), movie popularity log-normal.  ml-20m itself is not available
    offline, so the ratings are drawn.  Returns (SyntheticShard, b) with one budget per movie: half the number of users whose
    best-rated movie it is, plus a little (so that popular rows bind)."""
    device = torch.device(device)
    gen = torch.Generator(device=device).manual_seed(seed)
    deg = torch.exp(mean_log_deg + sigma_log_deg * torch.randn(n_users, device=device, generator=gen)).to(torch.int64)
    deg.clamp_(min=20, max=min(n_movies, 9254))  # ml-20m: 20 .. 9254 ratings per user
    pop = torch.exp(1.2 * torch.randn(n_movies, device=device, generator=gen, dtype=torch.float64))
    cdf = torch.cumsum(pop / pop.sum(), 0)
    cdf[-1] = 1.0
    draws = (deg.double() * 1.15).to(torch.int64) + 4  # oversample: duplicates within a column are dropped
    total = int(draws.sum().item())
    col = torch.repeat_interleave(torch.arange(n_users, device=device), draws, output_size=total)
    dest = torch.searchsorted(cdf, torch.rand(total, device=device, dtype=torch.float64, generator=gen)).clamp_(max=n_movies - 1)
    key, _ = torch.sort(col * n_movies + dest)
    keep = torch.ones(total, dtype=torch.bool, device=device)
    keep[1:] = key[1:] != key[:-1]
    key = key[keep]
    col = torch.div(key, n_movies, rounding_mode="floor")
    row = key - col * n_movies
    # trim every column to its target length (entries are sorted by row within a column: drop a random subset instead of a tail)
    r = torch.rand(key.numel(), device=device, generator=gen)
    have = torch.bincount(col, minlength=n_users)
    frac = (deg.double() / have.clamp(min=1).double()).clamp(max=1.0).float()
    sel = r <= frac[col]
    col, row = col[sel], row[sel]
    lens = torch.bincount(col, minlength=n_users)
    ccol = torch.zeros(n_users + 1, dtype=torch.int64, device=device)
    torch.cumsum(lens, 0, out=ccol[1:])
    E = int(row.numel())
    probs = torch.tensor([.01, .03, .02, .07, .05, .2, .12, .28, .08, .14], device=device)
    rating = 0.5 * (1 + torch.multinomial(probs, E, replacement=True, generator=gen).float())
    c = -rating
    a = torch.ones(E, dtype=torch.float32, device=device)
    cmin = torch.full((n_users,), float("inf"), device=device).scatter_reduce_(0, col, c, reduce="amin", include_self=True)
    best = c == cmin[col]
    first_best = torch.ones(E, dtype=torch.bool, device=device)
    pos = torch.arange(E, device=device)
    first_pos = torch.full((n_users,), E, dtype=torch.int64, device=device).scatter_reduce_(0, col[best], pos[best], reduce="amin",
                                                                                            include_self=True)
    first_best = pos == first_pos[col]
    load = torch.zeros(n_movies, dtype=torch.float64, device=device)
    load.index_add_(0, row[first_best], torch.ones(int(first_best.sum().item()), dtype=torch.float64, device=device))
    b = (0.5 * load + 0.05).float()
    return SyntheticShard(ccol, row, a, c, load, n_movies, 0, n_users), b
