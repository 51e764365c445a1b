"""Synthetic document pairs for tests and bench.py (SURVEY.md §8d generator).

A latent sequence of unit vectors U ~ N(0,1)^(n_units x D) is segmented independently on each
side into runs of 1-3 units (P(len>=2)=0.35, P(len=3)=0.175), 5 % of the segments are dropped, and
the overlap tensor is vecs[j, e] = normalise(sum of the units of segments e-j..e) + 0.25/sqrt(D) N(0,1),
zero rows where e < j — the (K, N, D) fp32 layout of make_doc_embedding
(svecalign/utils/embedding_utils.py:135-203).  The mix of 1-1 / 1-2 / 2-1 / 2-2 / deletions it
produces has on-path DP decision margins >= 1e-3.

`synth_pair` is numpy (seed-stable, used for fixtures/tests); `synth_pair_torch` draws on the
given torch device (bench: data born in HBM).
"""
from math import ceil

import numpy as np

P_LEN2, P_LEN3, P_DROP, NOISE = 0.35, 0.175, 0.05, 0.25


def _segments(rng_uniform, n, n_units):
    """start/end unit index of the first n kept segments (numpy arrays), or None if short."""
    m = int(n / (1.0 - P_DROP) * 1.15) + 16
    u, drop = rng_uniform(m), rng_uniform(m)
    length = 1 + (u < P_LEN2) + (u < P_LEN3)
    end = np.cumsum(length)
    start = end - length
    keep = (drop >= P_DROP) & (end <= n_units)
    idx = np.nonzero(keep)[0]
    if idx.size < n:
        return None
    idx = idx[:n]
    return start[idx], end[idx]


def synth_pair(n0, n1, k, dim=1024, seed=0):
    """-> (vecs0 (k,n0,dim), vecs1 (k,n1,dim)) float32, C-contiguous, unnormalised-noisy."""
    rng = np.random.default_rng(seed)
    n_units = int(ceil(max(n0, n1, 1) * 1.9)) + 32
    units = rng.standard_normal((n_units, dim), dtype=np.float32)
    prefix = np.zeros((n_units + 1, dim), dtype=np.float64)
    np.cumsum(units, axis=0, dtype=np.float64, out=prefix[1:])
    out = []
    for n in (n0, n1):
        seg = None
        while seg is None:
            seg = _segments(lambda m: rng.random(m), n, n_units)
        start, end = seg
        v = np.zeros((k, n, dim), dtype=np.float32)
        for j in range(min(k, n)):
            e = np.arange(j, n)
            # units of segments e-j .. e: kept segments only, so sum segment by segment
            tot = np.zeros((n - j, dim), dtype=np.float64)
            for d in range(j + 1):
                tot += prefix[end[e - d]] - prefix[start[e - d]]
            tot /= np.linalg.norm(tot, axis=1, keepdims=True) + 1e-12
            tot += (NOISE / np.sqrt(dim)) * rng.standard_normal((n - j, dim))
            v[j, j:] = tot.astype(np.float32)
        out.append(v)
    return out[0], out[1]


def synth_pair_torch(n0, n1, k, dim=1024, seed=0, device="cuda", out0=None, out1=None):
    """Same model drawn with torch on `device`; writes into out0/out1 ((k,n,dim) fp32) if given."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    host = np.random.default_rng(int(seed))
    n_units = int(ceil(max(n0, n1, 1) * 1.9)) + 32
    units = torch.randn((n_units, dim), generator=g, device=device, dtype=torch.float32)
    prefix = torch.zeros((n_units + 1, dim), device=device, dtype=torch.float64)
    torch.cumsum(units.double(), dim=0, out=prefix[1:])
    res = []
    for n, dst in ((n0, out0), (n1, out1)):
        seg = None
        while seg is None:
            seg = _segments(lambda m: host.random(m), n, n_units)
        start = torch.from_numpy(seg[0]).to(device)
        end = torch.from_numpy(seg[1]).to(device)
        v = dst if dst is not None else torch.empty((k, n, dim), device=device, dtype=torch.float32)
        v.zero_()
        for j in range(min(k, n)):
            e = torch.arange(j, n, device=device)
            tot = torch.zeros((n - j, dim), device=device, dtype=torch.float64)
            for d in range(j + 1):
                tot += prefix[end[e - d]] - prefix[start[e - d]]
            tot /= tot.norm(dim=1, keepdim=True) + 1e-12
            tot += (NOISE / np.sqrt(dim)) * torch.randn((n - j, dim), generator=g, device=device, dtype=torch.float64)
            v[j, j:] = tot.float()
        res.append(v)
    return res[0], res[1]


def batch_sizes(npairs, lo=200, hi=800, seed=1234):
    """BASELINE config 4 lengths: N0 ~ U{lo..hi}, N1 = clip(N0 * U(0.9, 1.1), lo, hi)."""
    rng = np.random.default_rng(seed)
    n0 = rng.integers(lo, hi + 1, size=npairs)
    n1 = np.clip(np.rint(n0 * rng.uniform(0.9, 1.1, size=npairs)), lo, hi).astype(np.int64)
    return n0.astype(np.int64), n1
