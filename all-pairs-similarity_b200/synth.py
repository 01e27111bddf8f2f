"""Synthetic power-law (Zipf) sparse TF-IDF-like vectors -- the workload generator G(N, D, z, s, seed)
of SURVEY.md 8(d).  Harness code (bench.py, tests); not on the scoring path.

Model: a document is k ~ Poisson(lambda) draws (with replacement) from a Zipf-Mandelbrot law over
ranks, rank -> dimension through a seeded permutation (hashed-TF-like: ids carry no frequency
order).  tf = multiplicity of the draw, idf(d) = ln((N+1)/(df(d)+1)) from the realised document
frequencies (Spark 1.2 IDF, PreprocessWithTFIDF.scala:50-51), value = tf*idf, then L2-normalised as
LoadGenerator.scala:35-37 does.  lambda is calibrated so that the expected number of DISTINCT dims
per vector is `nnz_mean`.  A fraction `dup_frac` of the vectors are near-duplicates of an earlier
vector (5-25 % of the draws replaced, weights jittered +-10 %), mirroring the duplicated mails of
data/maildir_small (sent vs sent_items), so that the answer set is not empty.

Runs on CPU (tests) or CUDA (bench at full size); everything is generated from `seed` and the
resulting arrays are what both the oracle and the GPU path consume, so parity never depends on the
two devices' RNG streams agreeing.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class SparseBatch:
    """CSR rows: indptr int64 [n+1], indices int32 (ascending per row), values float64."""
    indptr: torch.Tensor
    indices: torch.Tensor
    values: torch.Tensor
    dim: int

    @property
    def n(self):
        return self.indptr.numel() - 1

    @property
    def nnz(self):
        return self.indices.numel()

    def rows(self, lo, hi):
        a, b = int(self.indptr[lo]), int(self.indptr[hi])
        return SparseBatch(self.indptr[lo:hi + 1] - self.indptr[lo], self.indices[a:b], self.values[a:b], self.dim)

    def numpy(self):
        return (self.indptr.cpu().numpy(), self.indices.cpu().numpy(), self.values.cpu().numpy())

    def to(self, device):
        return SparseBatch(self.indptr.to(device), self.indices.to(device), self.values.to(device), self.dim)

    def pin(self):
        return SparseBatch(self.indptr.cpu().pin_memory(), self.indices.cpu().pin_memory(), self.values.cpu().pin_memory(), self.dim)


def zipf_pmf(D, s=1.0, q=0.0):
    r = np.arange(1, D + 1, dtype=np.float64)
    p = 1.0 / np.power(r + q, s)
    return p / p.sum()


def calibrate_lambda(pmf, nnz_mean):
    """lambda with sum_r (1 - exp(-lambda p_r)) = nnz_mean (Poissonised draws => independent dims)."""
    lo, hi = float(nnz_mean), float(nnz_mean) * 64.0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if np.sum(-np.expm1(-mid * pmf)) < nnz_mean:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def generate(N, D, nnz_mean, s=1.0, q=0.0, seed=0, dup_frac=0.10, device="cpu", chunk=1 << 17,
             heavy_tail=False, k_min=4) -> SparseBatch:
    dev = torch.device(device)
    pmf = zipf_pmf(D, s, q)
    lam = calibrate_lambda(pmf, nnz_mean)
    cdf = torch.from_numpy(np.cumsum(pmf)).to(dev)
    cdf[-1] = 1.0
    perm = torch.from_numpy(np.random.RandomState(seed ^ 0x5EED).permutation(D).astype(np.int64)).to(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))

    # draws per vector
    if heavy_tail:      # lognormal, mean lam, sigma 1, clipped (config 5)
        sigma = 1.0
        mu = math.log(lam) - 0.5 * sigma * sigma
        k = torch.exp(mu + sigma * torch.randn(N, generator=g, device=dev, dtype=torch.float64)).round().clamp(8, 4096).long()
    else:
        k = torch.poisson(torch.full((N,), lam, device=dev, dtype=torch.float64), generator=g).long().clamp(min=k_min)
    kmax = int(k.max().item())

    is_dup = torch.rand(N, generator=g, device=dev) < dup_frac
    is_dup[0] = False
    src = (torch.rand(N, generator=g, device=dev, dtype=torch.float64) * torch.arange(N, device=dev, dtype=torch.float64)).long()
    src = torch.where(is_dup, src, torch.arange(N, device=dev))
    k_eff = k[src]                                   # a near-duplicate keeps its source's length
    replace_p = 0.05 + 0.20 * torch.rand(N, generator=g, device=dev)

    # own draws for every vector, chunked (per-chunk seeds: keep `chunk` fixed for reproducibility)
    rows_l, dims_l, tf_l = [], [], []
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        n = hi - lo
        gc = torch.Generator(device=dev)
        gc.manual_seed(int(seed) * 1000003 + lo + 17)

        u = torch.rand((n, kmax), generator=gc, device=dev, dtype=torch.float64)
        own = perm[torch.searchsorted(cdf, u.reshape(-1)).clamp(max=D - 1)].reshape(n, kmax)
        del u
        rows_l.append(own)
    own_all = torch.cat(rows_l, 0) if len(rows_l) > 1 else rows_l[0]     # [N, kmax] int64
    del rows_l

    out_ptr = [torch.zeros(1, dtype=torch.int64, device=dev)]
    df = torch.zeros(D, dtype=torch.int64, device=dev)
    base = 0
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        n = hi - lo
        gc = torch.Generator(device=dev)
        gc.manual_seed(int(seed) * 7919 + lo + 3)
        sel = own_all[src[lo:hi]]                                        # source draws (self for originals)
        rep = torch.rand((n, kmax), generator=gc, device=dev) < replace_p[lo:hi, None]
        rep &= is_dup[lo:hi, None]
        mixed = torch.where(rep, own_all[lo:hi], sel)
        valid = torch.arange(kmax, device=dev)[None, :] < k_eff[lo:hi, None]
        key = torch.arange(n, device=dev, dtype=torch.int64)[:, None] * D + mixed
        key = key[valid]
        key, _ = torch.sort(key)
        uk, cnt = torch.unique_consecutive(key, return_counts=True)
        r = uk // D
        d = uk - r * D
        dims_l.append(d.to(torch.int32))
        tf_l.append(cnt.to(torch.int16))
        rc = torch.bincount(r, minlength=n)
        out_ptr.append(base + torch.cumsum(rc, 0))
        base += int(uk.numel())
        df += torch.bincount(d, minlength=D)
    del own_all
    indptr = torch.cat(out_ptr)
    indices = torch.cat(dims_l)
    tf = torch.cat(tf_l).to(torch.float64)
    del dims_l, tf_l

    idf = torch.from_numpy(np.log((N + 1.0) / (df.cpu().numpy().astype(np.float64) + 1.0))).to(dev)
    gj = torch.Generator(device=dev)
    gj.manual_seed(int(seed) + 99991)
    jitter = 0.9 + 0.2 * torch.rand(indices.numel(), generator=gj, device=dev, dtype=torch.float64)
    row_of = torch.repeat_interleave(torch.arange(N, device=dev), indptr[1:] - indptr[:-1])
    jitter = torch.where(is_dup[row_of], jitter, torch.ones_like(jitter))
    val = tf * idf[indices.long()] * jitter
    sq = torch.zeros(N, dtype=torch.float64, device=dev).index_add_(0, row_of, val * val)
    nrm = torch.sqrt(sq)
    nrm = torch.where(nrm > 0, nrm, torch.ones_like(nrm))
    val = val / nrm[row_of]
    return SparseBatch(indptr, indices, val, D)


class FlatShard:
    """Structure (rows, dims, term frequencies) of a generated shard before IDF weighting; lets the
    ranks of a multi-GPU job generate their own shards and all-reduce the document frequencies."""

    def __init__(self, indptr, indices, tf, dup_mask_rows, D, seed):
        self.indptr, self.indices, self.tf, self.is_dup, self.D, self.seed = indptr, indices, tf, dup_mask_rows, D, seed

    @property
    def n(self):
        return self.indptr.numel() - 1

    def df(self):
        return torch.bincount(self.indices.long(), minlength=self.D)

    def finalize(self, idf) -> SparseBatch:
        """value = tf * idf(d) (jittered +-10 % on planted near-duplicates), L2-normalised."""
        dev = self.indices.device
        N = self.n
        row_of = torch.repeat_interleave(torch.arange(N, device=dev), self.indptr[1:] - self.indptr[:-1])
        g = torch.Generator(device=dev); g.manual_seed(int(self.seed) + 99991)
        jitter = 0.9 + 0.2 * torch.rand(self.indices.numel(), generator=g, device=dev, dtype=torch.float64)
        jitter = torch.where(self.is_dup[row_of], jitter, torch.ones_like(jitter))
        val = self.tf.to(torch.float64) * idf.to(dev)[self.indices.long()] * jitter
        sq = torch.zeros(N, dtype=torch.float64, device=dev).index_add_(0, row_of, val * val)
        nrm = torch.sqrt(sq)
        nrm = torch.where(nrm > 0, nrm, torch.ones_like(nrm))
        return SparseBatch(self.indptr, self.indices, val / nrm[row_of], self.D)


def generate_flat(N, D, nnz_mean, s=1.0, q=0.0, seed=0, dup_frac=0.10, device="cpu", heavy_tail=False, k_min=4) -> FlatShard:
    """Variable-length ("flat") variant of generate() for heavy-tailed lengths (config C5) and per-rank
    shards: no [N, kmax] matrix.  A planted near-duplicate is a copy of an earlier row OF THE SAME CALL with
    ~10 % of its draws dropped (its weights are jittered in finalize())."""
    dev = torch.device(device)
    pmf = zipf_pmf(D, s, q)
    lam = calibrate_lambda(pmf, nnz_mean)
    cdf = torch.from_numpy(np.cumsum(pmf)).to(dev); cdf[-1] = 1.0
    perm = torch.from_numpy(np.random.RandomState(20260105 ^ 0x5EED).permutation(D).astype(np.int64)).to(dev)   # shared by all shards
    g = torch.Generator(device=dev); g.manual_seed(int(seed))
    if heavy_tail:
        sigma = 1.0
        mu = math.log(lam) - 0.5 * sigma * sigma
        k = torch.exp(mu + sigma * torch.randn(N, generator=g, device=dev, dtype=torch.float64)).round().clamp(8, 4096).long()
    else:
        k = torch.poisson(torch.full((N,), lam, device=dev, dtype=torch.float64), generator=g).long().clamp(min=k_min)
    is_dup = torch.rand(N, generator=g, device=dev) < dup_frac
    is_dup[0] = False
    src = (torch.rand(N, generator=g, device=dev, dtype=torch.float64) * torch.arange(N, device=dev, dtype=torch.float64)).long()
    src = torch.where(is_dup, src, torch.arange(N, device=dev))
    start = torch.cumsum(k, 0) - k                       # draw ranges of the originals
    total = int(k.sum())
    u = torch.rand(total, generator=g, device=dev, dtype=torch.float64)
    draws = perm[torch.searchsorted(cdf, u).clamp(max=D - 1)]
    del u
    # every row reads the draw range of its source row (itself for originals)
    k_eff = k[src]
    row_of = torch.repeat_interleave(torch.arange(N, device=dev), k_eff)
    off = torch.arange(int(k_eff.sum()), device=dev) - torch.repeat_interleave(torch.cumsum(k_eff, 0) - k_eff, k_eff)
    dims = draws[start[src][row_of] + off]
    keep = ~(is_dup[row_of] & (torch.rand(dims.numel(), generator=g, device=dev) < 0.10))
    key = (row_of * D + dims)[keep]
    del draws, dims, off, row_of
    key, _ = torch.sort(key)
    uk, cnt = torch.unique_consecutive(key, return_counts=True)
    del key
    r = uk // D
    indices = (uk - r * D).to(torch.int32)
    indptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(torch.bincount(r, minlength=N), 0)])
    return FlatShard(indptr, indices, cnt.to(torch.int16), is_dup, D, seed)


def idf_from_df(df, n_docs):
    """ln((N+1)/(df+1)) with libm's log on the host (Spark 1.2 IDF; see etl.idf_fit for why not numpy.log)."""
    d = df.cpu().numpy().astype(np.float64)
    return torch.from_numpy(np.log((n_docs + 1.0) / (d + 1.0)))


# the named workloads of BASELINE.json / SURVEY.md 8(d)
CONFIGS = {
    "C2": dict(N=100_000, D=1 << 16, nnz_mean=50, s=1.0, seed=20260102, threshold=0.8, batch=4096),
    "C3": dict(N=1_000_000, D=1 << 18, nnz_mean=100, s=1.0, seed=20260103, threshold=0.7, batch=16384),
    "C4": dict(N=5_000_000, D=1 << 18, nnz_mean=100, s=1.0, seed=20260104, threshold=0.7, batch=10000),
    "C5": dict(N=20_000_000, D=1 << 20, nnz_mean=200, s=1.0, seed=20260105, threshold=0.5, batch=16384, heavy_tail=True),
}


def generate_config(name, device="cpu", N=None):
    c = CONFIGS[name]
    return generate(N or c["N"], c["D"], c["nnz_mean"], s=c["s"], seed=c["seed"], device=device,
                    heavy_tail=c.get("heavy_tail", False))
