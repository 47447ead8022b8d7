"""Synthetic corpora for bench.py at BASELINE.json's sizes (SURVEY.md §8(d) generator).

The reference ships no data (its inputs are laptop paths, reference
cmu_ron/TrainAndPredict.java:203-205), so every config is generated: k_true topics with
phi_k ∝ Dirichlet(0.01) x Zipf(1.07), theta_d ~ Dirichlet(0.1), L_d ~ round(LogNormal(mu, 0.6)).
Generation runs on the GPU with torch (plumbing, not the product) because C4 has 738 M tokens;
documents are generated in fixed chunks with per-chunk seeds, so a shard's tokens do not depend
on how many GPUs the corpus is split over. The small CPU sample the baselines use comes from
oracle/corpus_gen.c (same generative model).
"""
from __future__ import annotations

import math

import numpy as np
import torch

WORKLOADS = {
    # name: D, V, mean_len, K, k_true, seed                     (BASELINE.json configs)
    "c1": dict(D=10_000, V=5_000, mean_len=100.0, K=20, k_true=20, seed=1,
               desc="synthetic 10k docs, V=5k, ~1M tokens, K=20"),
    "c2": dict(D=300_000, V=102_000, mean_len=333.0, K=100, k_true=100, seed=2,
               desc="synthetic NYTimes-shaped: 300k docs, V=102k, ~100M tokens, K=100"),
    "c3": dict(D=300_000, V=102_000, mean_len=333.0, K=1000, k_true=200, seed=2,
               desc="synthetic NYTimes-shaped: 300k docs, V=102k, ~100M tokens, K=1000"),
    "c4": dict(D=8_200_000, V=141_000, mean_len=90.0, K=1000, k_true=200, seed=3,
               desc="synthetic PubMed-shaped: 8.2M docs, V=141k, ~738M tokens, K=1000"),
}

CHUNK_DOCS = 65536
SIGMA = 0.6


def doc_lengths(D: int, mean_len: float, seed: int, device) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed * 7919 + 1)
    mu = math.log(mean_len) - 0.5 * SIGMA * SIGMA
    x = torch.randn(D, generator=g, device=device, dtype=torch.float64)
    return torch.exp(mu + SIGMA * x).round().clamp_(1, 65535).to(torch.int64)


def _flat_cdf(p: torch.Tensor) -> torch.Tensor:
    """Row-normalised cdf of p with row r shifted into (r, r+1], flattened (sorted overall)."""
    s = p.sum(dim=1, keepdim=True)
    cdf = torch.cumsum(p / s, dim=1)
    cdf[:, -1] = 1.0
    cdf.clamp_(max=1.0)
    return (cdf + torch.arange(p.shape[0], device=p.device, dtype=p.dtype)[:, None]).reshape(-1)


def phi_flat_cdf(V: int, k_true: int, seed: int, device) -> torch.Tensor:
    g = torch.Generator(device=device).manual_seed(seed * 7919 + 2)
    conc = torch.full((k_true, V), 0.01, device=device, dtype=torch.float64)
    gam = torch._standard_gamma(conc, generator=g)
    zipf = torch.arange(1, V + 1, device=device, dtype=torch.float64).pow_(-1.07)
    p = gam * zipf
    dead = p.sum(dim=1) <= 0
    if dead.any():
        p[dead] = zipf
    return _flat_cdf(p)


def generate_docs(doc_begin: int, doc_end: int, lengths: torch.Tensor, phi_cdf: torch.Tensor, V: int,
                  k_true: int, seed: int, device) -> torch.Tensor:
    """Word ids (int32, on `device`) of documents [doc_begin, doc_end), document order."""
    out = []
    c0 = (doc_begin // CHUNK_DOCS) * CHUNK_DOCS
    for cb in range(c0, doc_end, CHUNK_DOCS):
        ce = min(cb + CHUNK_DOCS, lengths.numel())
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + cb // CHUNK_DOCS + 17)
        lens = lengths[cb:ce]
        c = ce - cb
        conc = torch.full((c, k_true), 0.1, device=device, dtype=torch.float64)
        gam = torch._standard_gamma(conc, generator=g)
        dead = gam.sum(dim=1) <= 0
        if dead.any():
            gam[dead, 0] = 1.0
        theta_cdf = _flat_cdf(gam)
        n = int(lens.sum().item())
        doc_local = torch.repeat_interleave(torch.arange(c, device=device), lens, output_size=n)
        u = torch.rand(2, n, generator=g, device=device, dtype=torch.float64)
        idx = torch.searchsorted(theta_cdf, doc_local.to(torch.float64) + u[0], right=True)
        topic = (idx - doc_local * k_true).clamp_(0, k_true - 1)
        widx = torch.searchsorted(phi_cdf, topic.to(torch.float64) + u[1], right=True)
        word = (widx - topic * V).clamp_(0, V - 1).to(torch.int32)
        # keep only this call's documents
        lo = max(doc_begin, cb) - cb
        hi = min(doc_end, ce) - cb
        if lo > 0 or hi < c:
            cum = torch.zeros(c + 1, dtype=torch.int64, device=device)
            cum[1:] = torch.cumsum(lens, 0)
            word = word[int(cum[lo].item()):int(cum[hi].item())]
        out.append(word)
    if not out:
        return torch.empty(0, dtype=torch.int32, device=device)
    return torch.cat(out)


def cpu_sample(workload: str, num_docs: int):
    """Bounded CPU sample of a workload's shape (oracle generator): doc_ptr, tok_word, V, K."""
    from oracle import oracle as O
    w = WORKLOADS[workload]
    dp, tok = O.gen_corpus(num_docs, w["V"], w["mean_len"], w["k_true"], w["seed"])
    return dp, tok, w["V"], w["K"]
