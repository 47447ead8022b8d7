#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ (run from the repo root).

The reference (qianjinding/LDAGibbsSampling) ships no tests, fixtures or recorded outputs for
this path and cannot be executed here (Java + un-vendored Mallet jar, no JVM), so these vectors
are OUR goldens (SURVEY.md §8(c) "own golden vectors to create"):

  tiny_conditionals.json   D=8, V=12, K=4 corpus with the exact collapsed conditional of every
                           token computed HERE in rational arithmetic (fractions.Fraction),
                           independently of the C oracle: pins oracle_exact_conditional and, through
                           a fine uniform grid, the spec sampler's bucket boundaries.
  frozen_triples.npz       (corpus, counts via z, uniforms, expected z) triples of the sampling spec
                           for several K: regression pin for the oracle, parity target for the GPU.
  c1_ll_trajectory.json    BASELINE.json config 1 (10k docs, V=5k, ~1M tokens, K=20): LL/token of
                           the Mallet-faithful oracle at sweeps 1,10,50,100,500 for 3 seeds, from a
                           shared Philox init (the GPU chain must land within 1 %).
"""
import json
import os
import sys
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import oracle as O  # noqa: E402


def tiny():
    rng = np.random.default_rng(20261018)
    D, V, K = 8, 12, 4
    alpha = [Fraction(1, 10), Fraction(1, 5), Fraction(1, 10), Fraction(3, 10)]
    beta = Fraction(1, 100)
    lens = [3, 7, 1, 5, 9, 2, 6, 4]
    doc_ptr = np.concatenate([[0], np.cumsum(lens)]).astype(int)
    tok = rng.integers(0, V, doc_ptr[-1])
    z = rng.integers(0, K, doc_ptr[-1])
    nwk = np.zeros((V, K), int)
    nk = np.zeros(K, int)
    for w, k in zip(tok, z):
        nwk[w, k] += 1
        nk[k] += 1
    cond = []
    for d in range(D):
        ndk = np.bincount(z[doc_ptr[d]:doc_ptr[d + 1]], minlength=K)
        for i in range(doc_ptr[d], doc_ptr[d + 1]):
            w, o = int(tok[i]), int(z[i])
            p = []
            for k in range(K):
                ex = 1 if k == o else 0
                # n_k is NOT decremented (sampling spec: sweep-start snapshot of n_k)
                p.append((Fraction(int(nwk[w, k]) - ex) + beta) * (Fraction(int(ndk[k]) - ex) + alpha[k]) /
                         (Fraction(int(nk[k])) + V * beta))
            tot = sum(p)
            cond.append([float(x / tot) for x in p])
    out = {"D": D, "V": V, "K": K, "alpha": [float(a) for a in alpha], "beta": float(beta),
           "doc_ptr": doc_ptr.tolist(), "tok_word": tok.tolist(), "z": z.tolist(), "conditional": cond}
    json.dump(out, open(os.path.join(HERE, "tiny_conditionals.json"), "w"), indent=0)


def frozen():
    out = {}
    for name, (D, V, mean_len, kt, K) in {"k4": (40, 30, 12.0, 4, 4), "k20": (60, 80, 30.0, 8, 20),
                                          "k100": (40, 120, 90.0, 20, 100), "k1500": (12, 60, 220.0, 10, 1500),
                                          "k3000": (6, 60, 600.0, 10, 3000)}.items():  # k3000: rows wider than 8 tiles
        dp, tok = O.gen_corpus(D, V, mean_len, kt, 77)
        z = O.init_z(len(tok), K, 31)
        z = O.spec_sweeps(dp, tok, z, V, K, 0.1, 0.01, 31, 1, 2)  # a non-uniform snapshot
        u = np.random.default_rng(5).random(len(tok), dtype=np.float32)
        out[name + "_meta"] = np.array([D, V, K], np.int64)
        out[name + "_doc_ptr"] = dp
        out[name + "_tok"] = tok
        out[name + "_z"] = z
        out[name + "_u"] = u
        out[name + "_expected_u"] = O.spec_frozen(dp, tok, z, V, K, 0.1, 0.01, 31, 1, uniforms=u)
        out[name + "_expected_philox"] = O.spec_frozen(dp, tok, z, V, K, 0.1, 0.01, 31, 9)
    np.savez_compressed(os.path.join(HERE, "frozen_triples.npz"), **out)


def c1():
    D, V, K = 10000, 5000, 20
    dp, tok = O.gen_corpus(D, V, 100.0, 20, 1)
    z0 = O.init_z(len(tok), K, 7)
    marks = [1, 10, 50, 100, 500]
    traj = {}
    for seed in (1, 2, 3):
        m = O.MalletModel(K, 0.1 * K, 0.01, seed=seed)
        m.add_instances(dp, tok, V, z_init=z0)
        done, row = 0, []
        for mk in marks:
            m.estimate(mk - done)
            done = mk
            row.append(m.model_log_likelihood() / len(tok))
        traj[str(seed)] = row
        m.close()
    json.dump({"workload": "c1", "D": D, "V": V, "K": K, "alpha_k": 0.1, "beta": 0.01, "corpus_seed": 1,
               "mean_len": 100.0, "k_true": 20, "init": "oracle.init_z(N, K, seed=7)", "tokens": int(len(tok)),
               "sweeps": marks, "mallet_ll_per_token": traj,
               "ll_init": O.loglik(dp, tok, z0, V, K, 0.1, 0.01, True) / len(tok)},
              open(os.path.join(HERE, "c1_ll_trajectory.json"), "w"), indent=1)


def _c4s_run(args):
    threads, seed, marks = args
    import bench_corpus as BC
    dp, tok, V, K = BC.cpu_sample("c4", 20000)
    z0 = O.init_z(len(tok), K, 7)
    m = O.MalletModel(K, 0.1 * K, 0.01, seed=seed, threads=threads)
    m.add_instances(dp, tok, V, z_init=z0)
    done, row = 0, []
    for mk in marks:
        m.estimate(mk - done)
        done = mk
        row.append(m.model_log_likelihood() / len(tok))
    m.close()
    return threads, seed, row


def c4s():
    """C4-SHAPED sample both sides can run (20 000 documents of BASELINE.json config 4's generator,
    V = 141 000, K = 1000, alpha_k = 0.1, beta = 0.01 - the sample bench.py's CPU leg uses): LL/token of
    the Mallet-faithful oracle with 1, 2 and 4 worker threads at sweeps 25, 50, 100, 200 for 3 seeds
    each, from a shared Philox init. The GPU chains (LIVE, DEFERRED, 2 and 4 shards) must land
    within 1 % of every seed of the matching row (tests/test_gpu_ll_parity.py)."""
    import multiprocessing as mp
    import bench_corpus as BC
    marks = [25, 50, 100, 200]
    dp, tok, V, K = BC.cpu_sample("c4", 20000)
    z0 = O.init_z(len(tok), K, 7)
    out = {"workload": "c4 sample", "D": 20000, "V": V, "K": K, "alpha_k": 0.1, "beta": 0.01,
           "init": "oracle.init_z(N, K, seed=7)", "tokens": int(len(tok)), "sweeps": marks,
           "ll_init": O.loglik(dp, tok, z0, V, K, 0.1, 0.01, True) / len(tok), "mallet_ll_per_token": {}}
    for threads in (1, 2, 4):
        with mp.Pool(3 if threads < 4 else 2) as pool:
            res = pool.map(_c4s_run, [(threads, seed, marks) for seed in (1, 2, 3)])
        out["mallet_ll_per_token"][str(threads)] = {str(seed): row for _, seed, row in res}
        json.dump(out, open(os.path.join(HERE, "c4s_ll_trajectory.json"), "w"), indent=1)


if __name__ == "__main__":
    O.build()
    which = sys.argv[1:] or ["tiny", "frozen", "c1", "c4s"]
    for w in which:
        {"tiny": tiny, "frozen": frozen, "c1": c1, "c4s": c4s}[w]()
        print("wrote", w)
