#!/usr/bin/env python
"""Unloaded per-token latency of the sampling kernel: ONE document (one warp, nothing else on the
GPU), so a sweep's device time / tokens is the length of the per-token dependent chain."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=20000)
    ap.add_argument("--sweeps", type=int, default=10)
    a = ap.parse_args()
    import ldagibbssampling_b200 as L
    rng = np.random.default_rng(1)
    for K, V, alpha_k in ((20, 5000, 0.1), (1000, 141000, 0.1), (1000, 141000, 0.001)):
        for mode in (L.MODE_LIVE, L.MODE_DEFERRED):
            dp = np.array([0, a.tokens], np.int64)
            words = rng.integers(0, V, a.tokens).astype(np.int32)
            s = L.Sampler(K, V, alpha_k * K, 0.01, seed=1, mode=mode)
            s.load_corpus(dp, words)
            s.init_assignments(None)
            s.sweep(3)
            s.reset_stats()
            s.sweep(a.sweeps)
            st = s.stats()
            n = a.tokens * st["cum_sweeps"]
            ms = st["cum_sample_ms"] / st["cum_sweeps"]
            print(json.dumps({"K": K, "alpha_k": alpha_k, "mode": "live" if mode == L.MODE_LIVE else "deferred",
                              "us_per_token": 1e3 * ms / a.tokens, "kd": st["cum_doc_topics"] / n,
                              "moved": st["cum_tokens_moved"] / n, "prior": st["cum_prior_bucket"] / n}), flush=True)
            del s


if __name__ == "__main__":
    main()
