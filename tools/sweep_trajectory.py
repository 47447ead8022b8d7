#!/usr/bin/env python
"""Per-sweep trajectory of the sampling kernel on a workload shape: device time, tokens/s,
mean non-zero doc topics, moved / prior-bucket fractions and LL/token as the chain mixes."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4"); ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--topics", type=int, default=0)
    ap.add_argument("--sweeps", type=int, default=150); ap.add_argument("--every", type=int, default=10)
    ap.add_argument("--mode", default="live")
    ap.add_argument("--sort-words", action="store_true", help="order every document's tokens by word id before loading "
                    "(what an in-document word order would do to the n_wk gathers; the library keeps the caller's order)")
    a = ap.parse_args()
    import torch, bench_corpus as BC, ldagibbssampling_b200 as L
    dev = torch.device("cuda", 0)
    w = BC.WORKLOADS[a.workload]; D = a.docs or w["D"]; V = w["V"]; K = a.topics or w["K"]
    lengths = BC.doc_lengths(D, w["mean_len"], w["seed"], dev)
    dp = np.zeros(D + 1, np.int64); dp[1:] = torch.cumsum(lengths, 0).cpu().numpy()
    phi = BC.phi_flat_cdf(V, w["k_true"], w["seed"], dev)
    words = BC.generate_docs(0, D, lengths, phi, V, w["k_true"], w["seed"], dev)
    if a.sort_words:
        doc_id = torch.repeat_interleave(torch.arange(D, device=dev), lengths.to(torch.int64))
        key, _ = torch.sort(doc_id * V + words.to(torch.int64))
        words = (key % V).to(words.dtype); del key, doc_id
        print(json.dumps({"repeat_frac": float((words[1:] == words[:-1]).float().mean())}))
    words = words.cpu().numpy()
    del phi; torch.cuda.empty_cache()
    s = L.Sampler(K, V, 0.1 * K, 0.01, seed=1, mode=L.MODE_LIVE if a.mode == "live" else L.MODE_DEFERRED)
    s.load_corpus(dp, words); s.init_assignments(None)
    N = len(words); done = 0
    print(json.dumps({"workload": a.workload, "docs": D, "tokens": N, "K": K, "mode": a.mode, "stats0": {k: v for k, v in s.stats().items() if "slot" in k or "ctas" in k or "long" in k or "smem" in k}}))
    while done < a.sweeps:
        s.reset_stats(); s.sweep(a.every); done += a.every
        st = s.stats(); n = N * st["cum_sweeps"]
        print(json.dumps({"sweep": done, "sample_ms": st["cum_sample_ms"] / st["cum_sweeps"],
                          "tables_ms": st["cum_tables_ms"] / st["cum_sweeps"],
                          "tok_per_s": N / (st["cum_sample_ms"] / st["cum_sweeps"] / 1e3),
                          "kd": st["cum_doc_topics"] / n, "moved": st["cum_tokens_moved"] / n,
                          "prior": st["cum_prior_bucket"] / n, "ll_per_token": s.loglik() / N}), flush=True)

if __name__ == "__main__":
    main()
