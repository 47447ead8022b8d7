#!/usr/bin/env python
"""Wall-clock breakdown of the end-to-end step (host buffers in, host buffers out) per C-ABI call."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4"); ap.add_argument("--docs", type=int, default=2_000_000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch, bench_corpus as BC, ldagibbssampling_b200 as L
    dev = torch.device("cuda", 0)
    w = BC.WORKLOADS[a.workload]; D = a.docs or w["D"]; V = w["V"]; K = w["K"]
    lengths = BC.doc_lengths(D, w["mean_len"], w["seed"], dev)
    dp = torch.zeros(D + 1, dtype=torch.int64); dp[1:] = torch.cumsum(lengths, 0).cpu()
    phi = BC.phi_flat_cdf(V, w["k_true"], w["seed"], dev)
    words = BC.generate_docs(0, D, lengths, phi, V, w["k_true"], w["seed"], dev)
    h_dp = dp.pin_memory(); h_w = torch.empty(words.numel(), dtype=torch.int32, pin_memory=True); h_w.copy_(words)
    h_z = torch.empty(words.numel(), dtype=torch.int32, pin_memory=True)
    N = words.numel(); del phi, words; torch.cuda.empty_cache()
    s = L.Sampler(K, V, 0.1 * K, 0.01, seed=1)
    s.load_corpus_raw(D, h_dp.data_ptr(), h_w.data_ptr(), N); s.init_assignments(None); s.assignments_raw(h_z.data_ptr())
    for r in range(a.reps):
        t = [time.perf_counter()]
        s.load_corpus_raw(D, h_dp.data_ptr(), h_w.data_ptr(), N); t.append(time.perf_counter())
        s.init_assignments_raw(h_z.data_ptr()); t.append(time.perf_counter())
        s.sweep(1); t.append(time.perf_counter())
        s.assignments_raw(h_z.data_ptr()); t.append(time.perf_counter())
        d = np.diff(t) * 1e3
        print(json.dumps({"tokens": N, "load_corpus_ms": d[0], "init_assignments_ms": d[1], "sweep_ms": d[2],
                          "get_assignments_ms": d[3], "total_ms": float(d.sum()), "e2e_tok_per_s": N / d.sum() * 1e3}), flush=True)

if __name__ == "__main__":
    main()
