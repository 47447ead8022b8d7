#!/usr/bin/env python
"""Throughput of batched held-out inference (TopicInferencer.getSampledDistribution(inst, 100, 10, 10),
reference cmu_ron/TrainAndPredict.java:144) on the GPU vs the Mallet-port CPU oracle.
Model: the reference's own workload-A hyper-parameters (K=500, alphaSum=100, beta=1;
cmu_ron/TrainAndPredict.java:160) trained for a few sweeps on a synthetic corpus."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train-docs", type=int, default=20000); ap.add_argument("--vocab", type=int, default=20000)
    ap.add_argument("--heldout", type=int, default=20000); ap.add_argument("--cpu-docs", type=int, default=200)
    ap.add_argument("--topics", type=int, default=500)
    a = ap.parse_args()
    import ldagibbssampling_b200 as L
    from oracle import oracle as O
    K, V, alpha_sum, beta = a.topics, a.vocab, 100.0, 1.0
    dp, tok = O.gen_corpus(a.train_docs, V, 60.0, 50, 5)
    s = L.Sampler(K, V, alpha_sum, beta, seed=3)
    s.load_corpus(dp, tok); s.init_assignments(None); s.sweep(50)
    hd, htok = O.gen_corpus(a.heldout, V, 40.0, 50, 6)
    s.infer(hd[:101], htok[:hd[100]], 100, 10, 10, 1)  # warm
    t0 = time.perf_counter(); th = s.infer(hd, htok, 100, 10, 10, 7); t1 = time.perf_counter()
    gpu_docs_s = a.heldout / (t1 - t0); gpu_tok_iter_s = len(htok) * 100 / (t1 - t0)
    # CPU: the Mallet-port inferencer, one document at a time as the reference calls it
    m = O.MalletModel(K, alpha_sum, beta, seed=1); m.add_instances(dp, tok, V, z_init=s.assignments())
    n = a.cpu_docs; t0 = time.perf_counter()
    for d in range(n): m.infer(htok[hd[d]:hd[d + 1]], 100, 10, 10, seed=d)
    t1 = time.perf_counter()
    cpu_docs_s = n / (t1 - t0); cpu_tok_iter_s = int(hd[n]) * 100 / (t1 - t0)
    print(json.dumps({"what": "getSampledDistribution(inst, 100, 10, 10), batched on the GPU", "K": K, "V": V,
                      "heldout_docs": a.heldout, "heldout_tokens": int(len(htok)), "gpu_s": (len(htok) * 100) / gpu_tok_iter_s,
                      "gpu_docs_per_s": gpu_docs_s, "gpu_token_iterations_per_s": gpu_tok_iter_s,
                      "cpu_port_docs_per_s_1_thread": cpu_docs_s, "cpu_port_token_iterations_per_s": cpu_tok_iter_s,
                      "cpu_sample_docs": n, "theta_rows_sum_to_1": bool(np.allclose(th.sum(1), 1.0))}))

if __name__ == "__main__":
    main()
