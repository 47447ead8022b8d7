#!/usr/bin/env python
"""A small pass through every kernel of libb200lda.so for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py
Register classes 0-2, the wide class (shared-memory rows) and a row that moves from registers to shared
memory mid-visit, LIVE with table refreshers, DEFERRED, frozen, inference, log-likelihood, hyper-parameter
statistics, state blob, n_dk CSR, invariants."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ldagibbssampling_b200 as L

rng = np.random.default_rng(1)
V, K = 300, 600
lens = np.concatenate([rng.integers(5, 90, 120), rng.integers(100, 250, 30), rng.integers(300, 700, 6), [1200]]).astype(np.int64)
dp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
tok = rng.integers(0, V, int(dp[-1])).astype(np.int32)
for mode, refresh in ((L.MODE_LIVE, 4), (L.MODE_DEFERRED, 0)):
    s = L.Sampler(K, V, 0.5 * K, 0.01, seed=3, mode=mode, table_refresh=refresh)
    s.load_corpus(dp, tok)
    s.init_assignments(None)
    s.sweep(3)
    s.sample_frozen(None, sweep=7)
    s.infer(dp[:9], tok[:dp[8]], iterations=12, thinning=3, burn_in=3, seed=5)
    s.loglik()
    s.hyper_begin(int(lens.max()) + 1); s.hyper_collect(); s.hyper_get()
    blob = s.get_state(); s.set_state(blob); s.sweep(1)
    rp, t, c = s.ndk_csr()
    inv = s.check_invariants()
    assert inv[0] == inv[1] == inv[3] == len(tok) and inv[2] == 0, inv
    assert c.sum() == len(tok)
    s.theta(); s.phi(); s.close()
print("sanitize_small: ok")
