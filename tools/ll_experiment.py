#!/usr/bin/env python
"""LL/token of the LIVE chain on the C4-shaped 20k-document sample (tests/golden/c4s_ll_trajectory.json)
under the experiment knobs B200LDA_TABLE_REFRESH / B200LDA_MAX_CTAS: what each source of staleness costs."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import oracle as O
import bench_corpus as BC
import ldagibbssampling_b200 as L
g = json.load(open(os.path.join(ROOT, "tests/golden/c4s_ll_trajectory.json")))
dp, tok, V, K = BC.cpu_sample("c4", g["D"])
z0 = O.init_z(len(tok), K, 7)
marks = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "25,50").split(",")]
mode = L.MODE_DEFERRED if os.environ.get("MODE") == "deferred" else L.MODE_LIVE
s = L.Sampler(K, V, 0.1 * K, 0.01, seed=7, mode=mode, table_refresh=0)
s.load_corpus(dp, tok); s.init_assignments(z0)
done, out = 0, []
for m in marks:
    s.sweep(m - done); done = m
    out.append(round(s.loglik() / len(tok), 4))
st = s.stats()
print(json.dumps({"refresh": os.environ.get("B200LDA_TABLE_REFRESH", "0 (auto)"), "max_ctas": os.environ.get("B200LDA_MAX_CTAS", "0"),
                  "marks": marks, "ll": out, "ms_per_sweep": round(st["cum_sample_ms"] / st["cum_sweeps"], 3), "refresh_used": st["table_refresh_last"], "hot": st["hot_words"], "rows_last": st["rows_refreshed_last"],
                  "mallet_T1": [round(float(np.mean([g["mallet_ll_per_token"]["1"][k][g["sweeps"].index(m)] for k in "123"])), 4) for m in marks if m in g["sweeps"]]}))
