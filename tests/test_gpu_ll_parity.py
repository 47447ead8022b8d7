"""LL/token of the GPU chains against the Mallet-faithful oracle at the benchmark's K.

BASELINE.json's headline runs K = 1000 on V = 141 000; the approximations that distinguish the
GPU sampler from Mallet (prior bucket from a sweep-start table, sweep-start n_k, fp32 prefix sums)
grow with K, so the 1 % LL tolerance of `north_star` is checked HERE on a C4-SHAPED sample both
sides can run: 20 000 documents of config 4's generator (1.8 M tokens, V = 141 000, K = 1000,
alpha_k = 0.1, beta = 0.01 - the sample bench.py's CPU leg times), same corpus, same initial
topics. tests/golden/c4s_ll_trajectory.json holds the oracle's LL/token at sweeps 25, 50, 100, 200
for 3 seeds with 1, 2 and 4 worker threads (tests/golden/make_golden.py c4s); the reference call
sites are estimate() and modelLogLikelihood(), cmu_ron/TrainAndPredict.java:159-171,234.

Every assertion is against EACH oracle seed (not the closest one). Measured (profiles/r02_ll_parity.md):
on a corpus this small AD-LDA with T replicas mixes visibly slower than the single chain - Mallet
itself loses 4.7 % / 1.3 % LL at sweeps 25 / 200 going from 1 to 2 threads. The GPU's LIVE mode
(prior bucket from per-word tables that are rebuilt DURING the sweep - on a corpus this small: the
sweep runs in 8 segments with a full rebuild between them -, n_k from the sweep start) follows the
single chain. What is asserted:
  * LIVE, 1 shard: never behind Mallet with 2 threads (0.3 % slack), i.e. ahead of the reference's
    own configuration, setNumThreads(4); within 1.75 / 1 / 1 / 1 % of the SINGLE chain (measured over the round's runs: 1.1-1.35 / 0.64-0.73 / 0.39-0.44 / 0.30-0.36 %; LIVE is not bit-reproducible, the sweep-25 figure moves by ~0.2 points run to run) at sweeps
    25 / 50 / 100 / 200 (the gap closes with sweeps: the reference runs 1 000-10 000);
  * LIVE, G = 2, 4 shards (the reference's setNumThreads(4)): within 1.5 % of Mallet with G threads at
    sweep 25 and within 1 % from sweep 50 on (measured 0.8 / 0.5 / 0.25 / 0.2 % at G = 2 and
    0.4 / 0.15 / 0.1 / 0.1 % at G = 4);
  * DEFERRED is AD-LDA with one replica per DOCUMENT (every other document's counts are a sweep
    old): within 1 % of Mallet with 4 threads from sweep 100 on, never more than 3 % behind it.
The measured curves are written to gpurun_out/ll_parity_c4s.json when that directory exists.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
ALPHA, BETA = 0.1, 0.01


def _record(key, value):
    out = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out):
        return
    path = os.path.join(out, "ll_parity_c4s.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = value
    json.dump(data, open(path, "w"), indent=1)


@pytest.fixture(scope="module")
def c4s(oracle):
    import bench_corpus as BC
    g = json.load(open(os.path.join(GOLD, "c4s_ll_trajectory.json")))
    dp, tok, V, K = BC.cpu_sample("c4", g["D"])
    assert (V, K, len(tok)) == (g["V"], g["K"], g["tokens"])
    z0 = oracle.init_z(len(tok), K, 7)
    return g, dp, tok, V, K, z0


def _band(g, threads):
    return np.array([g["mallet_ll_per_token"][str(threads)][s] for s in ("1", "2", "3")])  # seeds x marks


@pytest.mark.parametrize("mode_name", ["LIVE", "DEFERRED"])
def test_single_shard_ll_within_one_percent_of_mallet_at_k1000(c4s, mode_name):
    import ldagibbssampling_b200 as L
    g, dp, tok, V, K, z0 = c4s
    ref = _band(g, 1)
    s = L.Sampler(K, V, ALPHA * K, BETA, seed=7, mode=getattr(L, "MODE_" + mode_name))
    s.load_corpus(dp, tok)
    s.init_assignments(z0)
    N = len(tok)
    assert abs(s.loglik() / N - g["ll_init"]) < 1e-6
    done, curve = 0, []
    for mark in g["sweeps"]:
        s.sweep(mark - done)
        done = mark
        curve.append(s.loglik() / N)
    s.close()
    _record(mode_name.lower() + "_1", curve)
    curve = np.array(curve)
    if mode_name == "LIVE":
        t2 = _band(g, 2)
        assert (curve[None, :] >= t2 * 1.003).all(), (curve.tolist(), t2.tolist())  # LL < 0: x1.003 is 0.3 % lower
        tol = np.array([0.0175, 0.01, 0.01, 0.01])
        rel = np.abs(curve[None, :] - ref) / np.abs(ref)
        assert (rel <= tol[None, :]).all(), (curve.tolist(), ref.tolist())
    else:
        t4 = _band(g, 4)
        rel = (t4 - curve[None, :]) / np.abs(t4)   # > 0: behind Mallet with 4 threads
        assert (rel <= 0.03).all(), (curve.tolist(), t4.tolist())
        assert (np.abs(rel[:, 2:]) <= 0.01).all(), (curve.tolist(), t4.tolist())


@pytest.mark.parametrize("world", [2, 4])
def test_live_shards_ll_within_one_percent_of_mallet_with_as_many_threads(c4s, world):
    """G contexts on one device, exchange buffers summed between sweep_begin and sweep_end (what
    NCCL does across GPUs): Mallet's setNumThreads(G), cmu_ron/TrainAndPredict.java:164."""
    import ldagibbssampling_b200 as L
    from test_gpu_model import _run_shards_marks
    g, dp, tok, V, K, z0 = c4s
    ref = _band(g, world)
    curve = _run_shards_marks(L, dp, tok, V, K, world, L.MODE_LIVE, seed=7, z0=z0, marks=g["sweeps"])
    _record(f"live_{world}", curve)
    for i, mark in enumerate(g["sweeps"]):
        rel = np.abs(curve[i] - ref[:, i]) / np.abs(ref[:, i])
        assert rel.max() <= {25: 0.015}.get(mark, 0.01), (world, mark, curve[i], ref[:, i].tolist())
