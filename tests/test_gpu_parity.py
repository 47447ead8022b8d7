"""GPU parity tests proper: the CUDA path, through the C ABI, against the CPU oracle.

Bit-exact (integer / index work): initial topics, n_wk / n_k / n_dk, the frozen-snapshot topic
index of every token, and the whole DEFERRED-mode chain. Floating point: log-likelihood within
1e-9 relative of the oracle's fp64 evaluation of the same counts.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ALPHA, BETA = 0.1, 0.01


def _sampler(K, V, seed, mode=None, **kw):
    import ldagibbssampling_b200 as L
    return L.Sampler(K, V, ALPHA * K, BETA, seed=seed,
                     mode=L.MODE_DEFERRED if mode is None else mode, **kw)


CASES = [
    # D, V, mean_len, k_true, K
    (300, 200, 40.0, 8, 4),
    (300, 500, 60.0, 10, 20),
    (200, 400, 120.0, 20, 100),
    (150, 300, 200.0, 30, 1000),
    (60, 100, 300.0, 10, 1500),
    (40, 80, 120.0, 10, 10000),   # tables too large for shared memory: read through L1
]


@pytest.mark.parametrize("D,V,mean_len,k_true,K", CASES)
def test_init_and_counts_match_oracle(oracle, D, V, mean_len, k_true, K):
    dp, tok = oracle.gen_corpus(D, V, mean_len, k_true, 11)
    s = _sampler(K, V, seed=5)
    s.load_corpus(dp, tok)
    s.init_assignments(None)
    z = s.assignments()
    assert np.array_equal(z, oracle.init_z(len(tok), K, 5))
    nwk, nk = oracle.count(dp, tok, z, V, K)
    assert np.array_equal(s.nwk(), nwk)
    assert np.array_equal(s.nk(), nk)
    rp, nnz, topic, cnt = oracle.ndk_csr(dp, z, K)
    g_rp, g_topic, g_cnt = s.ndk_csr()
    assert np.array_equal(np.diff(g_rp), nnz)
    for d in range(D):
        assert np.array_equal(g_topic[g_rp[d]:g_rp[d + 1]], topic[rp[d]:rp[d] + nnz[d]])
        assert np.array_equal(g_cnt[g_rp[d]:g_rp[d + 1]], cnt[rp[d]:rp[d] + nnz[d]])


@pytest.mark.parametrize("D,V,mean_len,k_true,K", CASES)
def test_frozen_snapshot_index_parity(oracle, D, V, mean_len, k_true, K):
    dp, tok = oracle.gen_corpus(D, V, mean_len, k_true, 12)
    z0 = oracle.init_z(len(tok), K, 9)
    s = _sampler(K, V, seed=9)
    s.load_corpus(dp, tok)
    s.init_assignments(z0)
    # Philox uniforms
    got = s.sample_frozen(None, sweep=3)
    want = oracle.spec_frozen(dp, tok, z0, V, K, ALPHA, BETA, 9, 3)
    assert np.array_equal(got, want)
    # caller-supplied uniforms
    u = np.random.default_rng(0).random(len(tok), dtype=np.float32)
    got = s.sample_frozen(u)
    want = oracle.spec_frozen(dp, tok, z0, V, K, ALPHA, BETA, 9, 1, uniforms=u)
    assert np.array_equal(got, want)
    # frozen mode moved nothing
    assert np.array_equal(s.assignments(), z0)


@pytest.mark.parametrize("D,V,mean_len,k_true,K", CASES[:4])
def test_deferred_chain_bit_exact(oracle, D, V, mean_len, k_true, K):
    dp, tok = oracle.gen_corpus(D, V, mean_len, k_true, 13)
    z0 = oracle.init_z(len(tok), K, 21)
    s = _sampler(K, V, seed=21)
    s.load_corpus(dp, tok)
    s.init_assignments(z0)
    s.sweep(4)
    want = oracle.spec_sweeps(dp, tok, z0, V, K, ALPHA, BETA, 21, 1, 4)
    got = s.assignments()
    assert np.array_equal(got, want)
    nwk, nk = oracle.count(dp, tok, want, V, K)
    assert np.array_equal(s.nwk(), nwk)
    assert np.array_equal(s.nk(), nk)


@pytest.mark.parametrize("mode_name", ["LIVE", "DEFERRED"])
def test_count_invariants_after_sweeps(oracle, mode_name):
    import ldagibbssampling_b200 as L
    D, V, K = 2000, 1500, 50
    dp, tok = oracle.gen_corpus(D, V, 80.0, 20, 14)
    s = _sampler(K, V, seed=3, mode=getattr(L, "MODE_" + mode_name))
    s.load_corpus(dp, tok)
    s.init_assignments(None)
    for _ in range(3):
        s.sweep(2)
        z = s.assignments()
        nwk, nk = oracle.count(dp, tok, z, V, K)
        g_nwk, g_nk = s.nwk(), s.nk()
        assert np.array_equal(g_nwk, nwk)
        assert np.array_equal(g_nk, nk)
        assert g_nwk.sum() == g_nk.sum() == len(tok)
        rp, topic, cnt = s.ndk_csr()
        assert cnt.sum() == len(tok)
        o_rp, o_nnz, o_topic, o_cnt = oracle.ndk_csr(dp, z, K)
        assert np.array_equal(np.diff(rp), o_nnz)
        d = D // 2
        assert np.array_equal(topic[rp[d]:rp[d + 1]], o_topic[o_rp[d]:o_rp[d] + o_nnz[d]])
        assert np.array_equal(cnt[rp[d]:rp[d + 1]], o_cnt[o_rp[d]:o_rp[d] + o_nnz[d]])


def test_loglik_theta_phi_match_oracle(oracle):
    D, V, K = 500, 800, 30
    dp, tok = oracle.gen_corpus(D, V, 70.0, 12, 15)
    s = _sampler(K, V, seed=4)
    s.load_corpus(dp, tok)
    s.init_assignments(None)
    s.sweep(3)
    z = s.assignments()
    want = oracle.loglik(dp, tok, z, V, K, ALPHA, BETA)
    got = s.loglik()
    assert abs(got - want) <= 1e-9 * abs(want)  # fp64, tolerance 1e-9 relative
    stirling = oracle.loglik(dp, tok, z, V, K, ALPHA, BETA, stirling=True)
    assert abs(got - stirling) <= 1e-6 * abs(stirling)  # Mallet's logGammaStirling (series error ~5e-6 per term)
    th = s.theta(0, D)
    for d in (0, 7, D - 1):
        assert np.allclose(th[d], oracle.theta(z[dp[d]:dp[d + 1]], K, ALPHA), rtol=1e-14, atol=0)
    nwk, nk = oracle.count(dp, tok, z, V, K)
    assert np.allclose(s.phi(), oracle.phi(nwk, nk, BETA), rtol=1e-14, atol=0)


def test_empty_and_ragged_documents(oracle):
    K, V = 16, 40
    lens = np.array([0, 1, 0, 0, 33, 1, 64, 0, 2, 700, 0], np.int64)
    dp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    rng = np.random.default_rng(3)
    tok = rng.integers(0, V, int(dp[-1])).astype(np.int32)
    z0 = oracle.init_z(len(tok), K, 2)
    s = _sampler(K, V, seed=2)
    s.load_corpus(dp, tok)
    s.init_assignments(z0)
    assert np.array_equal(s.sample_frozen(None, 1), oracle.spec_frozen(dp, tok, z0, V, K, ALPHA, BETA, 2, 1))
    s.sweep(3)
    assert np.array_equal(s.assignments(), oracle.spec_sweeps(dp, tok, z0, V, K, ALPHA, BETA, 2, 1, 3))
    # a corpus with no tokens at all
    s2 = _sampler(K, V, seed=2)
    s2.load_corpus(np.zeros(4, np.int64), np.zeros(0, np.int32))
    s2.init_assignments(None)
    s2.sweep(1)
    assert s2.nk().sum() == 0


def test_errors_are_reported_not_swallowed(oracle):
    import ldagibbssampling_b200 as L
    s = _sampler(8, 10, seed=1)
    with pytest.raises(L.B200LDAError) as e:
        s.sweep(1)
    assert e.value.code == -5  # ESTATE: no corpus
    with pytest.raises(L.B200LDAError) as e:
        s.load_corpus(np.array([0, 2], np.int64), np.array([1, 10], np.int32))
    assert e.value.code == -6  # ERANGE: word id >= V
    s.load_corpus(np.array([0, 2], np.int64), np.array([1, 9], np.int32))
    with pytest.raises(L.B200LDAError) as e:
        s.init_assignments(np.array([0, 8], np.int32))
    assert e.value.code == -6  # ERANGE: topic >= K
    with pytest.raises(L.B200LDAError):
        L.Sampler(0, 10, 1.0, 0.1)


def test_u16_assignment_paths_equal_the_int32_ones(oracle):
    """b200lda_init_assignments_u16 / b200lda_get_assignments_u16 (topics in the device's own width,
    what bench.py's end-to-end step uses) against the int32 entry points: same counts, same rows,
    same DEFERRED chain; an out-of-range topic is refused and leaves the context usable."""
    import ldagibbssampling_b200 as L
    D, V, K = 400, 300, 50
    dp, tok = oracle.gen_corpus(D, V, 70.0, 12, 21)
    z0 = oracle.init_z(len(tok), K, 9)
    a, b = _sampler(K, V, seed=4), _sampler(K, V, seed=4)
    a.load_corpus(dp, tok)
    b.load_corpus(dp, tok)
    a.init_assignments(z0.astype(np.int32))
    b.init_assignments(z0.astype(np.uint16))
    assert np.array_equal(a.nwk(), b.nwk()) and np.array_equal(a.nk(), b.nk())
    for x, y in zip(a.ndk_csr(), b.ndk_csr()):
        assert np.array_equal(x, y)
    a.sweep(3)
    b.sweep(3)
    za, zb = a.assignments(), b.assignments(np.uint16)
    assert zb.dtype == np.uint16 and np.array_equal(za, zb.astype(np.int32))
    assert np.array_equal(za, oracle.spec_sweeps(dp, tok, z0, V, K, ALPHA, BETA, 4, 1, 3))
    bad = z0.astype(np.uint16)
    bad[17] = K
    with pytest.raises(L.B200LDAError) as e:
        b.init_assignments(bad)
    assert e.value.code == -6  # ERANGE
    with pytest.raises(L.B200LDAError):
        b.sweep(1)  # no valid assignments any more
    b.init_assignments(z0.astype(np.uint16))
    assert np.array_equal(b.nwk(), oracle.count(dp, tok, z0, V, K)[0])
    a.close()
    b.close()


def test_word_order_csr_is_a_permutation_grouped_by_word(oracle):
    D, V, K = 700, 300, 8
    dp, tok = oracle.gen_corpus(D, V, 30.0, 6, 16)
    s = _sampler(K, V, seed=1)
    s.load_corpus(dp, tok)
    word_ptr, toks = s.word_order()
    assert word_ptr[0] == 0 and word_ptr[-1] == len(tok)
    assert np.array_equal(np.diff(word_ptr), np.bincount(tok, minlength=V))
    assert np.array_equal(np.sort(toks), np.arange(len(tok)))          # a bijection on token indices
    assert np.array_equal(tok[toks], np.repeat(np.arange(V), np.diff(word_ptr)))  # grouped by word


def test_corpus_validation_on_device(oracle):
    import ldagibbssampling_b200 as L
    s = _sampler(8, 10, seed=1)
    with pytest.raises(L.B200LDAError) as e:   # non-monotone CSR
        s._check(s._lib.b200lda_load_corpus(s._h, 2, np.array([0, 3, 2], np.int64).ctypes.data,
                                            np.array([1, 2, 3], np.int32).ctypes.data))
    assert e.value.code == -1
    long_doc = np.zeros(70000, np.int32)
    with pytest.raises(L.B200LDAError) as e:   # document longer than 65535 tokens
        s.load_corpus(np.array([0, 70000], np.int64), long_doc)
    assert e.value.code == -6
    s.load_corpus(np.array([0, 65535], np.int64), long_doc[:65535])   # the limit itself is fine
    s.init_assignments(None)
    s.sweep(1)
    assert s.nk().sum() == 65535


def test_live_mode_equals_sequential_oracle_when_documents_do_not_interact(oracle):
    """LIVE mode updates n_wk in place. Across documents that is a race (by design), but documents
    with disjoint vocabularies never touch the same n_wk rows, so the chain is deterministic and
    must equal the oracle's sequential rendering of LIVE mode (n_wk moves immediately, tables and
    n_k from the sweep start: table_refresh = 1 switches the in-sweep table rebuilds off) bit for bit."""
    import ldagibbssampling_b200 as L
    rng = np.random.default_rng(7)
    K, words_per_doc = 40, 50
    lens = [1, 700, 33, 2500, 64, 190]
    dp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok = np.concatenate([d * words_per_doc + rng.integers(0, words_per_doc, n) for d, n in enumerate(lens)]).astype(np.int32)
    V = words_per_doc * len(lens)
    z0 = oracle.init_z(len(tok), K, 19)
    s = _sampler(K, V, seed=19, mode=L.MODE_LIVE, table_refresh=1)
    s.load_corpus(dp, tok)
    s.init_assignments(z0)
    s.sweep(5)
    want = oracle.spec_sweeps(dp, tok, z0, V, K, ALPHA, BETA, 19, 1, 5, live=True)
    assert np.array_equal(s.assignments(), want)
    # with the rebuilds on (every row 16 times per sweep) the chain differs but the counts stay exact
    s = _sampler(K, V, seed=19, mode=L.MODE_LIVE, table_refresh=16)
    s.load_corpus(dp, tok)
    s.init_assignments(z0)
    s.sweep(5)
    z = s.assignments()
    nwk, nk = oracle.count(dp, tok, z, V, K)
    assert np.array_equal(s.nwk(), nwk) and np.array_equal(s.nk(), nk)
    assert oracle.loglik(dp, tok, z, V, K, ALPHA, BETA) > oracle.loglik(dp, tok, z0, V, K, ALPHA, BETA)


@pytest.mark.parametrize("K,V,lens", [
    (1, 5, [3, 40, 1]),                 # a single topic: every draw must return topic 0
    (2, 1, [100, 7]),                   # a single word type
    (3, 6, [500]),                      # one document much longer than K (row capacity = K, always full)
    (33, 9, [64, 65, 31, 32, 33, 200]), # K just past one tile: rows cross the 32-slot boundary both ways
    (64, 50, [64, 129, 300]),           # rows that fill their class capacity exactly
])
def test_edge_shapes_match_oracle(oracle, K, V, lens):
    rng = np.random.default_rng(K * 1000 + V)
    dp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok = rng.integers(0, V, int(dp[-1])).astype(np.int32)
    z0 = oracle.init_z(len(tok), K, 4)
    s = _sampler(K, V, seed=4)
    s.load_corpus(dp, tok)
    s.init_assignments(z0)
    assert np.array_equal(s.sample_frozen(None, 2), oracle.spec_frozen(dp, tok, z0, V, K, ALPHA, BETA, 4, 2))
    s.sweep(6)
    want = oracle.spec_sweeps(dp, tok, z0, V, K, ALPHA, BETA, 4, 1, 6)
    assert np.array_equal(s.assignments(), want)
    nwk, nk = oracle.count(dp, tok, want, V, K)
    assert np.array_equal(s.nwk(), nwk) and np.array_equal(s.nk(), nk)
    ll = oracle.loglik(dp, tok, want, V, K, ALPHA, BETA)
    assert abs(s.loglik() - ll) <= 1e-9 * abs(ll)


def test_launch_policy_does_not_change_deferred_results(oracle, monkeypatch):
    """Row-width classes run as bulk launches in sequence plus background launches whose grid is
    retuned every sweep from event timings. None of that may show in DEFERRED results: same chain,
    bit for bit, with the classes forked at full size, all in sequence, or under the default
    policy, and equal to the oracle. The corpus has a long-document tail (several classes)."""
    rng = np.random.default_rng(5)
    V, K = 600, 700
    lens = np.concatenate([rng.integers(20, 120, 900), rng.integers(300, 700, 12), [1500, 2200]]).astype(np.int64)
    rng.shuffle(lens)
    dp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok = rng.integers(0, V, int(dp[-1])).astype(np.int32)
    z0 = oracle.init_z(len(tok), K, 4)
    want = oracle.spec_sweeps(dp, tok, z0, V, K, ALPHA, BETA, 4, 1, 6)
    classes = None
    for policy in ("0", "1", "2"):
        monkeypatch.setenv("B200LDA_CLASS_STREAMS", policy)
        s = _sampler(K, V, seed=4)
        s.load_corpus(dp, tok)
        s.init_assignments(z0)
        s.sweep(6)
        assert np.array_equal(s.assignments(), want), f"policy {policy}"
        classes = s.stats()["row_classes"]
    assert classes >= 3


def test_row_growing_past_eight_tiles_moves_to_shared_memory(oracle):
    """Documents of ~300 tokens start a sweep with ~250 distinct topics (8 register tiles) and, at
    K = 3000 where nearly every draw lands on a topic new to the document, grow past 256 live slots
    during the visit: the row moves from registers to shared memory mid-visit. Frozen topics and the
    whole DEFERRED chain stay bit-exact with the oracle, before and after the move."""
    rng = np.random.default_rng(8)
    V, K, D = 400, 3000, 24
    lens = rng.integers(290, 330, D).astype(np.int64)
    dp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok = rng.integers(0, V, int(dp[-1])).astype(np.int32)
    z0 = np.empty(int(dp[-1]), np.int32)
    for d in range(D):  # ~250 distinct topics per document, the rest repeats
        L = int(lens[d])
        distinct = rng.choice(K, 250, replace=False)
        z0[dp[d]:dp[d + 1]] = np.concatenate([distinct, rng.choice(distinct, L - 250)])
    import ldagibbssampling_b200 as L
    alpha = 0.5  # alpha K = 1500 against 300 tokens: the prior bucket dominates and the rows grow
    s = L.Sampler(K, V, alpha * K, BETA, seed=6, mode=L.MODE_DEFERRED)
    s.load_corpus(dp, tok)
    s.init_assignments(z0)
    assert np.array_equal(s.sample_frozen(None, sweep=2), oracle.spec_frozen(dp, tok, z0, V, K, alpha, BETA, 6, 2))
    s.sweep(3)
    want = oracle.spec_sweeps(dp, tok, z0, V, K, alpha, BETA, 6, 1, 3)
    assert np.array_equal(s.assignments(), want)
    rp, topic, cnt = s.ndk_csr()
    assert np.diff(rp).max() > 256  # the rows did outgrow the register path
    nwk, nk = oracle.count(dp, tok, want, V, K)
    assert np.array_equal(s.nwk(), nwk) and np.array_equal(s.nk(), nk)


def test_live_mode_with_every_warp_on_the_same_word_rows(oracle):
    """LIVE mode's worst case for own-move visibility: 512 documents over only 12 word types, so the
    8 warps of every CTA (and every CTA) hammer the same 12 n_wk rows with atomics while reading them
    through L1. The atomics are exact whatever the reads saw: the counts must equal a recount from z,
    no count may be negative, and the chain must still improve the likelihood like DEFERRED does."""
    import ldagibbssampling_b200 as L
    rng = np.random.default_rng(11)
    D, V, K = 512, 12, 24
    lens = rng.integers(40, 200, D).astype(np.int64)
    dp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    tok = rng.integers(0, V, int(dp[-1])).astype(np.int32)
    z0 = oracle.init_z(len(tok), K, 2)
    ll = {}
    for name, mode, refresh in (("live", L.MODE_LIVE, 1), ("live_refresh", L.MODE_LIVE, 8), ("deferred", L.MODE_DEFERRED, 0)):
        s = _sampler(K, V, seed=2, mode=mode, table_refresh=refresh)
        s.load_corpus(dp, tok)
        s.init_assignments(z0)
        for _ in range(4):
            s.sweep(5)
            z = s.assignments()
            nwk, nk = oracle.count(dp, tok, z, V, K)
            g = s.nwk()
            assert g.min() >= 0 and np.array_equal(g, nwk) and np.array_equal(s.nk(), nk), name
            assert s.check_invariants() == (len(tok), len(tok), 0, len(tok)), name
        ll[name] = s.loglik() / len(tok)
        s.close()
    ll0 = oracle.loglik(dp, tok, z0, V, K, ALPHA, BETA) / len(tok)
    assert min(ll.values()) > ll0
    assert abs(ll["live"] - ll["deferred"]) < 0.02 * abs(ll["deferred"])
    assert abs(ll["live_refresh"] - ll["deferred"]) < 0.02 * abs(ll["deferred"])
