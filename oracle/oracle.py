"""ctypes binding of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE.

PARITY UNPINNED: the reference holds no golden vectors for this path and its arithmetic lives in
the absent cc.mallet:mallet:2.0.7 jar (reference pom.xml:107-111); see oracle/lda_oracle.h.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module. The product package (ldagibbssampling_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc, no reference sources)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    L.oracle_gen_corpus.restype = C.c_int64
    L.oracle_gen_corpus.argtypes = [C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_uint64,
                                    _i64p, C.c_void_p, C.c_int64]
    L.oracle_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
    L.oracle_java_random_ints.argtypes = [C.c_int64, C.c_int32, C.c_int32, _i32p]
    L.oracle_java_random_uniforms.argtypes = [C.c_int64, C.c_int32, _f64p]
    L.oracle_init_z_philox.argtypes = [C.c_int64, C.c_int32, C.c_uint64, C.c_int64, _i32p]
    L.oracle_count.argtypes = [C.c_int64, C.c_int32, C.c_int32, _i64p, _i32p, _i32p, _i32p, _i32p]
    L.oracle_ndk_csr.argtypes = [C.c_int64, C.c_int32, _i64p, _i32p, _i64p, _i32p, _i32p, _i32p]
    L.oracle_tile_scan_f32.argtypes = [_f32p, C.c_int64, _f32p]
    L.oracle_lane_strided_prefix_f32.argtypes = [_f32p, C.c_int64, _f32p, C.POINTER(C.c_float)]
    L.oracle_spec_tables.argtypes = [C.c_int32, C.c_int32, _i32p, _i32p, _f64p, C.c_double,
                                     _f32p, _f32p, _f32p, _f32p]
    L.oracle_spec_hsearch.restype = C.c_int32
    L.oracle_spec_hsearch.argtypes = [_f32p, C.c_int32, C.c_float]
    L.oracle_spec_select.restype = C.c_int32
    L.oracle_spec_select.argtypes = [C.c_int32, _i32p, _i32p, C.c_int32, _i32p, _f32p, _f32p,
                                     _f32p, C.c_float, C.c_float, C.c_int32, C.c_float]
    L.oracle_spec_frozen.argtypes = [C.c_int64, C.c_int32, C.c_int32, _i64p, _i32p, _i32p, _f64p,
                                     C.c_double, C.c_uint64, C.c_uint32, C.c_int64, C.c_void_p,
                                     _i32p]
    L.oracle_spec_sweeps.argtypes = [C.c_int64, C.c_int32, C.c_int32, _i64p, _i32p, _i32p, _f64p,
                                     C.c_double, C.c_uint64, C.c_uint32, C.c_int32, C.c_int64]
    L.oracle_spec_sweeps_mode.argtypes = L.oracle_spec_sweeps.argtypes + [C.c_int32]
    L.oracle_spec_sweep_given_counts.argtypes = [C.c_int64, C.c_int32, C.c_int32, _i64p, _i32p, _i32p, _i32p,
                                                 _i32p, _f64p, C.c_double, C.c_uint64, C.c_uint32, C.c_int64,
                                                 C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    L.oracle_spec_infer.argtypes = [C.c_int64, C.c_int32, C.c_int32, _i64p, _i32p, _i32p, _i32p, _f64p,
                                    C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, _f64p]
    L.oracle_exact_conditional.argtypes = [C.c_int32, C.c_int32, _i32p, _i32p, _i32p, _f64p,
                                           C.c_double, C.c_int32, _f64p]
    L.oracle_loglik.restype = C.c_double
    L.oracle_loglik.argtypes = [C.c_int64, C.c_int32, C.c_int32, _i64p, _i32p, _i32p, _f64p,
                                C.c_double, C.c_int32]
    L.oracle_log_gamma_stirling.restype = C.c_double
    L.oracle_log_gamma_stirling.argtypes = [C.c_double]
    L.oracle_theta.argtypes = [C.c_int32, _i32p, C.c_int64, _f64p, _f64p]
    L.oracle_phi.argtypes = [C.c_int32, C.c_int32, _i32p, _i32p, C.c_double, _f64p]
    L.mallet_create.restype = C.c_void_p
    L.mallet_create.argtypes = [C.c_int32, C.c_double, C.c_double]
    L.mallet_destroy.argtypes = [C.c_void_p]
    L.mallet_set_random_seed.argtypes = [C.c_void_p, C.c_int32]
    L.mallet_set_num_threads.argtypes = [C.c_void_p, C.c_int32]
    L.mallet_add_instances.restype = C.c_int
    L.mallet_add_instances.argtypes = [C.c_void_p, C.c_int64, C.c_int32, _i64p, _i32p, C.c_void_p]
    L.mallet_estimate.restype = C.c_int
    L.mallet_estimate.argtypes = [C.c_void_p, C.c_int32]
    L.mallet_get_timers.restype = None
    L.mallet_get_timers.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mallet_model_log_likelihood.restype = C.c_double
    L.mallet_model_log_likelihood.argtypes = [C.c_void_p]
    L.mallet_get_assignments.argtypes = [C.c_void_p, _i32p]
    L.mallet_get_counts.argtypes = [C.c_void_p, _i32p, _i32p]
    L.mallet_get_topic_probabilities.argtypes = [C.c_void_p, C.c_int64, _f64p]
    L.mallet_infer.argtypes = [C.c_void_p, _i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, _f64p]
    L.oracle_digamma.restype = C.c_double
    L.oracle_digamma.argtypes = [C.c_double]
    L.oracle_learn_parameters.restype = C.c_double
    L.oracle_learn_parameters.argtypes = [_f64p, C.c_int32, _i32p, C.c_int32, _i32p, C.c_double, C.c_double, C.c_int32]
    L.oracle_learn_symmetric_concentration.restype = C.c_double
    L.oracle_learn_symmetric_concentration.argtypes = [_i64p, C.c_int64, _i64p, C.c_int64, C.c_int32, C.c_double]
    L.mallet_num_tokens.restype = C.c_int64
    L.mallet_num_tokens.argtypes = [C.c_void_p]
    _lib = L
    return L


# ---- corpus ---------------------------------------------------------------------------------

def gen_corpus(D: int, V: int, mean_len: float, k_true: int, seed: int):
    """Seeded synthetic corpus. Returns (doc_ptr int64[D+1], tok_word int32[N])."""
    L = lib()
    doc_ptr = np.zeros(D + 1, np.int64)
    n = L.oracle_gen_corpus(D, V, mean_len, k_true, seed, doc_ptr, None, 0)
    tok = np.zeros(n, np.int32)
    got = L.oracle_gen_corpus(D, V, mean_len, k_true, seed, doc_ptr, tok.ctypes.data, n)
    assert got == n
    return doc_ptr, tok


def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().oracle_philox4x32_10(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
    return out


def java_ints(seed: int, n: int, bound: int = 0):
    out = np.zeros(n, np.int32)
    lib().oracle_java_random_ints(seed, n, bound, out)
    return out


def java_uniforms(seed: int, n: int):
    out = np.zeros(n, np.float64)
    lib().oracle_java_random_uniforms(seed, n, out)
    return out


# ---- spec sampler ---------------------------------------------------------------------------

def _alpha(alpha, K):
    a = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha, np.float64), (K,)))
    return a


def init_z(N, K, seed, global_off=0):
    z = np.zeros(N, np.int32)
    lib().oracle_init_z_philox(N, K, seed, global_off, z)
    return z


def count(doc_ptr, tok, z, V, K):
    nwk = np.zeros((V, K), np.int32)
    nk = np.zeros(K, np.int32)
    lib().oracle_count(len(doc_ptr) - 1, V, K, doc_ptr, tok, np.ascontiguousarray(z, np.int32),
                       nwk.reshape(-1), nk)
    return nwk, nk


def ndk_csr(doc_ptr, z, K):
    D = len(doc_ptr) - 1
    lens = np.diff(doc_ptr)
    cap = int(np.minimum(lens, K).sum())
    row_ptr = np.zeros(D + 1, np.int64)
    nnz = np.zeros(D, np.int32)
    topic = np.zeros(max(cap, 1), np.int32)
    cnt = np.zeros(max(cap, 1), np.int32)
    lib().oracle_ndk_csr(D, K, doc_ptr, np.ascontiguousarray(z, np.int32), row_ptr, nnz, topic, cnt)
    return row_ptr, nnz, topic, cnt


def tile_scan(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(x)
    lib().oracle_tile_scan_f32(x, len(x), out)
    return out


def lane_strided_prefix(x):
    """Doc-bucket prefix order of the spec; returns (prefix per slot, scanned total of lane 31)."""
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(x)
    total = C.c_float(0.0)
    lib().oracle_lane_strided_prefix_f32(x, len(x), out, C.byref(total))
    return out, np.float32(total.value)


def spec_tables(nwk, nk, alpha, beta):
    V, K = nwk.shape
    invden = np.zeros(K, np.float32)
    ab = np.zeros(K, np.float32)
    prior = np.zeros((V, K), np.float32)
    q = np.zeros(V, np.float32)
    lib().oracle_spec_tables(V, K, np.ascontiguousarray(nwk, np.int32).reshape(-1),
                             np.ascontiguousarray(nk, np.int32), _alpha(alpha, K), beta, invden, ab,
                             prior.reshape(-1), q)
    return invden, ab, prior, q


def hsearch(row, s):
    row = np.ascontiguousarray(row, np.float32)
    return int(lib().oracle_spec_hsearch(row, len(row), float(np.float32(s))))


def spec_select(K, slot_topic, slot_count, nwk_row, invden, ab, prior_row, q_w, beta, old, u):
    return int(lib().oracle_spec_select(
        K, np.ascontiguousarray(slot_topic, np.int32), np.ascontiguousarray(slot_count, np.int32),
        len(slot_topic), np.ascontiguousarray(nwk_row, np.int32), invden, ab,
        np.ascontiguousarray(prior_row, np.float32), float(q_w), float(np.float32(beta)), int(old),
        float(np.float32(u))))


def spec_frozen(doc_ptr, tok, z_in, V, K, alpha, beta, seed, sweep, global_off=0, uniforms=None):
    z_out = np.zeros(len(tok), np.int32)
    if uniforms is not None:
        uniforms = np.ascontiguousarray(uniforms, np.float32)
        up = uniforms.ctypes.data
    else:
        up = None
    lib().oracle_spec_frozen(len(doc_ptr) - 1, V, K, doc_ptr, tok, np.ascontiguousarray(z_in, np.int32),
                             _alpha(alpha, K), beta, seed, sweep, global_off, up, z_out)
    return z_out


def spec_sweeps(doc_ptr, tok, z, V, K, alpha, beta, seed, first_sweep, n_sweeps, global_off=0,
                live=False):
    """DEFERRED-mode chain (live=False, what the GPU reproduces bit for bit) or the sequential
    rendering of LIVE mode (live=True)."""
    z = np.array(z, np.int32, copy=True)
    lib().oracle_spec_sweeps_mode(len(doc_ptr) - 1, V, K, doc_ptr, tok, z, _alpha(alpha, K), beta,
                                  seed, first_sweep, n_sweeps, global_off, 1 if live else 0)
    return z


def spec_sweep_given_counts(doc_ptr, tok, z, nwk, nk, alpha, beta, seed, sweep, global_off=0, live=False):
    """One DEFERRED (or LIVE) sweep of a shard against the given global counts.
    Returns (z_new, delta_nwk, delta_nk); nwk is not modified unless live."""
    V, K = nwk.shape
    z = np.array(z, np.int32, copy=True)
    nwk_c = np.ascontiguousarray(nwk, np.int32) if not live else nwk
    d_nwk = np.zeros((V, K), np.int32)
    d_nk = np.zeros(K, np.int32)
    lib().oracle_spec_sweep_given_counts(len(doc_ptr) - 1, V, K, doc_ptr, tok, z, nwk_c.reshape(-1),
                                         np.ascontiguousarray(nk, np.int32), _alpha(alpha, K), beta, seed, sweep,
                                         global_off, 1 if live else 0, 1, d_nwk.ctypes.data, d_nk.ctypes.data)
    return z, d_nwk, d_nk


def spec_infer(doc_ptr, tok, nwk, nk, alpha, beta, iterations=100, thinning=10, burn_in=10, seed=0):
    V, K = nwk.shape
    theta = np.zeros((len(doc_ptr) - 1, K), np.float64)
    lib().oracle_spec_infer(len(doc_ptr) - 1, V, K, np.ascontiguousarray(doc_ptr, np.int64),
                            np.ascontiguousarray(tok, np.int32), np.ascontiguousarray(nwk, np.int32).reshape(-1),
                            np.ascontiguousarray(nk, np.int32), _alpha(alpha, K), beta, iterations, thinning,
                            burn_in, seed, theta.reshape(-1))
    return theta


def exact_conditional(ndk_dense, nwk_row, nk, alpha, beta, V, old):
    K = len(nk)
    p = np.zeros(K, np.float64)
    lib().oracle_exact_conditional(K, V, np.ascontiguousarray(ndk_dense, np.int32),
                                   np.ascontiguousarray(nwk_row, np.int32),
                                   np.ascontiguousarray(nk, np.int32), _alpha(alpha, K), beta, old, p)
    return p


def loglik(doc_ptr, tok, z, V, K, alpha, beta, stirling=False):
    return float(lib().oracle_loglik(len(doc_ptr) - 1, V, K, doc_ptr, tok,
                                     np.ascontiguousarray(z, np.int32), _alpha(alpha, K), beta,
                                     1 if stirling else 0))


def log_gamma_stirling(x):
    return float(lib().oracle_log_gamma_stirling(x))


def theta(z_doc, K, alpha):
    out = np.zeros(K, np.float64)
    z_doc = np.ascontiguousarray(z_doc, np.int32)
    lib().oracle_theta(K, z_doc, len(z_doc), _alpha(alpha, K), out)
    return out


def phi(nwk, nk, beta):
    V, K = nwk.shape
    out = np.zeros((K, V), np.float64)
    lib().oracle_phi(V, K, np.ascontiguousarray(nwk, np.int32).reshape(-1),
                     np.ascontiguousarray(nk, np.int32), beta, out.reshape(-1))
    return out


def digamma(x):
    return float(lib().oracle_digamma(x))


def learn_parameters(alpha, topic_doc_counts, doc_length_counts, shape=1.00001, scale=1.0, iterations=200):
    """Dirichlet.learnParameters on histograms; returns (new alpha, alphaSum)."""
    a = np.array(alpha, np.float64, copy=True)
    obs = np.ascontiguousarray(topic_doc_counts, np.int32)
    K, width = obs.shape
    lens = np.ascontiguousarray(doc_length_counts, np.int32)
    assert len(lens) == width
    s = lib().oracle_learn_parameters(a, K, obs.reshape(-1), width, lens, shape, scale, iterations)
    return a, float(s)


def learn_symmetric_concentration(count_histogram, topic_size_histogram, num_types, beta_sum):
    ch = np.ascontiguousarray(count_histogram, np.int64)
    th = np.ascontiguousarray(topic_size_histogram, np.int64)
    return float(lib().oracle_learn_symmetric_concentration(ch, len(ch), th, len(th), num_types, beta_sum))


# ---- Mallet-faithful model ------------------------------------------------------------------

class MalletModel:
    """Oracle stand-in for cc.mallet.topics.ParallelTopicModel as the reference drives it
    (cmu_ron/TrainAndPredict.java:159-171). Hyper-parameter optimisation pinned off."""

    def __init__(self, K, alpha_sum, beta, seed=None, threads=1):
        self._L = lib()
        self._h = self._L.mallet_create(K, alpha_sum, beta)
        self.K = K
        self.V = 0
        if seed is not None:
            self._L.mallet_set_random_seed(self._h, seed)
        self._L.mallet_set_num_threads(self._h, threads)

    def add_instances(self, doc_ptr, tok, V, z_init=None):
        zp = None
        if z_init is not None:
            z_init = np.ascontiguousarray(z_init, np.int32)
            zp = z_init.ctypes.data
        rc = self._L.mallet_add_instances(self._h, len(doc_ptr) - 1, V, doc_ptr, tok, zp)
        if rc != 0:
            raise ValueError("word id out of range")
        self.V = max(self.V, V)

    def estimate(self, iterations):
        self._L.mallet_estimate(self._h, iterations)

    def model_log_likelihood(self):
        return float(self._L.mallet_model_log_likelihood(self._h))

    def timers(self):
        """Cumulative wall-clock seconds of estimate(): (worker set-up, sampling phase, merge)."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._L.mallet_get_timers(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def num_tokens(self):
        return int(self._L.mallet_num_tokens(self._h))

    def assignments(self):
        z = np.zeros(self.num_tokens(), np.int32)
        self._L.mallet_get_assignments(self._h, z)
        return z

    def counts(self):
        nwk = np.zeros((self.V, self.K), np.int32)
        nk = np.zeros(self.K, np.int32)
        self._L.mallet_get_counts(self._h, nwk.reshape(-1), nk)
        return nwk, nk

    def topic_probabilities(self, doc):
        out = np.zeros(self.K, np.float64)
        self._L.mallet_get_topic_probabilities(self._h, doc, out)
        return out

    def infer(self, words, iters=100, thinning=10, burn_in=10, seed=0):
        out = np.zeros(self.K, np.float64)
        words = np.ascontiguousarray(words, np.int32)
        self._L.mallet_infer(self._h, words, len(words), iters, thinning, burn_in, seed, out)
        return out

    def close(self):
        if self._h:
            self._L.mallet_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
