"""ctypes binding of libb200lda.so — binds exactly the symbols include/b200lda.h declares.

This is the stand-in for the JVM side of the boundary (java/B200TopicModel.java binds the same
symbols through Panama FFM). There is no fallback of any kind: if the CUDA library is missing or
no sm_100 GPU is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200LDA_LIB: load another build of the same library (kernel-tuning experiments, tools/)
LIB_PATH = os.environ.get("B200LDA_LIB") or os.path.join(_HERE, "libb200lda.so")

MODE_LIVE = 0
MODE_DEFERRED = 1

OK, EINVAL, ENODEV, ENOMEM, ECUDA, ESTATE, ERANGE = 0, -1, -2, -3, -4, -5, -6
_STATUS_NAMES = {EINVAL: "EINVAL", ENODEV: "ENODEV", ENOMEM: "ENOMEM", ECUDA: "ECUDA",
                 ESTATE: "ESTATE", ERANGE: "ERANGE"}


class B200LDAError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"b200lda {_STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("num_topics", C.c_int32), ("num_types", C.c_int32),
        ("mode", C.c_int32), ("alpha_sum", C.c_double), ("beta", C.c_double), ("seed", C.c_uint64),
        ("device", C.c_int32), ("rank", C.c_int32), ("world_size", C.c_int32),
        ("table_refresh", C.c_int32), ("global_token_offset", C.c_int64),
        ("global_doc_offset", C.c_int64), ("stream", C.c_void_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("num_docs", C.c_int64), ("num_tokens", C.c_int64), ("sweeps_done", C.c_int64),
        ("kernel_launches", C.c_int64), ("tokens_sampled", C.c_int64),
        ("last_sweep_ms", C.c_double), ("last_tables_ms", C.c_double),
        ("last_sample_ms", C.c_double), ("last_finish_ms", C.c_double),
        ("mean_doc_topics", C.c_double), ("tokens_moved_last", C.c_int64),
        ("prior_bucket_last", C.c_int64), ("device_bytes", C.c_int64),
        ("smem_bytes_per_cta", C.c_int32), ("warps_per_cta", C.c_int32), ("ctas", C.c_int32),
        ("slot_capacity", C.c_int32),
        ("cum_sweeps", C.c_int64), ("cum_tables_ms", C.c_double), ("cum_sample_ms", C.c_double),
        ("cum_finish_ms", C.c_double), ("cum_tokens_moved", C.c_int64),
        ("cum_prior_bucket", C.c_int64), ("cum_doc_topics", C.c_int64),
        ("long_docs", C.c_int64), ("long_slot_capacity", C.c_int32), ("long_ctas", C.c_int32),
        ("row_classes", C.c_int32), ("table_refresh_last", C.c_int32),
        ("hot_words", C.c_int64), ("rows_refreshed_last", C.c_int64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# every symbol include/b200lda.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("b200lda_last_error", C.c_char_p, []),
    ("b200lda_abi_version", C.c_int, []),
    ("b200lda_device_count", C.c_int, []),
    ("b200lda_create", C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    ("b200lda_destroy", None, [_P]),
    ("b200lda_load_corpus", C.c_int, [_P, C.c_int64, _P, _P]),
    ("b200lda_init_assignments", C.c_int, [_P, _P]),
    ("b200lda_init_assignments_u16", C.c_int, [_P, _P]),
    ("b200lda_sweep", C.c_int, [_P, C.c_int32]),
    ("b200lda_sweep_begin", C.c_int, [_P]),
    ("b200lda_exchange_buffer", C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    ("b200lda_sweep_end", C.c_int, [_P]),
    ("b200lda_counts_sync_begin", C.c_int, [_P]),
    ("b200lda_counts_sync_end", C.c_int, [_P]),
    ("b200lda_infer", C.c_int, [_P, C.c_int64, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, _P]),
    ("b200lda_synchronize", C.c_int, [_P]),
    ("b200lda_get_stream", C.c_int, [_P, C.POINTER(_P)]),
    ("b200lda_group_allreduce", C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32]),
    ("b200lda_nccl_unique_id", C.c_int, [_P]),
    ("b200lda_comm_init", C.c_int, [_P, _P]),
    ("b200lda_group_comm_init", C.c_int, [C.POINTER(_P), C.c_int32]),
    ("b200lda_group_sync_counts", C.c_int, [C.POINTER(_P), C.c_int32]),
    ("b200lda_group_sweep", C.c_int, [C.POINTER(_P), C.c_int32, C.c_int32]),
    ("b200lda_sample_frozen", C.c_int, [_P, _P, C.c_uint32, _P]),
    ("b200lda_loglik", C.c_int, [_P, C.POINTER(C.c_double)]),
    ("b200lda_loglik_parts", C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("b200lda_check_invariants", C.c_int, [_P, _P]),
    ("b200lda_get_assignments", C.c_int, [_P, _P]),
    ("b200lda_get_assignments_u16", C.c_int, [_P, _P]),
    ("b200lda_get_nwk", C.c_int, [_P, _P]),
    ("b200lda_get_nk", C.c_int, [_P, _P]),
    ("b200lda_get_ndk_csr", C.c_int, [_P, _P, _P, _P]),
    ("b200lda_get_word_order", C.c_int, [_P, _P, _P]),
    ("b200lda_get_theta", C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    ("b200lda_get_phi", C.c_int, [_P, _P]),
    ("b200lda_set_alpha", C.c_int, [_P, _P]),
    ("b200lda_get_alpha", C.c_int, [_P, _P]),
    ("b200lda_set_beta", C.c_int, [_P, C.c_double]),
    ("b200lda_get_beta", C.c_int, [_P, C.POINTER(C.c_double)]),
    ("b200lda_hyper_begin", C.c_int, [_P, C.c_int32]),
    ("b200lda_hyper_collect", C.c_int, [_P]),
    ("b200lda_hyper_buffer", C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    ("b200lda_hyper_get", C.c_int, [_P, _P, _P]),
    ("b200lda_optimize_alpha", C.c_int, [_P]),
    ("b200lda_optimize_beta", C.c_int, [_P]),
    ("b200lda_state_size", C.c_int, [_P, C.POINTER(C.c_int64)]),
    ("b200lda_get_state", C.c_int, [_P, _P, C.c_int64]),
    ("b200lda_set_state", C.c_int, [_P, _P, C.c_int64]),
    ("b200lda_set_sweep_counter", C.c_int, [_P, C.c_int64]),
    ("b200lda_get_stats", C.c_int, [_P, C.POINTER(Stats)]),
    ("b200lda_reset_stats", C.c_int, [_P]),
    ("b200lda_host_alloc", C.c_int, [C.POINTER(_P), C.c_size_t]),
    ("b200lda_host_free", C.c_int, [_P]),
]

_lib = None


def load_library():
    """Load libb200lda.so. Raises if it has not been built — there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m ldagibbssampling_b200.build` "
            "(nvcc, sm_100a). ldagibbssampling_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Sampler:
    """One context = one GPU = one AD-LDA shard. Thin, 1:1 over the C ABI; host numpy in/out."""

    def __init__(self, num_topics, num_types, alpha_sum, beta, seed=0, mode=MODE_LIVE, device=0,
                 rank=0, world_size=1, global_token_offset=0, global_doc_offset=0, stream=None, table_refresh=0):
        self._lib = load_library()
        self._h = C.c_void_p()
        cfg = Config(struct_size=C.sizeof(Config), num_topics=num_topics, num_types=num_types,
                     mode=mode, alpha_sum=alpha_sum, beta=beta, seed=seed, device=device, rank=rank,
                     world_size=world_size, table_refresh=table_refresh, global_token_offset=global_token_offset,
                     global_doc_offset=global_doc_offset, stream=stream)
        self.K, self.V = num_topics, num_types
        self.device = device
        self.num_docs = 0
        self.num_tokens = 0
        self._check(self._lib.b200lda_create(C.byref(cfg), C.byref(self._h)))

    def _check(self, rc):
        if rc != OK:
            raise B200LDAError(rc, self._lib.b200lda_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.b200lda_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- corpus / state in ------------------------------------------------------------------
    def load_corpus(self, doc_ptr, tok_word):
        doc_ptr = np.ascontiguousarray(doc_ptr, np.int64)
        tok_word = np.ascontiguousarray(tok_word, np.int32)
        if doc_ptr.ndim != 1 or len(doc_ptr) < 1:
            raise ValueError("doc_ptr must have D+1 entries")
        if len(tok_word) != int(doc_ptr[-1]):
            raise ValueError("doc_ptr[-1] must equal len(tok_word)")
        self._check(self._lib.b200lda_load_corpus(self._h, len(doc_ptr) - 1, _ptr(doc_ptr), _ptr(tok_word)))
        self.num_docs = len(doc_ptr) - 1
        self.num_tokens = len(tok_word)

    def load_corpus_raw(self, num_docs, doc_ptr_addr, tok_word_addr, num_tokens):
        """Same call with raw host addresses (pinned staging buffers)."""
        self._check(self._lib.b200lda_load_corpus(self._h, num_docs, doc_ptr_addr, tok_word_addr))
        self.num_docs, self.num_tokens = num_docs, num_tokens

    def init_assignments(self, z=None):
        """z: None (Philox draw on the device), an int32 array (the JVM's layout) or a uint16 array
        (the device's own width: half the bytes over the bus)."""
        if z is not None and getattr(z, "dtype", None) == np.uint16:
            z = np.ascontiguousarray(z)
            if len(z) != self.num_tokens:
                raise ValueError("z must have one entry per token")
            self._check(self._lib.b200lda_init_assignments_u16(self._h, _ptr(z)))
            return
        if z is not None:
            z = np.ascontiguousarray(z, np.int32)
            if len(z) != self.num_tokens:
                raise ValueError("z must have one entry per token")
        self._check(self._lib.b200lda_init_assignments(self._h, _ptr(z)))

    def init_assignments_u16_raw(self, z_addr):
        self._check(self._lib.b200lda_init_assignments_u16(self._h, z_addr))

    def init_assignments_raw(self, z_addr):
        self._check(self._lib.b200lda_init_assignments(self._h, z_addr))

    # -- sampling ---------------------------------------------------------------------------
    def sweep(self, n=1):
        self._check(self._lib.b200lda_sweep(self._h, n))

    def sweep_begin(self):
        self._check(self._lib.b200lda_sweep_begin(self._h))

    def sweep_end(self):
        self._check(self._lib.b200lda_sweep_end(self._h))

    def counts_sync_begin(self):
        self._check(self._lib.b200lda_counts_sync_begin(self._h))

    def counts_sync_end(self):
        self._check(self._lib.b200lda_counts_sync_end(self._h))

    def infer(self, doc_ptr, tok_word, iterations=100, thinning=10, burn_in=10, seed=0):
        """TopicInferencer.getSampledDistribution for a batch of held-out documents: D x K thetas."""
        doc_ptr = np.ascontiguousarray(doc_ptr, np.int64)
        tok_word = np.ascontiguousarray(tok_word, np.int32)
        nd = len(doc_ptr) - 1
        theta = np.empty((nd, self.K), np.float64)
        self._check(self._lib.b200lda_infer(self._h, nd, _ptr(doc_ptr), _ptr(tok_word), iterations, thinning,
                                            burn_in, seed, _ptr(theta)))
        return theta

    def synchronize(self):
        self._check(self._lib.b200lda_synchronize(self._h))

    def exchange_buffer(self):
        """(device address, int32 element count) of the per-sweep delta to all-reduce."""
        p, n = C.c_void_p(), C.c_int64()
        self._check(self._lib.b200lda_exchange_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def stream(self):
        p = C.c_void_p()
        self._check(self._lib.b200lda_get_stream(self._h, C.byref(p)))
        return p.value or 0

    def sample_frozen(self, uniforms=None, sweep=1):
        z_out = np.empty(self.num_tokens, np.int32)
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, np.float32)
            if len(uniforms) != self.num_tokens:
                raise ValueError("uniforms must have one entry per token")
        self._check(self._lib.b200lda_sample_frozen(self._h, _ptr(uniforms), sweep, _ptr(z_out)))
        return z_out

    # -- state out --------------------------------------------------------------------------
    def loglik(self):
        out = C.c_double()
        self._check(self._lib.b200lda_loglik(self._h, C.byref(out)))
        return out.value

    def loglik_parts(self):
        a, b = C.c_double(), C.c_double()
        self._check(self._lib.b200lda_loglik_parts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def check_invariants(self):
        """(sum n_k, sum n_wk, #topics with column sum != n_k, sum of this shard's n_dk), computed on the device."""
        out = np.zeros(4, np.int64)
        self._check(self._lib.b200lda_check_invariants(self._h, _ptr(out)))
        return tuple(int(x) for x in out)

    def assignments(self, dtype=np.int32):
        z = np.empty(self.num_tokens, dtype)
        if np.dtype(dtype) == np.uint16:
            self._check(self._lib.b200lda_get_assignments_u16(self._h, _ptr(z)))
        else:
            self._check(self._lib.b200lda_get_assignments(self._h, _ptr(z)))
        return z

    def assignments_u16_raw(self, z_addr):
        self._check(self._lib.b200lda_get_assignments_u16(self._h, z_addr))

    # -- resumable state ---------------------------------------------------------------------
    def get_state(self) -> bytes:
        n = C.c_int64()
        self._check(self._lib.b200lda_state_size(self._h, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        self._check(self._lib.b200lda_get_state(self._h, _ptr(buf), n.value))
        return buf.tobytes()

    def set_state(self, blob: bytes):
        buf = np.frombuffer(blob, np.uint8)
        self._check(self._lib.b200lda_set_state(self._h, _ptr(buf), len(buf)))

    # -- NCCL inside the library -----------------------------------------------------------------
    def comm_init(self, unique_id: bytes):
        buf = np.frombuffer(unique_id, np.uint8)
        if len(buf) != NCCL_ID_BYTES:
            raise ValueError("an NCCL unique id has 128 bytes")
        self._check(self._lib.b200lda_comm_init(self._h, _ptr(buf)))

    def assignments_raw(self, z_addr):
        self._check(self._lib.b200lda_get_assignments(self._h, z_addr))

    def nwk(self):
        out = np.empty((self.V, self.K), np.int32)
        self._check(self._lib.b200lda_get_nwk(self._h, _ptr(out)))
        return out

    def nk(self):
        out = np.empty(self.K, np.int32)
        self._check(self._lib.b200lda_get_nk(self._h, _ptr(out)))
        return out

    def ndk_csr(self):
        row_ptr = np.zeros(self.num_docs + 1, np.int64)
        self._check(self._lib.b200lda_get_ndk_csr(self._h, _ptr(row_ptr), None, None))
        n = int(row_ptr[-1])
        topic = np.zeros(max(n, 1), np.int32)
        count = np.zeros(max(n, 1), np.int32)
        self._check(self._lib.b200lda_get_ndk_csr(self._h, _ptr(row_ptr), _ptr(topic), _ptr(count)))
        return row_ptr, topic[:n], count[:n]

    def word_order(self):
        """(word_ptr int64[V+1], word_tokens int64[N]): the word -> token CSR order."""
        word_ptr = np.zeros(self.V + 1, np.int64)
        toks = np.zeros(max(self.num_tokens, 1), np.int64)
        self._check(self._lib.b200lda_get_word_order(self._h, _ptr(word_ptr), _ptr(toks)))
        return word_ptr, toks[:self.num_tokens]

    def theta(self, doc_begin=0, doc_end=None):
        doc_end = self.num_docs if doc_end is None else doc_end
        out = np.empty((max(doc_end - doc_begin, 0), self.K), np.float64)
        self._check(self._lib.b200lda_get_theta(self._h, doc_begin, doc_end, _ptr(out)))
        return out

    def phi(self):
        out = np.empty((self.K, self.V), np.float64)
        self._check(self._lib.b200lda_get_phi(self._h, _ptr(out)))
        return out

    def set_alpha(self, alpha):
        alpha = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha, np.float64), (self.K,)))
        self._check(self._lib.b200lda_set_alpha(self._h, _ptr(alpha)))

    def alpha(self):
        out = np.empty(self.K, np.float64)
        self._check(self._lib.b200lda_get_alpha(self._h, _ptr(out)))
        return out

    def set_beta(self, beta):
        self._check(self._lib.b200lda_set_beta(self._h, beta))

    def beta(self):
        out = C.c_double()
        self._check(self._lib.b200lda_get_beta(self._h, C.byref(out)))
        return out.value

    def hyper_begin(self, width):
        self._check(self._lib.b200lda_hyper_begin(self._h, width))
        self._hyper_width = width

    def hyper_collect(self):
        self._check(self._lib.b200lda_hyper_collect(self._h))

    def hyper_buffer(self):
        p, n = C.c_void_p(), C.c_int64()
        self._check(self._lib.b200lda_hyper_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def hyper_get(self):
        w = self._hyper_width
        tdc = np.zeros((self.K, w), np.int32)
        dlc = np.zeros(w, np.int32)
        self._check(self._lib.b200lda_hyper_get(self._h, _ptr(tdc), _ptr(dlc)))
        return tdc, dlc

    def optimize_alpha(self):
        self._check(self._lib.b200lda_optimize_alpha(self._h))

    def optimize_beta(self):
        self._check(self._lib.b200lda_optimize_beta(self._h))

    def set_sweep_counter(self, n):
        self._check(self._lib.b200lda_set_sweep_counter(self._h, n))

    def stats(self):
        s = Stats()
        self._check(self._lib.b200lda_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        self._check(self._lib.b200lda_reset_stats(self._h))


def group_allreduce(samplers, which=0):
    """In-process all-reduce(sum) of the samplers' exchange (which=0) or hyper (which=1) buffers."""
    lib = load_library()
    arr = (C.c_void_p * len(samplers))(*[s._h for s in samplers])
    rc = lib.b200lda_group_allreduce(arr, len(samplers), which)
    if rc != OK:
        raise B200LDAError(rc, lib.b200lda_last_error().decode())


NCCL_ID_BYTES = 128


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = np.zeros(NCCL_ID_BYTES, np.uint8)
    rc = lib.b200lda_nccl_unique_id(_ptr(buf))
    if rc != OK:
        raise B200LDAError(rc, lib.b200lda_last_error().decode())
    return buf.tobytes()


def _group_call(name, samplers, *args):
    lib = load_library()
    arr = (C.c_void_p * len(samplers))(*[s._h for s in samplers])
    rc = getattr(lib, name)(arr, len(samplers), *args)
    if rc != OK:
        raise B200LDAError(rc, lib.b200lda_last_error().decode())


def group_comm_init(samplers):
    """One NCCL communicator set for n contexts of this process (one per GPU)."""
    _group_call("b200lda_group_comm_init", samplers)


def group_sync_counts(samplers):
    _group_call("b200lda_group_sync_counts", samplers)


def group_sweep(samplers, sweeps=1):
    _group_call("b200lda_group_sweep", samplers, sweeps)


def device_count() -> int:
    return int(load_library().b200lda_device_count())
