// b200lda.cu — host runtime + C ABI of libb200lda.so (include/b200lda.h).
//
// One context = one GPU = one AD-LDA shard. The runtime owns the device-resident corpus (CSR
// doc->token and word->token orders), the count state (n_wk dense int32 V x K, n_k, packed n_dk
// rows), the per-sweep tables and one CUDA stream on which every kernel is enqueued.
// There is no CPU fallback anywhere in this file: without an sm_100 device create() fails.
#include "../../include/b200lda.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges show up in nsys / ncu timelines when a tool is attached

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "device_common.cuh"
#include "hyper_kernels.cuh"
#include "loglik_kernels.cuh"
#include "pack_kernels.cuh"
#include "sweep_kernel.cuh"
#include "table_kernels.cuh"

using namespace b200lda;

namespace {

thread_local std::string g_err;

struct NvtxRange {  // NVTX range around a phase of the sweep (host side: enqueue time of its launches)
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(expr)                                                                          \
  do {                                                                                    \
    cudaError_t e__ = (expr);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return fail(e__ == cudaErrorMemoryAllocation ? B200LDA_ENOMEM : B200LDA_ECUDA,      \
                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define TRY(expr)                        \
  do {                                   \
    int rc__ = (expr);                   \
    if (rc__ != B200LDA_OK) return rc__; \
  } while (0)

// NCCL is loaded at run time (dlopen), never linked: a process that already holds a libnccl (torch
// bundles its own) keeps using that one, a host without NCCL still runs single-GPU models.
struct NcclApi {
  void* handle = nullptr;
  bool tried = false;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  // optional (NCCL >= 2.19): the replica allocated by NCCL and registered with the communicator, so
  // that the NVSwitch reduction (NVLS) reads and writes it in place instead of staging it
  ncclResult_t (*MemAlloc)(void**, size_t) = nullptr;
  ncclResult_t (*MemFree)(void*) = nullptr;
  ncclResult_t (*CommRegister)(const ncclComm_t, void*, size_t, void**) = nullptr;
  ncclResult_t (*CommDeregister)(const ncclComm_t, void*) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return B200LDA_OK;
  if (!g_nccl.tried) {
    g_nccl.tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy the process already uses
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
#define B200LDA_NCCL_SYM(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name))
      B200LDA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
      B200LDA_NCCL_SYM(CommInitRank, "ncclCommInitRank");
      B200LDA_NCCL_SYM(CommInitAll, "ncclCommInitAll");
      B200LDA_NCCL_SYM(CommDestroy, "ncclCommDestroy");
      B200LDA_NCCL_SYM(AllReduce, "ncclAllReduce");
      B200LDA_NCCL_SYM(GroupStart, "ncclGroupStart");
      B200LDA_NCCL_SYM(GroupEnd, "ncclGroupEnd");
      B200LDA_NCCL_SYM(GetErrorString, "ncclGetErrorString");
      B200LDA_NCCL_SYM(MemAlloc, "ncclMemAlloc");
      B200LDA_NCCL_SYM(MemFree, "ncclMemFree");
      B200LDA_NCCL_SYM(CommRegister, "ncclCommRegister");
      B200LDA_NCCL_SYM(CommDeregister, "ncclCommDeregister");
#undef B200LDA_NCCL_SYM
      if (g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommInitAll && g_nccl.CommDestroy && g_nccl.AllReduce &&
          g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.GetErrorString)
        g_nccl.handle = h;
    }
  }
  if (!g_nccl.handle) return fail(B200LDA_ENODEV, "libnccl.so.2 not found (or too old): multi-GPU exchange needs NCCL");
  return B200LDA_OK;
}

#define NCCL(expr)                                                                              \
  do {                                                                                          \
    ncclResult_t r__ = (expr);                                                                  \
    if (r__ != ncclSuccess)                                                                     \
      return fail(B200LDA_ECUDA, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

constexpr int kExchangeSlabs = 8;  // the all-reduce goes in (at most) this many slabs; slab i is applied while slab i+1 is reduced
constexpr size_t kMaxSmemPerCta = 227 * 1024;
// d_counters: [0] scheduler (long class), [1..3] last sweep {moved, prior draws, nnz sum},
// [4] nnz(n_wk) scratch, [5..7] cumulative {same three}, [8] scheduler (short class)
constexpr int kCounters = 16;  // [13] corpus checksum scratch;  // [9..11] sink for the stats of inference passes, [12] prior rows rebuilt in the last sweep
constexpr int kMaxClasses = 16;
constexpr int kMaxSegments = 32;  // chunk ranges a small corpus's sweep is cut into (full table rebuild between them)
constexpr int kMaxRefresh = 64;  // table rebuilds per LIVE sweep (b200lda_ctx::table_refresh)
constexpr int kEventPool = 256;  // sweeps whose device times can be pending before a resolve
constexpr int kPartial = 1184;   // 148 SMs x 8 blocks: fixed so the LL reduction order is fixed

int round_up32(int x) { return (x + 31) & ~31; }

PriorLayout make_layout(int K) {
  PriorLayout L{};
  int sizes[5], n = 0;
  sizes[n++] = K;
  while (sizes[n - 1] > 32 && n < 5) {
    sizes[n] = (sizes[n - 1] + 31) / 32;
    ++n;
  }
  L.nlev = n;
  int off = 0;
  for (int i = 0; i < n; ++i) {
    L.off[i] = off;
    L.size[i] = sizes[i];
    off += round_up32(sizes[i]);
  }
  L.stride = off;
  return L;
}

// Launch shape of the sampling kernel for one document class.
struct SweepShape {
  int slot_cap = 95;   // longest row (min(K, document length)) the class holds
  int rc = 0;          // kernel instance (sweep_kernel.cuh: ROWCLASS)
  int cap_tiles = 0;   // wide class: tiles of the shared-memory row
  int warps_per_cta = 8, ctas = 0, doc_chunk = 4;
  size_t smem = 0;
  bool tables_in_smem = true;
};

// A packed, device-resident set of documents: the training corpus of a context, or a batch of
// held-out documents during inference.
struct DeviceCorpus {
  int64_t D = 0, N = 0;
  int64_t* d_doc_ptr = nullptr;   // [D+1]
  int32_t* d_tok_word = nullptr;  // [N]   doc -> token order
  uint16_t* d_z = nullptr;        // [N]
  int64_t* d_row_ptr = nullptr;   // [D+1] packed n_dk row offsets (capacity min(K, L_d))
  int32_t* d_row_nnz = nullptr;   // [D]
  uint32_t* d_rows = nullptr;     // [cap_rows]
  int32_t* d_doc_order = nullptr; // [D]   long-row class first, longest first
  long long* d_word_ptr = nullptr;  // [V+1] word -> token CSR (training corpus only)
  int64_t* d_wtok = nullptr;        // [N]
  int64_t cap_docs = 0, cap_tokens = 0, cap_rows = 0, rows_total = 0;
  int max_doc_len = 0;
  bool word_order_built = false;
  // Row-width classes, widest first: class i holds the documents whose packed row needs at most
  // classes[i].shape.slot_cap slots; each class is one launch with shared memory sized to it.
  struct DocClass {
    int64_t begin = 0, end = 0;  // slice of doc_order
    int64_t tokens = 0;          // tokens of those documents
    int side_ctas = 0;           // > 0: background class, launched with this many CTAs on its own stream
    SweepShape shape;
  };
  std::vector<DocClass> classes;
  int tune_waits = 0;  // sweeps after which the host still waits for the timings (retune_background)
};

}  // namespace

struct b200lda_ctx {
  b200lda_config cfg{};
  int K = 0, V = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 0;
  int64_t device_bytes = 0;
  int64_t launches = 0;

  DeviceCorpus corp;  // training documents
  std::vector<SweepShape> shape_cache;  // launch shapes by (row class, tiles): computed once per context
  int class_streams = 0;  // see launch_sweep
  int max_ctas = 0;       // B200LDA_MAX_CTAS (experiments): cap on the sampling kernel's grid, 0 = none
  int exchange_slabs = kExchangeSlabs, apply_ctas = 1 << 20;  // B200LDA_EXCHANGE_SLABS / B200LDA_APPLY_CTAS (experiments)
  int table_refresh = 0;  // LIVE mode: table rebuilds per sweep; 0 = auto (auto_table_refresh)
  int last_refresh = 1;   // what the last sweep used
  int64_t rows_rebuilt_by_host = 0;  // rows rebuilt by the full passes between a small corpus's segments (last sweep)

  // counts + tables
  // d_nwk: [V K | K] the counts (+ in multi-shard models a K-cell tail: this shard's n_k moves);
  // d_nwk_b, same shape: DEFERRED single shard: the copy the moves go to; multi-shard: the global
  // counts of the sweep start (what the in-place exchange subtracts, and DEFERRED's frozen read copy)
  int32_t *d_nwk = nullptr, *d_nwk_b = nullptr, *d_nk = nullptr, *d_nk_delta = nullptr;
  ncclComm_t comm = nullptr;        // set by b200lda_comm_init / b200lda_group_comm_init
  bool nwk_from_nccl = false;       // d_nwk came from ncclMemAlloc (multi-shard contexts, when NCCL offers it)
  void* nwk_reg = nullptr;          // its registration with comm
  cudaStream_t apply_stream = nullptr;
  cudaEvent_t ev_slab[kExchangeSlabs] = {}, ev_applied = nullptr;
  float *d_invden = nullptr, *d_ab = nullptr, *d_prior = nullptr, *d_q = nullptr, *d_alpha_f = nullptr;
  int32_t* d_prior_sel = nullptr;   // [V] LIVE mode: which of the two copies of word w's prior row / Q_w is current
  unsigned* d_row_cursor = nullptr; // next hot word whose row the sampling warps rebuild
  int32_t* d_hot_words = nullptr;   // [V] words carrying ~90 % of the tokens, [hot_count] valid
  int hot_count = -1;               // -1: not built for the loaded corpus yet
  double *d_alpha = nullptr, *d_lg_alpha = nullptr;
  std::vector<double> alpha;
  double alpha_sum = 0.0, beta = 0.0;
  PriorLayout layout{};

  // scratch
  unsigned long long* d_counters = nullptr;
  unsigned long long* d_sched = nullptr;  // [kMaxClasses] document scheduler counter per class launch
  // side streams so the row-width classes of one sweep run concurrently (fork/join by events)
  cudaStream_t side[kMaxClasses] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxClasses] = {}, ev_bulk = nullptr;
  cudaStream_t refresh_stream = nullptr;  // LIVE mode: the table refreshers run beside the bulk launches
  cudaEvent_t ev_rfork = nullptr, ev_rjoin = nullptr;
  bool refresher_pending = false;
  int stream_priority = 0;
  const void* timed_corpus = nullptr;  // corpus whose last sweep left ev_fork / ev_bulk / ev_join to read
  int* d_bad = nullptr;
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  double* d_partial = nullptr;  // [2 * kPartial + 2]
  uint32_t* d_hist_scratch = nullptr;
  size_t hist_scratch_bytes = 0;
  int32_t* d_hyper = nullptr;  // [(K + 1) * hyper_width] topicDocCounts rows + docLengthCounts row
  int hyper_width = 0, hyper_samples = 0;
  std::vector<unsigned long long> h_len_hist;  // pack_corpus: host sides of its async copies
  std::vector<long long> h_len_start;

  // state
  bool corpus_loaded = false, assigned = false, in_sweep = false, in_sync = false;
  bool snapshot_valid = false;  // multi-shard: d_nwk_b holds the global counts of the sweep start
  int64_t sweeps_done = 0, tokens_sampled = 0;
  // per-sweep device timing: 4 events per sweep (begin, tables done, sample done, end), resolved
  // lazily at b200lda_get_stats so no sweep ever synchronises for bookkeeping
  std::vector<cudaEvent_t> ev_pool;
  int ev_pending = 0;
  cudaEvent_t* ev = nullptr;
  double cum_tables_ms = 0, cum_sample_ms = 0, cum_finish_ms = 0;
  double last_tables_ms = 0, last_sample_ms = 0, last_finish_ms = 0;
  int64_t cum_sweeps = 0;
};

namespace {

int dev_alloc(b200lda_ctx* c, void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(B200LDA_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
  }
  c->device_bytes += (int64_t)bytes;
  return B200LDA_OK;
}
template <typename T>
int dev_alloc_t(b200lda_ctx* c, T** p, size_t count) {
  return dev_alloc(c, reinterpret_cast<void**>(p), count * sizeof(T));
}
template <typename T>
void dev_free(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

int ensure_stage(b200lda_ctx* c, size_t bytes) {
  if (bytes <= c->stage_bytes) return B200LDA_OK;
  if (c->d_stage) {
    cudaFree(c->d_stage);
    c->device_bytes -= (int64_t)c->stage_bytes;
    c->d_stage = nullptr;
    c->stage_bytes = 0;
  }
  TRY(dev_alloc(c, &c->d_stage, bytes));
  c->stage_bytes = bytes;
  return B200LDA_OK;
}

int grid_for(b200lda_ctx* c, int64_t work_items, int block, int per_sm = 8) {
  int64_t blocks = (work_items + block - 1) / block;
  int64_t cap = (int64_t)c->sm_count * per_sm;
  return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

int enter(b200lda_ctx* c) {
  if (!c) return fail(B200LDA_EINVAL, "null context");
  CU(cudaSetDevice(c->cfg.device));
  return B200LDA_OK;
}

int push_alpha(b200lda_ctx* c) {
  std::vector<float> af(c->K);
  std::vector<double> lga(c->K);
  c->alpha_sum = 0.0;
  for (int k = 0; k < c->K; ++k) {
    af[k] = (float)c->alpha[k];
    lga[k] = std::lgamma(c->alpha[k]);
    c->alpha_sum += c->alpha[k];
  }
  CU(cudaMemcpyAsync(c->d_alpha_f, af.data(), sizeof(float) * c->K, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_alpha, c->alpha.data(), sizeof(double) * c->K, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_lg_alpha, lga.data(), sizeof(double) * c->K, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return B200LDA_OK;
}

// ---- sampling-kernel launch shapes ----------------------------------------------------------

template <int MODE, bool LIVE, bool TS, int RC>
int sweep_occupancy_rc(int threads, size_t smem, int* occ) {
  CU(cudaFuncSetAttribute(k_gibbs_sweep<MODE, LIVE, TS, RC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)kMaxSmemPerCta));
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_gibbs_sweep<MODE, LIVE, TS, RC>, threads, smem));
  return B200LDA_OK;
}
template <int MODE, bool LIVE, bool TS>
int sweep_occupancy(int rc, int threads, size_t smem, int* occ) {
  switch (rc) {
    case 0: return sweep_occupancy_rc<MODE, LIVE, TS, 0>(threads, smem, occ);
    case 1: return sweep_occupancy_rc<MODE, LIVE, TS, 1>(threads, smem, occ);
    case 2: return sweep_occupancy_rc<MODE, LIVE, TS, 2>(threads, smem, occ);
    default: return sweep_occupancy_rc<MODE, LIVE, TS, 3>(threads, smem, occ);
  }
}

// which kernel instance serves rows of up to this many slots (sweep_kernel.cuh: ROWCLASS)
int rowclass_for(int max_row) {
  for (int rc = 0; rc < kWideClass; ++rc)
    if (max_row <= rowclass_max_len(rc)) return rc;
  return kWideClass;
}

int occupancy_of(bool ts, int rc, int threads, size_t smem, int* out) {
  int occ[4] = {0, 0, 0, 0};
  if (ts) {
    TRY((sweep_occupancy<MODE_UPDATE, true, true>(rc, threads, smem, &occ[0])));
    TRY((sweep_occupancy<MODE_UPDATE, false, true>(rc, threads, smem, &occ[1])));
    TRY((sweep_occupancy<MODE_FROZEN, false, true>(rc, threads, smem, &occ[2])));
    TRY((sweep_occupancy<MODE_INFER, false, true>(rc, threads, smem, &occ[3])));
  } else {
    TRY((sweep_occupancy<MODE_UPDATE, true, false>(rc, threads, smem, &occ[0])));
    TRY((sweep_occupancy<MODE_UPDATE, false, false>(rc, threads, smem, &occ[1])));
    TRY((sweep_occupancy<MODE_FROZEN, false, false>(rc, threads, smem, &occ[2])));
    TRY((sweep_occupancy<MODE_INFER, false, false>(rc, threads, smem, &occ[3])));
  }
  *out = std::min(std::min(occ[0], occ[1]), std::min(occ[2], occ[3]));
  return B200LDA_OK;
}

// Picks, among {tables in shared memory, tables read through L1} x {8, 4, 2, 1 warps per CTA}, the
// shape with the most resident warps per SM (ties: shared-memory tables, then wider CTAs). At
// K = 1000 that is smem tables at 8 warps per CTA; at K = 10 000 the 120 KB of tables would leave
// one CTA per SM, so they stay in global memory. Shapes depend only on (K, row class, tiles), so
// they are computed once per context (the occupancy queries cost ~10 us each, 64 per shape).
int shape_for(b200lda_ctx* c, int max_row, int doc_chunk, int longest, SweepShape* out) {
  const int rc = rowclass_for(max_row);
  const int cap_tiles = rc == kWideClass ? max_row / 32 + 1 : 0;
  for (const SweepShape& s : c->shape_cache)
    if (s.rc == rc && s.cap_tiles == cap_tiles) {
      *out = s;
      out->slot_cap = max_row;
      out->doc_chunk = doc_chunk;
      return B200LDA_OK;
    }
  const size_t tab = sizeof(uint32_t) * (size_t)sweep_table_words(c->K);  // invden, ab, per-CTA n_k delta
  const size_t per_warp = sizeof(uint32_t) * (size_t)sweep_warp_words(c->K, cap_tiles);
  SweepShape best;
  int best_warps = 0;  // in quarter-warps: tables in global memory count 3/4 (every move then
                       // pays two global atomics on the K topic totals instead of two shared ones)
  for (int ts = 1; ts >= 0; --ts) {
    for (int wpc = 8; wpc >= 1; wpc >>= 1) {
      const size_t need = (ts ? tab : 0) + per_warp * wpc;
      if (need > kMaxSmemPerCta) continue;
      int occ = 0;
      TRY(occupancy_of(ts != 0, rc, wpc * 32, need, &occ));
      if (occ * wpc * (ts ? 4 : 3) > best_warps) {
        best_warps = occ * wpc * (ts ? 4 : 3);
        best.rc = rc;
        best.cap_tiles = cap_tiles;
        best.tables_in_smem = ts != 0;
        best.warps_per_cta = wpc;
        best.smem = need;
        best.ctas = c->sm_count * occ;  // persistent grid: every CTA resident, documents fetched dynamically
      }
    }
  }
  if (best_warps == 0)
    return fail(B200LDA_ERANGE, "document rows of %d slots do not fit shared memory (K=%d, longest doc=%d)",
                max_row, c->K, longest);
  c->shape_cache.push_back(best);
  *out = best;
  out->slot_cap = max_row;
  out->doc_chunk = doc_chunk;
  return B200LDA_OK;
}

// Row-width classes: rows of up to 95 / 159 / 255 slots (min(K, document length)) live in registers
// (3 / 5 / 8 tiles, one kernel instance each so the narrow classes keep their register allocation);
// longer rows live in shared memory, in classes of 511, 1023, ... slots up to the widest row, so
// the long tail does not dictate everyone's shared memory and occupancy.
// len_ge[L] = number of documents with at least L tokens (documents are ordered longest first).
int configure_sweep(b200lda_ctx* c, DeviceCorpus& cp, const std::vector<int64_t>& len_ge) {
  cp.classes.clear();
  cp.tune_waits = 8;
  if (c->timed_corpus == &cp) c->timed_corpus = nullptr;  // the classes the pending timings describe are gone
  const int widest = std::min(c->K, std::max(1, cp.max_doc_len));
  std::vector<int> caps;
  for (int rc = 0; rc < kWideClass && rowclass_max_len(rc) < widest; ++rc) caps.push_back(rowclass_max_len(rc));
  for (int cap = 511; cap < widest; cap = 2 * cap + 1) caps.push_back(cap);  // 511 slots = 16 tiles still run 4 CTAs per SM
  caps.push_back(widest);
  auto docs_with_row_above = [&](int cap) -> int64_t {  // rows are min(len, K) slots wide
    if (cap >= c->K || cap + 1 >= (int)len_ge.size()) return 0;
    return len_ge[(size_t)cap + 1];
  };
  // tokens in documents of at least L tokens = L * len_ge[L] + sum_{M > L} len_ge[M]
  std::vector<int64_t> suffix(len_ge.size() + 1, 0);
  for (int L = (int)len_ge.size() - 1; L >= 0; --L) suffix[(size_t)L] = suffix[(size_t)L + 1] + len_ge[(size_t)L];
  auto tokens_with_row_above = [&](int cap) -> int64_t {
    if (cap >= c->K || cap + 1 >= (int)len_ge.size()) return 0;
    return (int64_t)(cap + 1) * len_ge[(size_t)cap + 1] + suffix[(size_t)cap + 2];
  };
  const int64_t all_tokens = suffix[1];
  for (int i = (int)caps.size() - 1; i >= 0; --i) {
    DeviceCorpus::DocClass dc;
    dc.begin = docs_with_row_above(caps[i]);
    dc.end = i == 0 ? cp.D : docs_with_row_above(caps[i - 1]);
    if (dc.end <= dc.begin) continue;
    dc.tokens = (i == 0 ? all_tokens : tokens_with_row_above(caps[i - 1])) - tokens_with_row_above(caps[i]);
    TRY(shape_for(c, caps[i], caps[i] > 255 ? 1 : 4, cp.max_doc_len, &dc.shape));
    cp.classes.push_back(dc);
  }
  // Background classes. A class with too few documents to keep every resident warp busy several
  // times over (the long tail: each document is one warp's serial chain) is bound by its longest
  // documents, not by throughput. Alone it would idle the GPU; on a full-size grid beside the bulk
  // it would take a register-file slot of 8 bulk warps on every SM for 2 of its own. So it runs on
  // its own stream with just enough warps to finish in about half the time the bulk needs:
  //   warps = 2 * (per-warp token latency ~2.5 us) * (bulk rate ~2.4e9 tokens/s) * share of tokens
  // as a first guess (doubled below), corrected sweep by sweep from the measured finish times.
  // Measured on the C4 shape (profiles/r01_tuning.md): full-size grids on forked streams 330 ms,
  // everything on one stream 287 ms, at 738 M tokens; 36.1 vs 38.7 ms at 90 M tokens.
  const size_t nc = cp.classes.size();
  for (size_t i = 0; i < nc; ++i) {
    DeviceCorpus::DocClass& dc = cp.classes[i];
    const int64_t resident_warps = (int64_t)dc.shape.ctas * dc.shape.warps_per_cta;
    const bool background = i + 1 < nc && (dc.end - dc.begin) < 4 * resident_warps;
    if (!background) continue;
    const double share = all_tokens > 0 ? (double)dc.tokens / (double)all_tokens : 1.0;
    const int64_t warps = (int64_t)std::ceil(24000.0 * share);  // first guess, 2x margin; retune_background follows the clock
    const int64_t ctas = (warps + dc.shape.warps_per_cta - 1) / dc.shape.warps_per_cta;
    dc.side_ctas = (int)std::max<int64_t>(c->sm_count / 8, std::min<int64_t>(dc.shape.ctas, ctas));
  }
  return B200LDA_OK;
}

// ---- corpus packing ---------------------------------------------------------------------------

void free_corpus(DeviceCorpus& cp) {
  dev_free(cp.d_doc_ptr);
  dev_free(cp.d_tok_word);
  dev_free(cp.d_z);
  dev_free(cp.d_row_ptr);
  dev_free(cp.d_row_nnz);
  dev_free(cp.d_rows);
  dev_free(cp.d_doc_order);
  dev_free(cp.d_word_ptr);
  dev_free(cp.d_wtok);
  cp.cap_docs = cp.cap_tokens = cp.cap_rows = 0;
}

// Uploads a CSR document set and plans it ON THE DEVICE: CSR validation, document-length
// histogram, packed n_dk row offsets (capacity min(K, L_d), two-level scan), visiting order
// (longest first, counting sort) and word-id validation. Only the 64 K-bin length histogram comes
// back to the host, which derives the row-width classes from it.
int pack_corpus(b200lda_ctx* c, DeviceCorpus& cp, int64_t num_docs, const int64_t* doc_ptr, const int32_t* tok_word) {
  if (num_docs < 0 || !doc_ptr) return fail(B200LDA_EINVAL, "bad corpus arguments");
  if (doc_ptr[0] != 0) return fail(B200LDA_EINVAL, "doc_ptr[0] must be 0");
  const int64_t N = doc_ptr[num_docs];
  if (N < 0) return fail(B200LDA_EINVAL, "doc_ptr[num_docs] is negative");
  if (N > 0 && !tok_word) return fail(B200LDA_EINVAL, "tok_word is null");
  cp.word_order_built = false;
  if (num_docs > cp.cap_docs) {
    dev_free(cp.d_doc_ptr);
    dev_free(cp.d_row_ptr);
    dev_free(cp.d_row_nnz);
    dev_free(cp.d_doc_order);
    TRY(dev_alloc_t(c, &cp.d_doc_order, (size_t)num_docs));
    TRY(dev_alloc_t(c, &cp.d_doc_ptr, (size_t)num_docs + 1));
    TRY(dev_alloc_t(c, &cp.d_row_ptr, (size_t)num_docs + 1));
    TRY(dev_alloc_t(c, &cp.d_row_nnz, (size_t)num_docs));
    cp.cap_docs = num_docs;
  }
  if (N > cp.cap_tokens) {
    dev_free(cp.d_tok_word);
    dev_free(cp.d_z);
    dev_free(cp.d_wtok);
    TRY(dev_alloc_t(c, &cp.d_tok_word, (size_t)N));
    TRY(dev_alloc_t(c, &cp.d_z, (size_t)N));
    cp.cap_tokens = N;
  }
  cp.D = num_docs;
  cp.N = N;
  // The plan needs only doc_ptr: it goes first, and the token stream (the bulk of the bytes) is
  // copied while the host derives the row-width classes from the length histogram.
  CU(cudaMemcpyAsync(cp.d_doc_ptr, doc_ptr, sizeof(int64_t) * (num_docs + 1), cudaMemcpyHostToDevice, c->stream));

  // scratch: [len_hist 65537 | cursor 65537 | len_start 65537 | bad_doc | block sums/offsets ...]
  constexpr int kBins = 65537;
  const int64_t nblocks = (num_docs + kScanBlock - 1) / kScanBlock;
  const size_t words64 = 3 * (size_t)kBins + 1 + 2 * (size_t)(nblocks + 1);
  TRY(ensure_stage(c, sizeof(unsigned long long) * words64));
  unsigned long long* d_hist = reinterpret_cast<unsigned long long*>(c->d_stage);
  unsigned long long* d_cursor = d_hist + kBins;
  long long* d_len_start = reinterpret_cast<long long*>(d_cursor + kBins);
  long long* d_bad_doc = d_len_start + kBins;
  unsigned long long* d_block_sum = reinterpret_cast<unsigned long long*>(d_bad_doc + 1);
  long long* d_block_off = reinterpret_cast<long long*>(d_block_sum + nblocks + 1);
  CU(cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * 2 * kBins, c->stream));
  const long long no_bad = 0x7fffffffffffffffLL;
  CU(cudaMemcpyAsync(d_bad_doc, &no_bad, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(c->d_bad, 0, sizeof(int), c->stream));
  std::vector<unsigned long long>& h_hist = c->h_len_hist;
  h_hist.assign((size_t)kBins, 0);
  long long bad_doc = no_bad;
  int bad_word = 0;
  if (num_docs > 0) {
    k_doc_lengths<<<grid_for(c, num_docs, 256), 256, 0, c->stream>>>(num_docs, cp.d_doc_ptr, d_hist, d_bad_doc);
    k_row_block_sums<<<(unsigned)nblocks, kScanBlock, 0, c->stream>>>(num_docs, c->K, cp.d_doc_ptr, d_block_sum);
    k_exclusive_scan_u64<<<1, 1024, 0, c->stream>>>((int)nblocks, d_block_sum, d_block_off);
    k_row_ptr_apply<<<(unsigned)nblocks, kScanBlock, 0, c->stream>>>(num_docs, c->K, cp.d_doc_ptr, d_block_off, cp.d_row_ptr);
    c->launches += 4;
  } else {
    CU(cudaMemsetAsync(cp.d_row_ptr, 0, sizeof(int64_t), c->stream));
  }
  CU(cudaMemcpyAsync(h_hist.data(), d_hist, sizeof(unsigned long long) * kBins, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&bad_doc, d_bad_doc, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));  // the plan's inputs (a few hundred KB); the token stream has not moved yet
  if (bad_doc != no_bad) {
    const int64_t len = doc_ptr[bad_doc + 1] - doc_ptr[bad_doc];
    if (len < 0) return fail(B200LDA_EINVAL, "doc_ptr is not monotone at document %lld", bad_doc);
    return fail(B200LDA_ERANGE, "document %lld has %lld tokens (limit 65535)", bad_doc, (long long)len);
  }
  if (N > 0) {
    CU(cudaMemcpyAsync(cp.d_tok_word, tok_word, sizeof(int32_t) * N, cudaMemcpyHostToDevice, c->stream));
    k_validate_words<<<grid_for(c, N, 256), 256, 0, c->stream>>>(N, c->V, cp.d_tok_word, c->d_bad);
    c->launches += 1;
    CU(cudaMemcpyAsync(&bad_word, c->d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  }

  // host (while the token stream is copied): 64 K bins -> longest document, row capacity, len_ge,
  // class boundaries, order offsets
  int max_len = 0;
  int64_t rows_total = 0;
  for (int L = 0; L < kBins; ++L)
    if (h_hist[(size_t)L]) {
      max_len = L;
      rows_total += (int64_t)h_hist[(size_t)L] * std::min(L, c->K);
    }
  cp.max_doc_len = max_len;
  cp.rows_total = rows_total;
  std::vector<int64_t> len_ge((size_t)max_len + 2, 0);
  for (int L = max_len; L >= 0; --L) len_ge[(size_t)L] = len_ge[(size_t)L + 1] + (int64_t)h_hist[(size_t)L];
  std::vector<long long>& len_start = c->h_len_start;  // documents longer than L come first; outlives the async copy below
  len_start.assign((size_t)kBins, 0);
  for (int L = 0; L <= max_len; ++L) len_start[(size_t)L] = len_ge[(size_t)L + 1];
  if (rows_total > cp.cap_rows) {
    dev_free(cp.d_rows);
    TRY(dev_alloc_t(c, &cp.d_rows, (size_t)rows_total));
    cp.cap_rows = rows_total;
  }
  if (num_docs > 0) {
    CU(cudaMemcpyAsync(d_len_start, len_start.data(), sizeof(long long) * kBins, cudaMemcpyHostToDevice, c->stream));
    k_doc_order_scatter<<<grid_for(c, num_docs, 256), 256, 0, c->stream>>>(num_docs, cp.d_doc_ptr, d_len_start, d_cursor,
                                                                         cp.d_doc_order);
    c->launches += 1;
  }
  CU(cudaGetLastError());
  TRY(configure_sweep(c, cp, len_ge));
  CU(cudaStreamSynchronize(c->stream));  // the caller's buffers are borrowed for the duration of the call only
  if (bad_word) return fail(B200LDA_ERANGE, "tok_word holds a word id outside [0, %d)", c->V);
  return B200LDA_OK;
}

// word -> token CSR of a packed corpus (device counting sort), built on demand: the sampler itself
// never needs it (n_wk is counted straight from the doc order), b200lda_get_word_order does.
int build_word_order(b200lda_ctx* c, DeviceCorpus& cp) {
  if (cp.word_order_built) return B200LDA_OK;
  if (!cp.d_word_ptr) TRY(dev_alloc_t(c, &cp.d_word_ptr, (size_t)c->V + 1));
  if (!cp.d_wtok && cp.N > 0) TRY(dev_alloc_t(c, &cp.d_wtok, (size_t)cp.cap_tokens));
  TRY(ensure_stage(c, sizeof(unsigned long long) * (size_t)c->V));
  unsigned long long* d_wcount = reinterpret_cast<unsigned long long*>(c->d_stage);
  CU(cudaMemsetAsync(d_wcount, 0, sizeof(unsigned long long) * c->V, c->stream));
  CU(cudaMemsetAsync(c->d_bad, 0, sizeof(int), c->stream));
  if (cp.N > 0) k_word_hist<<<grid_for(c, cp.N, 256), 256, 0, c->stream>>>(cp.N, c->V, cp.d_tok_word, d_wcount, c->d_bad);
  k_exclusive_scan_u64<<<1, 1024, 0, c->stream>>>(c->V, d_wcount, cp.d_word_ptr);
  CU(cudaMemsetAsync(d_wcount, 0, sizeof(unsigned long long) * c->V, c->stream));
  if (cp.N > 0) k_word_scatter<<<grid_for(c, cp.N, 256), 256, 0, c->stream>>>(cp.N, cp.d_tok_word, cp.d_word_ptr, d_wcount, cp.d_wtok);
  c->launches += 3;
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  cp.word_order_built = true;
  return B200LDA_OK;
}

// Packed ascending-topic n_dk rows from cp.d_z.
int build_doc_rows(b200lda_ctx* c, DeviceCorpus& cp) {
  if (cp.D == 0) return B200LDA_OK;
  const int wpc = 8;
  const size_t hist_smem = sizeof(uint32_t) * (size_t)c->K * wpc;
  const int grid = grid_for(c, cp.D * 32, 256, hist_smem <= 48 * 1024 ? 8 : 2);
  if (hist_smem <= 96 * 1024) {
    CU(cudaFuncSetAttribute(k_build_doc_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
    k_build_doc_rows<<<grid, wpc * 32, hist_smem, c->stream>>>(cp.D, c->K, cp.d_doc_ptr, cp.d_z, cp.d_row_ptr,
                                                             cp.d_row_nnz, cp.d_rows, nullptr);
  } else {
    const size_t need = sizeof(uint32_t) * (size_t)c->K * wpc * grid;
    if (need > c->hist_scratch_bytes) {
      if (c->d_hist_scratch) c->device_bytes -= (int64_t)c->hist_scratch_bytes;
      dev_free(c->d_hist_scratch);
      TRY(dev_alloc(c, reinterpret_cast<void**>(&c->d_hist_scratch), need));
      c->hist_scratch_bytes = need;
    }
    k_build_doc_rows<<<grid, wpc * 32, 0, c->stream>>>(cp.D, c->K, cp.d_doc_ptr, cp.d_z, cp.d_row_ptr, cp.d_row_nnz,
                                                     cp.d_rows, c->d_hist_scratch);
  }
  c->launches += 1;
  CU(cudaGetLastError());
  return B200LDA_OK;
}

// ---- per-sweep pieces ---------------------------------------------------------------------------

// use_sel: LIVE training sweeps keep two copies of every word's row (d_prior_sel says which is
// current); the build writes the other copy and flips. Frozen / inference passes build copy 0.
int build_tables(b200lda_ctx* c, bool use_sel = false) {
  const float beta_f = (float)c->beta;
  const float vbeta = (float)c->V * beta_f;
  k_topic_tables<<<(c->K + 255) / 256, 256, 0, c->stream>>>(c->K, c->d_nk, c->d_alpha_f, vbeta, c->d_invden, c->d_ab);
  int32_t* sel = use_sel ? c->d_prior_sel : nullptr;
  if (!use_sel && c->d_prior_sel) CU(cudaMemsetAsync(c->d_prior_sel, 0, sizeof(int32_t) * c->V, c->stream));
  k_prior_rows<<<grid_for(c, (int64_t)c->V * 32, 256), 256, 0, c->stream>>>(c->V, c->K, c->d_nwk, c->d_ab, beta_f,
                                                                            c->layout, c->d_prior, c->d_q, sel);
  c->launches += 2;
  CU(cudaGetLastError());
  return B200LDA_OK;
}

SweepParams sweep_params(b200lda_ctx* c, const DeviceCorpus& cp, const int32_t* nwk_read, int32_t* nwk_write,
                         uint32_t sweep) {
  SweepParams p{};
  p.doc_order = cp.d_doc_order;
  p.doc_ptr = cp.d_doc_ptr;
  p.tok_word = cp.d_tok_word;
  p.z = cp.d_z;
  p.z_out = nullptr;
  p.row_ptr = cp.d_row_ptr;
  p.row_nnz = cp.d_row_nnz;
  p.rows = cp.d_rows;
  p.nwk_read = nwk_read;
  p.nwk_write = nwk_write;
  p.nk_delta = c->d_nk_delta;
  p.invden = c->d_invden;
  p.ab = c->d_ab;
  p.prior = c->d_prior;
  p.q = c->d_q;
  p.uniforms = nullptr;
  p.prior_sel = nullptr;
  p.hot_words = nullptr;
  p.hot_count = 0;
  p.row_cursor = c->d_row_cursor;
  p.refresh_rows = 0u;
  p.sampler_warps = 0;
  p.V = c->V;
  p.layout = c->layout;
  p.K = c->K;
  p.beta_f = (float)c->beta;
  p.seed = c->cfg.seed;
  p.sweep = sweep;
  p.global_tok_off = c->cfg.global_token_offset;
  p.stats = c->d_counters + 1;
  p.refresh_count = c->d_counters + 12;
  return p;
}

// LIVE mode reads the prior bucket (49 % of the draws at K = 1000, alpha_k = 0.1) from per-word
// prefix tables. Built once per sweep the chain mixes like Mallet with twice as many threads
// (LL/token 4 % behind the single chain at sweep 25 on the C4-shaped 20 k-document sample,
// profiles/r02_ll_parity.md). So in LIVE mode a small kernel beside every bulk launch (launch_class;
// sweep_kernel.cuh: k_prior_refresher) keeps rebuilding the rows of the HOT words (those that carry
// ~90 % of the tokens) from the live counts while the sweep runs, each `table_refresh` times per
// sweep in all. A rebuilding warp is busy ~40 us per row (it waits on DRAM for the word's n_wk
// row), i.e. the refreshers need  rows/s x 40 us  warps.  Auto: up to 16 rebuilds per sweep, as many
// as 3 % of the grid's CTAs manage; small corpora take the segmented form instead (launch_sweep).
constexpr double kRefreshRowSeconds = 40.0e-6;
constexpr double kSweepTokensPerSecond = 3.0e9;
int auto_table_refresh(const b200lda_ctx* c, const DeviceCorpus& cp) {
  if (c->hot_count <= 0 || cp.N <= 0 || cp.D <= 0) return 1;
  // The tables only matter for the draws that land in the prior bucket: about
  // alpha_sum / (alpha_sum + mean document length) of them (measured 0.49 / 0.22 / 0.03 / 0.02 on
  // C4 / C3 / C2 / C1 against 0.53 / 0.23 / 0.03 / 0.02 from this formula). Below a tenth the age
  // of the tables is invisible in the chain and the rebuilds are not worth their cost.
  const double mean_len = (double)cp.N / (double)cp.D;
  if (c->alpha_sum / (c->alpha_sum + mean_len) < 0.1) return 1;
  return 16;  // the refreshers' share of the grid (large corpora) / 8 segments (small ones) bound what is reached
}

// The hot words of the loaded corpus: the most frequent words that together carry 90 % of the
// tokens (at most V/4 of them). Word frequencies = row sums of n_wk; the threshold comes from a
// 32-bin log2 histogram on the host.
int build_hot_words(b200lda_ctx* c) {
  if (c->hot_count >= 0) return B200LDA_OK;
  if (!c->d_hot_words) TRY(dev_alloc_t(c, &c->d_hot_words, (size_t)c->V));
  TRY(ensure_stage(c, sizeof(int32_t) * (size_t)c->V));
  int32_t* d_cnt = reinterpret_cast<int32_t*>(c->d_stage);
  k_word_counts<<<grid_for(c, (int64_t)c->V * 32, 256), 256, 0, c->stream>>>(c->V, c->K, c->d_nwk, d_cnt);
  std::vector<int32_t> cnt((size_t)c->V);
  CU(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int32_t) * c->V, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  int64_t mass[33] = {0}, words[33] = {0}, total = 0;
  for (int32_t v : cnt) {
    int b = 0;
    while (b < 32 && (1ll << (b + 1)) <= (int64_t)v) ++b;  // 2^b <= v < 2^(b+1)
    if (v > 0) {
      mass[b] += v;
      words[b] += 1;
      total += v;
    }
  }
  int threshold = 1;
  int64_t m = 0, n = 0;
  for (int b = 32; b >= 0; --b) {
    if (words[b] == 0) continue;
    if (n > 0 && (n + words[b] > c->V / 4 || m >= (total * 9) / 10)) break;
    m += mass[b];
    n += words[b];
    threshold = (int)std::min<int64_t>(1ll << b, 0x7fffffff);
  }
  CU(cudaMemsetAsync(c->d_bad, 0, sizeof(int), c->stream));
  k_hot_words<<<(c->V + 255) / 256, 256, 0, c->stream>>>(c->V, d_cnt, threshold, c->d_hot_words, c->d_bad);
  c->launches += 2;
  int h = 0;
  CU(cudaMemcpyAsync(&h, c->d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  c->hot_count = h;
  return B200LDA_OK;
}

template <int MODE, bool LIVE>
int launch_class(b200lda_ctx* c, SweepParams p, const SweepShape& sh, int64_t begin, int64_t end,
                 unsigned long long* counter, cudaStream_t stream, int max_ctas = 0, int refresh_passes = 0,
                 unsigned* cursor = nullptr, int seg = 0, int nseg = 1) {
  if (end <= begin) return B200LDA_OK;
  p.order_begin = begin;
  p.order_end = end;
  p.cap_tiles = sh.cap_tiles;
  p.doc_chunk = sh.doc_chunk;
  p.doc_counter = counter;
  // seg / nseg: the launch covers the seg-th of nseg equal ranges of the class's scheduler chunks
  // (chunks are strided through the longest-first order, so every range is a cross-section)
  const int64_t all_chunks = (end - begin + sh.doc_chunk - 1) / sh.doc_chunk;
  p.chunk_begin = (unsigned long long)(all_chunks * seg / nseg);
  p.chunk_end = (unsigned long long)(all_chunks * (seg + 1) / nseg);
  if (p.chunk_end <= p.chunk_begin) return B200LDA_OK;
  const int64_t warps_needed = (int64_t)(p.chunk_end - p.chunk_begin);
  if (c->max_ctas > 0) max_ctas = max_ctas > 0 ? std::min(max_ctas, c->max_ctas) : c->max_ctas;
  const int grid_cap = max_ctas > 0 ? std::min(max_ctas, sh.ctas) : sh.ctas;
  int ctas = (int)std::max<int64_t>(
      1, std::min<int64_t>(grid_cap, (warps_needed + sh.warps_per_cta - 1) / sh.warps_per_cta));
  if (LIVE && MODE == MODE_UPDATE && refresh_passes > 0 && c->hot_count > 0 && c->corp.N > 0) {
    // This launch's share of the sweep's row rebuilds and the refresher CTAs (8 warps each) that
    // keep up with it. A refresher CTA takes the place of a sampler CTA, so its cost is its share
    // of the grid: the automatic choice spends 3 % of the CTAs and lets the refreshers do what
    // they can of 16 passes; an explicit table_refresh sizes them to reach it (a quarter of the
    // grid at most). A sampling grid that leaves half the GPU empty (small corpora) has room.
    const bool automatic = c->table_refresh == 0;
    const double share = (double)(end - begin) / (double)std::max<int64_t>(1, c->corp.D);
    const double rows = (double)refresh_passes * (double)c->hot_count * share;
    const double launch_s = share * (double)c->corp.N / kSweepTokensPerSecond;
    const double row_s = kRefreshRowSeconds * std::max(1.0, (double)c->K / 1000.0);
    const int want = (int)std::ceil(rows * row_s / std::max(launch_s, 1e-6) / 8.0);
    const int room = std::min(grid_cap - ctas, grid_cap / 4);
    const int cap = automatic ? std::max(room, (grid_cap * 3 + 99) / 100) : std::max(room, grid_cap / 4);
    const int refreshers = std::max(1, std::min(want, cap));
    if (rows >= 1.0 && grid_cap - refreshers >= 1) {
      SweepParams pr = p;
      pr.refresh_rows = (unsigned)std::min(rows, 4.0e9);
      pr.row_cursor = cursor;
      ctas = std::min(ctas, grid_cap - refreshers);
      pr.sampler_warps = ctas * sh.warps_per_cta;
      CU(cudaEventRecord(c->ev_rfork, stream));  // after the previous class's samplers, beside this class's
      CU(cudaStreamWaitEvent(c->refresh_stream, c->ev_rfork, 0));
      k_prior_refresher<<<refreshers, 256, 0, c->refresh_stream>>>(pr, (unsigned long long)warps_needed);
      c->launches += 1;
      c->refresher_pending = true;
      CU(cudaGetLastError());
    }
  }
  const int threads = sh.warps_per_cta * 32;
#define B200LDA_LAUNCH(TS, RC) k_gibbs_sweep<MODE, LIVE, TS, RC><<<ctas, threads, sh.smem, stream>>>(p)
  switch (sh.rc * 2 + (sh.tables_in_smem ? 1 : 0)) {
    case 0: B200LDA_LAUNCH(false, 0); break;
    case 1: B200LDA_LAUNCH(true, 0); break;
    case 2: B200LDA_LAUNCH(false, 1); break;
    case 3: B200LDA_LAUNCH(true, 1); break;
    case 4: B200LDA_LAUNCH(false, 2); break;
    case 5: B200LDA_LAUNCH(true, 2); break;
    case 6: B200LDA_LAUNCH(false, 3); break;
    default: B200LDA_LAUNCH(true, 3); break;
  }
#undef B200LDA_LAUNCH
  c->launches += 1;
  CU(cudaGetLastError());
  return B200LDA_OK;
}

// Background grids follow the clock: after a sweep, compare when each background class finished
// with when the bulk did (events of that sweep, read only if already complete: no synchronisation)
// and double a grid that finished late, shrink one that finished in under a third of the bulk's
// time. The first guess (configure_sweep) assumes K ~ 1000; at K = 10 000 a token of a long document
// costs several times more (three dependent table levels) and the guess is 8x too small.
// Sweeps are enqueued without host synchronisation, so the previous sweep's events are normally
// still pending here; for the first tune_waits sweeps of a corpus the host waits for them (the
// grids settle within a handful of sweeps), afterwards it only uses timings that happen to be ready.
void retune_background(b200lda_ctx* c, DeviceCorpus& cp) {
  if (c->timed_corpus != &cp || c->class_streams != 0) return;
  c->timed_corpus = nullptr;
  const size_t n = cp.classes.size();
  bool any_background = false;
  for (size_t i = 0; i + 1 < n; ++i) any_background = any_background || cp.classes[i].side_ctas > 0;
  if (!any_background) return;
  if (cudaEventQuery(c->ev_bulk) != cudaSuccess) {
    (void)cudaGetLastError();
    if (cp.tune_waits <= 0) return;
    --cp.tune_waits;
    if (cudaEventSynchronize(c->ev_bulk) != cudaSuccess) return;
    for (size_t i = 0; i + 1 < n; ++i)
      if (cp.classes[i].side_ctas > 0 && cudaEventSynchronize(c->ev_join[i]) != cudaSuccess) return;
  }
  float t_bulk = 0.0f;
  if (cudaEventElapsedTime(&t_bulk, c->ev_fork, c->ev_bulk) != cudaSuccess) return;
  for (size_t i = 0; i + 1 < n; ++i) {
    DeviceCorpus::DocClass& dc = cp.classes[i];
    if (dc.side_ctas <= 0) continue;
    float t_bg = 0.0f;
    if (cudaEventQuery(c->ev_join[i]) != cudaSuccess ||
        cudaEventElapsedTime(&t_bg, c->ev_fork, c->ev_join[i]) != cudaSuccess)
      continue;
    if (t_bg > 0.9f * t_bulk)
      dc.side_ctas = std::min(dc.shape.ctas, std::max(dc.side_ctas + 1, dc.side_ctas * 2));
    else if (t_bg < 0.35f * t_bulk)
      dc.side_ctas = std::max(std::max(1, c->sm_count / 8), dc.side_ctas * 3 / 4);
  }
  (void)cudaGetLastError();
}

// One pass over a corpus = one launch per row-width class. The bulk classes run one after the
// other on the context's stream, each filling the GPU. The long-tail classes (few, long documents:
// each a long serial chain on one warp) run in the background on side streams forked off the
// context's stream with a reduced grid, and the sweep joins them: a long document's latency is
// hidden behind the bulk, and its CTAs do not take bulk CTAs' register-file slots on every SM.
template <int MODE, bool LIVE>
int launch_sweep(b200lda_ctx* c, DeviceCorpus& cp, const SweepParams& p, int refresh_passes = 0) {
  retune_background(c, cp);
  if (MODE != MODE_INFER) CU(cudaMemsetAsync(c->d_counters, 0, sizeof(unsigned long long) * 4, c->stream));
  if (MODE == MODE_UPDATE) CU(cudaMemsetAsync(c->d_counters + 12, 0, sizeof(unsigned long long), c->stream));
  CU(cudaMemsetAsync(c->d_sched, 0, sizeof(unsigned long long) * kMaxClasses * kMaxSegments, c->stream));
  CU(cudaMemsetAsync(c->d_row_cursor, 0, sizeof(unsigned) * kMaxClasses, c->stream));
  const size_t n = cp.classes.size();
  if (n == 0) return B200LDA_OK;
  // LIVE table refresh, two forms. Large corpora: the refresher kernel beside every bulk launch
  // (launch_class). Small corpora (a sweep of a few ms: launches too short for a second kernel to be
  // reliably scheduled beside them, and a full table rebuild costs next to nothing): the bulk classes
  // run in `segments` chunk ranges with a full rebuild of the tables from the live counts between
  // them - the rebuild writes every word's other copy and flips its selector, so the background
  // classes still running on their streams keep reading consistent rows.
  int segments = 1;
  if (LIVE && MODE == MODE_UPDATE && refresh_passes > 0 && (double)cp.N / kSweepTokensPerSecond < 5.0e-3) {
    segments = std::min(refresh_passes + 1, c->table_refresh > 0 ? kMaxSegments : 8);
    refresh_passes = 0;
  }
  if (n > 1) CU(cudaEventRecord(c->ev_fork, c->stream));
  // class_streams (B200LDA_CLASS_STREAMS, experiments): 0 = the policy above (default);
  // 1 = everything in sequence; 2 = every class forked at full size.
  std::vector<bool> forked(n, false);
  for (size_t i = 0; i + 1 < n; ++i) {
    const DeviceCorpus::DocClass& dc = cp.classes[i];
    forked[i] = c->class_streams == 2 || (c->class_streams == 0 && dc.side_ctas > 0);
    if (!forked[i]) continue;
    if (!c->side[i]) CU(cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, c->stream_priority));
    CU(cudaStreamWaitEvent(c->side[i], c->ev_fork, 0));
    TRY((launch_class<MODE, LIVE>(c, p, dc.shape, dc.begin, dc.end, c->d_sched + i * kMaxSegments, c->side[i],
                                  c->class_streams == 0 ? dc.side_ctas : 0)));
    CU(cudaEventRecord(c->ev_join[i], c->side[i]));
  }
  for (int seg = 0; seg < segments; ++seg) {
    if (seg > 0) {
      TRY(build_tables(c, true));
      if (MODE == MODE_UPDATE) c->rows_rebuilt_by_host += c->V;
    }
    for (size_t i = 0; i < n; ++i) {
      if (i + 1 < n && forked[i]) continue;
      const DeviceCorpus::DocClass& dc = cp.classes[i];
      TRY((launch_class<MODE, LIVE>(c, p, dc.shape, dc.begin, dc.end, c->d_sched + i * kMaxSegments + seg, c->stream, 0,
                                    refresh_passes, c->d_row_cursor + i, seg, segments)));
    }
  }
  if (n > 1) {
    CU(cudaEventRecord(c->ev_bulk, c->stream));
    c->timed_corpus = &cp;
  }
  for (size_t i = 0; i + 1 < n; ++i)
    if (forked[i]) CU(cudaStreamWaitEvent(c->stream, c->ev_join[i], 0));
  if (c->refresher_pending) {  // the refreshers leave right after their samplers: join them
    CU(cudaEventRecord(c->ev_rjoin, c->refresh_stream));
    CU(cudaStreamWaitEvent(c->stream, c->ev_rjoin, 0));
    c->refresher_pending = false;
  }
  return B200LDA_OK;
}

// Fold every finished sweep's event quadruple into the running sums (stream must be idle).
int resolve_events(b200lda_ctx* c) {
  CU(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < c->ev_pending; ++i) {
    cudaEvent_t* e = c->ev_pool.data() + 4 * i;
    float t01 = 0, t12 = 0, t23 = 0;
    CU(cudaEventElapsedTime(&t01, e[0], e[1]));
    CU(cudaEventElapsedTime(&t12, e[1], e[2]));
    CU(cudaEventElapsedTime(&t23, e[2], e[3]));
    c->cum_tables_ms += t01;
    c->cum_sample_ms += t12;
    c->cum_finish_ms += t23;
    c->last_tables_ms = t01;
    c->last_sample_ms = t12;
    c->last_finish_ms = t23;
    c->cum_sweeps += 1;
  }
  c->ev_pending = 0;
  return B200LDA_OK;
}

int next_event_quad(b200lda_ctx* c) {
  if (c->ev_pending == kEventPool) TRY(resolve_events(c));
  const size_t need = 4 * (size_t)(c->ev_pending + 1);
  while (c->ev_pool.size() < need) {
    cudaEvent_t e;
    CU(cudaEventCreate(&e));
    c->ev_pool.push_back(e);
  }
  c->ev = c->ev_pool.data() + 4 * c->ev_pending;
  return B200LDA_OK;
}

int need_ready(b200lda_ctx* c) {
  if (!c->corpus_loaded) return fail(B200LDA_ESTATE, "no corpus loaded (call b200lda_load_corpus)");
  if (!c->assigned) return fail(B200LDA_ESTATE, "no topic assignments (call b200lda_init_assignments)");
  if (c->in_sweep) return fail(B200LDA_ESTATE, "b200lda_sweep_begin without b200lda_sweep_end");
  if (c->in_sync) return fail(B200LDA_ESTATE, "b200lda_counts_sync_begin without b200lda_counts_sync_end");
  return B200LDA_OK;
}

}  // namespace

extern "C" {

const char* b200lda_last_error(void) { return g_err.c_str(); }
int b200lda_abi_version(void) { return B200LDA_ABI_VERSION; }

int b200lda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) ++ok;
  }
  return ok;
}

int b200lda_create(const b200lda_config* cfg, b200lda_ctx** out) {
  if (!cfg || !out) return fail(B200LDA_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->struct_size != (int32_t)sizeof(b200lda_config))
    return fail(B200LDA_EINVAL, "b200lda_config.struct_size=%d, library expects %zu", cfg->struct_size,
                sizeof(b200lda_config));
  if (cfg->num_topics < 1 || cfg->num_topics > 65536) return fail(B200LDA_EINVAL, "num_topics must be in 1..65536");
  if (cfg->num_types < 1) return fail(B200LDA_EINVAL, "num_types must be >= 1");
  if (!(cfg->alpha_sum > 0.0) || !(cfg->beta > 0.0)) return fail(B200LDA_EINVAL, "alpha_sum and beta must be > 0");
  if (cfg->mode != B200LDA_MODE_LIVE && cfg->mode != B200LDA_MODE_DEFERRED) return fail(B200LDA_EINVAL, "bad mode");
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size)
    return fail(B200LDA_EINVAL, "bad rank/world_size");
  if ((double)cfg->num_types * (double)cfg->num_topics > 8.0e9)
    return fail(B200LDA_ERANGE, "V*K too large for a dense n_wk replica");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(B200LDA_ENODEV, "no CUDA device visible; libb200lda has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(B200LDA_ENODEV, "device %d out of range (%d visible)", cfg->device, ndev);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10)
    return fail(B200LDA_ENODEV, "device %d is sm_%d%d; libb200lda is built for sm_100a only", cfg->device, prop.major,
                prop.minor);
  CU(cudaSetDevice(cfg->device));

  b200lda_ctx* c = new (std::nothrow) b200lda_ctx();
  if (!c) return fail(B200LDA_ENOMEM, "out of host memory");
  c->cfg = *cfg;
  c->K = cfg->num_topics;
  c->V = cfg->num_types;
  c->beta = cfg->beta;
  c->sm_count = prop.multiProcessorCount;
  c->layout = make_layout(c->K);
  if (const char* e = std::getenv("B200LDA_CLASS_STREAMS")) c->class_streams = atoi(e);  // tuning knob for experiments
  if (const char* e = std::getenv("B200LDA_MAX_CTAS")) c->max_ctas = atoi(e);
  if (const char* e = std::getenv("B200LDA_EXCHANGE_SLABS")) c->exchange_slabs = std::max(1, std::min(atoi(e), kExchangeSlabs));
  if (const char* e = std::getenv("B200LDA_APPLY_CTAS")) c->apply_ctas = std::max(1, atoi(e));
  c->table_refresh = std::max(0, std::min((int)cfg->table_refresh, kMaxRefresh));
  if (const char* e = std::getenv("B200LDA_TABLE_REFRESH")) c->table_refresh = std::max(0, std::min(atoi(e), kMaxRefresh));
  c->alpha.assign(c->K, cfg->alpha_sum / c->K);
  int rc = B200LDA_OK;
  auto bail = [&](int code) {
    b200lda_destroy(c);
    return code;
  };
  if (cfg->stream) {
    c->stream = (cudaStream_t)cfg->stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess)
      return bail(fail(B200LDA_ECUDA, "cudaStreamCreate failed"));
    c->own_stream = true;
  }
  c->ev_pool.reserve(4 * kEventPool);
  {
    int lo = 0, hi = 0;  // wide-row classes get the higher priority: their chains are the critical path
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    bool ok = cudaEventCreate(&c->ev_fork) == cudaSuccess && cudaEventCreate(&c->ev_bulk) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_rfork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_rjoin, cudaEventDisableTiming) == cudaSuccess &&
              cudaStreamCreateWithPriority(&c->refresh_stream, cudaStreamNonBlocking, hi) == cudaSuccess;
    // the side streams of the background classes are created when a corpus needs them: a process
    // has few hardware queues, and streams that share one are serialised (the refreshers must run
    // BESIDE the samplers)
    c->stream_priority = hi;
    for (int i = 0; i < kMaxClasses && ok; ++i) ok = cudaEventCreate(&c->ev_join[i]) == cudaSuccess;
    if (!ok) return bail(fail(B200LDA_ECUDA, "creating side streams failed"));
  }
  const size_t VK = (size_t)c->V * c->K;
  const bool multi = cfg->world_size > 1;
  // multi-shard: the replica (= the exchange buffer) comes from ncclMemAlloc and is registered with
  // the communicator (measured at 8 GPUs, K = 1000: exchange 3.08 -> 2.63 ms per sweep);
  // B200LDA_NCCL_REGISTER=0 keeps cudaMalloc
  const char* reg_env = std::getenv("B200LDA_NCCL_REGISTER");
  if (multi && !(reg_env && atoi(reg_env) == 0) && load_nccl() == B200LDA_OK && g_nccl.MemAlloc && g_nccl.MemFree) {
    void* ptr = nullptr;
    if (g_nccl.MemAlloc(&ptr, sizeof(int32_t) * (VK + c->K)) == ncclSuccess && ptr) {
      c->d_nwk = static_cast<int32_t*>(ptr);
      c->nwk_from_nccl = true;
      c->device_bytes += (int64_t)(sizeof(int32_t) * (VK + c->K));
    }
  }
  if ((!c->d_nwk && (rc = dev_alloc_t(c, &c->d_nwk, VK + c->K))) || (rc = dev_alloc_t(c, &c->d_nk, c->K)) ||
      (rc = dev_alloc_t(c, &c->d_nk_delta, c->K)) || (rc = dev_alloc_t(c, &c->d_invden, c->K)) ||
      (rc = dev_alloc_t(c, &c->d_ab, c->K)) || (rc = dev_alloc_t(c, &c->d_alpha_f, c->K)) ||
      (rc = dev_alloc_t(c, &c->d_alpha, c->K)) || (rc = dev_alloc_t(c, &c->d_lg_alpha, c->K)) ||
      (rc = dev_alloc_t(c, &c->d_prior, (size_t)(cfg->mode == B200LDA_MODE_LIVE ? 2 : 1) * c->V * c->layout.stride)) ||
      (rc = dev_alloc_t(c, &c->d_q, (size_t)(cfg->mode == B200LDA_MODE_LIVE ? 2 : 1) * c->V)) ||
      (rc = dev_alloc_t(c, &c->d_row_cursor, kMaxClasses)) ||
      (rc = dev_alloc_t(c, &c->d_counters, kCounters)) || (rc = dev_alloc_t(c, &c->d_sched, kMaxClasses * kMaxSegments)) || (rc = dev_alloc_t(c, &c->d_bad, 1)) ||
      (rc = dev_alloc_t(c, &c->d_partial, 2 * kPartial + 2)))
    return bail(rc);
  if (cfg->mode == B200LDA_MODE_DEFERRED || multi)
    if ((rc = dev_alloc_t(c, &c->d_nwk_b, VK + c->K))) return bail(rc);
  if (cfg->mode == B200LDA_MODE_LIVE) {
    if ((rc = dev_alloc_t(c, &c->d_prior_sel, c->V))) return bail(rc);
    if (cudaMemsetAsync(c->d_prior_sel, 0, sizeof(int32_t) * c->V, c->stream) != cudaSuccess ||
        cudaMemsetAsync(c->d_row_cursor, 0, sizeof(unsigned) * kMaxClasses, c->stream) != cudaSuccess)
      return bail(fail(B200LDA_ECUDA, "cudaMemset failed"));
  }
  if (multi) {
    bool ok = cudaStreamCreateWithFlags(&c->apply_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_applied, cudaEventDisableTiming) == cudaSuccess &&
              cudaMemsetAsync(c->d_nwk_b, 0, sizeof(int32_t) * (VK + c->K), c->stream) == cudaSuccess;
    for (int i = 0; i < kExchangeSlabs && ok; ++i)
      ok = cudaEventCreateWithFlags(&c->ev_slab[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) return bail(fail(B200LDA_ECUDA, "creating the exchange streams failed"));
  }
  if (cudaMemsetAsync(c->d_nk_delta, 0, sizeof(int32_t) * c->K, c->stream) != cudaSuccess ||
      cudaMemsetAsync(c->d_nwk, 0, sizeof(int32_t) * (VK + c->K), c->stream) != cudaSuccess ||
      cudaMemsetAsync(c->d_nk, 0, sizeof(int32_t) * c->K, c->stream) != cudaSuccess ||
      cudaMemsetAsync(c->d_counters, 0, sizeof(unsigned long long) * kCounters, c->stream) != cudaSuccess)
    return bail(fail(B200LDA_ECUDA, "cudaMemset failed"));
  if ((rc = push_alpha(c))) return bail(rc);
  *out = c;
  return B200LDA_OK;
}

void b200lda_destroy(b200lda_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  free_corpus(c->corp);
  // the registration goes before the communicator, the communicator before the memory it used
  if (c->comm && c->nwk_reg && g_nccl.CommDeregister) g_nccl.CommDeregister(c->comm, c->nwk_reg);
  c->nwk_reg = nullptr;
  if (c->comm && g_nccl.handle) g_nccl.CommDestroy(c->comm);
  c->comm = nullptr;
  if (c->nwk_from_nccl && c->d_nwk) {
    g_nccl.MemFree(c->d_nwk);
    c->d_nwk = nullptr;
  }
  dev_free(c->d_nwk);
  dev_free(c->d_nwk_b);
  dev_free(c->d_nk);
  dev_free(c->d_nk_delta);
  if (c->apply_stream) cudaStreamDestroy(c->apply_stream);
  if (c->ev_applied) cudaEventDestroy(c->ev_applied);
  for (int i = 0; i < kExchangeSlabs; ++i)
    if (c->ev_slab[i]) cudaEventDestroy(c->ev_slab[i]);
  dev_free(c->d_invden);
  dev_free(c->d_ab);
  dev_free(c->d_prior);
  dev_free(c->d_q);
  dev_free(c->d_prior_sel);
  dev_free(c->d_row_cursor);
  dev_free(c->d_hot_words);
  dev_free(c->d_alpha_f);
  dev_free(c->d_alpha);
  dev_free(c->d_lg_alpha);
  dev_free(c->d_counters);
  dev_free(c->d_sched);
  dev_free(c->d_bad);
  dev_free(c->d_partial);
  dev_free(c->d_hist_scratch);
  dev_free(c->d_hyper);
  if (c->d_stage) cudaFree(c->d_stage);
  for (auto& e : c->ev_pool)
    if (e) cudaEventDestroy(e);
  if (c->ev_rfork) cudaEventDestroy(c->ev_rfork);
  if (c->ev_rjoin) cudaEventDestroy(c->ev_rjoin);
  if (c->refresh_stream) cudaStreamDestroy(c->refresh_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_bulk) cudaEventDestroy(c->ev_bulk);
  for (int i = 0; i < kMaxClasses; ++i) {
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    if (c->side[i]) cudaStreamDestroy(c->side[i]);
  }
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  cudaGetLastError();
  delete c;
}

int b200lda_load_corpus(b200lda_ctx* c, int64_t num_docs, const int64_t* doc_ptr, const int32_t* tok_word) {
  TRY(enter(c));
  if (c->in_sweep || c->in_sync) return fail(B200LDA_ESTATE, "a sweep or count sync is open");
  c->corpus_loaded = false;
  c->assigned = false;
  TRY(pack_corpus(c, c->corp, num_docs, doc_ptr, tok_word));
  c->corpus_loaded = true;
  return B200LDA_OK;
}

namespace {
// z32 / z16: the caller's topics in either width (at most one non-null); both null: Philox draw.
int init_assignments_impl(b200lda_ctx* c, const int32_t* z32, const uint16_t* z16) {
  TRY(enter(c));
  if (!c->corpus_loaded) return fail(B200LDA_ESTATE, "no corpus loaded (call b200lda_load_corpus)");
  if (c->in_sweep || c->in_sync) return fail(B200LDA_ESTATE, "a sweep or count sync is open");
  c->assigned = false;
  c->hot_count = -1;
  c->snapshot_valid = false;
  DeviceCorpus& cp = c->corp;
  const int64_t N = cp.N;
  int bad = 0;
  if (N > 0) {
    CU(cudaMemsetAsync(c->d_bad, 0, sizeof(int), c->stream));
    if (z32) {
      TRY(ensure_stage(c, sizeof(int32_t) * (size_t)N));
      CU(cudaMemcpyAsync(c->d_stage, z32, sizeof(int32_t) * N, cudaMemcpyHostToDevice, c->stream));
      k_narrow_z<<<grid_for(c, N, 256), 256, 0, c->stream>>>(N, c->K, reinterpret_cast<const int32_t*>(c->d_stage),
                                                          cp.d_z, c->d_bad);
    } else if (z16) {  // the device's own width: no staging, half the bytes over the bus
      CU(cudaMemcpyAsync(cp.d_z, z16, sizeof(uint16_t) * N, cudaMemcpyHostToDevice, c->stream));
      k_validate_z<<<grid_for(c, N, 256), 256, 0, c->stream>>>(N, c->K, cp.d_z, c->d_bad);
    } else {
      k_init_z<<<grid_for(c, N, 256), 256, 0, c->stream>>>(N, c->K, c->cfg.seed, c->cfg.global_token_offset, cp.d_z);
    }
    if (z32 || z16) CU(cudaMemcpyAsync(&bad, c->d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    c->launches += 1;
  }
  // n_wk by integer atomics straight from the doc -> token order, n_k as its column sums
  // (an out-of-range topic was clamped to 0 above and is reported below: nothing reads out of bounds)
  const size_t VK = (size_t)c->V * c->K;
  CU(cudaMemsetAsync(c->d_nwk, 0, sizeof(int32_t) * VK, c->stream));
  CU(cudaMemsetAsync(c->d_nk, 0, sizeof(int32_t) * c->K, c->stream));
  CU(cudaMemsetAsync(c->d_nk_delta, 0, sizeof(int32_t) * c->K, c->stream));
  if (N > 0) {
    k_count_direct<<<grid_for(c, N, 256, 16), 256, 0, c->stream>>>(N, c->K, cp.d_tok_word, cp.d_z, c->d_nwk);
    const int rows_per_block = 256;
    dim3 grid((c->K + 255) / 256, (c->V + rows_per_block - 1) / rows_per_block);
    k_col_sums<<<grid, 256, 0, c->stream>>>(c->V, c->K, rows_per_block, c->d_nwk, c->d_nk);
    c->launches += 2;
  }
  TRY(build_doc_rows(c, cp));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));  // the one synchronisation of the call (the caller's z is borrowed)
  if (bad) return fail(B200LDA_ERANGE, "z holds a topic outside [0, %d)", c->K);
  c->assigned = true;
  return B200LDA_OK;
}
}  // namespace

int b200lda_init_assignments(b200lda_ctx* c, const int32_t* z) { return init_assignments_impl(c, z, nullptr); }
int b200lda_init_assignments_u16(b200lda_ctx* c, const uint16_t* z) {
  if (!z) return fail(B200LDA_EINVAL, "z is null");
  return init_assignments_impl(c, nullptr, z);
}

int b200lda_counts_sync_begin(b200lda_ctx* c) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (c->cfg.world_size <= 1) return fail(B200LDA_ESTATE, "count sync needs world_size > 1");
  const size_t VK = (size_t)c->V * c->K;
  // the exchange buffer is d_nwk itself: this shard's counts, its n_k in the tail
  CU(cudaMemcpyAsync(c->d_nwk + VK, c->d_nk, sizeof(int32_t) * c->K, cudaMemcpyDeviceToDevice, c->stream));
  c->in_sync = true;
  return B200LDA_OK;
}

int b200lda_counts_sync_end(b200lda_ctx* c) {
  TRY(enter(c));
  if (!c->in_sync) return fail(B200LDA_ESTATE, "b200lda_counts_sync_end without b200lda_counts_sync_begin");
  const size_t VK = (size_t)c->V * c->K;
  k_install_nk_tail<<<(c->K + 255) / 256, 256, 0, c->stream>>>(c->K, c->d_nk, c->d_nwk + VK, c->d_nwk_b + VK);
  CU(cudaMemcpyAsync(c->d_nwk_b, c->d_nwk, sizeof(int32_t) * VK, cudaMemcpyDeviceToDevice, c->stream));
  c->launches += 1;
  CU(cudaGetLastError());
  c->snapshot_valid = true;
  c->hot_count = -1;
  c->in_sync = false;
  return B200LDA_OK;
}

int b200lda_sweep_begin(b200lda_ctx* c) {
  TRY(enter(c));
  TRY(need_ready(c));
  const bool multi = c->cfg.world_size > 1;
  const bool deferred = c->cfg.mode == B200LDA_MODE_DEFERRED;
  const size_t VK = (size_t)c->V * c->K;
  if (!deferred && c->table_refresh != 1) TRY(build_hot_words(c));  // once per corpus
  if (multi && !c->snapshot_valid) {  // counts were installed without a count sync (restored state, one-shard tests)
    CU(cudaMemcpyAsync(c->d_nwk_b, c->d_nwk, sizeof(int32_t) * VK, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemsetAsync(c->d_nwk + VK, 0, sizeof(int32_t) * c->K, c->stream));
    CU(cudaMemsetAsync(c->d_nwk_b + VK, 0, sizeof(int32_t) * c->K, c->stream));
    c->snapshot_valid = true;
  }
  TRY(next_event_quad(c));
  CU(cudaEventRecord(c->ev[0], c->stream));
  nvtxRangePushA("b200lda:tables");
  const int rc_tables = build_tables(c, !deferred);
  nvtxRangePop();
  TRY(rc_tables);
  CU(cudaEventRecord(c->ev[1], c->stream));
  NvtxRange sample_range("b200lda:sample");
  if (deferred && !multi)
    CU(cudaMemcpyAsync(c->d_nwk_b, c->d_nwk, sizeof(int32_t) * VK, cudaMemcpyDeviceToDevice, c->stream));
  // one shard: LIVE reads and writes d_nwk; DEFERRED reads the frozen d_nwk, writes the copy d_nwk_b
  //            (swapped in at the end of the sweep).
  // shards:    d_nwk_b = the global counts of the sweep start on every shard. LIVE reads and writes
  //            d_nwk; DEFERRED reads d_nwk_b and writes d_nwk. Either way d_nwk ends the sweep as
  //            "start + this shard's moves" with the n_k moves in its tail: the exchange buffer.
  const int32_t* rd = (deferred && multi) ? c->d_nwk_b : c->d_nwk;
  int32_t* wr = (deferred && !multi) ? c->d_nwk_b : c->d_nwk;
  SweepParams p = sweep_params(c, c->corp, rd, wr, (uint32_t)(c->sweeps_done + 1));
  if (multi) p.nk_delta = c->d_nwk + VK;
  if (deferred) {
    TRY((launch_sweep<MODE_UPDATE, false>(c, c->corp, p)));
  } else {
    // every hot word's prior row is rebuilt table_refresh times per sweep, the first of them by
    // build_tables above, the others by the refreshers beside the bulk launches
    const int refresh = c->table_refresh > 0 ? c->table_refresh : auto_table_refresh(c, c->corp);
    c->last_refresh = refresh;
    c->rows_rebuilt_by_host = 0;
    p.prior_sel = c->d_prior_sel;
    p.hot_words = c->d_hot_words;
    p.hot_count = c->hot_count;
    TRY((launch_sweep<MODE_UPDATE, true>(c, c->corp, p, refresh - 1)));
  }
  k_accumulate_stats<<<1, 32, 0, c->stream>>>(c->d_counters + 1, c->d_counters + 5);
  CU(cudaEventRecord(c->ev[2], c->stream));
  c->in_sweep = true;
  return B200LDA_OK;
}

int b200lda_exchange_buffer(b200lda_ctx* c, void** d_buf, int64_t* count) {
  if (!c || !d_buf || !count) return fail(B200LDA_EINVAL, "null argument");
  const bool multi = c->cfg.world_size > 1;
  *d_buf = multi ? c->d_nwk : nullptr;
  *count = multi ? (int64_t)((size_t)c->V * c->K + c->K) : 0;
  return B200LDA_OK;
}

namespace {
// bookkeeping shared by every way a sweep ends
int sweep_close(b200lda_ctx* c) {
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev[3], c->stream));
  c->ev_pending += 1;
  c->in_sweep = false;
  c->sweeps_done += 1;
  c->tokens_sampled += c->corp.N;
  return B200LDA_OK;
}

// Slab s of the exchange buffer: [*i0, *i1) cells of the V x K part.
void exchange_slab(const b200lda_ctx* c, int s, size_t* i0, size_t* i1) {
  const size_t VK = (size_t)c->V * c->K;
  const size_t per = ((VK + c->exchange_slabs - 1) / c->exchange_slabs + 3) & ~(size_t)3;
  *i0 = std::min(VK, per * (size_t)s);
  *i1 = std::min(VK, per * (size_t)(s + 1));
}

// The exchange done by the library over NCCL, for `n` contexts driven by this thread (n = 1: one
// process per GPU). The K-cell tail goes first (n_k), then the V x K cells in slabs: slab s is
// applied on the context's second stream while NCCL reduces slab s + 1.
int exchange_nccl(b200lda_ctx** ctxs, int n) {
  NvtxRange range("b200lda:exchange");
  TRY(load_nccl());
  for (int i = 0; i < n; ++i)
    if (!ctxs[i]->comm) return fail(B200LDA_ESTATE, "context %d has no communicator (b200lda_comm_init / b200lda_group_comm_init)", i);
  const size_t VK = (size_t)ctxs[0]->V * ctxs[0]->K;
  const int K = ctxs[0]->K;
  const int nm1 = ctxs[0]->cfg.world_size - 1;
  NCCL(g_nccl.GroupStart());
  for (int i = 0; i < n; ++i) {
    b200lda_ctx* c = ctxs[i];
    NCCL(g_nccl.AllReduce(c->d_nwk + VK, c->d_nwk + VK, (size_t)K, ncclInt32, ncclSum, c->comm, c->stream));
  }
  NCCL(g_nccl.GroupEnd());
  for (int i = 0; i < n; ++i) {
    b200lda_ctx* c = ctxs[i];
    CU(cudaSetDevice(c->cfg.device));
    k_apply_nk_tail<<<(K + 255) / 256, 256, 0, c->stream>>>(K, c->d_nk, c->d_nwk + VK, c->d_nwk_b + VK);
    c->launches += 1;
  }
  for (int s = 0; s < ctxs[0]->exchange_slabs; ++s) {
    size_t i0 = 0, i1 = 0;
    exchange_slab(ctxs[0], s, &i0, &i1);
    if (i1 <= i0) continue;
    NCCL(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
      b200lda_ctx* c = ctxs[i];
      NCCL(g_nccl.AllReduce(c->d_nwk + i0, c->d_nwk + i0, i1 - i0, ncclInt32, ncclSum, c->comm, c->stream));
    }
    NCCL(g_nccl.GroupEnd());
    for (int i = 0; i < n; ++i) {
      b200lda_ctx* c = ctxs[i];
      CU(cudaSetDevice(c->cfg.device));
      CU(cudaEventRecord(c->ev_slab[s], c->stream));
      CU(cudaStreamWaitEvent(c->apply_stream, c->ev_slab[s], 0));
      // the apply leaves SMs to NCCL's kernels (the next slab's all-reduce is running beside it)
      const int apply_grid = std::min(grid_for(c, (int64_t)((i1 - i0) / 4 + 1), 256), c->apply_ctas);
      k_apply_sum<<<apply_grid, 256, 0, c->apply_stream>>>(i0, i1, nm1, c->d_nwk, c->d_nwk_b);
      c->launches += 1;
      CU(cudaGetLastError());
    }
  }
  for (int i = 0; i < n; ++i) {
    b200lda_ctx* c = ctxs[i];
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaEventRecord(c->ev_applied, c->apply_stream));
    CU(cudaStreamWaitEvent(c->stream, c->ev_applied, 0));
  }
  return B200LDA_OK;
}
}  // namespace

int b200lda_sweep_end(b200lda_ctx* c) {
  TRY(enter(c));
  if (!c->in_sweep) return fail(B200LDA_ESTATE, "b200lda_sweep_end without b200lda_sweep_begin");
  const bool multi = c->cfg.world_size > 1;
  const bool deferred = c->cfg.mode == B200LDA_MODE_DEFERRED;
  const size_t VK = (size_t)c->V * c->K;
  if (multi) {  // the caller has summed the exchange buffer over the shards
    k_apply_nk_tail<<<(c->K + 255) / 256, 256, 0, c->stream>>>(c->K, c->d_nk, c->d_nwk + VK, c->d_nwk_b + VK);
    k_apply_sum<<<grid_for(c, (int64_t)(VK / 4 + 1), 256), 256, 0, c->stream>>>(0, VK, c->cfg.world_size - 1, c->d_nwk, c->d_nwk_b);
    c->launches += 2;
  } else {
    if (deferred) std::swap(c->d_nwk, c->d_nwk_b);
    k_apply_nk<<<(c->K + 255) / 256, 256, 0, c->stream>>>(c->K, c->d_nk, c->d_nk_delta);
    c->launches += 1;
  }
  return sweep_close(c);
}

int b200lda_synchronize(b200lda_ctx* c) {
  TRY(enter(c));
  CU(cudaStreamSynchronize(c->stream));
  return B200LDA_OK;
}

int b200lda_get_stream(b200lda_ctx* c, void** stream) {
  if (!c || !stream) return fail(B200LDA_EINVAL, "null argument");
  *stream = (void*)c->stream;
  return B200LDA_OK;
}

int b200lda_sweep(b200lda_ctx* c, int32_t n) {
  TRY(enter(c));
  if (n < 0) return fail(B200LDA_EINVAL, "negative sweep count");
  if (c->cfg.world_size > 1 && !c->comm)
    return fail(B200LDA_ESTATE, "a shard sweeps through b200lda_comm_init + b200lda_sweep, or sweep_begin / all-reduce / sweep_end");
  TRY(need_ready(c));
  for (int32_t i = 0; i < n; ++i) {
    TRY(b200lda_sweep_begin(c));
    if (c->cfg.world_size > 1) {
      TRY(exchange_nccl(&c, 1));
      TRY(sweep_close(c));
    } else {
      TRY(b200lda_sweep_end(c));
    }
  }
  CU(cudaStreamSynchronize(c->stream));
  return B200LDA_OK;
}

int b200lda_sample_frozen(b200lda_ctx* c, const float* uniforms, uint32_t sweep, int32_t* z_out) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!z_out) return fail(B200LDA_EINVAL, "z_out is null");
  const int64_t N = c->corp.N;
  if (N == 0) return B200LDA_OK;
  TRY(ensure_stage(c, (size_t)N * 8));
  int32_t* d_zout = reinterpret_cast<int32_t*>(c->d_stage);
  float* d_u = reinterpret_cast<float*>(c->d_stage) + N;
  if (uniforms) CU(cudaMemcpyAsync(d_u, uniforms, sizeof(float) * N, cudaMemcpyHostToDevice, c->stream));
  TRY(build_tables(c));
  SweepParams p = sweep_params(c, c->corp, c->d_nwk, nullptr, sweep);  // MODE_FROZEN never writes counts
  p.z_out = d_zout;
  p.uniforms = uniforms ? d_u : nullptr;
  TRY((launch_sweep<MODE_FROZEN, false>(c, c->corp, p)));
  CU(cudaMemcpyAsync(z_out, d_zout, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return B200LDA_OK;
}

int b200lda_infer(b200lda_ctx* c, int64_t num_docs, const int64_t* doc_ptr, const int32_t* tok_word, int32_t iterations,
                  int32_t thinning, int32_t burn_in, uint64_t seed, double* theta) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (num_docs < 0 || !doc_ptr || (num_docs > 0 && !theta)) return fail(B200LDA_EINVAL, "bad inference arguments");
  if (iterations < 0 || thinning < 1 || burn_in < 0) return fail(B200LDA_EINVAL, "bad iterations/thinning/burn-in");
  if (num_docs == 0) return B200LDA_OK;
  // held-out documents are packed like a corpus of their own; the trained counts stay frozen
  DeviceCorpus cp;
  const int64_t bytes_before = c->device_bytes;
  int rc = pack_corpus(c, cp, num_docs, doc_ptr, tok_word);
  cp.tune_waits = 0;  // one pass over a throw-away corpus: nothing to tune, never wait for timings
  int32_t* d_acc = nullptr;
  double* d_theta = nullptr;
  auto cleanup = [&](int code) {
    cudaStreamSynchronize(c->stream);
    if (c->timed_corpus == &cp) c->timed_corpus = nullptr;
    free_corpus(cp);
    dev_free(d_acc);
    dev_free(d_theta);
    c->device_bytes = bytes_before;
    return code;
  };
  if (rc) return cleanup(rc);
  const size_t DK = (size_t)num_docs * c->K;
  if ((rc = dev_alloc_t(c, &d_acc, DK)) || (rc = dev_alloc_t(c, &d_theta, DK))) return cleanup(rc);
  if (cudaMemsetAsync(d_acc, 0, sizeof(int32_t) * DK, c->stream) != cudaSuccess) return cleanup(fail(B200LDA_ECUDA, "memset failed"));
  if (cp.N > 0)  // TopicInferencer: uniform random initial topics (own Philox stream: sweep 0, seed-keyed)
    k_init_z<<<grid_for(c, cp.N, 256), 256, 0, c->stream>>>(cp.N, c->K, seed ^ 0x9E3779B97F4A7C15ull, 0, cp.d_z);
  if ((rc = build_doc_rows(c, cp))) return cleanup(rc);
  if ((rc = build_tables(c))) return cleanup(rc);
  int samples = 0;
  for (int32_t it = 1; it <= iterations; ++it)
    if (it > burn_in && (it - burn_in) % thinning == 0) ++samples;
  if (iterations >= 1) {
    // n_wk / n_k are frozen, so every document is an independent chain: one launch set in which a
    // warp runs all iterations of its document (samples added to d_acc as they are reached)
    // replaces `iterations` launch sets + accumulate passes. Philox keys: sweep = iteration.
    SweepParams p = sweep_params(c, cp, c->d_nwk, nullptr, 0u);  // MODE_INFER: the held-out tokens are not part of n_wk / n_k
    p.stats = c->d_counters + 9;  // inference leaves the training chain's sweep statistics alone
    p.seed = seed;
    p.global_tok_off = 0;
    p.infer_iters = iterations;
    p.infer_burn_in = burn_in;
    p.infer_thinning = thinning;
    p.infer_samples = samples;
    p.infer_acc = d_acc;
    if ((rc = launch_sweep<MODE_INFER, false>(c, cp, p))) return cleanup(rc);
    if (samples == 0) samples = 1;  // Mallet: no sample saved -> the final state (added by the kernel)
  } else {  // no iterations at all: the (random) initial state is the sample
    k_infer_accumulate<<<grid_for(c, num_docs * 32, 256), 256, 0, c->stream>>>(num_docs, c->K, cp.d_row_ptr, cp.d_row_nnz,
                                                                           cp.d_rows, d_acc);
    samples = 1;
  }
  k_infer_theta<<<grid_for(c, (int64_t)DK, 256), 256, 0, c->stream>>>(num_docs, c->K, samples, cp.d_doc_ptr, d_acc, c->d_alpha,
                                                                     c->alpha_sum, d_theta);
  c->launches += 2;
  if (cudaGetLastError() != cudaSuccess) return cleanup(fail(B200LDA_ECUDA, "inference kernel launch failed"));
  if (cudaMemcpyAsync(theta, d_theta, sizeof(double) * DK, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
      cudaStreamSynchronize(c->stream) != cudaSuccess)
    return cleanup(fail(B200LDA_ECUDA, "inference copy-back failed: %s", cudaGetErrorString(cudaGetLastError())));
  // inference must not disturb the training chain's n_k delta accumulator or stats of the last sweep
  return cleanup(B200LDA_OK);
}

int b200lda_loglik_parts(b200lda_ctx* c, double* doc_part, double* word_part) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!doc_part || !word_part) return fail(B200LDA_EINVAL, "null argument");
  const DeviceCorpus& cp = c->corp;
  const size_t VK = (size_t)c->V * c->K;
  double* p_docs = c->d_partial;
  double* p_words = c->d_partial + kPartial;
  double* d_out = c->d_partial + 2 * kPartial;
  unsigned long long* d_nz = c->d_counters + 4;
  CU(cudaMemsetAsync(d_nz, 0, sizeof(unsigned long long), c->stream));
  k_loglik_docs<<<kPartial, 256, 0, c->stream>>>(cp.D, cp.d_doc_ptr, cp.d_row_ptr, cp.d_row_nnz, cp.d_rows, c->d_alpha,
                                                c->d_lg_alpha, c->alpha_sum, p_docs);
  k_loglik_words<<<kPartial, 256, 0, c->stream>>>(VK, c->d_nwk, c->beta, p_words, d_nz);
  k_loglik_final<<<1, 256, 0, c->stream>>>(kPartial, p_docs, 0, c->d_nk, 0.0, d_out + 0);
  k_loglik_final<<<1, 256, 0, c->stream>>>(kPartial, p_words, c->K, c->d_nk, c->beta * (double)c->V, d_out + 1);
  c->launches += 4;
  CU(cudaGetLastError());
  double h[2];
  unsigned long long nz = 0;
  CU(cudaMemcpyAsync(h, d_out, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(&nz, d_nz, sizeof(nz), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  *doc_part = h[0] + (double)cp.D * std::lgamma(c->alpha_sum);
  *word_part = h[1] + (double)c->K * std::lgamma(c->beta * (double)c->V) - (double)nz * std::lgamma(c->beta);
  return B200LDA_OK;
}

int b200lda_loglik(b200lda_ctx* c, double* out) {
  if (!out) return fail(B200LDA_EINVAL, "null argument");
  double a = 0.0, b = 0.0;
  TRY(b200lda_loglik_parts(c, &a, &b));
  *out = a + b;
  return B200LDA_OK;
}

int b200lda_check_invariants(b200lda_ctx* c, int64_t* out) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!out) return fail(B200LDA_EINVAL, "null argument");
  const DeviceCorpus& cp = c->corp;
  const size_t VK = (size_t)c->V * c->K;
  TRY(ensure_stage(c, sizeof(int32_t) * (size_t)c->K + 4 * sizeof(unsigned long long) + 16));
  unsigned long long* d_out = reinterpret_cast<unsigned long long*>(c->d_stage);
  int32_t* d_col = reinterpret_cast<int32_t*>(d_out + 4);
  CU(cudaMemsetAsync(c->d_stage, 0, 4 * sizeof(unsigned long long) + sizeof(int32_t) * (size_t)c->K, c->stream));
  const int rows_per_block = 256;
  dim3 grid((c->K + 255) / 256, (c->V + rows_per_block - 1) / rows_per_block);
  k_col_sums<<<grid, 256, 0, c->stream>>>(c->V, c->K, rows_per_block, c->d_nwk, d_col);
  k_check_nk<<<(c->K + 255) / 256, 256, 0, c->stream>>>(c->K, c->d_nk, d_col, d_out);
  k_sum_i32<<<grid_for(c, (int64_t)VK, 256), 256, 0, c->stream>>>(VK, c->d_nwk, d_out + 1);
  if (cp.D > 0)
    k_sum_rows<<<grid_for(c, cp.D * 32, 256), 256, 0, c->stream>>>(cp.D, cp.d_row_ptr, cp.d_row_nnz, cp.d_rows, d_out + 3);
  c->launches += 4;
  CU(cudaGetLastError());
  unsigned long long h[4];
  CU(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 4; ++i) out[i] = (int64_t)h[i];
  return B200LDA_OK;
}

int b200lda_get_assignments(b200lda_ctx* c, int32_t* z) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!z) return fail(B200LDA_EINVAL, "z is null");
  const int64_t N = c->corp.N;
  if (N == 0) return B200LDA_OK;
  TRY(ensure_stage(c, sizeof(int32_t) * (size_t)N));
  int32_t* d_wide = reinterpret_cast<int32_t*>(c->d_stage);
  k_widen_z<<<grid_for(c, N, 256), 256, 0, c->stream>>>(N, c->corp.d_z, d_wide);
  c->launches += 1;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(z, d_wide, sizeof(int32_t) * N, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return B200LDA_OK;
}

int b200lda_get_assignments_u16(b200lda_ctx* c, uint16_t* z) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!z) return fail(B200LDA_EINVAL, "z is null");
  if (c->corp.N == 0) return B200LDA_OK;
  CU(cudaMemcpyAsync(z, c->corp.d_z, sizeof(uint16_t) * c->corp.N, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return B200LDA_OK;
}

int b200lda_get_nwk(b200lda_ctx* c, int32_t* nwk) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!nwk) return fail(B200LDA_EINVAL, "null argument");
  CU(cudaMemcpyAsync(nwk, c->d_nwk, sizeof(int32_t) * (size_t)c->V * c->K, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return B200LDA_OK;
}

int b200lda_get_nk(b200lda_ctx* c, int32_t* nk) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!nk) return fail(B200LDA_EINVAL, "null argument");
  CU(cudaMemcpyAsync(nk, c->d_nk, sizeof(int32_t) * c->K, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return B200LDA_OK;
}

int b200lda_get_ndk_csr(b200lda_ctx* c, int64_t* row_ptr, int32_t* topic, int32_t* count) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!row_ptr) return fail(B200LDA_EINVAL, "row_ptr is null");
  const DeviceCorpus& cp = c->corp;
  // compact on the device: exclusive scan of the rows' nnz (two-level, as the row planner does),
  // then one pass that unpacks (topic << 16 | count) into the two output arrays
  const int64_t nblocks = (cp.D + kScanBlock - 1) / kScanBlock;
  int64_t* d_ptr = nullptr;
  int64_t total = 0;
  if (cp.D > 0) {
    TRY(ensure_stage(c, sizeof(int64_t) * ((size_t)cp.D + 1 + 2 * (size_t)(nblocks + 1))));
    d_ptr = reinterpret_cast<int64_t*>(c->d_stage);
    unsigned long long* d_bsum = reinterpret_cast<unsigned long long*>(d_ptr + cp.D + 1);
    long long* d_boff = reinterpret_cast<long long*>(d_bsum + nblocks + 1);
    k_nnz_block_sums<<<(unsigned)nblocks, kScanBlock, 0, c->stream>>>(cp.D, cp.d_row_nnz, d_bsum);
    k_exclusive_scan_u64<<<1, 1024, 0, c->stream>>>((int)nblocks, d_bsum, d_boff);
    k_nnz_ptr_apply<<<(unsigned)nblocks, kScanBlock, 0, c->stream>>>(cp.D, cp.d_row_nnz, d_boff, d_ptr);
    c->launches += 3;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(row_ptr, d_ptr, sizeof(int64_t) * (cp.D + 1), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    total = row_ptr[cp.D];
  } else {
    row_ptr[0] = 0;
  }
  if (!topic || !count || total == 0) return B200LDA_OK;
  int32_t* d_tc = nullptr;  // [topic | count]
  TRY(dev_alloc_t(c, &d_tc, 2 * (size_t)total));
  auto done = [&](int code) {
    c->device_bytes -= (int64_t)(sizeof(int32_t) * 2 * (size_t)total);
    dev_free(d_tc);
    return code;
  };
  k_unpack_rows<<<grid_for(c, cp.D * 32, 256), 256, 0, c->stream>>>(cp.D, cp.d_row_ptr, cp.d_row_nnz, cp.d_rows, d_ptr, d_tc,
                                                                   d_tc + total);
  c->launches += 1;
  if (cudaGetLastError() != cudaSuccess ||
      cudaMemcpyAsync(topic, d_tc, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
      cudaMemcpyAsync(count, d_tc + total, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
      cudaStreamSynchronize(c->stream) != cudaSuccess)
    return done(fail(B200LDA_ECUDA, "n_dk compaction failed: %s", cudaGetErrorString(cudaGetLastError())));
  return done(B200LDA_OK);
}

int b200lda_get_word_order(b200lda_ctx* c, int64_t* word_ptr, int64_t* word_tokens) {
  TRY(enter(c));
  if (!c->corpus_loaded) return fail(B200LDA_ESTATE, "no corpus loaded (call b200lda_load_corpus)");
  if (!word_ptr) return fail(B200LDA_EINVAL, "word_ptr is null");
  DeviceCorpus& cp = c->corp;
  TRY(build_word_order(c, cp));
  static_assert(sizeof(long long) == sizeof(int64_t), "word_ptr layout");
  CU(cudaMemcpyAsync(word_ptr, cp.d_word_ptr, sizeof(int64_t) * ((size_t)c->V + 1), cudaMemcpyDeviceToHost, c->stream));
  if (word_tokens && cp.N > 0)
    CU(cudaMemcpyAsync(word_tokens, cp.d_wtok, sizeof(int64_t) * cp.N, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return B200LDA_OK;
}

int b200lda_get_theta(b200lda_ctx* c, int64_t doc_begin, int64_t doc_end, double* theta) {
  TRY(enter(c));
  TRY(need_ready(c));
  const DeviceCorpus& cp = c->corp;
  if (doc_begin < 0 || doc_end > cp.D || doc_begin > doc_end) return fail(B200LDA_EINVAL, "bad document range");
  if (doc_begin == doc_end) return B200LDA_OK;
  if (!theta) return fail(B200LDA_EINVAL, "theta is null");
  const int64_t chunk = std::max<int64_t>(1, (int64_t)(256u << 20) / (int64_t)(sizeof(double) * c->K));
  for (int64_t d0 = doc_begin; d0 < doc_end; d0 += chunk) {
    const int64_t d1 = std::min(doc_end, d0 + chunk);
    TRY(ensure_stage(c, sizeof(double) * (size_t)(d1 - d0) * c->K));
    double* d_theta = reinterpret_cast<double*>(c->d_stage);
    k_theta<<<grid_for(c, (d1 - d0) * 32, 256), 256, 0, c->stream>>>(d0, d1, c->K, cp.d_doc_ptr, cp.d_row_ptr, cp.d_row_nnz,
                                                                    cp.d_rows, c->d_alpha, c->alpha_sum, d_theta);
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(theta + (size_t)(d0 - doc_begin) * c->K, d_theta, sizeof(double) * (size_t)(d1 - d0) * c->K,
                       cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return B200LDA_OK;
}

int b200lda_get_phi(b200lda_ctx* c, double* phi) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!phi) return fail(B200LDA_EINVAL, "phi is null");
  const int kchunk = std::max<int>(32, (int)(((int64_t)(256u << 20) / (int64_t)(sizeof(double) * c->V)) / 32 * 32));
  for (int k0 = 0; k0 < c->K; k0 += kchunk) {
    const int k1 = std::min(c->K, k0 + kchunk);
    TRY(ensure_stage(c, sizeof(double) * (size_t)(k1 - k0) * c->V));
    double* d_phi = reinterpret_cast<double*>(c->d_stage);
    dim3 grid((c->V + 31) / 32, (k1 - k0 + 31) / 32);
    k_phi<<<grid, 256, 0, c->stream>>>(c->V, c->K, k0, k1, c->d_nwk, c->d_nk, c->beta, d_phi);
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(phi + (size_t)k0 * c->V, d_phi, sizeof(double) * (size_t)(k1 - k0) * c->V, cudaMemcpyDeviceToHost,
                       c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return B200LDA_OK;
}

int b200lda_set_alpha(b200lda_ctx* c, const double* alpha) {
  TRY(enter(c));
  if (!alpha) return fail(B200LDA_EINVAL, "alpha is null");
  if (c->in_sweep) return fail(B200LDA_ESTATE, "b200lda_sweep_begin without b200lda_sweep_end");
  for (int k = 0; k < c->K; ++k)
    if (!(alpha[k] > 0.0)) return fail(B200LDA_EINVAL, "alpha[%d] must be > 0", k);
  c->alpha.assign(alpha, alpha + c->K);
  return push_alpha(c);
}

int b200lda_get_alpha(b200lda_ctx* c, double* alpha) {
  if (!c || !alpha) return fail(B200LDA_EINVAL, "null argument");
  std::copy(c->alpha.begin(), c->alpha.end(), alpha);
  return B200LDA_OK;
}

int b200lda_set_beta(b200lda_ctx* c, double beta) {
  if (!c) return fail(B200LDA_EINVAL, "null context");
  if (!(beta > 0.0)) return fail(B200LDA_EINVAL, "beta must be > 0");
  if (c->in_sweep) return fail(B200LDA_ESTATE, "b200lda_sweep_begin without b200lda_sweep_end");
  c->beta = beta;
  return B200LDA_OK;
}

int b200lda_set_sweep_counter(b200lda_ctx* c, int64_t sweeps_done) {
  if (!c) return fail(B200LDA_EINVAL, "null context");
  if (sweeps_done < 0) return fail(B200LDA_EINVAL, "negative sweep counter");
  c->sweeps_done = sweeps_done;
  return B200LDA_OK;
}

// ---- hyper-parameter optimisation --------------------------------------------------------------

int b200lda_hyper_begin(b200lda_ctx* c, int32_t width) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (width <= c->corp.max_doc_len)
    return fail(B200LDA_EINVAL, "histogram width %d must exceed the longest document (%d tokens)", width,
                c->corp.max_doc_len);
  const size_t cells = ((size_t)c->K + 1) * (size_t)width;
  if (width != c->hyper_width) {
    dev_free(c->d_hyper);
    TRY(dev_alloc_t(c, &c->d_hyper, cells));
    c->hyper_width = width;
  }
  CU(cudaMemsetAsync(c->d_hyper, 0, sizeof(int32_t) * cells, c->stream));
  c->hyper_samples = 0;
  return B200LDA_OK;
}

int b200lda_hyper_collect(b200lda_ctx* c) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!c->d_hyper) return fail(B200LDA_ESTATE, "call b200lda_hyper_begin first");
  const DeviceCorpus& cp = c->corp;
  if (cp.D > 0) {
    k_hyper_collect<<<grid_for(c, cp.D * 32, 256), 256, 0, c->stream>>>(cp.D, c->K, c->hyper_width, cp.d_doc_ptr,
                                                                       cp.d_row_ptr, cp.d_row_nnz, cp.d_rows, c->d_hyper);
    c->launches += 1;
    CU(cudaGetLastError());
  }
  c->hyper_samples += 1;
  return B200LDA_OK;
}

int b200lda_hyper_buffer(b200lda_ctx* c, void** d_buf, int64_t* count) {
  if (!c || !d_buf || !count) return fail(B200LDA_EINVAL, "null argument");
  *d_buf = c->d_hyper;
  *count = c->d_hyper ? (int64_t)(((size_t)c->K + 1) * (size_t)c->hyper_width) : 0;
  return B200LDA_OK;
}

int b200lda_hyper_get(b200lda_ctx* c, int32_t* topic_doc_counts, int32_t* doc_length_counts) {
  TRY(enter(c));
  if (!c->d_hyper) return fail(B200LDA_ESTATE, "call b200lda_hyper_begin first");
  const size_t w = (size_t)c->hyper_width;
  if (topic_doc_counts)
    CU(cudaMemcpyAsync(topic_doc_counts, c->d_hyper, sizeof(int32_t) * (size_t)c->K * w, cudaMemcpyDeviceToHost, c->stream));
  if (doc_length_counts)
    CU(cudaMemcpyAsync(doc_length_counts, c->d_hyper + (size_t)c->K * w, sizeof(int32_t) * w, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return B200LDA_OK;
}

// ParallelTopicModel.optimizeAlpha: Dirichlet.learnParameters(alpha, topicDocCounts, docLengthCounts)
// = Minka's fixed point on the histograms with a Gamma(shape 1.00001, scale 1) prior, 200 rounds.
int b200lda_optimize_alpha(b200lda_ctx* c) {
  TRY(enter(c));
  TRY(need_ready(c));
  if (!c->d_hyper) return fail(B200LDA_ESTATE, "call b200lda_hyper_begin / b200lda_hyper_collect first");
  const int K = c->K, W = c->hyper_width;
  std::vector<int32_t> hist(((size_t)K + 1) * (size_t)W);
  CU(cudaMemcpyAsync(hist.data(), c->d_hyper, sizeof(int32_t) * hist.size(), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  const int32_t* lengths = hist.data() + (size_t)K * W;
  const double shape = 1.00001, scale = 1.0;
  std::vector<int> limit((size_t)K, -1);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < W; ++n)
      if (hist[(size_t)k * W + n] > 0) limit[(size_t)k] = n;
  std::vector<double> a = c->alpha;
  double sum = 0.0;
  for (double v : a) sum += v;
  for (int it = 0; it < 200; ++it) {
    double denom = 0.0, dg = 0.0;
    for (int i = 1; i < W; ++i) {
      dg += 1.0 / (sum + i - 1);
      denom += lengths[i] * dg;
    }
    denom -= 1.0 / scale;
    sum = 0.0;
    for (int k = 0; k < K; ++k) {
      const double old = a[(size_t)k];
      const int32_t* h = hist.data() + (size_t)k * W;
      double acc = 0.0;
      dg = 0.0;
      for (int i = 1; i <= limit[(size_t)k]; ++i) {
        dg += 1.0 / (old + i - 1);
        acc += h[i] * dg;
      }
      a[(size_t)k] = old * (acc + shape) / denom;
      sum += a[(size_t)k];
    }
  }
  for (int k = 0; k < K; ++k)
    if (!(a[(size_t)k] > 0.0) || !std::isfinite(a[(size_t)k]))
      return fail(B200LDA_ERANGE, "alpha optimisation diverged at topic %d (no statistics collected?)", k);
  c->alpha = a;
  TRY(push_alpha(c));
  CU(cudaMemsetAsync(c->d_hyper, 0, sizeof(int32_t) * hist.size(), c->stream));  // Mallet clears the histograms
  c->hyper_samples = 0;
  return B200LDA_OK;
}

// ParallelTopicModel.optimizeBeta: Dirichlet.learnSymmetricConcentration on the histogram of cell
// values of n_wk and the topic sizes n_k; 200 fixed-point rounds, each one streaming pass over n_wk.
int b200lda_optimize_beta(b200lda_ctx* c) {
  TRY(enter(c));
  TRY(need_ready(c));
  const size_t VK = (size_t)c->V * c->K;
  std::vector<int32_t> nk((size_t)c->K);
  CU(cudaMemcpyAsync(nk.data(), c->d_nk, sizeof(int32_t) * c->K, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  auto digamma = [](double z) {  // recurrence + asymptotic series
    double psi = 0.0;
    while (z < 10.0) {
      psi -= 1.0 / z;
      z += 1.0;
    }
    const double iz = 1.0 / z, iz2 = iz * iz;
    return psi + std::log(z) - 0.5 * iz -
           iz2 * (1.0 / 12 - iz2 * (1.0 / 120 - iz2 * (1.0 / 252 - iz2 * (1.0 / 240 - iz2 * (1.0 / 132)))));
  };
  double beta_sum = c->beta * (double)c->V;
  double* d_part = c->d_partial;  // kPartial block partials, summed on the host in a fixed order
  std::vector<double> part((size_t)kPartial);
  for (int it = 1; it <= 200; ++it) {
    const double param = beta_sum / (double)c->V;
    k_beta_numerator<<<kPartial, 256, 0, c->stream>>>(VK, c->d_nwk, param, d_part);
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(part.data(), d_part, sizeof(double) * kPartial, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    double numerator = 0.0;
    for (double v : part) numerator += v;
    const double base = digamma(beta_sum);
    double denominator = 0.0;
    for (int k = 0; k < c->K; ++k)
      if (nk[(size_t)k] > 0) denominator += digamma(beta_sum + (double)nk[(size_t)k]) - base;
    if (!(denominator > 0.0) || !(numerator > 0.0)) return fail(B200LDA_ERANGE, "beta optimisation has no statistics");
    const double next = param * numerator / denominator;
    const bool settled = std::fabs(next - beta_sum) <= 1e-10 * std::fabs(beta_sum);  // Mallet runs all 200 rounds; past
    beta_sum = next;                                                                  // this point they change nothing
    if (settled) break;
  }
  if (!(beta_sum > 0.0) || !std::isfinite(beta_sum)) return fail(B200LDA_ERANGE, "beta optimisation diverged");
  c->beta = beta_sum / (double)c->V;
  return B200LDA_OK;
}

int b200lda_get_beta(b200lda_ctx* c, double* beta) {
  if (!c || !beta) return fail(B200LDA_EINVAL, "null argument");
  *beta = c->beta;
  return B200LDA_OK;
}

int b200lda_get_stats(b200lda_ctx* c, b200lda_stats* out) {
  TRY(enter(c));
  if (!out) return fail(B200LDA_EINVAL, "null argument");
  memset(out, 0, sizeof(*out));
  CU(cudaStreamSynchronize(c->stream));
  out->num_docs = c->corp.D;
  out->num_tokens = c->corp.N;
  out->sweeps_done = c->sweeps_done;
  out->kernel_launches = c->launches;
  out->tokens_sampled = c->tokens_sampled;
  out->device_bytes = c->device_bytes;
  if (!c->corp.classes.empty()) {
    const SweepShape& bulk = c->corp.classes.back().shape;    // narrowest rows
    const SweepShape& tail = c->corp.classes.front().shape;   // widest rows
    out->smem_bytes_per_cta = (int32_t)bulk.smem;
    out->warps_per_cta = bulk.warps_per_cta;
    out->ctas = bulk.ctas;
    out->slot_capacity = bulk.slot_cap;
    out->long_docs = c->corp.classes.size() > 1 ? c->corp.classes.front().end - c->corp.classes.front().begin : 0;
    out->long_slot_capacity = tail.slot_cap;
    out->long_ctas = tail.ctas;
    out->row_classes = (int32_t)c->corp.classes.size();
    out->table_refresh_last = c->last_refresh;
    out->hot_words = std::max(0, c->hot_count);
  }
  if (c->in_sweep) return B200LDA_OK;  // timings of an open sweep are not resolvable yet
  TRY(resolve_events(c));
  out->last_tables_ms = c->last_tables_ms;
  out->last_sample_ms = c->last_sample_ms;
  out->last_finish_ms = c->last_finish_ms;
  out->last_sweep_ms = c->last_tables_ms + c->last_sample_ms + c->last_finish_ms;
  out->cum_sweeps = c->cum_sweeps;
  out->cum_tables_ms = c->cum_tables_ms;
  out->cum_sample_ms = c->cum_sample_ms;
  out->cum_finish_ms = c->cum_finish_ms;
  unsigned long long h[kCounters];
  CU(cudaMemcpy(h, c->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
  out->tokens_moved_last = (int64_t)h[1];
  out->prior_bucket_last = (int64_t)h[2];
  out->mean_doc_topics = c->corp.N > 0 ? (double)h[3] / (double)c->corp.N : 0.0;
  out->cum_tokens_moved = (int64_t)h[5];
  out->cum_prior_bucket = (int64_t)h[6];
  out->cum_doc_topics = (int64_t)h[7];
  out->rows_refreshed_last = (int64_t)h[12] + c->rows_rebuilt_by_host;
  return B200LDA_OK;
}

int b200lda_reset_stats(b200lda_ctx* c) {
  TRY(enter(c));
  if (c->in_sweep) return fail(B200LDA_ESTATE, "b200lda_sweep_begin without b200lda_sweep_end");
  TRY(resolve_events(c));
  c->cum_tables_ms = c->cum_sample_ms = c->cum_finish_ms = 0.0;
  c->cum_sweeps = 0;
  CU(cudaMemset(c->d_counters + 5, 0, sizeof(unsigned long long) * 3));
  return B200LDA_OK;
}

// In-process all-reduce(sum) of one buffer kind over n contexts (one per GPU, or several per GPU):
// the functional equivalent of the NCCL all-reduce for hosts that keep all shards in one process
// (a JVM with setNumThreads(4), the C++ mirror). Gathers onto the first context's device with peer
// copies, adds there, copies the sum back. Blocking.
int b200lda_group_allreduce(b200lda_ctx** ctxs, int32_t n, int32_t which) {
  if (!ctxs || n < 1) return fail(B200LDA_EINVAL, "bad context list");
  if (which != B200LDA_BUFFER_EXCHANGE && which != B200LDA_BUFFER_HYPER) return fail(B200LDA_EINVAL, "bad buffer kind");
  std::vector<int32_t*> bufs((size_t)n);
  size_t count = 0;
  bool all_comm = true;
  for (int32_t i = 0; i < n; ++i) {
    b200lda_ctx* c = ctxs[i];
    if (!c) return fail(B200LDA_EINVAL, "null context");
    int32_t* b = which == B200LDA_BUFFER_EXCHANGE ? (c->cfg.world_size > 1 ? c->d_nwk : nullptr) : c->d_hyper;
    const size_t cnt = which == B200LDA_BUFFER_EXCHANGE ? (size_t)c->V * c->K + c->K
                                                        : ((size_t)c->K + 1) * (size_t)c->hyper_width;
    if (!b) return fail(B200LDA_ESTATE, "context %d has no such buffer (world_size == 1, or hyper_begin not called)", i);
    if (i > 0 && cnt != count) return fail(B200LDA_EINVAL, "contexts disagree on the buffer size");
    count = cnt;
    bufs[(size_t)i] = b;
    all_comm = all_comm && c->comm != nullptr;
  }
  if (n == 1 && !(all_comm && ctxs[0]->cfg.world_size > 1)) return B200LDA_OK;
  if (all_comm) {  // one NCCL all-reduce per context, grouped, each on its context's stream: not blocking
    TRY(load_nccl());
    NCCL(g_nccl.GroupStart());
    for (int32_t i = 0; i < n; ++i)
      NCCL(g_nccl.AllReduce(bufs[(size_t)i], bufs[(size_t)i], count, ncclInt32, ncclSum, ctxs[i]->comm, ctxs[i]->stream));
    NCCL(g_nccl.GroupEnd());
    return B200LDA_OK;
  }
  // No communicators (several contexts on ONE device, which NCCL refuses; tests): gather onto the
  // first context's device with peer copies, add there, copy the sum back. Blocking.
  for (int32_t i = 0; i < n; ++i) {
    CU(cudaSetDevice(ctxs[i]->cfg.device));
    CU(cudaStreamSynchronize(ctxs[i]->stream));
  }
  b200lda_ctx* root = ctxs[0];
  CU(cudaSetDevice(root->cfg.device));
  TRY(ensure_stage(root, sizeof(int32_t) * count));
  int32_t* tmp = reinterpret_cast<int32_t*>(root->d_stage);
  for (int32_t i = 1; i < n; ++i) {
    CU(cudaMemcpyPeerAsync(tmp, root->cfg.device, bufs[(size_t)i], ctxs[i]->cfg.device, sizeof(int32_t) * count, root->stream));
    k_add_i32<<<grid_for(root, (int64_t)(count / 4 + 1), 256), 256, 0, root->stream>>>(count, bufs[0], tmp);
    root->launches += 1;
  }
  CU(cudaGetLastError());
  for (int32_t i = 1; i < n; ++i)
    CU(cudaMemcpyPeerAsync(bufs[(size_t)i], ctxs[i]->cfg.device, bufs[0], root->cfg.device, sizeof(int32_t) * count, root->stream));
  CU(cudaStreamSynchronize(root->stream));
  return B200LDA_OK;
}

// ---- NCCL inside the library ------------------------------------------------------------------

namespace {
// A replica that NCCL allocated is registered with the communicator (zero-copy collectives on
// NVSwitch); failure only means the default staged path.
void register_replica(b200lda_ctx* c) {
  if (!c->nwk_from_nccl || !g_nccl.CommRegister || c->nwk_reg) return;
  const size_t bytes = sizeof(int32_t) * ((size_t)c->V * c->K + c->K);
  if (g_nccl.CommRegister(c->comm, c->d_nwk, bytes, &c->nwk_reg) != ncclSuccess) c->nwk_reg = nullptr;
}
}  // namespace

int b200lda_nccl_unique_id(void* id) {
  if (!id) return fail(B200LDA_EINVAL, "null argument");
  TRY(load_nccl());
  static_assert(sizeof(ncclUniqueId) == B200LDA_NCCL_ID_BYTES, "ncclUniqueId size");
  NCCL(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(id)));
  return B200LDA_OK;
}

int b200lda_comm_init(b200lda_ctx* c, const void* id) {
  TRY(enter(c));
  if (!id) return fail(B200LDA_EINVAL, "null argument");
  if (c->cfg.world_size <= 1) return fail(B200LDA_ESTATE, "a communicator needs world_size > 1");
  if (c->comm) return fail(B200LDA_ESTATE, "the context already has a communicator");
  TRY(load_nccl());
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  NCCL(g_nccl.CommInitRank(&c->comm, c->cfg.world_size, uid, c->cfg.rank));
  register_replica(c);
  return B200LDA_OK;
}

int b200lda_group_comm_init(b200lda_ctx** ctxs, int32_t n) {
  if (!ctxs || n < 2) return fail(B200LDA_EINVAL, "bad context list");
  TRY(load_nccl());
  std::vector<int> devs((size_t)n);
  for (int32_t i = 0; i < n; ++i) {
    if (!ctxs[i]) return fail(B200LDA_EINVAL, "null context");
    if (ctxs[i]->cfg.world_size != n || ctxs[i]->cfg.rank != i)
      return fail(B200LDA_EINVAL, "context %d must be rank %d of %d", i, i, n);
    if (ctxs[i]->comm) return fail(B200LDA_ESTATE, "context %d already has a communicator", i);
    devs[(size_t)i] = ctxs[i]->cfg.device;
    for (int32_t j = 0; j < i; ++j)
      if (devs[(size_t)j] == devs[(size_t)i])
        return fail(B200LDA_EINVAL, "contexts %d and %d share device %d: NCCL needs one GPU per shard", j, i, devs[(size_t)i]);
  }
  std::vector<ncclComm_t> comms((size_t)n);
  NCCL(g_nccl.CommInitAll(comms.data(), n, devs.data()));
  for (int32_t i = 0; i < n; ++i) {
    ctxs[i]->comm = comms[(size_t)i];
    cudaSetDevice(ctxs[i]->cfg.device);
    register_replica(ctxs[i]);
  }
  return B200LDA_OK;
}

int b200lda_group_sync_counts(b200lda_ctx** ctxs, int32_t n) {
  if (!ctxs || n < 1) return fail(B200LDA_EINVAL, "bad context list");
  for (int32_t i = 0; i < n; ++i)
    if (!ctxs[i]) return fail(B200LDA_EINVAL, "null context");
  if (n == 1 && ctxs[0]->cfg.world_size == 1) return B200LDA_OK;
  if (n < ctxs[0]->cfg.world_size)  // the other shards live in other processes: only NCCL reaches them
    for (int32_t i = 0; i < n; ++i)
      if (!ctxs[i]->comm)
        return fail(B200LDA_ESTATE, "%d of %d shards given and context %d has no communicator (b200lda_comm_init)", n,
                    ctxs[0]->cfg.world_size, i);
  for (int32_t i = 0; i < n; ++i) {
    const int rc = b200lda_counts_sync_begin(ctxs[i]);
    if (rc != B200LDA_OK) {  // nobody stays inside a half-opened sync
      for (int32_t j = 0; j < i; ++j) ctxs[j]->in_sync = false;
      return rc;
    }
  }
  // b200lda_group_allreduce: grouped NCCL all-reduces when every context has a communicator, peer copies otherwise
  const int rc = b200lda_group_allreduce(ctxs, n, B200LDA_BUFFER_EXCHANGE);
  if (rc != B200LDA_OK) {
    for (int32_t i = 0; i < n; ++i) ctxs[i]->in_sync = false;
    return rc;
  }
  for (int32_t i = 0; i < n; ++i) TRY(b200lda_counts_sync_end(ctxs[i]));
  return B200LDA_OK;
}

int b200lda_group_sweep(b200lda_ctx** ctxs, int32_t n, int32_t sweeps) {
  if (!ctxs || n < 1 || sweeps < 0) return fail(B200LDA_EINVAL, "bad arguments");
  for (int32_t i = 0; i < n; ++i)
    if (!ctxs[i] || ctxs[i]->cfg.world_size != ctxs[0]->cfg.world_size) return fail(B200LDA_EINVAL, "bad context list");
  const bool multi = ctxs[0]->cfg.world_size > 1;
  bool all_comm = true;
  for (int32_t i = 0; i < n; ++i) all_comm = all_comm && ctxs[i]->comm != nullptr;
  for (int32_t it = 0; it < sweeps; ++it) {
    for (int32_t i = 0; i < n; ++i) TRY(b200lda_sweep_begin(ctxs[i]));
    if (multi && all_comm) {
      TRY(exchange_nccl(ctxs, n));
      for (int32_t i = 0; i < n; ++i) {
        CU(cudaSetDevice(ctxs[i]->cfg.device));
        TRY(sweep_close(ctxs[i]));
      }
    } else {
      if (multi) TRY(b200lda_group_allreduce(ctxs, n, B200LDA_BUFFER_EXCHANGE));
      for (int32_t i = 0; i < n; ++i) TRY(b200lda_sweep_end(ctxs[i]));
    }
  }
  for (int32_t i = 0; i < n; ++i) TRY(b200lda_synchronize(ctxs[i]));
  return B200LDA_OK;
}

// ---- resumable state ----------------------------------------------------------------------------

namespace {
struct StateHeader {
  char magic[8];
  int32_t version, K, V, mode;
  int64_t D, N, sweeps_done;
  uint64_t seed, corpus_hash;
  double beta;
  int32_t rank, world_size;
  int64_t global_token_offset;
};
constexpr char kStateMagic[8] = {'B', '2', '0', '0', 'L', 'D', 'A', '1'};

int corpus_hash(b200lda_ctx* c, uint64_t* out) {
  const DeviceCorpus& cp = c->corp;
  unsigned long long* d_h = c->d_counters + 13;
  CU(cudaMemsetAsync(d_h, 0, sizeof(unsigned long long), c->stream));
  if (cp.D > 0) k_hash_i64<<<grid_for(c, cp.D + 1, 256), 256, 0, c->stream>>>(cp.D + 1, cp.d_doc_ptr, d_h);
  if (cp.N > 0) k_hash_i32<<<grid_for(c, cp.N, 256), 256, 0, c->stream>>>(cp.N, cp.d_tok_word, d_h);
  c->launches += 2;
  unsigned long long h = 0;
  CU(cudaMemcpyAsync(&h, d_h, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  *out = (uint64_t)h;
  return B200LDA_OK;
}
}  // namespace

int b200lda_state_size(b200lda_ctx* c, int64_t* bytes) {
  if (!c || !bytes) return fail(B200LDA_EINVAL, "null argument");
  *bytes = (int64_t)(sizeof(StateHeader) + sizeof(double) * (size_t)c->K + sizeof(uint16_t) * (size_t)c->corp.N);
  return B200LDA_OK;
}

int b200lda_get_state(b200lda_ctx* c, void* buf, int64_t bytes) {
  TRY(enter(c));
  TRY(need_ready(c));
  int64_t need = 0;
  TRY(b200lda_state_size(c, &need));
  if (!buf || bytes < need) return fail(B200LDA_EINVAL, "state buffer too small (%lld bytes needed)", (long long)need);
  StateHeader h{};
  memcpy(h.magic, kStateMagic, 8);
  h.version = 1;
  h.K = c->K;
  h.V = c->V;
  h.mode = c->cfg.mode;
  h.D = c->corp.D;
  h.N = c->corp.N;
  h.sweeps_done = c->sweeps_done;
  h.seed = c->cfg.seed;
  h.beta = c->beta;
  h.rank = c->cfg.rank;
  h.world_size = c->cfg.world_size;
  h.global_token_offset = c->cfg.global_token_offset;
  TRY(corpus_hash(c, &h.corpus_hash));
  char* out = static_cast<char*>(buf);
  memcpy(out, &h, sizeof(h));
  memcpy(out + sizeof(h), c->alpha.data(), sizeof(double) * (size_t)c->K);
  if (c->corp.N > 0) {
    CU(cudaMemcpyAsync(out + sizeof(h) + sizeof(double) * (size_t)c->K, c->corp.d_z, sizeof(uint16_t) * c->corp.N,
                       cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  return B200LDA_OK;
}

int b200lda_set_state(b200lda_ctx* c, const void* buf, int64_t bytes) {
  TRY(enter(c));
  if (!c->corpus_loaded) return fail(B200LDA_ESTATE, "no corpus loaded (call b200lda_load_corpus)");
  if (c->in_sweep || c->in_sync) return fail(B200LDA_ESTATE, "a sweep or count sync is open");
  if (!buf || bytes < (int64_t)sizeof(StateHeader)) return fail(B200LDA_EINVAL, "not a state blob");
  StateHeader h;
  memcpy(&h, buf, sizeof(h));
  if (memcmp(h.magic, kStateMagic, 8) != 0 || h.version != 1) return fail(B200LDA_EINVAL, "not a b200lda state blob (magic / version)");
  if (h.K != c->K || h.V != c->V) return fail(B200LDA_EINVAL, "state is for K=%d V=%d, context has K=%d V=%d", h.K, h.V, c->K, c->V);
  if (h.D != c->corp.D || h.N != c->corp.N)
    return fail(B200LDA_EINVAL, "state is for %lld documents / %lld tokens, the loaded corpus has %lld / %lld", (long long)h.D,
                (long long)h.N, (long long)c->corp.D, (long long)c->corp.N);
  const int64_t need = (int64_t)(sizeof(StateHeader) + sizeof(double) * (size_t)c->K + sizeof(uint16_t) * (size_t)h.N);
  if (bytes < need) return fail(B200LDA_EINVAL, "state blob truncated (%lld of %lld bytes)", (long long)bytes, (long long)need);
  uint64_t hash = 0;
  TRY(corpus_hash(c, &hash));
  if (hash != h.corpus_hash) return fail(B200LDA_EINVAL, "state belongs to another corpus (checksum of doc_ptr / tok_word differs)");
  const char* in = static_cast<const char*>(buf);
  std::vector<double> alpha((size_t)c->K);
  memcpy(alpha.data(), in + sizeof(h), sizeof(double) * (size_t)c->K);
  TRY(b200lda_set_alpha(c, alpha.data()));
  TRY(b200lda_set_beta(c, h.beta));
  c->cfg.seed = h.seed;  // the Philox key is part of the chain's identity
  c->cfg.global_token_offset = h.global_token_offset;
  TRY(init_assignments_impl(c, nullptr, reinterpret_cast<const uint16_t*>(in + sizeof(h) + sizeof(double) * (size_t)c->K)));
  c->sweeps_done = h.sweeps_done;
  return B200LDA_OK;
}

int b200lda_host_alloc(void** out, size_t bytes) {
  if (!out) return fail(B200LDA_EINVAL, "null argument");
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(B200LDA_ENOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  }
  return B200LDA_OK;
}

int b200lda_host_free(void* p) {
  if (!p) return B200LDA_OK;
  if (cudaFreeHost(p) != cudaSuccess) {
    cudaGetLastError();
    return fail(B200LDA_ECUDA, "cudaFreeHost failed");
  }
  return B200LDA_OK;
}

}  // extern "C"
