// index.cu — host side of libcqs_b200.so: the C ABI declared in
// include/cqs_b200.h.  Owns device memory, streams and the per-index mutex;
// all arithmetic lives in the kernels (scan_single.cu, scan_batch.cu,
// sparse_fuse.cu).  There is deliberately NO CPU fallback: every search either
// launches the CUDA kernels or returns an error.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/cqs_b200.h"
#include "internal.h"
#include "peer_host.h"

namespace cqs {
std::atomic<uint64_t> g_kernel_launches{0};
}
using namespace cqs;

static thread_local std::string t_last_error;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}

// t_last_error itself may be what failed to allocate: fall back to a static message
static int api_exception(int code, const char* msg) noexcept {
  try {
    t_last_error = msg;
  } catch (...) {
  }
  return code;
}

// used by the other translation units (peer.cu) to report through the same thread-local slot
__attribute__((visibility("hidden"))) int cqs_b200_internal_fail(int code, const char* msg) {
  return fail(code, "%s", msg);
}

namespace {

struct Shard {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  uint8_t* d_rows = nullptr;
  uint8_t* d_rows16 = nullptr;  // STORAGE_BF16_F32: bf16 shadow of the f32 master rows
  uint64_t n_rows = 0, cap_rows = 0;
  uint64_t first_row = 0;  // offset of this shard's row 0 inside the index
  float* d_stage = nullptr;  // staging for row ingest
  uint64_t stage_rows = 0;
  float* h_query = nullptr;   // pinned
  uint32_t* h_sq = nullptr;   // pinned staging of a sparse query: [kSpMaxQ] tokens | [kSpMaxQ] weights
  uint32_t* d_bitset = nullptr;
  uint64_t bitset_words = 0;
  // Scratch of one dense scan launch: per-CTA partial lists, CTA ticket + tile counter, the
  // zero-padded query.  kLanes sets, used round-robin, so that a caller that rotates over
  // kLanes streams gets the tail of launch i (list merge, cross-shard exchange) overlapped with
  // the streaming phase of the next launches; a launch only waits for the launch kLanes back.
  struct ScanScratch {
    ckey_t* d_partial = nullptr;
    uint32_t* d_partial_cnt = nullptr;
    uint32_t* d_done = nullptr;
    ckey_t* d_col = nullptr;       // [kMaxGrid], zero between launches
    float* d_query = nullptr;      // [ld], zero padded beyond dim
    cudaEvent_t ev = nullptr;      // recorded after the last launch that used this set
    cudaStream_t stream = nullptr;
    bool used = false;
  } scr[kLanes];
  uint32_t next_scr = 0;
  // extra launch lanes of the many-query entry points (lane 0 is the caller's stream)
  cudaStream_t lane_stream[kLanes - 1] = {};
  cudaEvent_t ev_fork = nullptr;
  cudaEvent_t ev_join[kLanes - 1] = {};
  // dense result / pool
  float* d_out_scores = nullptr;
  uint64_t* d_out_rows = nullptr;
  uint32_t* d_out_n = nullptr;
  // sparse pool
  float* d_sp_scores = nullptr;
  uint64_t* d_sp_rows = nullptr;
  uint32_t* d_sp_n = nullptr;
  float* d_spm_scores = nullptr;   // cross-shard merged sparse pool (sharded hybrid)
  uint64_t* d_spm_rows = nullptr;
  uint32_t* d_spm_n = nullptr;
  ckey_t* d_sp_partial = nullptr;
  uint32_t* d_sp_partial_cnt = nullptr;
  uint32_t* d_sp_done = nullptr;
  uint32_t* d_q_tok = nullptr;
  float* d_q_w = nullptr;
  uint32_t* d_bounds = nullptr;   // sparse pass-1 scratch, grown on demand
  size_t bounds_bytes = 0;
  void* d_sp_block = nullptr;     // accumulate -> select hand-over: per-doc scores + touched bits, grown on demand
  size_t sp_block_bytes = 0;
  uint32_t* d_sp_claim = nullptr; // block counter of the accumulate kernel
  // fused output
  uint64_t* d_f_rows = nullptr;
  float* d_f_fused = nullptr;
  float* d_f_dense = nullptr;
  float* d_f_sraw = nullptr;
  uint8_t* d_f_present = nullptr;
  uint32_t* d_f_n = nullptr;
  // pinned host mirror for results (sized for the fused output, the largest)
  uint8_t* h_out = nullptr;     // pinned + mapped: [scores kMaxK][rows kMaxK][n][flag] (+ fused layout)
  uint8_t* d_hout = nullptr;    // device alias of h_out
  uint32_t seq = 0;             // completion sequence number of the latency path
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_b0 = nullptr, ev_b1 = nullptr;  // whole batch pipeline (cqs_b200_set_timing)
  unsigned long long* d_trace = nullptr;  // development aid (CQS_B200_TRACE=1)
  // batched tensor-core path (bf16 storage), allocated on first use
  float* d_bq = nullptr;          // [kBatchMaxQ][ld] padded f32 queries
  void* d_bscratch = nullptr;
  float* d_bout_scores = nullptr; // [kBatchMaxQ][kMaxK]
  uint64_t* d_bout_rows = nullptr;
  uint32_t* d_bout_n = nullptr;   // [kBatchMaxQ]
  uint32_t* d_bflags = nullptr;   // [kBatchMaxQ]
  float* d_bm_scores = nullptr;   // [kBatchMaxQ][kMaxK] cross-shard merged results (sharded batch)
  uint64_t* d_bm_rows = nullptr;
  uint32_t* d_bm_n = nullptr;
  float* d_maxnorm = nullptr;     // [1]
  float max_row_norm = 0.f;
  float max_row_delta = 0.f;      // STORAGE_BF16_F32: max |f32 row - bf16 shadow row|_2
  // structured filter / per-row signals (cqs_b200_set_row_meta / _signals)
  uint8_t* d_ctype = nullptr;
  uint8_t* d_lang = nullptr;
  float* d_note_boost = nullptr;
  float* d_importance = nullptr;
  SparseDev sparse;
};

constexpr uint32_t kSpMaxQ = 1024;
constexpr size_t kHostOutBytes = kMaxK * (8 + 4 + 4 + 4 + 1) + 64;
constexpr size_t kOffHybridFlag = kMaxK * (8 + 4 + 4 + 4 + 1) + 8;   // completion word of the fused-pool layout

}  // namespace

struct cqs_b200_index {
  std::mutex mu;
  uint32_t dim = 0;
  RowLayout layout{};    // master rows (what single-query scans read)
  RowLayout layout16{};  // bf16 shadow (STORAGE_BF16_F32 only)
  int metric = 0, storage = 0;
  bool finalized = false;
  std::atomic<int> poisoned{0};
  uint64_t row_base = 0;
  uint64_t n_rows = 0, reserved = 0, rows_per_shard = 0;
  float last_kernel_ms = 0.f;
  bool timing = false;   // record CUDA events around the dominant kernel (cqs_b200_set_timing)
  float max_note_boost = 1.f, max_importance = 1.f;
  uint64_t shadow_reruns = 0;       // single-query shadow scans re-run on the f32 master (cumulative)
  uint32_t batch_reruns_total = 0;  // queries the batched path sent to the exact kernel (cumulative)
  uint32_t last_batch_reruns = 0;   // ... by the most recent batch call
  // Device time of the most recent tensor-core batch call, CUDA events on the shard's stream from
  // "queries resident" to "last kernel done": candidate scan + rescoring + the exact re-runs of
  // unproven queries + (sharded) the gather/merge exchange, summed over the call's <= 1024-query chunks.
  float last_batch_ms = 0.f;
  std::vector<Shard> shards;
};

// Every extern "C" entry point is a function-try-block closed by API_CATCH: a host allocation
// failure (std::vector / std::string) or any other C++ exception becomes an error code instead
// of unwinding through the C ABI into the Rust caller ("never throws, never aborts").
#define API_CATCH                                                                          \
  catch (const std::bad_alloc&) {                                                          \
    return api_exception(CQS_B200_ERR_OOM, "host allocation failed");                      \
  }                                                                                        \
  catch (...) {                                                                            \
    return api_exception(CQS_B200_ERR_INVALID, "unexpected C++ exception");                \
  }

#define CK(ix, expr)                                                                       \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      if (ix) (ix)->poisoned.store(1);                                                     \
      cudaGetLastError();                                                                  \
      return fail(_e == cudaErrorMemoryAllocation ? CQS_B200_ERR_OOM : CQS_B200_ERR_CUDA,  \
                  "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,        \
                  __LINE__);                                                               \
    }                                                                                      \
  } while (0)

static size_t row_bytes(const cqs_b200_index* ix) {
  return (size_t)ix->layout.ld * (ix->layout.mode == 0 ? 4 : 2);
}

static void free_shard(Shard& s) {
  cudaSetDevice(s.device);
  if (s.stream) cudaStreamSynchronize(s.stream);
  cudaFree(s.d_rows); cudaFree(s.d_rows16); cudaFree(s.d_stage); cudaFree(s.d_bitset);
  for (auto& c : s.scr) {
    cudaFree(c.d_partial); cudaFree(c.d_partial_cnt); cudaFree(c.d_done); cudaFree(c.d_query); cudaFree(c.d_col);
    if (c.ev) cudaEventDestroy(c.ev);
  }
  cudaFree(s.d_out_scores); cudaFree(s.d_out_rows); cudaFree(s.d_out_n);
  cudaFree(s.d_sp_scores); cudaFree(s.d_sp_rows); cudaFree(s.d_sp_n);
  cudaFree(s.d_spm_scores); cudaFree(s.d_spm_rows); cudaFree(s.d_spm_n);
  cudaFree(s.d_sp_partial); cudaFree(s.d_sp_partial_cnt); cudaFree(s.d_sp_done);
  cudaFree(s.d_q_tok); cudaFree(s.d_q_w); cudaFree(s.d_bounds); cudaFree(s.d_sp_block); cudaFree(s.d_sp_claim);
  cudaFree(s.d_f_rows); cudaFree(s.d_f_fused); cudaFree(s.d_f_dense); cudaFree(s.d_f_sraw);
  cudaFree(s.d_f_present); cudaFree(s.d_f_n); cudaFree(s.d_trace);
  cudaFree(s.d_bq); cudaFree(s.d_bscratch); cudaFree(s.d_bout_scores); cudaFree(s.d_bout_rows);
  cudaFree(s.d_bout_n); cudaFree(s.d_bflags); cudaFree(s.d_maxnorm);
  cudaFree(s.d_bm_scores); cudaFree(s.d_bm_rows); cudaFree(s.d_bm_n);
  cudaFree(s.d_ctype); cudaFree(s.d_lang); cudaFree(s.d_note_boost); cudaFree(s.d_importance);
  free_sparse(s.sparse);
  if (s.h_query) cudaFreeHost(s.h_query);
  if (s.h_sq) cudaFreeHost(s.h_sq);
  if (s.h_out) cudaFreeHost(s.h_out);
  if (s.ev0) cudaEventDestroy(s.ev0);
  if (s.ev1) cudaEventDestroy(s.ev1);
  if (s.ev_b0) cudaEventDestroy(s.ev_b0);
  if (s.ev_b1) cudaEventDestroy(s.ev_b1);
  if (s.ev_fork) cudaEventDestroy(s.ev_fork);
  for (auto& e : s.ev_join) if (e) cudaEventDestroy(e);
  for (auto& st : s.lane_stream) if (st) cudaStreamDestroy(st);
  if (s.stream) cudaStreamDestroy(s.stream);
  s = Shard();
}

static int init_shard(cqs_b200_index* ix, Shard& s, int device) {
  s.device = device;
  CK(ix, cudaSetDevice(device));
  CK(ix, cudaDeviceGetAttribute(&s.num_sms, cudaDevAttrMultiProcessorCount, device));
  CK(ix, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  for (auto& st : s.lane_stream) CK(ix, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(ix, cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
  for (auto& e : s.ev_join) CK(ix, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(ix, cudaHostAlloc((void**)&s.h_query, sizeof(float) * ix->layout.ld,
                       cudaHostAllocDefault));
  CK(ix, cudaHostAlloc((void**)&s.h_sq, sizeof(uint32_t) * 2 * 1024, cudaHostAllocDefault));
  for (auto& c : s.scr) {
    CK(ix, cudaMalloc((void**)&c.d_query, sizeof(float) * ix->layout.ld));
    CK(ix, cudaMalloc((void**)&c.d_partial, sizeof(ckey_t) * kMaxGrid * kMaxK));
    CK(ix, cudaMalloc((void**)&c.d_partial_cnt, sizeof(uint32_t) * kMaxGrid));
    CK(ix, cudaMalloc((void**)&c.d_done, 4 * sizeof(uint32_t)));
    CK(ix, cudaMemset(c.d_done, 0, 4 * sizeof(uint32_t)));
    CK(ix, cudaMalloc((void**)&c.d_col, sizeof(ckey_t) * 2 * kMaxGrid));
    CK(ix, cudaMemset(c.d_col, 0, sizeof(ckey_t) * 2 * kMaxGrid));
    CK(ix, cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming));
  }
  CK(ix, cudaMalloc((void**)&s.d_out_scores, sizeof(float) * kMaxK));
  CK(ix, cudaMalloc((void**)&s.d_out_rows, sizeof(uint64_t) * kMaxK));
  CK(ix, cudaMalloc((void**)&s.d_out_n, sizeof(uint32_t)));
  CK(ix, cudaMalloc((void**)&s.d_sp_scores, sizeof(float) * kMaxK));
  CK(ix, cudaMalloc((void**)&s.d_sp_rows, sizeof(uint64_t) * kMaxK));
  CK(ix, cudaMalloc((void**)&s.d_sp_n, sizeof(uint32_t)));
  CK(ix, cudaMalloc((void**)&s.d_sp_partial, sizeof(ckey_t) * kMaxGrid * kMaxK));
  CK(ix, cudaMalloc((void**)&s.d_sp_partial_cnt, sizeof(uint32_t) * kMaxGrid));
  CK(ix, cudaMalloc((void**)&s.d_sp_done, sizeof(uint32_t)));
  CK(ix, cudaMemset(s.d_sp_done, 0, sizeof(uint32_t)));
  CK(ix, cudaMalloc((void**)&s.d_sp_claim, sizeof(uint32_t)));
  CK(ix, cudaMemset(s.d_sp_claim, 0, sizeof(uint32_t)));
  CK(ix, cudaMalloc((void**)&s.d_q_tok, sizeof(uint32_t) * kSpMaxQ));
  CK(ix, cudaMalloc((void**)&s.d_q_w, sizeof(float) * kSpMaxQ));
  CK(ix, cudaMalloc((void**)&s.d_f_rows, sizeof(uint64_t) * kMaxK));
  CK(ix, cudaMalloc((void**)&s.d_f_fused, sizeof(float) * kMaxK));
  CK(ix, cudaMalloc((void**)&s.d_f_dense, sizeof(float) * kMaxK));
  CK(ix, cudaMalloc((void**)&s.d_f_sraw, sizeof(float) * kMaxK));
  CK(ix, cudaMalloc((void**)&s.d_f_present, kMaxK));
  CK(ix, cudaMalloc((void**)&s.d_f_n, sizeof(uint32_t)));
  CK(ix, cudaHostAlloc((void**)&s.h_out, kHostOutBytes, cudaHostAllocMapped));
  memset(s.h_out, 0, kHostOutBytes);
  CK(ix, cudaHostGetDevicePointer((void**)&s.d_hout, s.h_out, 0));
  CK(ix, cudaEventCreate(&s.ev0));
  CK(ix, cudaEventCreate(&s.ev1));
  CK(ix, cudaEventCreate(&s.ev_b0));
  CK(ix, cudaEventCreate(&s.ev_b1));
  if (ix->shards.size() == 1) {
    // Query / result buffers of the batch entry points (27 MB), allocated up front so that no
    // search call allocates between its kernels (a device allocation can serialise against
    // kernels that are already waiting for this rank's part of an exchange).  Only the 250 MB
    // scratch of the tensor-core path stays lazy (first bf16 batch).
    const uint32_t ld = ix->layout.ld;
    CK(ix, cudaMalloc((void**)&s.d_bq, sizeof(float) * (size_t)kBatchMaxQ * ld));
    CK(ix, cudaMalloc((void**)&s.d_bout_scores, sizeof(float) * (size_t)kBatchMaxQ * kMaxK));
    CK(ix, cudaMalloc((void**)&s.d_bout_rows, sizeof(uint64_t) * (size_t)kBatchMaxQ * kMaxK));
    CK(ix, cudaMalloc((void**)&s.d_bout_n, sizeof(uint32_t) * kBatchMaxQ));
    CK(ix, cudaMalloc((void**)&s.d_bm_scores, sizeof(float) * (size_t)kBatchMaxQ * kMaxK));
    CK(ix, cudaMalloc((void**)&s.d_bm_rows, sizeof(uint64_t) * (size_t)kBatchMaxQ * kMaxK));
    CK(ix, cudaMalloc((void**)&s.d_bm_n, sizeof(uint32_t) * kBatchMaxQ));
    CK(ix, cudaMalloc((void**)&s.d_spm_scores, sizeof(float) * kMaxK));
    CK(ix, cudaMalloc((void**)&s.d_spm_rows, sizeof(uint64_t) * kMaxK));
    CK(ix, cudaMalloc((void**)&s.d_spm_n, sizeof(uint32_t)));
  }
  if (getenv("CQS_B200_TRACE")) {
    CK(ix, cudaMalloc((void**)&s.d_trace, sizeof(unsigned long long) * kMaxGrid * 8));
    CK(ix, cudaMemset(s.d_trace, 0, sizeof(unsigned long long) * kMaxGrid * 8));
  }
  return 0;
}

static int grow_shard(cqs_b200_index* ix, Shard& s, uint64_t want_rows) {
  if (want_rows <= s.cap_rows) return 0;
  uint64_t cap = std::max<uint64_t>(want_rows, s.cap_rows + s.cap_rows / 2);
  cap = std::max<uint64_t>(cap, 1024);
  CK(ix, cudaSetDevice(s.device));
  uint8_t* nd = nullptr;
  size_t rb = row_bytes(ix);
  CK(ix, cudaMalloc((void**)&nd, cap * rb));
  if (s.n_rows) CK(ix, cudaMemcpyAsync(nd, s.d_rows, s.n_rows * rb, cudaMemcpyDeviceToDevice, s.stream));
  uint8_t* nd16 = nullptr;
  if (ix->storage == CQS_B200_STORAGE_BF16_F32) {
    const size_t rb16 = (size_t)ix->layout16.ld * 2;
    CK(ix, cudaMalloc((void**)&nd16, cap * rb16));
    if (s.n_rows) CK(ix, cudaMemcpyAsync(nd16, s.d_rows16, s.n_rows * rb16, cudaMemcpyDeviceToDevice, s.stream));
  }
  CK(ix, cudaStreamSynchronize(s.stream));
  cudaFree(s.d_rows);
  cudaFree(s.d_rows16);
  s.d_rows = nd;
  s.d_rows16 = nd16;
  s.cap_rows = cap;
  return 0;
}

extern "C" {

int cqs_b200_create(const int* device_ids, int n_dev, uint32_t dim, int metric, int storage,
                    cqs_b200_index** out) try {
  if (!out) return fail(CQS_B200_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (n_dev < 1 || n_dev > 64) return fail(CQS_B200_ERR_INVALID, "n_dev must be in 1..64");
  if (metric != CQS_B200_METRIC_COSINE && metric != CQS_B200_METRIC_DOT)
    return fail(CQS_B200_ERR_INVALID, "unknown metric %d", metric);
  if (storage != CQS_B200_STORAGE_F32 && storage != CQS_B200_STORAGE_BF16 &&
      storage != CQS_B200_STORAGE_BF16_F32)
    return fail(CQS_B200_ERR_INVALID, "unknown storage %d", storage);
  RowLayout lay, lay16{};
  if (!choose_layout(dim, storage == CQS_B200_STORAGE_BF16 ? 1 : 0, &lay))
    return fail(CQS_B200_ERR_INVALID, "unsupported dim %u (1..2048)", dim);
  if (storage == CQS_B200_STORAGE_BF16_F32 && (!choose_layout(dim, 1, &lay16) || lay16.ld != lay.ld))
    return fail(CQS_B200_ERR_UNSUPPORTED, "no common row stride for dim %u", dim);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(CQS_B200_ERR_CUDA, "no CUDA device: %s (this backend has no CPU fallback)",
                cudaGetErrorString(e));
  }
  cqs_b200_index* ix = new (std::nothrow) cqs_b200_index();
  if (!ix) return fail(CQS_B200_ERR_OOM, "host allocation failed");
  ix->dim = dim;
  ix->layout = lay;
  ix->layout16 = lay16;
  ix->metric = metric;
  ix->storage = storage;
  ix->shards.resize(n_dev);
  for (int i = 0; i < n_dev; ++i) {
    int dev = device_ids ? device_ids[i] : i;
    if (dev < 0 || dev >= count) {
      cqs_b200_destroy(ix);
      return fail(CQS_B200_ERR_INVALID, "device id %d out of range (%d devices)", dev, count);
    }
    int rc = init_shard(ix, ix->shards[i], dev);
    if (rc) {
      cqs_b200_destroy(ix);
      return rc;
    }
  }
  *out = ix;
  return CQS_B200_OK;
} API_CATCH

void cqs_b200_destroy(cqs_b200_index* ix) {
  if (!ix) return;
  for (auto& s : ix->shards) free_shard(s);
  delete ix;
}

int cqs_b200_reserve(cqs_b200_index* ix, uint64_t n_rows) try {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  std::lock_guard<std::mutex> g(ix->mu);
  if (ix->finalized) return fail(CQS_B200_ERR_INVALID, "index is finalized");
  if (ix->n_rows) return fail(CQS_B200_ERR_INVALID, "reserve must precede the first append");
  uint64_t nd = ix->shards.size();
  uint64_t rps = (n_rows + nd - 1) / nd;
  rps = (rps + 31) / 32 * 32;  // shard boundaries on bitset word boundaries
  if (rps > 0xFFFFFFFFull) return fail(CQS_B200_ERR_INVALID, "more than 2^32-1 rows per device");
  ix->reserved = n_rows;
  if (nd == 1) {
    // one device: a capacity hint only — the shard regrows on demand, so reopen + append
    // (the TieredIndex::extend analogue) keeps working on built and loaded indexes
    return grow_shard(ix, ix->shards[0], n_rows);
  }
  // several devices: the shard boundaries are fixed here; appends beyond nd * rps are
  // rejected (re-shard by rebuilding the index with a larger reserve)
  ix->rows_per_shard = rps;
  for (uint64_t i = 0; i < nd; ++i) {
    ix->shards[i].first_row = i * rps;
    int rc = grow_shard(ix, ix->shards[i], rps);
    if (rc) return rc;
  }
  return CQS_B200_OK;
} API_CATCH

int cqs_b200_set_row_base(cqs_b200_index* ix, uint64_t row_base) try {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  std::lock_guard<std::mutex> g(ix->mu);
  ix->row_base = row_base;
  return CQS_B200_OK;
} API_CATCH

static int append_impl(cqs_b200_index* ix, const float* rows, uint64_t n_rows, bool on_device) {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  if (n_rows == 0) return CQS_B200_OK;
  if (!rows) return fail(CQS_B200_ERR_INVALID, "rows is NULL");
  std::lock_guard<std::mutex> g(ix->mu);
  if (ix->poisoned.load()) return fail(CQS_B200_ERR_POISONED, "index is poisoned");
  if (ix->finalized) return fail(CQS_B200_ERR_INVALID, "index is finalized (call cqs_b200_reopen)");
  const uint64_t nd = ix->shards.size();
  if (nd > 1 && !ix->rows_per_shard)
    return fail(CQS_B200_ERR_INVALID, "n_dev > 1 requires cqs_b200_reserve before append");
  if (on_device && nd > 1)
    return fail(CQS_B200_ERR_UNSUPPORTED, "device-pointer append needs a single-device index");
  if (ix->rows_per_shard && ix->n_rows + n_rows > ix->rows_per_shard * nd)
    return fail(CQS_B200_ERR_INVALID, "append exceeds the reserved size");
  const size_t rb = row_bytes(ix);
  const bool direct = (ix->layout.mode == 0 && ix->layout.ld == ix->dim);
  uint64_t done = 0;
  while (done < n_rows) {
    uint64_t grow0 = ix->n_rows + done;  // row index inside the index
    uint64_t si = ix->rows_per_shard ? grow0 / ix->rows_per_shard : 0;
    Shard& s = ix->shards[si];
    uint64_t local = grow0 - s.first_row;
    uint64_t room = ix->rows_per_shard ? ix->rows_per_shard - local : (n_rows - done);
    uint64_t take = std::min<uint64_t>(n_rows - done, room);
    if (local + take > 0xFFFFFFFFull) return fail(CQS_B200_ERR_INVALID, "more than 2^32-1 rows per device");
    CK(ix, cudaSetDevice(s.device));
    // a device-side source may still be being produced on another stream of the
    // caller (e.g. a framework's default stream): the build path simply waits.
    if (on_device) CK(ix, cudaDeviceSynchronize());
    int rc = grow_shard(ix, s, local + take);
    if (rc) return rc;
    const float* src = rows + done * ix->dim;
    const bool shadow = ix->storage == CQS_B200_STORAGE_BF16_F32;
    if (direct && !shadow) {
      CK(ix, cudaMemcpyAsync(s.d_rows + local * rb, src, take * rb,
                             on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s.stream));
    } else if (on_device) {
      CK(ix, launch_convert_rows(src, ix->dim, take, s.d_rows, local, ix->layout, s.stream));
      if (shadow) CK(ix, launch_convert_rows(src, ix->dim, take, s.d_rows16, local, ix->layout16, s.stream));
    } else {
      const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / (sizeof(float) * ix->dim));
      if (s.stage_rows < std::min(chunk, take)) {
        cudaFree(s.d_stage);
        s.d_stage = nullptr;
        s.stage_rows = std::min(chunk, std::max<uint64_t>(take, 1024));
        CK(ix, cudaMalloc((void**)&s.d_stage, s.stage_rows * sizeof(float) * ix->dim));
      }
      for (uint64_t c = 0; c < take; c += s.stage_rows) {
        uint64_t m = std::min(s.stage_rows, take - c);
        CK(ix, cudaMemcpyAsync(s.d_stage, src + c * ix->dim, m * sizeof(float) * ix->dim,
                               cudaMemcpyHostToDevice, s.stream));
        CK(ix, launch_convert_rows(s.d_stage, ix->dim, m, s.d_rows, local + c, ix->layout, s.stream));
        if (shadow)
          CK(ix, launch_convert_rows(s.d_stage, ix->dim, m, s.d_rows16, local + c, ix->layout16, s.stream));
        CK(ix, cudaStreamSynchronize(s.stream));
      }
    }
    CK(ix, cudaStreamSynchronize(s.stream));
    s.n_rows = local + take;
    done += take;
  }
  ix->n_rows += n_rows;
  return CQS_B200_OK;
}

int cqs_b200_append_rows_f32(cqs_b200_index* ix, const float* rows, uint64_t n_rows) try {
  return append_impl(ix, rows, n_rows, false);
} API_CATCH
int cqs_b200_append_rows_f32_device(cqs_b200_index* ix, const float* d_rows, uint64_t n_rows) try {
  return append_impl(ix, d_rows, n_rows, true);
} API_CATCH

int cqs_b200_finalize(cqs_b200_index* ix) try {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  std::lock_guard<std::mutex> g(ix->mu);
  if (ix->poisoned.load()) return fail(CQS_B200_ERR_POISONED, "index is poisoned");
  for (auto& s : ix->shards) {
    CK(ix, cudaSetDevice(s.device));
    cudaFree(s.d_stage);
    s.d_stage = nullptr;
    s.stage_rows = 0;
    uint64_t words = (s.n_rows + 31) / 32;
    if (words > s.bitset_words) {
      cudaFree(s.d_bitset);
      s.d_bitset = nullptr;
      CK(ix, cudaMalloc((void**)&s.d_bitset, std::max<uint64_t>(words, 1) * 4));
      s.bitset_words = words;
    }
    if (ix->storage != CQS_B200_STORAGE_F32 && s.n_rows) {
      // exactness bound of the batched tensor-core path needs max |row|
      if (!s.d_maxnorm) CK(ix, cudaMalloc((void**)&s.d_maxnorm, sizeof(float)));
      CK(ix, launch_max_row_norm(s.d_rows, s.n_rows, ix->layout, s.d_maxnorm, s.stream));
      CK(ix, cudaMemcpyAsync(&s.max_row_norm, s.d_maxnorm, sizeof(float), cudaMemcpyDeviceToHost,
                             s.stream));
      s.max_row_delta = 0.f;
      if (ix->storage == CQS_B200_STORAGE_BF16_F32) {
        CK(ix, cudaStreamSynchronize(s.stream));
        CK(ix, launch_max_row_delta(s.d_rows, s.d_rows16, s.n_rows, ix->layout.ld, s.d_maxnorm, s.stream));
        CK(ix, cudaMemcpyAsync(&s.max_row_delta, s.d_maxnorm, sizeof(float), cudaMemcpyDeviceToHost,
                               s.stream));
      }
    }
    CK(ix, cudaStreamSynchronize(s.stream));
  }
  ix->finalized = true;
  return CQS_B200_OK;
} API_CATCH

int cqs_b200_reopen(cqs_b200_index* ix) try {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  std::lock_guard<std::mutex> g(ix->mu);
  ix->finalized = false;
  // everything that is aligned row-by-row with the dense matrix is now stale: the sparse
  // postings, the type/language codes and the score signals must be attached again after
  // the next finalize
  for (auto& s : ix->shards) {
    cudaSetDevice(s.device);
    cudaStreamSynchronize(s.stream);
    free_sparse(s.sparse);
    cudaFree(s.d_ctype); cudaFree(s.d_lang); cudaFree(s.d_note_boost); cudaFree(s.d_importance);
    s.d_ctype = s.d_lang = nullptr;
    s.d_note_boost = s.d_importance = nullptr;
  }
  ix->max_note_boost = ix->max_importance = 1.f;
  return CQS_B200_OK;
} API_CATCH

static bool query_is_finite(const float* q, uint32_t n) {
  for (uint32_t i = 0; i < n; ++i)
    if (!isfinite(q[i])) return false;
  return true;
}

// Host merge of per-device results (in-process multi-GPU): same order rule.
struct HostCand {
  uint32_t key;
  uint64_t row;
  float score;
};
static uint32_t host_ordered(float f) {
  uint32_t b;
  memcpy(&b, &f, 4);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Upload query + (slice of) bitset to a shard and launch the dense scan into the
// shard's dense pool buffers.  Asynchronous on s.stream.
// Make `st` wait for the last launch that used this shard's scratch on another stream.
static int order_after_last(cqs_b200_index* ix, Shard& s, cudaStream_t st) {
  for (auto& c : s.scr)
    if (c.used && c.stream != st) CK(ix, cudaStreamWaitEvent(st, c.ev, 0));
  return 0;
}
static int mark_last(cqs_b200_index* ix, Shard& s, cudaStream_t st) {
  for (auto& c : s.scr) {
    CK(ix, cudaEventRecord(c.ev, st));
    c.stream = st;
    c.used = true;
  }
  return 0;
}
// Dense scans: take the next scratch set; `st` only has to wait for the launch that used
// that set last (kLanes launches back), so launches issued round-robin on kLanes streams overlap.
static int acquire_scratch(cqs_b200_index* ix, Shard& s, cudaStream_t st, Shard::ScanScratch** out) {
  Shard::ScanScratch& c = s.scr[s.next_scr];
  s.next_scr = (s.next_scr + 1) % kLanes;
  if (c.used && c.stream != st) CK(ix, cudaStreamWaitEvent(st, c.ev, 0));
  *out = &c;
  return 0;
}
static int release_scratch(cqs_b200_index* ix, Shard::ScanScratch* c, cudaStream_t st) {
  CK(ix, cudaEventRecord(c->ev, st));
  c->stream = st;
  c->used = true;
  return 0;
}

constexpr size_t kOffScores = 0, kOffRows = sizeof(float) * kMaxK,
                 kOffN = (sizeof(float) + sizeof(uint64_t)) * kMaxK,
                 kOffFlag = kMaxK * (8 + 4 + 4 + 4 + 1) + 16;   // completion words live past both result layouts

// Which rows a single-query scan streams.  STORAGE_BF16_F32 and a plain search (no score fold):
// the bf16 shadow (2 B/elem) with an over-fetch of k' candidates that the kernel tail re-scores on
// the f32 master rows — f32-exact answers at bf16 bandwidth; a pool that cannot be proven complete
// comes back with kUnprovenBit and is re-run on the master (force_master).  Everything else: the
// master rows.
static void fill_scan_rows(const cqs_b200_index* ix, const Shard& s, ScanArgs& a, uint32_t k, bool signals,
                           bool force_master) {
  a.d_rows = s.d_rows;
  a.layout = ix->layout;
  a.k = k;
  const uint32_t kp = shadow_kprime(k);
  if (ix->storage == CQS_B200_STORAGE_BF16_F32 && !signals && !force_master && kp != 0 && s.d_rows16) {
    a.d_rows = s.d_rows16;
    a.layout = ix->layout16;
    a.k = kp;
    a.k_out = k;
    a.d_exact_rows = s.d_rows;
    a.exact_nv = ix->layout.nv;
    a.max_row_delta = s.max_row_delta;
    a.max_row_norm = s.max_row_norm + s.max_row_delta;
  }
}

// to_host: the kernel writes the result into the host-mapped buffer and publishes a
// completion word (the latency path of cqs_b200_search); otherwise into the device pool.
static int launch_dense(cqs_b200_index* ix, Shard& s, const float* query, uint32_t k,
                        const uint32_t* bitset, const ScanSignals* sig = nullptr,
                        bool to_host = false, const PeerCtx* peer = nullptr, bool force_master = false) {
  CK(ix, cudaSetDevice(s.device));
  Shard::ScanScratch* scr = nullptr;
  if (int rc = acquire_scratch(ix, s, s.stream, &scr)) return rc;
  memset(s.h_query, 0, sizeof(float) * ix->layout.ld);
  memcpy(s.h_query, query, sizeof(float) * ix->dim);
  CK(ix, cudaMemcpyAsync(scr->d_query, s.h_query, sizeof(float) * ix->layout.ld,
                         cudaMemcpyHostToDevice, s.stream));
  const uint32_t* d_bits = nullptr;
  if (bitset) {
    uint64_t words = (s.n_rows + 31) / 32;
    CK(ix, cudaMemcpyAsync(s.d_bitset, bitset + s.first_row / 32, words * 4,
                           cudaMemcpyHostToDevice, s.stream));
    d_bits = s.d_bitset;
  }
  ScanArgs a;
  fill_scan_rows(ix, s, a, k, sig != nullptr, force_master);
  a.n_rows = s.n_rows; a.d_query = scr->d_query;
  a.d_bitset = d_bits; a.row_base = ix->row_base + s.first_row;
  a.d_partial = scr->d_partial; a.d_partial_cnt = scr->d_partial_cnt; a.d_done = scr->d_done; a.d_col = scr->d_col;
  a.d_out_scores = s.d_out_scores; a.d_out_rows = s.d_out_rows; a.d_out_n = s.d_out_n;
  if (to_host) {
    a.d_out_scores = (float*)(s.d_hout + kOffScores);
    a.d_out_rows = (uint64_t*)(s.d_hout + kOffRows);
    a.d_out_n = (uint32_t*)(s.d_hout + kOffN);
    a.d_host_flag = (uint32_t*)(s.d_hout + kOffFlag);
    a.seq = ++s.seq;
    if (a.seq == 0) a.seq = s.seq = 1;
  }
  a.d_trace = s.d_trace;
  a.signals = sig;
  a.peer = peer;
  if (ix->timing) CK(ix, cudaEventRecord(s.ev0, s.stream));
  CK(ix, launch_scan_single(a, s.num_sms, s.stream));
  if (ix->timing) CK(ix, cudaEventRecord(s.ev1, s.stream));
  return release_scratch(ix, scr, s.stream);
}

// Poll the completion word the kernel writes into mapped host memory; fall back to the
// stream status every so often so a faulted kernel cannot hang the caller.
static int wait_host_flag(cqs_b200_index* ix, Shard& s, size_t flag_off = kOffFlag) {
  // acquire load: the result words in h_out that the caller reads next must not be
  // speculated ahead of the flag (matters on weakly ordered hosts, e.g. Grace)
  const uint32_t* flag = (const uint32_t*)(s.h_out + flag_off);
  for (uint64_t spin = 0; __atomic_load_n(flag, __ATOMIC_ACQUIRE) != s.seq; ++spin) {
    if ((spin & 0xFFF) == 0xFFF) {
      CK(ix, cudaSetDevice(s.device));
      cudaError_t e = cudaStreamQuery(s.stream);
      if (e == cudaSuccess) {
        if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) != s.seq) {
          ix->poisoned.store(1);
          return fail(CQS_B200_ERR_CUDA, "scan finished without publishing its result");
        }
        break;
      }
      if (e != cudaErrorNotReady) CK(ix, e);
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  return 0;
}

static int check_searchable(cqs_b200_index* ix) {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  if (ix->poisoned.load()) return fail(CQS_B200_ERR_POISONED, "index is poisoned; rebuild it");
  if (!ix->finalized) return fail(CQS_B200_ERR_INVALID, "index is not finalized");
  return 0;
}

// sig_template (nullable): structured filter / signal fold; the per-shard device arrays are
// filled in here.
static int search_impl(cqs_b200_index* ix, const float* query, uint32_t k, const uint32_t* bitset,
                       const ScanSignals* sig_template, uint64_t* out_rows, float* out_scores,
                       uint32_t* out_n) {
  if (out_n) *out_n = 0;
  int rc = check_searchable(ix);
  if (rc) return rc;
  if (!query || !out_rows || !out_scores || !out_n)
    return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (k > kMaxK) return fail(CQS_B200_ERR_INVALID, "k=%u exceeds max_k=%u", k, kMaxK);
  if (k == 0 || ix->n_rows == 0) return CQS_B200_OK;      // src/cagra.rs:445
  if (!query_is_finite(query, ix->dim)) return CQS_B200_OK;  // src/cagra.rs:458-470
  std::lock_guard<std::mutex> g(ix->mu);
  std::vector<Shard*> live;
  for (auto& s : ix->shards)
    if (s.n_rows) live.push_back(&s);
  for (Shard* s : live) {
    ScanSignals sig;
    if (sig_template) {
      sig = *sig_template;
      sig.d_ctype = s->d_ctype;
      sig.d_lang = s->d_lang;
      sig.d_note_boost = sig.pipeline ? s->d_note_boost : nullptr;
      sig.d_importance = (sig.pipeline && sig_template->d_importance) ? s->d_importance : nullptr;
      if (!sig.d_importance) sig.max_importance = 1.f;
    }
    rc = launch_dense(ix, *s, query, k, bitset, sig_template ? &sig : nullptr, /*to_host=*/true);
    if (rc) return rc;
  }
  float kms = 0.f;
  for (Shard* s : live) {
    if (int rcw = wait_host_flag(ix, *s)) return rcw;
    if (*(uint32_t*)(s->h_out + kOffN) & kUnprovenBit) {
      // bf16 shadow scan whose candidate pool could not be proven complete: the f32 master decides
      ++ix->shadow_reruns;
      rc = launch_dense(ix, *s, query, k, bitset, nullptr, /*to_host=*/true, nullptr, /*force_master=*/true);
      if (rc) return rc;
      if (int rcw = wait_host_flag(ix, *s)) return rcw;
    }
    if (ix->timing) {
      CK(ix, cudaSetDevice(s->device));
      CK(ix, cudaEventSynchronize(s->ev1));
      float ms = 0.f;
      CK(ix, cudaEventElapsedTime(&ms, s->ev0, s->ev1));
      kms = std::max(kms, ms);
    }
  }
  if (ix->timing) ix->last_kernel_ms = kms;
  if (live.size() == 1) {
    Shard* s = live[0];
    uint32_t n = *(uint32_t*)(s->h_out + (sizeof(float) + sizeof(uint64_t)) * kMaxK);
    n = std::min(n, k);
    memcpy(out_scores, s->h_out, sizeof(float) * n);
    memcpy(out_rows, s->h_out + sizeof(float) * kMaxK, sizeof(uint64_t) * n);
    *out_n = n;
    return CQS_B200_OK;
  }
  std::vector<HostCand> all;
  for (Shard* s : live) {
    const float* hs = (const float*)s->h_out;
    const uint64_t* hr = (const uint64_t*)(s->h_out + sizeof(float) * kMaxK);
    uint32_t n = std::min(*(uint32_t*)(s->h_out + (sizeof(float) + sizeof(uint64_t)) * kMaxK), k);
    for (uint32_t i = 0; i < n; ++i) all.push_back({host_ordered(hs[i]), hr[i], hs[i]});
  }
  std::sort(all.begin(), all.end(), [](const HostCand& a, const HostCand& b) {
    return a.key > b.key || (a.key == b.key && a.row < b.row);
  });
  uint32_t n = (uint32_t)std::min<size_t>(all.size(), k);
  for (uint32_t i = 0; i < n; ++i) {
    out_rows[i] = all[i].row;
    out_scores[i] = all[i].score;
  }
  *out_n = n;
  return CQS_B200_OK;
}

int cqs_b200_search(cqs_b200_index* ix, const float* query, uint32_t k, const uint32_t* bitset,
                    uint64_t* out_rows, float* out_scores, uint32_t* out_n) try {
  return search_impl(ix, query, k, bitset, nullptr, out_rows, out_scores, out_n);
} API_CATCH

static int upload_rows_u8(cqs_b200_index* ix, uint8_t* Shard::*field, const uint8_t* src, uint64_t n) {
  for (auto& s : ix->shards) {
    CK(ix, cudaSetDevice(s.device));
    cudaFree(s.*field);
    s.*field = nullptr;
    if (!src || s.n_rows == 0) continue;
    CK(ix, cudaMalloc((void**)&(s.*field), s.n_rows));
    CK(ix, cudaMemcpy(s.*field, src + s.first_row, s.n_rows, cudaMemcpyHostToDevice));
  }
  (void)n;
  return CQS_B200_OK;
}
static int upload_rows_f32(cqs_b200_index* ix, float* Shard::*field, const float* src) {
  for (auto& s : ix->shards) {
    CK(ix, cudaSetDevice(s.device));
    cudaFree(s.*field);
    s.*field = nullptr;
    if (!src || s.n_rows == 0) continue;
    CK(ix, cudaMalloc((void**)&(s.*field), s.n_rows * sizeof(float)));
    CK(ix, cudaMemcpy(s.*field, src + s.first_row, s.n_rows * sizeof(float), cudaMemcpyHostToDevice));
  }
  return CQS_B200_OK;
}

int cqs_b200_set_row_meta(cqs_b200_index* ix, const uint8_t* chunk_type, const uint8_t* lang,
                          uint64_t n_rows) try {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  std::lock_guard<std::mutex> g(ix->mu);
  if (!ix->finalized) return fail(CQS_B200_ERR_INVALID, "index is not finalized");
  if (n_rows != ix->n_rows) return fail(CQS_B200_ERR_INVALID, "n_rows != len");
  if ((chunk_type == nullptr) != (lang == nullptr))
    return fail(CQS_B200_ERR_INVALID, "chunk_type and lang must both be given or both be NULL");
  int rc = upload_rows_u8(ix, &Shard::d_ctype, chunk_type, n_rows);
  if (rc) return rc;
  return upload_rows_u8(ix, &Shard::d_lang, lang, n_rows);
} API_CATCH

int cqs_b200_set_row_signals(cqs_b200_index* ix, const float* note_boost, const float* importance,
                             uint64_t n_rows) try {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  std::lock_guard<std::mutex> g(ix->mu);
  if (!ix->finalized) return fail(CQS_B200_ERR_INVALID, "index is not finalized");
  if (n_rows != ix->n_rows) return fail(CQS_B200_ERR_INVALID, "n_rows != len");
  float mb = 1.f, mi = 1.f;
  for (uint64_t r = 0; r < n_rows; ++r) {
    if (note_boost) {
      if (!(note_boost[r] >= 0.f) || !isfinite(note_boost[r]))
        return fail(CQS_B200_ERR_INVALID, "note_boost[%llu] must be finite and >= 0", (unsigned long long)r);
      mb = std::max(mb, note_boost[r]);
    }
    if (importance) {
      if (!(importance[r] >= 0.f) || !isfinite(importance[r]))
        return fail(CQS_B200_ERR_INVALID, "importance[%llu] must be finite and >= 0", (unsigned long long)r);
      mi = std::max(mi, importance[r]);
    }
  }
  ix->max_note_boost = mb;
  ix->max_importance = mi;
  int rc = upload_rows_f32(ix, &Shard::d_note_boost, note_boost);
  if (rc) return rc;
  return upload_rows_f32(ix, &Shard::d_importance, importance);
} API_CATCH

int cqs_b200_search_filtered(cqs_b200_index* ix, const float* query, uint32_t limit,
                             float threshold, const uint64_t* type_mask, const uint64_t* lang_mask,
                             int enable_demotion, uint64_t* out_rows, float* out_scores,
                             uint32_t* out_n) try {
  if (out_n) *out_n = 0;
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  const bool have_meta = !ix->shards.empty() && ix->shards[0].d_ctype != nullptr;
  if ((type_mask || lang_mask) && !have_meta && ix->n_rows)
    return fail(CQS_B200_ERR_INVALID, "type/language filter requested but cqs_b200_set_row_meta was not called");
  ScanSignals sig;
  if (type_mask) memcpy(sig.type_mask, type_mask, 32);
  if (lang_mask) memcpy(sig.lang_mask, lang_mask, 32);
  sig.pipeline = 1;
  sig.threshold = threshold;
  sig.max_note_boost = ix->max_note_boost;
  sig.max_importance = ix->max_importance;
  // non-null marker: search_impl substitutes the per-shard array when demotion is on
  sig.d_importance = enable_demotion ? reinterpret_cast<const float*>(1) : nullptr;
  int rc = search_impl(ix, query, limit, nullptr, &sig, out_rows, out_scores, out_n);
  return rc;
} API_CATCH

int cqs_b200_search_device(cqs_b200_index* ix, const float* d_query, uint32_t k,
                           const uint32_t* d_bitset, float* d_out_scores, uint64_t* d_out_rows,
                           uint32_t* d_out_n, void* stream) try {
  int rc = check_searchable(ix);
  if (rc) return rc;
  if (!d_query || !d_out_scores || !d_out_rows || !d_out_n)
    return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (k == 0 || k > kMaxK) return fail(CQS_B200_ERR_INVALID, "k=%u out of range 1..%u", k, kMaxK);
  if (ix->shards.size() != 1)
    return fail(CQS_B200_ERR_UNSUPPORTED, "search_device needs a single-device index");
  std::lock_guard<std::mutex> g(ix->mu);
  Shard& s = ix->shards[0];
  if (s.n_rows == 0) return fail(CQS_B200_ERR_INVALID, "empty index");
  CK(ix, cudaSetDevice(s.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : s.stream;
  Shard::ScanScratch* scr = nullptr;
  if (int rc2 = acquire_scratch(ix, s, st, &scr)) return rc2;
  const float* qp = d_query;
  if (ix->layout.ld != ix->dim) {
    // the query must be zero padded to the row stride: stage it through d_query
    CK(ix, cudaMemsetAsync(scr->d_query, 0, sizeof(float) * ix->layout.ld, st));
    CK(ix, cudaMemcpyAsync(scr->d_query, d_query, sizeof(float) * ix->dim, cudaMemcpyDeviceToDevice, st));
    qp = scr->d_query;
  }
  ScanArgs a;
  fill_scan_rows(ix, s, a, k, false, false);
  a.n_rows = s.n_rows; a.d_query = qp;
  a.d_bitset = d_bitset; a.row_base = ix->row_base + s.first_row;
  a.d_partial = scr->d_partial; a.d_partial_cnt = scr->d_partial_cnt; a.d_done = scr->d_done; a.d_col = scr->d_col;
  a.d_out_scores = d_out_scores; a.d_out_rows = d_out_rows; a.d_out_n = d_out_n;
  a.d_trace = s.d_trace;
  CK(ix, launch_scan_single(a, s.num_sms, st));
  return release_scratch(ix, scr, st);
} API_CATCH

int cqs_b200_merge_topk_device(int device, const float* d_scores, const uint64_t* d_rows,
                               uint32_t n_lists, uint32_t n_queries, uint32_t k,
                               float* d_out_scores, uint64_t* d_out_rows, uint32_t* d_out_n,
                               void* stream) try {
  if (!d_scores || !d_rows || !d_out_scores || !d_out_rows || !d_out_n)
    return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (k == 0 || (uint64_t)n_lists * k > 8192)
    return fail(CQS_B200_ERR_INVALID, "n_lists*k must be in 1..8192");
  cqs_b200_index* none = nullptr;
  CK(none, cudaSetDevice(device));
  MergeArgs a{d_scores, d_rows, n_lists, n_queries, k, d_out_scores, d_out_rows, d_out_n};
  CK(none, launch_merge_topk(a, (cudaStream_t)stream));
  return CQS_B200_OK;
} API_CATCH

// ---- row-sharded corpus over NVLink peer memory (peer.cu / peer.cuh) ----------
// A timed-out exchange leaves the ranks' sequence numbers out of step: every later sharded search
// on this group fails.  VectorIndex::is_poisoned() is the reference's only recovery hook
// (src/index.rs:203-205 -> the daemon drops and rebuilds the index, cli/batch/view.rs:737-767), so
// the failure is surfaced there: the index that ran the search is poisoned together with the group.
static void peer_failed(cqs_b200_index* ix, cqs_b200_peer* peer) {
  peer->failed.store(1);
  ix->poisoned.store(1);
}

static int check_peer(cqs_b200_index* ix, cqs_b200_peer* peer) {
  if (!peer) return fail(CQS_B200_ERR_INVALID, "peer is NULL");
  if (ix->shards.size() != 1)
    return fail(CQS_B200_ERR_UNSUPPORTED, "a sharded search needs a single-device index per rank");
  if (peer->device != ix->shards[0].device)
    return fail(CQS_B200_ERR_INVALID, "peer group and index live on different devices");
  if (!peer->connected) return fail(CQS_B200_ERR_INVALID, "peer group is not connected");
  if (peer->failed.load()) {
    ix->poisoned.store(1);
    return fail(CQS_B200_ERR_POISONED, "peer group failed earlier; rebuild it");
  }
  if (ix->shards[0].n_rows == 0) return fail(CQS_B200_ERR_INVALID, "empty shard");
  return 0;
}

int cqs_b200_search_sharded_device(cqs_b200_index* ix, cqs_b200_peer* peer, const float* d_query,
                                   uint32_t k, const uint32_t* d_bitset, float* d_out_scores,
                                   uint64_t* d_out_rows, uint32_t* d_out_n, void* stream) try {
  int rc = check_searchable(ix);
  if (rc) return rc;
  if (!d_query || !d_out_scores || !d_out_rows || !d_out_n)
    return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (k == 0 || k > kMaxK) return fail(CQS_B200_ERR_INVALID, "k=%u out of range 1..%u", k, kMaxK);
  std::lock_guard<std::mutex> g(ix->mu);
  if ((rc = check_peer(ix, peer))) return rc;
  std::lock_guard<std::mutex> gp(peer->mu);
  Shard& s = ix->shards[0];
  CK(ix, cudaSetDevice(s.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : s.stream;
  Shard::ScanScratch* scr = nullptr;
  if (int rc2 = acquire_scratch(ix, s, st, &scr)) return rc2;
  const float* qp = d_query;
  if (ix->layout.ld != ix->dim) {
    CK(ix, cudaMemsetAsync(scr->d_query, 0, sizeof(float) * ix->layout.ld, st));
    CK(ix, cudaMemcpyAsync(scr->d_query, d_query, sizeof(float) * ix->dim, cudaMemcpyDeviceToDevice, st));
    qp = scr->d_query;
  }
  PeerCtx pc;
  CK(ix, peer_begin(peer, st, &pc, /*exclusive=*/false));
  ScanArgs a;
  fill_scan_rows(ix, s, a, k, false, false);
  a.n_rows = s.n_rows; a.d_query = qp;
  a.d_bitset = d_bitset; a.row_base = ix->row_base + s.first_row;
  a.d_partial = scr->d_partial; a.d_partial_cnt = scr->d_partial_cnt; a.d_done = scr->d_done; a.d_col = scr->d_col;
  a.d_out_scores = d_out_scores; a.d_out_rows = d_out_rows; a.d_out_n = d_out_n;
  a.peer = &pc;
  CK(ix, launch_scan_single(a, s.num_sms, st));
  CK(ix, peer_mark(peer, st, /*exclusive=*/false));
  return release_scratch(ix, scr, st);
} API_CATCH

// Device buffers of the batch entry points, allocated on first use (each one on its own: the
// exact path and the tensor-core path share the query and result buffers).
static int ensure_batch_buffers(cqs_b200_index* ix, Shard& s, bool tensor_path) {
  const uint32_t ld = ix->layout.ld;
  if (!s.d_bq) CK(ix, cudaMalloc((void**)&s.d_bq, sizeof(float) * (size_t)kBatchMaxQ * ld));
  if (!s.d_bout_scores) CK(ix, cudaMalloc((void**)&s.d_bout_scores, sizeof(float) * (size_t)kBatchMaxQ * kMaxK));
  if (!s.d_bout_rows) CK(ix, cudaMalloc((void**)&s.d_bout_rows, sizeof(uint64_t) * (size_t)kBatchMaxQ * kMaxK));
  if (!s.d_bout_n) CK(ix, cudaMalloc((void**)&s.d_bout_n, sizeof(uint32_t) * kBatchMaxQ));
  if (tensor_path) {
    if (!s.d_bscratch) CK(ix, cudaMalloc(&s.d_bscratch, batch_scratch_bytes(kBatchMaxQ, ld)));
    if (!s.d_bflags) CK(ix, cudaMalloc((void**)&s.d_bflags, sizeof(uint32_t) * kBatchMaxQ));
  }
  return 0;
}

// nq single-query scans, one launch each, issued round-robin on kLanes lanes (`st` and the
// shard's own lane streams) so the tail of launch i — list merge and, with `peer`, the
// cross-shard exchange — overlaps the streaming phase of the following launches.  Everything stays on
// the device; `st` is joined with the other lanes before returning.  d_queries: f32
// [nq][q_stride]; `skip` (nullable): queries that are not launched (non-finite).  Caller
// holds ix->mu (and peer->mu).
static int launch_scan_lanes(cqs_b200_index* ix, Shard& s, cqs_b200_peer* peer, const float* d_queries,
                             uint32_t q_stride, uint32_t nq, uint32_t k, const uint32_t* d_bitset,
                             float* d_out_scores, uint64_t* d_out_rows, uint32_t* d_out_n,
                             const uint8_t* skip, cudaStream_t st, bool force_master = false) {
  const bool staged = q_stride < ix->layout.ld;  // query must be zero padded to the row stride
  cudaStream_t lanes[kLanes];
  lanes[0] = st;
  for (uint32_t l = 1; l < kLanes; ++l) lanes[l] = s.lane_stream[l - 1];
  bool forked[kLanes] = {};
  // fork point = everything already queued on `st` (the queries' upload); recorded BEFORE the
  // first launch so that the other lanes do not wait for lane 0's first kernel
  if (nq > 1) CK(ix, cudaEventRecord(s.ev_fork, st));
  uint32_t li = 0;
  for (uint32_t i = 0; i < nq; ++i) {
    if (skip && skip[i]) continue;
    const uint32_t l = li % kLanes;
    cudaStream_t ln = lanes[l];
    if (l && !forked[l]) {
      CK(ix, cudaStreamWaitEvent(ln, s.ev_fork, 0));
      forked[l] = true;
    }
    ++li;
    Shard::ScanScratch* scr = nullptr;
    if (int rc = acquire_scratch(ix, s, ln, &scr)) return rc;
    const float* qp = d_queries + (size_t)i * q_stride;
    if (staged) {
      CK(ix, cudaMemsetAsync(scr->d_query, 0, sizeof(float) * ix->layout.ld, ln));
      CK(ix, cudaMemcpyAsync(scr->d_query, qp, sizeof(float) * ix->dim, cudaMemcpyDeviceToDevice, ln));
      qp = scr->d_query;
    }
    PeerCtx pc;
    if (peer) CK(ix, peer_begin(peer, ln, &pc, /*exclusive=*/false));
    ScanArgs a;
    fill_scan_rows(ix, s, a, k, false, force_master);
    a.n_rows = s.n_rows; a.d_query = qp;
    a.d_bitset = d_bitset; a.row_base = ix->row_base + s.first_row;
    a.d_partial = scr->d_partial; a.d_partial_cnt = scr->d_partial_cnt; a.d_done = scr->d_done; a.d_col = scr->d_col;
    a.d_out_scores = d_out_scores + (size_t)i * k;
    a.d_out_rows = d_out_rows + (size_t)i * k;
    a.d_out_n = d_out_n + i;
    a.peer = peer ? &pc : nullptr;
    CK(ix, launch_scan_single(a, s.num_sms, ln));
    if (peer) CK(ix, peer_mark(peer, ln, /*exclusive=*/false));
    if (int rc = release_scratch(ix, scr, ln)) return rc;
  }
  for (uint32_t l = 1; l < kLanes; ++l)
    if (forked[l]) {
      CK(ix, cudaEventRecord(s.ev_join[l - 1], lanes[l]));
      CK(ix, cudaStreamWaitEvent(st, s.ev_join[l - 1], 0));
    }
  return 0;
}

int cqs_b200_search_many_device(cqs_b200_index* ix, cqs_b200_peer* peer, const float* d_queries,
                                uint32_t nq, uint32_t k, const uint32_t* d_bitset,
                                float* d_out_scores, uint64_t* d_out_rows, uint32_t* d_out_n,
                                void* stream) try {
  int rc = check_searchable(ix);
  if (rc) return rc;
  if (nq == 0) return CQS_B200_OK;
  if (!d_queries || !d_out_scores || !d_out_rows || !d_out_n)
    return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (k == 0 || k > kMaxK) return fail(CQS_B200_ERR_INVALID, "k=%u out of range 1..%u", k, kMaxK);
  if (ix->shards.size() != 1)
    return fail(CQS_B200_ERR_UNSUPPORTED, "search_many_device needs a single-device index");
  std::lock_guard<std::mutex> g(ix->mu);
  if (peer && (rc = check_peer(ix, peer))) return rc;
  std::unique_lock<std::mutex> gp;
  if (peer) gp = std::unique_lock<std::mutex>(peer->mu);
  Shard& s = ix->shards[0];
  if (s.n_rows == 0) return fail(CQS_B200_ERR_INVALID, "empty index");
  CK(ix, cudaSetDevice(s.device));
  cudaStream_t st = stream ? (cudaStream_t)stream : s.stream;
  return launch_scan_lanes(ix, s, peer, d_queries, ix->dim, nq, k, d_bitset, d_out_scores, d_out_rows,
                           d_out_n, nullptr, st);
} API_CATCH

// Exact (CUDA-core) batch: host queries in, host results out, one H2D, nq pipelined launches,
// one D2H.  Used by cqs_b200_search_batch on f32 storage / small batches and by its sharded
// twin; results are those of nq cqs_b200_search calls.  <= kBatchMaxQ queries per call.
static int search_batch_exact(cqs_b200_index* ix, cqs_b200_peer* peer, const float* queries, uint32_t nq,
                              uint32_t k, const uint32_t* bitset, uint64_t* out_rows, float* out_scores,
                              uint32_t* out_n) {
  std::lock_guard<std::mutex> g(ix->mu);
  std::unique_lock<std::mutex> gp;
  if (peer) gp = std::unique_lock<std::mutex>(peer->mu);
  Shard& s = ix->shards[0];
  CK(ix, cudaSetDevice(s.device));
  const uint32_t ld = ix->layout.ld;
  if (int rcb = ensure_batch_buffers(ix, s, /*tensor_path=*/false)) return rcb;
  std::vector<float> padded((size_t)nq * ld, 0.f);
  std::vector<uint8_t> bad(nq, 0);
  for (uint32_t i = 0; i < nq; ++i) {
    const float* q = queries + (size_t)i * ix->dim;
    if (query_is_finite(q, ix->dim)) memcpy(&padded[(size_t)i * ld], q, sizeof(float) * ix->dim);
    else bad[i] = 1;  // empty result (src/cagra.rs:458-470); not launched — on every rank alike
  }
  if (int rc = order_after_last(ix, s, s.stream)) return rc;
  CK(ix, cudaMemcpyAsync(s.d_bq, padded.data(), sizeof(float) * padded.size(), cudaMemcpyHostToDevice, s.stream));
  const uint32_t* d_bits = nullptr;
  if (bitset) {
    CK(ix, cudaMemcpyAsync(s.d_bitset, bitset, ((s.n_rows + 31) / 32) * 4, cudaMemcpyHostToDevice, s.stream));
    d_bits = s.d_bitset;
  }
  CK(ix, cudaMemsetAsync(s.d_bout_n, 0, sizeof(uint32_t) * nq, s.stream));
  if (int rc = launch_scan_lanes(ix, s, peer, s.d_bq, ld, nq, k, d_bits, s.d_bout_scores, s.d_bout_rows,
                                 s.d_bout_n, bad.data(), s.stream))
    return rc;
  std::vector<uint32_t> ns(nq);
  uint32_t status = 0;
  CK(ix, cudaMemcpyAsync(out_scores, s.d_bout_scores, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, s.stream));
  CK(ix, cudaMemcpyAsync(out_rows, s.d_bout_rows, sizeof(uint64_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, s.stream));
  CK(ix, cudaMemcpyAsync(ns.data(), s.d_bout_n, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, s.stream));
  if (peer) CK(ix, cudaMemcpyAsync(&status, peer->d_status, sizeof status, cudaMemcpyDeviceToHost, s.stream));
  CK(ix, cudaStreamSynchronize(s.stream));
  if (status) {
    peer_failed(ix, peer);
    return fail(CQS_B200_ERR_CUDA, "peer exchange timed out (a rank did not take part in this batch)");
  }
  // STORAGE_BF16_F32: shadow scans whose pool could not be proven complete are re-run on the f32
  // master rows (the flag travels with the exchange, so every rank re-runs the same queries)
  std::vector<uint8_t> skip2(nq, 1);
  uint32_t n_unproven = 0;
  for (uint32_t i = 0; i < nq; ++i)
    if (!bad[i] && (ns[i] & kUnprovenBit)) {
      skip2[i] = 0;
      ++n_unproven;
    }
  if (n_unproven) {
    ix->shadow_reruns += n_unproven;
    if (int rc = launch_scan_lanes(ix, s, peer, s.d_bq, ld, nq, k, d_bits, s.d_bout_scores, s.d_bout_rows,
                                   s.d_bout_n, skip2.data(), s.stream, /*force_master=*/true))
      return rc;
    CK(ix, cudaMemcpyAsync(out_scores, s.d_bout_scores, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, s.stream));
    CK(ix, cudaMemcpyAsync(out_rows, s.d_bout_rows, sizeof(uint64_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, s.stream));
    CK(ix, cudaMemcpyAsync(ns.data(), s.d_bout_n, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, s.stream));
    if (peer) CK(ix, cudaMemcpyAsync(&status, peer->d_status, sizeof status, cudaMemcpyDeviceToHost, s.stream));
    CK(ix, cudaStreamSynchronize(s.stream));
    if (status) {
      peer_failed(ix, peer);
      return fail(CQS_B200_ERR_CUDA, "peer exchange timed out (a rank did not take part in this batch)");
    }
  }
  for (uint32_t i = 0; i < nq; ++i) out_n[i] = bad[i] ? 0 : std::min(ns[i] & ~kUnprovenBit, k);
  return CQS_B200_OK;
}

int cqs_b200_search_sharded(cqs_b200_index* ix, cqs_b200_peer* peer, const float* query, uint32_t k,
                            const uint32_t* bitset, uint64_t* out_rows, float* out_scores,
                            uint32_t* out_n) try {
  if (out_n) *out_n = 0;
  int rc = check_searchable(ix);
  if (rc) return rc;
  if (!query || !out_rows || !out_scores || !out_n) return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (k == 0 || k > kMaxK) return fail(CQS_B200_ERR_INVALID, "k=%u out of range 1..%u", k, kMaxK);
  // a non-finite query gives an empty result on EVERY rank (src/cagra.rs:458-470); the ranks
  // all see the same query, so skipping the exchange keeps their sequence numbers aligned
  if (!query_is_finite(query, ix->dim)) return CQS_B200_OK;
  std::lock_guard<std::mutex> g(ix->mu);
  if ((rc = check_peer(ix, peer))) return rc;
  std::lock_guard<std::mutex> gp(peer->mu);
  Shard& s = ix->shards[0];
  CK(ix, cudaSetDevice(s.device));
  PeerCtx pc;
  CK(ix, peer_begin(peer, s.stream, &pc, /*exclusive=*/false));
  if ((rc = launch_dense(ix, s, query, k, bitset, nullptr, /*to_host=*/true, &pc))) return rc;
  CK(ix, peer_mark(peer, s.stream, /*exclusive=*/false));
  if ((rc = wait_host_flag(ix, s))) return rc;
  if (*(uint32_t*)(s.h_out + kOffN) & kUnprovenBit) {
    // some shard's bf16 shadow scan could not prove its list; the flag reached every rank with the
    // exchange, so every rank repeats this search on its f32 master rows (one more exchange each)
    ++ix->shadow_reruns;
    CK(ix, peer_begin(peer, s.stream, &pc, /*exclusive=*/false));
    if ((rc = launch_dense(ix, s, query, k, bitset, nullptr, /*to_host=*/true, &pc, /*force_master=*/true))) return rc;
    CK(ix, peer_mark(peer, s.stream, /*exclusive=*/false));
    if ((rc = wait_host_flag(ix, s))) return rc;
  }
  uint32_t n = std::min(*(uint32_t*)(s.h_out + kOffN), k);
  if (n == 0) {
    // empty can also mean "a peer never answered": the kernel raised the sticky status word
    uint32_t st = 0;
    CK(ix, cudaMemcpy(&st, peer->d_status, sizeof st, cudaMemcpyDeviceToHost));
    if (st) {
      peer_failed(ix, peer);
      return fail(CQS_B200_ERR_CUDA, "peer exchange timed out (a rank did not take part in this search)");
    }
  }
  memcpy(out_scores, s.h_out + kOffScores, sizeof(float) * n);
  memcpy(out_rows, s.h_out + kOffRows, sizeof(uint64_t) * n);
  *out_n = n;
  return CQS_B200_OK;
} API_CATCH

// Tensor-core path for one chunk of <= kBatchMaxQ queries on a single-device bf16 index.
static int search_batch_tc(cqs_b200_index* ix, const float* queries, uint32_t nq, uint32_t k,
                           const uint32_t* bitset, uint64_t* out_rows, float* out_scores,
                           uint32_t* out_n, std::vector<uint32_t>* rerun) {
  std::lock_guard<std::mutex> g(ix->mu);
  Shard& s = ix->shards[0];
  CK(ix, cudaSetDevice(s.device));
  const uint32_t ld = ix->layout.ld;
  if (int rcb = ensure_batch_buffers(ix, s, /*tensor_path=*/true)) return rcb;
  // pad queries to the row stride; a non-finite query yields an empty result
  // (src/cagra.rs:458-470): it is scanned as a zero vector and blanked afterwards
  std::vector<float> padded((size_t)nq * ld, 0.f);
  std::vector<uint8_t> bad(nq, 0);
  for (uint32_t i = 0; i < nq; ++i) {
    const float* q = queries + (size_t)i * ix->dim;
    if (query_is_finite(q, ix->dim)) memcpy(&padded[(size_t)i * ld], q, sizeof(float) * ix->dim);
    else bad[i] = 1;
  }
  CK(ix, cudaMemcpyAsync(s.d_bq, padded.data(), sizeof(float) * padded.size(), cudaMemcpyHostToDevice, s.stream));
  const uint32_t* d_bits = nullptr;
  if (bitset) {
    CK(ix, cudaMemcpyAsync(s.d_bitset, bitset, ((s.n_rows + 31) / 32) * 4, cudaMemcpyHostToDevice, s.stream));
    d_bits = s.d_bitset;
  }
  BatchArgs a;
  const bool shadow = ix->storage == CQS_B200_STORAGE_BF16_F32;
  a.d_rows = shadow ? s.d_rows16 : s.d_rows;
  a.layout = shadow ? ix->layout16 : ix->layout;
  a.d_exact_rows = s.d_rows;
  a.exact_layout = ix->layout;
  // exactness bound inputs: the measured row rounding distance (0 when the bf16 rows are the
  // corpus) and a bound on the norm of the scanned bf16 rows (final_select_kernel)
  a.max_row_delta = shadow ? s.max_row_delta : 0.f;
  a.n_rows = s.n_rows; a.d_queries = s.d_bq; a.nq = nq;
  a.k = k; a.d_bitset = d_bits; a.row_base = ix->row_base + s.first_row;
  a.max_row_norm = s.max_row_norm + a.max_row_delta; a.d_scratch = s.d_bscratch;
  a.d_out_scores = s.d_bout_scores; a.d_out_rows = s.d_bout_rows; a.d_out_n = s.d_bout_n;
  a.d_flags = s.d_bflags;
  if (int rc2 = order_after_last(ix, s, s.stream)) return rc2;
  if (ix->timing) CK(ix, cudaEventRecord(s.ev_b0, s.stream));
  CK(ix, launch_scan_batch(a, s.num_sms, s.stream));
  if (int rc2 = mark_last(ix, s, s.stream)) return rc2;
  std::vector<uint32_t> flags(nq), ns(nq);
  CK(ix, cudaMemcpyAsync(out_scores, s.d_bout_scores, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, s.stream));
  CK(ix, cudaMemcpyAsync(out_rows, s.d_bout_rows, sizeof(uint64_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, s.stream));
  CK(ix, cudaMemcpyAsync(ns.data(), s.d_bout_n, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, s.stream));
  CK(ix, cudaMemcpyAsync(flags.data(), s.d_bflags, sizeof(uint32_t) * nq, cudaMemcpyDeviceToHost, s.stream));
  CK(ix, cudaStreamSynchronize(s.stream));
  for (uint32_t i = 0; i < nq; ++i) {
    out_n[i] = bad[i] ? 0 : std::min(ns[i], k);
    if (flags[i] && !bad[i]) rerun->push_back(i);
  }
  return CQS_B200_OK;
}

// Closes the bracket search_batch_tc opened (ev_b0) once the chunk's last kernel is queued.
static int finish_batch_timing(cqs_b200_index* ix) {
  if (!ix->timing) return 0;
  std::lock_guard<std::mutex> g(ix->mu);
  Shard& s = ix->shards[0];
  CK(ix, cudaSetDevice(s.device));
  CK(ix, cudaEventRecord(s.ev_b1, s.stream));
  CK(ix, cudaEventSynchronize(s.ev_b1));
  float ms = 0.f;
  CK(ix, cudaEventElapsedTime(&ms, s.ev_b0, s.ev_b1));
  ix->last_batch_ms += ms;
  ix->last_kernel_ms = ix->last_batch_ms;
  return 0;
}

int cqs_b200_search_batch(cqs_b200_index* ix, const float* queries, uint32_t nq, uint32_t k,
                          const uint32_t* bitset, uint64_t* out_rows, float* out_scores,
                          uint32_t* out_n) try {
  int rc = check_searchable(ix);
  if (rc) return rc;
  if (nq && (!queries || !out_rows || !out_scores || !out_n))
    return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (k > kMaxK) return fail(CQS_B200_ERR_INVALID, "k=%u exceeds max_k=%u", k, kMaxK);
  for (uint32_t i = 0; i < nq; ++i) out_n[i] = 0;
  if (k == 0 || ix->n_rows == 0 || nq == 0) return CQS_B200_OK;
  const bool tensor_path = ix->storage != CQS_B200_STORAGE_F32 && ix->shards.size() == 1 &&
                           nq >= 8 && ix->n_rows < (1ull << 31);
  if (!tensor_path) {
    if (ix->shards.size() == 1 && nq > 1) {
      // exact scans, pipelined: one H2D, nq launches on kLanes (4) launch lanes, one D2H
      for (uint32_t q0 = 0; q0 < nq; q0 += kBatchMaxQ) {
        const uint32_t m = std::min(kBatchMaxQ, nq - q0);
        rc = search_batch_exact(ix, nullptr, queries + (size_t)q0 * ix->dim, m, k, bitset,
                                out_rows + (size_t)q0 * k, out_scores + (size_t)q0 * k, out_n + q0);
        if (rc) return rc;
      }
      return CQS_B200_OK;
    }
    for (uint32_t i = 0; i < nq; ++i) {
      rc = cqs_b200_search(ix, queries + (size_t)i * ix->dim, k, bitset, out_rows + (size_t)i * k,
                           out_scores + (size_t)i * k, out_n + i);
      if (rc) return rc;
    }
    return CQS_B200_OK;
  }
  ix->last_batch_ms = 0.f;
  ix->last_batch_reruns = 0;
  for (uint32_t q0 = 0; q0 < nq; q0 += kBatchMaxQ) {
    const uint32_t m = std::min(kBatchMaxQ, nq - q0);
    std::vector<uint32_t> rerun;
    rc = search_batch_tc(ix, queries + (size_t)q0 * ix->dim, m, k, bitset, out_rows + (size_t)q0 * k,
                         out_scores + (size_t)q0 * k, out_n + q0, &rerun);
    if (rc) return rc;
    // queries whose candidate pool could not be proven complete: exact single-query scan
    for (uint32_t i : rerun) {
      rc = cqs_b200_search(ix, queries + (size_t)(q0 + i) * ix->dim, k, bitset,
                           out_rows + (size_t)(q0 + i) * k, out_scores + (size_t)(q0 + i) * k,
                           out_n + q0 + i);
      if (rc) return rc;
    }
    ix->last_batch_reruns += (uint32_t)rerun.size();
    ix->batch_reruns_total += (uint32_t)rerun.size();
    if ((rc = finish_batch_timing(ix))) return rc;
  }
  return CQS_B200_OK;
} API_CATCH

int cqs_b200_search_batch_sharded(cqs_b200_index* ix, cqs_b200_peer* peer, const float* queries,
                                  uint32_t nq, uint32_t k, const uint32_t* bitset,
                                  uint64_t* out_rows, float* out_scores, uint32_t* out_n) try {
  int rc = check_searchable(ix);
  if (rc) return rc;
  if (nq && (!queries || !out_rows || !out_scores || !out_n))
    return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (k == 0 || k > kMaxK) return fail(CQS_B200_ERR_INVALID, "k=%u out of range 1..%u", k, kMaxK);
  for (uint32_t i = 0; i < nq; ++i) out_n[i] = 0;
  if (nq == 0) return CQS_B200_OK;
  {
    std::lock_guard<std::mutex> g(ix->mu);
    if ((rc = check_peer(ix, peer))) return rc;
  }
  const bool tensor_path = ix->storage != CQS_B200_STORAGE_F32 && nq >= 8 && ix->n_rows < (1ull << 31);
  if (!tensor_path) {
    // every rank takes this branch together (same nq, same storage): one fused
    // scan + exchange per query, pipelined on kLanes (4) launch lanes
    for (uint32_t q0 = 0; q0 < nq; q0 += kBatchMaxQ) {
      const uint32_t m = std::min(kBatchMaxQ, nq - q0);
      rc = search_batch_exact(ix, peer, queries + (size_t)q0 * ix->dim, m, k, bitset,
                              out_rows + (size_t)q0 * k, out_scores + (size_t)q0 * k, out_n + q0);
      if (rc) return rc;
    }
    return CQS_B200_OK;
  }
  ix->last_batch_ms = 0.f;
  ix->last_batch_reruns = 0;
  // exchange granularity: as many queries as fit the mailbox (identical on every rank)
  const uint32_t q_per_x = std::min<uint32_t>(std::min(kBatchMaxQ, kPeerMaxQ), peer->cap / k);
  for (uint32_t q0 = 0; q0 < nq; q0 += q_per_x) {
    const uint32_t m = std::min(q_per_x, nq - q0);
    const float* qs = queries + (size_t)q0 * ix->dim;
    uint64_t* orow = out_rows + (size_t)q0 * k;
    float* osc = out_scores + (size_t)q0 * k;
    std::vector<uint32_t> rerun;
    // 1. local shard: tensor-core candidate scan + exact rescoring; lists stay in d_bout_*
    if ((rc = search_batch_tc(ix, qs, m, k, bitset, orow, osc, out_n + q0, &rerun))) return rc;
    std::vector<uint8_t> bad(m, 0);
    for (uint32_t i = 0; i < m; ++i) bad[i] = !query_is_finite(qs + (size_t)i * ix->dim, ix->dim);
    // 2. queries whose candidate pool could not be proven complete: exact scan, patched into the lists
    for (uint32_t i : rerun) {
      rc = cqs_b200_search(ix, qs + (size_t)i * ix->dim, k, bitset, orow + (size_t)i * k,
                           osc + (size_t)i * k, out_n + q0 + i);
      if (rc) return rc;
    }
    ix->last_batch_reruns += (uint32_t)rerun.size();
    ix->batch_reruns_total += (uint32_t)rerun.size();
    std::lock_guard<std::mutex> g(ix->mu);
    std::lock_guard<std::mutex> gp(peer->mu);
    Shard& s = ix->shards[0];
    CK(ix, cudaSetDevice(s.device));
    for (uint32_t i : rerun) {
      CK(ix, cudaMemcpyAsync(s.d_bout_scores + (size_t)i * k, osc + (size_t)i * k, sizeof(float) * out_n[q0 + i],
                             cudaMemcpyHostToDevice, s.stream));
      CK(ix, cudaMemcpyAsync(s.d_bout_rows + (size_t)i * k, orow + (size_t)i * k, sizeof(uint64_t) * out_n[q0 + i],
                             cudaMemcpyHostToDevice, s.stream));
      CK(ix, cudaMemcpyAsync(s.d_bout_n + i, out_n + q0 + i, sizeof(uint32_t), cudaMemcpyHostToDevice, s.stream));
    }
    // 3. push the lists to every peer, wait for theirs, merge (one kernel, no collective call)
    PeerCtx pc;
    CK(ix, peer_begin(peer, s.stream, &pc, /*exclusive=*/true));
    PeerGatherArgs ga{s.d_bout_scores, s.d_bout_rows, s.d_bout_n, m, k,
                      s.d_bm_scores, s.d_bm_rows, s.d_bm_n, peer->d_ticket};
    CK(ix, launch_peer_gather_merge(pc, ga, s.num_sms, s.stream));
    CK(ix, peer_mark(peer, s.stream, /*exclusive=*/true));
    if (ix->timing) CK(ix, cudaEventRecord(s.ev_b1, s.stream));
    std::vector<uint32_t> ns(m);
    uint32_t status = 0;
    CK(ix, cudaMemcpyAsync(osc, s.d_bm_scores, sizeof(float) * (size_t)m * k, cudaMemcpyDeviceToHost, s.stream));
    CK(ix, cudaMemcpyAsync(orow, s.d_bm_rows, sizeof(uint64_t) * (size_t)m * k, cudaMemcpyDeviceToHost, s.stream));
    CK(ix, cudaMemcpyAsync(ns.data(), s.d_bm_n, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, s.stream));
    CK(ix, cudaMemcpyAsync(&status, peer->d_status, sizeof status, cudaMemcpyDeviceToHost, s.stream));
    CK(ix, cudaStreamSynchronize(s.stream));
    if (ix->timing) {
      float ms = 0.f;
      CK(ix, cudaEventElapsedTime(&ms, s.ev_b0, s.ev_b1));
      ix->last_batch_ms += ms;
      ix->last_kernel_ms = ix->last_batch_ms;
    }
    if (status) {
      peer_failed(ix, peer);
      for (uint32_t i = 0; i < m; ++i) out_n[q0 + i] = 0;
      return fail(CQS_B200_ERR_CUDA, "peer exchange timed out (a rank did not take part in this batch)");
    }
    for (uint32_t i = 0; i < m; ++i) out_n[q0 + i] = bad[i] ? 0 : std::min(ns[i], k);
  }
  return CQS_B200_OK;
} API_CATCH

// ---- sparse ------------------------------------------------------------------

// Static block index over the longest posting lists (SparseDev): lists longer than 256 postings,
// longest first, as many as fit in a budget of 3 bytes per posting (25 % of the postings' own
// footprint).  The per-query bounds pass then only has to look at the short lists.
static int sparse_build_block_index(cqs_b200_index* ix, Shard& s) {
  SparseDev& sp = s.sparse;
  const uint32_t vocab = sp.vocab;
  const uint32_t stride = (uint32_t)((s.n_rows + kSparseDocsPerBlock - 1) / kSparseDocsPerBlock) + 1;
  std::vector<uint64_t> tptr((size_t)vocab + 1);
  CK(ix, cudaMemcpy(tptr.data(), sp.d_tptr, sizeof(uint64_t) * tptr.size(), cudaMemcpyDeviceToHost));
  std::vector<std::pair<uint64_t, uint32_t>> lens;  // (length, token)
  for (uint32_t t = 0; t < vocab; ++t)
    if (tptr[t + 1] - tptr[t] > 256) lens.push_back({tptr[t + 1] - tptr[t], t});
  std::sort(lens.begin(), lens.end(), [](const auto& a, const auto& b) { return a.first > b.first || (a.first == b.first && a.second < b.second); });
  const uint64_t budget = std::max<uint64_t>(16ull << 20, 3 * sp.nnz);
  const uint64_t max_slots = budget / ((uint64_t)stride * sizeof(uint32_t));
  const uint32_t n_slots = (uint32_t)std::min<uint64_t>(lens.size(), max_slots);
  if (n_slots == 0) return CQS_B200_OK;
  std::vector<int32_t> slot_of(vocab, -1);
  std::vector<uint32_t> toks(n_slots);
  for (uint32_t i = 0; i < n_slots; ++i) {
    slot_of[lens[i].second] = (int32_t)i;
    toks[i] = lens[i].second;
  }
  uint32_t* d_toks = nullptr;
  cudaError_t e = cudaMalloc((void**)&sp.d_slot_of, sizeof(int32_t) * vocab);
  if (e == cudaSuccess) e = cudaMalloc((void**)&sp.d_block_index, sizeof(uint32_t) * (size_t)n_slots * stride);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_toks, sizeof(uint32_t) * n_slots);
  if (e == cudaSuccess) e = cudaMemcpyAsync(sp.d_slot_of, slot_of.data(), sizeof(int32_t) * vocab, cudaMemcpyHostToDevice, s.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_toks, toks.data(), sizeof(uint32_t) * n_slots, cudaMemcpyHostToDevice, s.stream);
  sp.n_slots = n_slots;
  sp.index_stride = stride;
  for (uint32_t r0 = 0; e == cudaSuccess && r0 < n_slots; r0 += 1024)
    e = launch_sparse_block_index(sp, d_toks + r0, std::min<uint32_t>(1024, n_slots - r0), r0, s.n_rows, s.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
  cudaFree(d_toks);
  if (e != cudaSuccess) {
    // the index is an accelerator only: drop it and keep the postings usable
    cudaGetLastError();
    cudaFree(sp.d_slot_of); cudaFree(sp.d_block_index);
    sp.d_slot_of = nullptr; sp.d_block_index = nullptr; sp.n_slots = 0; sp.index_stride = 0;
    if (e != cudaErrorMemoryAllocation) CK(ix, e);
  }
  return CQS_B200_OK;
}

// Build the token-major postings from a doc-major CSR that already sits on the device
// (SpladeIndex::build, src/splade/index.rs:177-221, as a stable counting sort — sparse_build.cu).
static int sparse_build_on_device(cqs_b200_index* ix, Shard& s, const uint64_t* d_indptr, const uint32_t* d_tok,
                                  const float* d_w, uint64_t nnz, uint32_t vocab) {
  if ((size_t)vocab * sizeof(uint32_t) > 220 * 1024)
    return fail(CQS_B200_ERR_UNSUPPORTED, "vocab %u too large for the on-device build (max 56320)", vocab);
  SparseDev sp;
  void* d_scratch = nullptr;
  uint32_t* d_err = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_scratch); cudaFree(d_err);
  };
  auto drop = [&]() {
    cudaFree(sp.d_tptr); cudaFree(sp.d_doc); cudaFree(sp.d_post);
    cleanup();
  };
#define CKB(expr)                                                                      \
  do {                                                                                 \
    cudaError_t _eb = (expr);                                                          \
    if (_eb != cudaSuccess) {                                                          \
      drop();                                                                          \
      CK(ix, _eb);                                                                     \
    }                                                                                  \
  } while (0)
  CKB(cudaMalloc((void**)&sp.d_tptr, sizeof(uint64_t) * ((size_t)vocab + 1)));
  CKB(cudaMalloc((void**)&sp.d_doc, sizeof(uint32_t) * std::max<uint64_t>(nnz, 1)));
  CKB(cudaMalloc((void**)&sp.d_post, sizeof(uint2) * std::max<uint64_t>(nnz, 1)));
  CKB(cudaMalloc(&d_scratch, sparse_build_scratch_bytes(vocab, s.num_sms)));
  CKB(cudaMalloc((void**)&d_err, sizeof(uint32_t)));
  CKB(cudaMemsetAsync(d_err, 0, sizeof(uint32_t), s.stream));
  SparseBuildArgs a{d_indptr, d_tok, d_w, s.n_rows, nnz, vocab, sp.d_tptr, sp.d_doc, sp.d_post, d_scratch, d_err};
  CKB(launch_sparse_build(a, s.num_sms, s.stream));
  uint32_t err = 0;
  CKB(cudaMemcpyAsync(&err, d_err, sizeof err, cudaMemcpyDeviceToHost, s.stream));
  CKB(cudaStreamSynchronize(s.stream));
#undef CKB
  if (err) {
    drop();
    return fail(CQS_B200_ERR_INVALID, err == 1 ? "a token id is >= vocab %u" : "a doc lists a token twice (vocab %u)", vocab);
  }
  cleanup();
  free_sparse(s.sparse);
  sp.vocab = vocab;
  sp.nnz = nnz;
  s.sparse = sp;
  return sparse_build_block_index(ix, s);
}

int cqs_b200_sparse_attach(cqs_b200_index* ix, const uint64_t* indptr, const uint32_t* tok,
                           const float* w, uint32_t vocab) try {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  if (!indptr || vocab == 0) return fail(CQS_B200_ERR_INVALID, "NULL indptr / zero vocab");
  std::lock_guard<std::mutex> g(ix->mu);
  if (ix->poisoned.load()) return fail(CQS_B200_ERR_POISONED, "index is poisoned");
  if (ix->shards.size() != 1)
    return fail(CQS_B200_ERR_UNSUPPORTED, "sparse leg needs a single-device index per process");
  Shard& s = ix->shards[0];
  const uint64_t n = s.n_rows;
  if (indptr[0] != 0) return fail(CQS_B200_ERR_INVALID, "indptr[0] must be 0");
  for (uint64_t d = 0; d < n; ++d)
    if (indptr[d + 1] < indptr[d]) return fail(CQS_B200_ERR_INVALID, "indptr not monotone at %llu", (unsigned long long)d);
  const uint64_t nnz = indptr[n];
  if (nnz && (!tok || !w)) return fail(CQS_B200_ERR_INVALID, "NULL tok / w");
  // upload the doc-major rows as they come out of `sparse_vectors` (src/store/sparse.rs:342);
  // the transposition into posting lists happens on the device
  CK(ix, cudaSetDevice(s.device));
  uint64_t* d_indptr = nullptr;
  uint32_t* d_tok = nullptr;
  float* d_w = nullptr;
  auto free_in = [&]() { cudaFree(d_indptr); cudaFree(d_tok); cudaFree(d_w); };
  cudaError_t e = cudaMalloc((void**)&d_indptr, sizeof(uint64_t) * (n + 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_tok, sizeof(uint32_t) * std::max<uint64_t>(nnz, 1));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_w, sizeof(float) * std::max<uint64_t>(nnz, 1));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_indptr, indptr, sizeof(uint64_t) * (n + 1), cudaMemcpyHostToDevice, s.stream);
  if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(d_tok, tok, sizeof(uint32_t) * nnz, cudaMemcpyHostToDevice, s.stream);
  if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(d_w, w, sizeof(float) * nnz, cudaMemcpyHostToDevice, s.stream);
  if (e != cudaSuccess) {
    free_in();
    CK(ix, e);
  }
  int rc = sparse_build_on_device(ix, s, d_indptr, d_tok, d_w, nnz, vocab);
  free_in();
  return rc;
} API_CATCH

int cqs_b200_sparse_attach_device(cqs_b200_index* ix, const uint64_t* d_indptr, const uint32_t* d_tok,
                                  const float* d_w, uint64_t nnz, uint32_t vocab) try {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  if (!d_indptr || vocab == 0 || (nnz && (!d_tok || !d_w))) return fail(CQS_B200_ERR_INVALID, "NULL argument / zero vocab");
  std::lock_guard<std::mutex> g(ix->mu);
  if (ix->poisoned.load()) return fail(CQS_B200_ERR_POISONED, "index is poisoned");
  if (ix->shards.size() != 1)
    return fail(CQS_B200_ERR_UNSUPPORTED, "sparse leg needs a single-device index per process");
  Shard& s = ix->shards[0];
  CK(ix, cudaSetDevice(s.device));
  // the caller vouches for a monotone indptr with indptr[n_rows] == nnz; the two ends are checked
  uint64_t ends[2] = {1, 0};
  CK(ix, cudaMemcpy(&ends[0], d_indptr, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  CK(ix, cudaMemcpy(&ends[1], d_indptr + s.n_rows, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (ends[0] != 0 || ends[1] != nnz) return fail(CQS_B200_ERR_INVALID, "indptr[0] != 0 or indptr[n_rows] != nnz");
  return sparse_build_on_device(ix, s, d_indptr, d_tok, d_w, nnz, vocab);
} API_CATCH

// ---- SPLADE persistence (SpladeIndex::save / ::load, src/splade/index.rs:308-560) ----
// Same contract as the reference's "SPDX" file: a 64-byte header carrying a format version, the
// store's `splade_generation` counter, the chunk count and a body checksum; a load with a stale
// generation, a different chunk count or a damaged body fails and the caller rebuilds from
// `sparse_vectors`.  The body is this library's token-major layout (tptr, then (doc, weight)
// pairs); chunk ids stay in the flat-matrix sidecar.  Written to a temp file and renamed.
namespace {
struct SparseFileHeader {
  char magic[8];         // "CQSB2SPX"
  uint32_t version;      // 1
  uint32_t vocab;
  uint64_t generation;
  uint64_t n_docs;
  uint64_t nnz;
  uint64_t checksum;     // of header[0..40) + body
  uint8_t pad[16];
};
static_assert(sizeof(SparseFileHeader) == 64, "header must be 64 bytes");
uint64_t checksum_update(uint64_t h, const uint8_t* p, size_t n);
}  // namespace

int cqs_b200_sparse_save(cqs_b200_index* ix, const char* path, uint64_t generation) try {
  if (!ix || !path) return fail(CQS_B200_ERR_INVALID, "NULL argument");
  std::lock_guard<std::mutex> g(ix->mu);
  if (ix->shards.size() != 1) return fail(CQS_B200_ERR_UNSUPPORTED, "single-device index only");
  Shard& s = ix->shards[0];
  if (!s.sparse.d_tptr) return fail(CQS_B200_ERR_INVALID, "no sparse index attached");
  CK(ix, cudaSetDevice(s.device));
  const uint32_t vocab = s.sparse.vocab;
  const uint64_t nnz = s.sparse.nnz;
  std::vector<uint64_t> tptr((size_t)vocab + 1);
  std::vector<uint2> post(std::max<uint64_t>(nnz, 1));
  CK(ix, cudaMemcpy(tptr.data(), s.sparse.d_tptr, sizeof(uint64_t) * tptr.size(), cudaMemcpyDeviceToHost));
  if (nnz) CK(ix, cudaMemcpy(post.data(), s.sparse.d_post, sizeof(uint2) * nnz, cudaMemcpyDeviceToHost));
  SparseFileHeader h{};
  memcpy(h.magic, "CQSB2SPX", 8);
  h.version = 1; h.vocab = vocab; h.generation = generation; h.n_docs = s.n_rows; h.nnz = nnz;
  uint64_t ck = checksum_update(0xcbf29ce484222325ull, (const uint8_t*)&h, 40);
  ck = checksum_update(ck, (const uint8_t*)tptr.data(), sizeof(uint64_t) * tptr.size());
  ck = checksum_update(ck, (const uint8_t*)post.data(), sizeof(uint2) * nnz);
  h.checksum = ck;
  std::string tmp = std::string(path) + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return fail(CQS_B200_ERR_INVALID, "cannot open %s for writing", tmp.c_str());
  bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(tptr.data(), sizeof(uint64_t), tptr.size(), f) == tptr.size() &&
            (nnz == 0 || fwrite(post.data(), sizeof(uint2), nnz, f) == nnz);
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) {
    remove(tmp.c_str());
    return fail(CQS_B200_ERR_INVALID, "writing %s failed", path);
  }
  return CQS_B200_OK;
} API_CATCH

int cqs_b200_sparse_load(cqs_b200_index* ix, const char* path, uint64_t expected_generation) try {
  if (!ix || !path) return fail(CQS_B200_ERR_INVALID, "NULL argument");
  std::lock_guard<std::mutex> g(ix->mu);
  if (ix->poisoned.load()) return fail(CQS_B200_ERR_POISONED, "index is poisoned");
  if (ix->shards.size() != 1) return fail(CQS_B200_ERR_UNSUPPORTED, "single-device index only");
  Shard& s = ix->shards[0];
  FILE* f = fopen(path, "rb");
  if (!f) return fail(CQS_B200_ERR_INVALID, "cannot open %s", path);
  SparseFileHeader h{};
  auto bad = [&](const char* why) {
    fclose(f);
    return fail(CQS_B200_ERR_INVALID, "%s: %s (rebuild from sparse_vectors)", path, why);
  };
  if (fread(&h, sizeof h, 1, f) != 1) return bad("short header");
  if (memcmp(h.magic, "CQSB2SPX", 8) != 0 || h.version != 1) return bad("bad magic / version");
  if (h.generation != expected_generation) return bad("stale generation");
  if (h.n_docs != s.n_rows) return bad("chunk count differs from the index");
  // the header is untrusted until the checksum passes: bound it and match it against the file
  // size BEFORE anything is sized from it (a damaged vocab / nnz must not drive a 500 GB vector)
  if (h.vocab == 0 || h.vocab > 56320 || h.nnz > (1ull << 36)) return bad("implausible header");
  {
    if (fseek(f, 0, SEEK_END) != 0) return bad("cannot seek");
    const long fsize = ftell(f);
    const uint64_t want = sizeof h + 8ull * ((uint64_t)h.vocab + 1) + 8ull * h.nnz;
    if (fsize < 0 || (uint64_t)fsize != want) return bad("size does not match the header");
    if (fseek(f, (long)sizeof h, SEEK_SET) != 0) return bad("cannot seek");
  }
  std::vector<uint64_t> tptr((size_t)h.vocab + 1);
  std::vector<uint2> post(std::max<uint64_t>(h.nnz, 1));
  if (fread(tptr.data(), sizeof(uint64_t), tptr.size(), f) != tptr.size()) return bad("truncated");
  if (h.nnz && fread(post.data(), sizeof(uint2), h.nnz, f) != h.nnz) return bad("truncated");
  uint64_t ck = checksum_update(0xcbf29ce484222325ull, (const uint8_t*)&h, 40);
  ck = checksum_update(ck, (const uint8_t*)tptr.data(), sizeof(uint64_t) * tptr.size());
  ck = checksum_update(ck, (const uint8_t*)post.data(), sizeof(uint2) * h.nnz);
  if (ck != h.checksum) return bad("checksum mismatch");
  if (tptr[0] != 0 || tptr[h.vocab] != h.nnz) return bad("inconsistent offsets");
  for (uint32_t t = 0; t < h.vocab; ++t)
    if (tptr[t + 1] < tptr[t]) return bad("inconsistent offsets");
  fclose(f);
  std::vector<uint32_t> doc(std::max<uint64_t>(h.nnz, 1));
  for (uint64_t i = 0; i < h.nnz; ++i) {
    if (post[i].x >= s.n_rows) return fail(CQS_B200_ERR_INVALID, "%s: posting refers to chunk %u >= %llu", path, post[i].x, (unsigned long long)s.n_rows);
    doc[i] = post[i].x;
  }
  CK(ix, cudaSetDevice(s.device));
  SparseDev sp;
  cudaError_t e = cudaMalloc((void**)&sp.d_tptr, sizeof(uint64_t) * tptr.size());
  if (e == cudaSuccess) e = cudaMalloc((void**)&sp.d_doc, sizeof(uint32_t) * doc.size());
  if (e == cudaSuccess) e = cudaMalloc((void**)&sp.d_post, sizeof(uint2) * post.size());
  if (e == cudaSuccess) e = cudaMemcpy(sp.d_tptr, tptr.data(), sizeof(uint64_t) * tptr.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(sp.d_doc, doc.data(), sizeof(uint32_t) * doc.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(sp.d_post, post.data(), sizeof(uint2) * post.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(sp.d_tptr); cudaFree(sp.d_doc); cudaFree(sp.d_post);
    CK(ix, e);
  }
  free_sparse(s.sparse);
  sp.vocab = h.vocab;
  sp.nnz = h.nnz;
  s.sparse = sp;
  return sparse_build_block_index(ix, s);
} API_CATCH

// test hook: download the built postings (tptr [vocab+1], doc [nnz], weight [nnz])
int cqs_b200_debug_sparse_postings(cqs_b200_index* ix, uint64_t* tptr, uint32_t* doc, float* w) try {
  if (!ix || ix->shards.size() != 1 || !ix->shards[0].sparse.d_tptr) return CQS_B200_ERR_INVALID;
  Shard& s = ix->shards[0];
  cudaSetDevice(s.device);
  const uint64_t nnz = s.sparse.nnz;
  if (tptr) cudaMemcpy(tptr, s.sparse.d_tptr, sizeof(uint64_t) * ((size_t)s.sparse.vocab + 1), cudaMemcpyDeviceToHost);
  if (nnz && doc && w) {
    std::vector<uint2> post(nnz);
    cudaMemcpy(post.data(), s.sparse.d_post, sizeof(uint2) * nnz, cudaMemcpyDeviceToHost);
    std::vector<uint32_t> d2(nnz);
    cudaMemcpy(d2.data(), s.sparse.d_doc, sizeof(uint32_t) * nnz, cudaMemcpyDeviceToHost);
    for (uint64_t i = 0; i < nnz; ++i) {
      if (d2[i] != post[i].x) return CQS_B200_ERR_INVALID;   // the two arrays must agree
      doc[i] = post[i].x;
      memcpy(&w[i], &post[i].y, 4);
    }
  }
  return cudaGetLastError() == cudaSuccess ? CQS_B200_OK : CQS_B200_ERR_CUDA;
} API_CATCH

// uploads the sparse query on `st`, sizes the per-query scratch and fills the launch arguments
static int stage_sparse(cqs_b200_index* ix, Shard& s, const uint32_t* q_tok, const float* q_w,
                        uint32_t q_nnz, uint32_t k, const uint32_t* d_bits, cudaStream_t st, SparseArgs& a) {
  // through pinned staging: a pageable source would make these copies synchronous with the host
  memcpy(s.h_sq, q_tok, sizeof(uint32_t) * q_nnz);
  memcpy(s.h_sq + kSpMaxQ, q_w, sizeof(float) * q_nnz);
  CK(ix, cudaMemcpyAsync(s.d_q_tok, s.h_sq, sizeof(uint32_t) * q_nnz, cudaMemcpyHostToDevice, st));
  CK(ix, cudaMemcpyAsync(s.d_q_w, s.h_sq + kSpMaxQ, sizeof(float) * q_nnz, cudaMemcpyHostToDevice, st));
  const size_t need = sparse_bounds_bytes(s.n_rows, q_nnz);
  if (need > s.bounds_bytes) {
    CK(ix, cudaStreamSynchronize(st));
    cudaFree(s.d_bounds);
    s.d_bounds = nullptr;
    s.bounds_bytes = 0;
    CK(ix, cudaMalloc((void**)&s.d_bounds, need + need / 2));
    s.bounds_bytes = need + need / 2;
  }
  const size_t need_blk = sparse_block_scratch_bytes(s.n_rows);
  if (need_blk > s.sp_block_bytes) {
    CK(ix, cudaStreamSynchronize(st));
    cudaFree(s.d_sp_block);
    s.d_sp_block = nullptr;
    s.sp_block_bytes = 0;
    CK(ix, cudaMalloc(&s.d_sp_block, need_blk));
    s.sp_block_bytes = need_blk;
  }
  a.d_block_scratch = s.d_sp_block;
  a.d_claim = s.d_sp_claim;
  a.d_bounds = s.d_bounds;
  a.d_trace = s.d_trace ? s.d_trace + 512 * 8 : nullptr;   // rows 512.. of the trace buffer (the scan uses rows 0..147 and 1023)
  a.sp = s.sparse; a.n_docs = s.n_rows; a.d_q_tok = s.d_q_tok; a.d_q_w = s.d_q_w; a.q_nnz = q_nnz;
  a.d_bitset = d_bits; a.k = k; a.row_base = ix->row_base + s.first_row;
  a.d_partial = s.d_sp_partial; a.d_partial_cnt = s.d_sp_partial_cnt; a.d_done = s.d_sp_done;
  a.d_out_scores = s.d_sp_scores; a.d_out_rows = s.d_sp_rows; a.d_out_n = s.d_sp_n;
  return 0;
}

static int launch_sparse(cqs_b200_index* ix, Shard& s, const uint32_t* q_tok, const float* q_w,
                         uint32_t q_nnz, uint32_t k, const uint32_t* d_bits) {
  SparseArgs a;
  int rc = stage_sparse(ix, s, q_tok, q_w, q_nnz, k, d_bits, s.stream, a);
  if (rc) return rc;
  CK(ix, launch_sparse_search(a, s.stream));
  return 0;
}

int cqs_b200_search_sparse(cqs_b200_index* ix, const uint32_t* q_tok, const float* q_w,
                           uint32_t q_nnz, uint32_t k, const uint32_t* bitset, uint64_t* out_rows,
                           float* out_scores, uint32_t* out_n) try {
  if (out_n) *out_n = 0;
  int rc = check_searchable(ix);
  if (rc) return rc;
  if (!out_rows || !out_scores || !out_n) return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (k > kMaxK) return fail(CQS_B200_ERR_INVALID, "k=%u exceeds max_k=%u", k, kMaxK);
  if (q_nnz > kSpMaxQ) return fail(CQS_B200_ERR_INVALID, "query nnz %u exceeds %u", q_nnz, kSpMaxQ);
  if (ix->shards.size() != 1) return fail(CQS_B200_ERR_UNSUPPORTED, "single-device index required");
  std::lock_guard<std::mutex> g(ix->mu);
  Shard& s = ix->shards[0];
  if (!s.sparse.d_tptr) return fail(CQS_B200_ERR_INVALID, "no sparse index attached");
  if (k == 0 || q_nnz == 0 || s.n_rows == 0) return CQS_B200_OK;  // index.rs:236-238
  if (!q_tok || !q_w) return fail(CQS_B200_ERR_INVALID, "NULL query");
  CK(ix, cudaSetDevice(s.device));
  const uint32_t* d_bits = nullptr;
  if (bitset) {
    CK(ix, cudaMemcpyAsync(s.d_bitset, bitset, ((s.n_rows + 31) / 32) * 4, cudaMemcpyHostToDevice, s.stream));
    d_bits = s.d_bitset;
  }
  if (ix->timing) CK(ix, cudaEventRecord(s.ev0, s.stream));
  rc = launch_sparse(ix, s, q_tok, q_w, q_nnz, k, d_bits);
  if (rc) return rc;
  if (ix->timing) CK(ix, cudaEventRecord(s.ev1, s.stream));
  float* hs = (float*)s.h_out;
  uint64_t* hr = (uint64_t*)(s.h_out + sizeof(float) * kMaxK);
  uint32_t* hn = (uint32_t*)(s.h_out + (sizeof(float) + sizeof(uint64_t)) * kMaxK);
  CK(ix, cudaMemcpyAsync(hs, s.d_sp_scores, sizeof(float) * k, cudaMemcpyDeviceToHost, s.stream));
  CK(ix, cudaMemcpyAsync(hr, s.d_sp_rows, sizeof(uint64_t) * k, cudaMemcpyDeviceToHost, s.stream));
  CK(ix, cudaMemcpyAsync(hn, s.d_sp_n, sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
  CK(ix, cudaStreamSynchronize(s.stream));
  if (ix->timing) CK(ix, cudaEventElapsedTime(&ix->last_kernel_ms, s.ev0, s.ev1));
  uint32_t n = std::min(*hn, k);
  memcpy(out_scores, hs, sizeof(float) * n);
  memcpy(out_rows, hr, sizeof(uint64_t) * n);
  *out_n = n;
  return CQS_B200_OK;
} API_CATCH

// peer == nullptr: this index alone.  Otherwise the corpus is row-sharded: the dense leg's scan
// exchanges + merges in its tail (GLOBAL dense pool on every rank), the sparse leg's per-shard pool
// goes through one gather+merge kernel (GLOBAL sparse pool), and every rank fuses the same two
// pools — max_sparse is the merged pool's top-1 — so all ranks return the unsharded answer
// (SURVEY.md §8e).
static int search_hybrid_impl(cqs_b200_index* ix, cqs_b200_peer* peer, const float* query,
                              const uint32_t* q_tok, const float* q_w, uint32_t q_nnz, float alpha,
                              uint32_t pool_k, const uint32_t* bitset, uint64_t* out_rows,
                              float* out_fused, float* out_dense, float* out_sparse_raw,
                              uint8_t* out_present, uint32_t* out_n) {
  if (out_n) *out_n = 0;
  int rc = check_searchable(ix);
  if (rc) return rc;
  if (!query || !out_rows || !out_fused || !out_n) return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (pool_k > kMaxK) return fail(CQS_B200_ERR_INVALID, "pool_k=%u exceeds max_k=%u", pool_k, kMaxK);
  if (q_nnz > kSpMaxQ) return fail(CQS_B200_ERR_INVALID, "query nnz %u exceeds %u", q_nnz, kSpMaxQ);
  if (ix->shards.size() != 1) return fail(CQS_B200_ERR_UNSUPPORTED, "single-device index required");
  if (pool_k == 0 || ix->n_rows == 0) return CQS_B200_OK;
  std::lock_guard<std::mutex> g(ix->mu);
  if (peer && (rc = check_peer(ix, peer))) return rc;
  std::unique_lock<std::mutex> gp;
  if (peer) gp = std::unique_lock<std::mutex>(peer->mu);
  Shard& s = ix->shards[0];
  if (!s.sparse.d_tptr) return fail(CQS_B200_ERR_INVALID, "no sparse index attached");
  CK(ix, cudaSetDevice(s.device));
  // dense leg: a malformed query yields an EMPTY dense pool (src/cagra.rs:458-470);
  // the sparse leg still runs (search_hybrid_inner calls both unconditionally).
  const bool dense_ok = query_is_finite(query, ix->dim);
  if (ix->timing) CK(ix, cudaEventRecord(s.ev_b0, s.stream));
  CK(ix, cudaMemsetAsync(s.d_out_n, 0, 4, s.stream));
  CK(ix, cudaMemsetAsync(s.d_sp_n, 0, 4, s.stream));
  if (peer) CK(ix, cudaMemsetAsync(s.d_spm_n, 0, 4, s.stream));
  // The legs are independent until the fusion.  Without a filter or a peer group (one staging buffer,
  // one ordered exchange stream) the sparse leg's accumulate + select go to a second stream: their CTAs
  // cannot share an SM with the scan's, so they start as the scan's CTAs retire and run while the scan's
  // LAST CTA merges the per-CTA lists (tens of microseconds on one SM at k = 500) — the dense leg's tail
  // is hidden behind the sparse leg instead of preceding it.  The query upload and the bounds pass go
  // BEFORE the scan (a few microseconds on an idle device instead of after the scan).
  const bool overlap = !peer && !bitset && dense_ok && q_nnz != 0 && q_tok && q_w;
  SparseArgs sa;
  if (overlap) {
    if ((rc = stage_sparse(ix, s, q_tok, q_w, q_nnz, pool_k, nullptr, s.stream, sa))) return rc;
    CK(ix, launch_sparse_stage(sa, kSparseBounds, s.stream));
  }
  CK(ix, cudaEventRecord(s.ev_fork, s.stream));   // fork point of the sparse leg's streams
  const uint32_t* d_bits = nullptr;
  if (dense_ok) {
    PeerCtx pc;
    if (peer) CK(ix, peer_begin(peer, s.stream, &pc, /*exclusive=*/false));
    // (the fusion kernel reads the pool's length on the device: always the f32 master here)
    rc = launch_dense(ix, s, query, pool_k, bitset, nullptr, false, peer ? &pc : nullptr, /*force_master=*/true);
    if (rc) return rc;
    if (peer) CK(ix, peer_mark(peer, s.stream, /*exclusive=*/false));
    if (bitset) d_bits = s.d_bitset;
  } else if (bitset) {
    CK(ix, cudaMemcpyAsync(s.d_bitset, bitset, ((s.n_rows + 31) / 32) * 4, cudaMemcpyHostToDevice, s.stream));
    d_bits = s.d_bitset;
  }
  if (q_nnz) {
    if (!q_tok || !q_w) return fail(CQS_B200_ERR_INVALID, "NULL sparse query");
    if (overlap) {
      cudaStream_t l0 = s.lane_stream[0];
      CK(ix, cudaStreamWaitEvent(l0, s.ev_fork, 0));
      CK(ix, launch_sparse_stage(sa, kSparseAccum | kSparseSelect, l0));
      CK(ix, cudaEventRecord(s.ev_join[0], l0));
      CK(ix, cudaStreamWaitEvent(s.stream, s.ev_join[0], 0));
    } else {
      rc = launch_sparse(ix, s, q_tok, q_w, q_nnz, pool_k, d_bits);
      if (rc) return rc;
    }
    if (peer) {
      PeerCtx pc;
      CK(ix, peer_begin(peer, s.stream, &pc, /*exclusive=*/true));
      PeerGatherArgs ga{s.d_sp_scores, s.d_sp_rows, s.d_sp_n, 1, pool_k,
                        s.d_spm_scores, s.d_spm_rows, s.d_spm_n, peer->d_ticket};
      CK(ix, launch_peer_gather_merge(pc, ga, s.num_sms, s.stream));
      CK(ix, peer_mark(peer, s.stream, /*exclusive=*/true));
    }
  }
  FuseArgs f;
  f.d_dense_rows = s.d_out_rows; f.d_dense_scores = s.d_out_scores; f.d_n_dense = s.d_out_n;
  f.d_sparse_rows = peer ? s.d_spm_rows : s.d_sp_rows;
  f.d_sparse_scores = peer ? s.d_spm_scores : s.d_sp_scores;
  f.d_n_sparse = peer ? s.d_spm_n : s.d_sp_n;
  f.alpha = alpha; f.pool_k = pool_k;
  // the fused pool is written straight into the host-mapped result buffer (layout of copy_fused_out)
  // and a completion word is polled, as on the dense latency path
  {
    uint8_t* d = s.d_hout;
    f.d_out_rows = (uint64_t*)d;
    f.d_out_fused = (float*)(d + 8 * kMaxK);
    f.d_out_dense = f.d_out_fused + kMaxK;
    f.d_out_sparse_raw = f.d_out_dense + kMaxK;
    f.d_out_present = (uint8_t*)(f.d_out_sparse_raw + kMaxK);
    f.d_out_n = (uint32_t*)(f.d_out_present + kMaxK);
    f.d_host_flag = (uint32_t*)(d + kOffHybridFlag);
    f.seq = ++s.seq;
    if (f.seq == 0) f.seq = s.seq = 1;
  }
  CK(ix, launch_fuse_pools(f, s.stream));
  if (ix->timing) CK(ix, cudaEventRecord(s.ev_b1, s.stream));
  if ((rc = wait_host_flag(ix, s, kOffHybridFlag))) return rc;
  {
    const uint8_t* h = s.h_out;
    const uint32_t n = std::min(*(const uint32_t*)(h + (8 + 4 + 4 + 4 + 1) * kMaxK), pool_k);
    memcpy(out_rows, h, 8 * (size_t)n);
    memcpy(out_fused, h + 8 * kMaxK, 4 * (size_t)n);
    if (out_dense) memcpy(out_dense, h + 12 * kMaxK, 4 * (size_t)n);
    if (out_sparse_raw) memcpy(out_sparse_raw, h + 16 * kMaxK, 4 * (size_t)n);
    if (out_present) memcpy(out_present, h + 20 * kMaxK, n);
    *out_n = n;
  }
  if (ix->timing) CK(ix, cudaEventSynchronize(s.ev_b1));
  // device time of the whole hybrid pipeline (both legs + fusion), read like the batch timer
  if (ix->timing) CK(ix, cudaEventElapsedTime(&ix->last_batch_ms, s.ev_b0, s.ev_b1));
  if (peer) {
    uint32_t st = 0;
    CK(ix, cudaMemcpy(&st, peer->d_status, sizeof st, cudaMemcpyDeviceToHost));
    if (st) {
      peer_failed(ix, peer);
      *out_n = 0;
      return fail(CQS_B200_ERR_CUDA, "peer exchange timed out (a rank did not take part in this search)");
    }
  }
  if (dense_ok && ix->timing) CK(ix, cudaEventElapsedTime(&ix->last_kernel_ms, s.ev0, s.ev1));
  return CQS_B200_OK;
}

int cqs_b200_search_hybrid(cqs_b200_index* ix, const float* query, const uint32_t* q_tok,
                           const float* q_w, uint32_t q_nnz, float alpha, uint32_t pool_k,
                           const uint32_t* bitset, uint64_t* out_rows, float* out_fused,
                           float* out_dense, float* out_sparse_raw, uint8_t* out_present,
                           uint32_t* out_n) try {
  return search_hybrid_impl(ix, nullptr, query, q_tok, q_w, q_nnz, alpha, pool_k, bitset, out_rows,
                            out_fused, out_dense, out_sparse_raw, out_present, out_n);
} API_CATCH

int cqs_b200_search_hybrid_sharded(cqs_b200_index* ix, cqs_b200_peer* peer, const float* query,
                                   const uint32_t* q_tok, const float* q_w, uint32_t q_nnz,
                                   float alpha, uint32_t pool_k, const uint32_t* bitset,
                                   uint64_t* out_rows, float* out_fused, float* out_dense,
                                   float* out_sparse_raw, uint8_t* out_present, uint32_t* out_n) try {
  if (!peer) return fail(CQS_B200_ERR_INVALID, "peer is NULL");
  return search_hybrid_impl(ix, peer, query, q_tok, q_w, q_nnz, alpha, pool_k, bitset, out_rows,
                            out_fused, out_dense, out_sparse_raw, out_present, out_n);
} API_CATCH

int cqs_b200_fuse_pools(int device, const uint64_t* dense_rows, const float* dense_scores,
                        uint32_t n_dense, const uint64_t* sparse_rows, const float* sparse_scores,
                        uint32_t n_sparse, float alpha, uint32_t pool_k, uint64_t* out_rows,
                        float* out_fused, float* out_dense, float* out_sparse_raw,
                        uint8_t* out_present, uint32_t* out_n) try {
  if (out_n) *out_n = 0;
  if (!out_rows || !out_fused || !out_n) return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (n_dense > kMaxK || n_sparse > kMaxK || pool_k > 2 * kMaxK)
    return fail(CQS_B200_ERR_INVALID, "pool larger than %u", kMaxK);
  if ((n_dense && (!dense_rows || !dense_scores)) || (n_sparse && (!sparse_rows || !sparse_scores)))
    return fail(CQS_B200_ERR_INVALID, "NULL pool");
  uint32_t cap = std::min(pool_k, n_dense + n_sparse);
  if (cap == 0) return CQS_B200_OK;
  {
    // the fusion sort packs (row - min_row) into 32 bits
    uint64_t lo = ~0ull, hi = 0;
    for (uint32_t j = 0; j < n_dense; ++j) if (dense_rows[j] != ~0ull) { lo = std::min(lo, dense_rows[j]); hi = std::max(hi, dense_rows[j]); }
    for (uint32_t j = 0; j < n_sparse; ++j) if (sparse_rows[j] != ~0ull) { lo = std::min(lo, sparse_rows[j]); hi = std::max(hi, sparse_rows[j]); }
    if (hi > lo && hi - lo > 0xFFFFFFFEull) return fail(CQS_B200_ERR_UNSUPPORTED, "row ids span more than 2^32");
  }
  cqs_b200_index* none = nullptr;
  CK(none, cudaSetDevice(device));
  // one scratch allocation: [dense rows | sparse rows | dense sc | sparse sc | counts | outputs]
  const size_t K = kMaxK, OUTK = 2 * kMaxK;
  size_t bytes = 8 * K * 2 + 4 * K * 2 + 16 + OUTK * (8 + 4 + 4 + 4 + 1) + 16;
  uint8_t* d = nullptr;
  CK(none, cudaMalloc((void**)&d, bytes));
  uint64_t* d_dr = (uint64_t*)d;
  uint64_t* d_sr = d_dr + K;
  float* d_ds = (float*)(d_sr + K);
  float* d_ss = d_ds + K;
  uint32_t* d_cnt = (uint32_t*)(d_ss + K);
  uint64_t* o_rows = (uint64_t*)(d_cnt + 4);
  float* o_f = (float*)(o_rows + OUTK);
  float* o_d = o_f + OUTK;
  float* o_s = o_d + OUTK;
  uint8_t* o_p = (uint8_t*)(o_s + OUTK);
  uint32_t cnt[2] = {n_dense, n_sparse};
  int rc = CQS_B200_OK;
  auto body = [&]() -> int {
    if (n_dense) {
      CK(none, cudaMemcpy(d_dr, dense_rows, 8 * n_dense, cudaMemcpyHostToDevice));
      CK(none, cudaMemcpy(d_ds, dense_scores, 4 * n_dense, cudaMemcpyHostToDevice));
    }
    if (n_sparse) {
      CK(none, cudaMemcpy(d_sr, sparse_rows, 8 * n_sparse, cudaMemcpyHostToDevice));
      CK(none, cudaMemcpy(d_ss, sparse_scores, 4 * n_sparse, cudaMemcpyHostToDevice));
    }
    CK(none, cudaMemcpy(d_cnt, cnt, 8, cudaMemcpyHostToDevice));
    FuseArgs f;
    f.d_dense_rows = d_dr; f.d_dense_scores = d_ds; f.d_n_dense = d_cnt;
    f.d_sparse_rows = d_sr; f.d_sparse_scores = d_ss; f.d_n_sparse = d_cnt + 1;
    f.alpha = alpha; f.pool_k = cap;
    f.d_out_rows = o_rows; f.d_out_fused = o_f; f.d_out_dense = o_d; f.d_out_sparse_raw = o_s;
    f.d_out_present = o_p; f.d_out_n = d_cnt + 2;
    CK(none, launch_fuse_pools(f, 0));
    uint32_t n = 0;
    CK(none, cudaMemcpy(&n, d_cnt + 2, 4, cudaMemcpyDeviceToHost));
    n = std::min(n, cap);
    CK(none, cudaMemcpy(out_rows, o_rows, 8 * n, cudaMemcpyDeviceToHost));
    CK(none, cudaMemcpy(out_fused, o_f, 4 * n, cudaMemcpyDeviceToHost));
    if (out_dense) CK(none, cudaMemcpy(out_dense, o_d, 4 * n, cudaMemcpyDeviceToHost));
    if (out_sparse_raw) CK(none, cudaMemcpy(out_sparse_raw, o_s, 4 * n, cudaMemcpyDeviceToHost));
    if (out_present) CK(none, cudaMemcpy(out_present, o_p, n, cudaMemcpyDeviceToHost));
    *out_n = n;
    return CQS_B200_OK;
  };
  rc = body();
  cudaFree(d);
  return rc;
} API_CATCH

int cqs_b200_rrf_fuse(int device, const uint64_t* ids, const uint32_t* list_len, uint32_t n_lists,
                      float k, uint32_t limit, uint64_t* out_ids, float* out_scores, uint32_t* out_n) try {
  if (out_n) *out_n = 0;
  if (!out_ids || !out_scores || !out_n) return fail(CQS_B200_ERR_INVALID, "NULL argument");
  if (n_lists == 0 || limit == 0) return CQS_B200_OK;
  if (!ids || !list_len) return fail(CQS_B200_ERR_INVALID, "NULL lists");
  std::vector<uint32_t> off(n_lists + 1, 0);
  for (uint32_t l = 0; l < n_lists; ++l) off[l + 1] = off[l] + list_len[l];
  const uint32_t total = off[n_lists];
  if (total == 0) return CQS_B200_OK;
  if (total > kRrfMaxEntries) return fail(CQS_B200_ERR_INVALID, "more than %u ids in total", kRrfMaxEntries);
  uint64_t lo = ~0ull, hi = 0;
  for (uint32_t j = 0; j < total; ++j) if (ids[j] != ~0ull) { lo = std::min(lo, ids[j]); hi = std::max(hi, ids[j]); }
  if (hi > lo && hi - lo > 0xFFFFFFFEull) return fail(CQS_B200_ERR_UNSUPPORTED, "ids span more than 2^32");
  cqs_b200_index* none = nullptr;
  CK(none, cudaSetDevice(device));
  const uint32_t cap = std::min(limit, total);
  uint8_t* d = nullptr;
  const size_t bytes = 8 * (size_t)total + 4 * (n_lists + 1) + 16 + (8 + 4) * (size_t)cap + 64;
  CK(none, cudaMalloc((void**)&d, bytes));
  uint64_t* d_ids = (uint64_t*)d;
  uint64_t* d_out = d_ids + total;
  float* d_sc = (float*)(d_out + cap);
  uint32_t* d_off = (uint32_t*)(d_sc + cap);
  uint32_t* d_n = d_off + n_lists + 1;
  auto body = [&]() -> int {
    CK(none, cudaMemcpy(d_ids, ids, 8 * (size_t)total, cudaMemcpyHostToDevice));
    CK(none, cudaMemcpy(d_off, off.data(), 4 * (n_lists + 1), cudaMemcpyHostToDevice));
    CK(none, launch_rrf_fuse(d_ids, d_off, n_lists, k, cap, d_out, d_sc, d_n, 0));
    uint32_t n = 0;
    CK(none, cudaMemcpy(&n, d_n, 4, cudaMemcpyDeviceToHost));
    n = std::min(n, cap);
    CK(none, cudaMemcpy(out_ids, d_out, 8 * (size_t)n, cudaMemcpyDeviceToHost));
    CK(none, cudaMemcpy(out_scores, d_sc, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    *out_n = n;
    return CQS_B200_OK;
  };
  int rc = body();
  cudaFree(d);
  return rc;
} API_CATCH

int cqs_b200_search_typed(cqs_b200_index* ix, const float* query, uint32_t k, const uint64_t* type_mask,
                          const uint64_t* lang_mask, uint64_t* out_rows, float* out_scores, uint32_t* out_n) try {
  if (out_n) *out_n = 0;
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  if (!type_mask && !lang_mask) return search_impl(ix, query, k, nullptr, nullptr, out_rows, out_scores, out_n);
  if (ix->n_rows && (ix->shards.empty() || !ix->shards[0].d_ctype))
    return fail(CQS_B200_ERR_INVALID, "type/language filter requested but cqs_b200_set_row_meta was not called");
  ScanSignals sig;
  if (type_mask) memcpy(sig.type_mask, type_mask, 32);
  if (lang_mask) memcpy(sig.lang_mask, lang_mask, 32);
  sig.pipeline = 0;  // raw cosine, filter only
  return search_impl(ix, query, k, nullptr, &sig, out_rows, out_scores, out_n);
} API_CATCH

int cqs_b200_route_centroids(int device, const float* centroids, uint32_t n_c, uint32_t dim,
                             const float* queries, uint32_t nq, float threshold, int32_t* out_cat,
                             float* out_margin) try {
  if (nq == 0) return CQS_B200_OK;
  if (!centroids || !queries || !out_cat || !out_margin || n_c == 0 || dim == 0)
    return fail(CQS_B200_ERR_INVALID, "NULL / empty argument");
  cqs_b200_index* none = nullptr;
  CK(none, cudaSetDevice(device));
  float *d_c = nullptr, *d_q = nullptr, *d_m = nullptr;
  int32_t* d_cat = nullptr;
  int rc = CQS_B200_OK;
  auto body = [&]() -> int {
    CK(none, cudaMalloc((void**)&d_c, sizeof(float) * (size_t)n_c * dim));
    CK(none, cudaMalloc((void**)&d_q, sizeof(float) * (size_t)nq * dim));
    CK(none, cudaMalloc((void**)&d_m, sizeof(float) * nq));
    CK(none, cudaMalloc((void**)&d_cat, sizeof(int32_t) * nq));
    CK(none, cudaMemcpy(d_c, centroids, sizeof(float) * (size_t)n_c * dim, cudaMemcpyHostToDevice));
    CK(none, cudaMemcpy(d_q, queries, sizeof(float) * (size_t)nq * dim, cudaMemcpyHostToDevice));
    CK(none, launch_route_centroids(d_c, n_c, dim, d_q, nq, threshold, d_cat, d_m, 0));
    CK(none, cudaMemcpy(out_cat, d_cat, sizeof(int32_t) * nq, cudaMemcpyDeviceToHost));
    CK(none, cudaMemcpy(out_margin, d_m, sizeof(float) * nq, cudaMemcpyDeviceToHost));
    return CQS_B200_OK;
  };
  rc = body();
  cudaFree(d_c); cudaFree(d_q); cudaFree(d_m); cudaFree(d_cat);
  return rc;
} API_CATCH

// ---- persistence -----------------------------------------------------------------------

namespace {
struct FileHeader {
  char magic[8];
  uint32_t version, dim, storage, metric;
  uint64_t n_rows, checksum, reserved[3];
};
static_assert(sizeof(FileHeader) == 64, "header must be 64 bytes");
// 64-bit multiply-rotate checksum over 8-byte words (tail bytes zero padded)
uint64_t checksum_update(uint64_t h, const uint8_t* p, size_t n) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    h = (h ^ w) * 0x9E3779B97F4A7C15ull;
    h = (h << 31) | (h >> 33);
  }
  if (i < n) {
    uint64_t w = 0;
    memcpy(&w, p + i, n - i);
    h = (h ^ w) * 0x9E3779B97F4A7C15ull;
    h = (h << 31) | (h >> 33);
  }
  return h;
}
}  // namespace

int cqs_b200_save(cqs_b200_index* ix, const char* path) try {
  if (!ix || !path) return fail(CQS_B200_ERR_INVALID, "NULL argument");
  std::lock_guard<std::mutex> g(ix->mu);
  if (ix->poisoned.load()) return fail(CQS_B200_ERR_POISONED, "index is poisoned");
  if (!ix->finalized) return fail(CQS_B200_ERR_INVALID, "index is not finalized");
  const std::string tmp = std::string(path) + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return fail(CQS_B200_ERR_INVALID, "cannot open %s for writing", tmp.c_str());
  FileHeader hd{};
  memcpy(hd.magic, "CQSB2001", 8);
  hd.version = 1; hd.dim = ix->dim; hd.storage = (uint32_t)ix->storage; hd.metric = (uint32_t)ix->metric;
  hd.n_rows = ix->n_rows;
  const size_t esz = ix->layout.mode == 0 ? 4 : 2;
  const size_t row_out = (size_t)ix->dim * esz, row_in = (size_t)ix->layout.ld * esz;
  const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / row_out);
  std::vector<uint8_t> buf(chunk * row_out);
  uint64_t h = 0x243F6A8885A308D3ull;
  bool ok = fwrite(&hd, sizeof hd, 1, f) == 1;
  for (auto& s : ix->shards) {
    if (cudaSetDevice(s.device) != cudaSuccess) ok = false;
    for (uint64_t r = 0; ok && r < s.n_rows; r += chunk) {
      const uint64_t m = std::min(chunk, s.n_rows - r);
      if (cudaMemcpy2D(buf.data(), row_out, s.d_rows + r * row_in, row_in, row_out, m,
                       cudaMemcpyDeviceToHost) != cudaSuccess) {
        ix->poisoned.store(1);
        ok = false;
        break;
      }
      h = checksum_update(h, buf.data(), m * row_out);
      ok = fwrite(buf.data(), 1, m * row_out, f) == m * row_out;
    }
  }
  hd.checksum = h;
  ok = ok && fseek(f, 0, SEEK_SET) == 0 && fwrite(&hd, sizeof hd, 1, f) == 1;
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) {
    remove(tmp.c_str());
    return fail(CQS_B200_ERR_INVALID, "writing %s failed", path);
  }
  return CQS_B200_OK;
} API_CATCH

int cqs_b200_load(const char* path, const int* device_ids, int n_dev, cqs_b200_index** out) try {
  if (!out) return fail(CQS_B200_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!path) return fail(CQS_B200_ERR_INVALID, "path is NULL");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(CQS_B200_ERR_INVALID, "cannot open %s", path);
  FileHeader hd{};
  if (fread(&hd, sizeof hd, 1, f) != 1 || memcmp(hd.magic, "CQSB2001", 8) != 0 || hd.version != 1 ||
      hd.storage > 2 || hd.dim == 0 || hd.dim > 2048) {
    fclose(f);
    return fail(CQS_B200_ERR_INVALID, "%s: not a cqs_b200 index file (bad magic / header)", path);
  }
  const size_t esz = hd.storage == CQS_B200_STORAGE_BF16 ? 2 : 4;
  const size_t row = (size_t)hd.dim * esz;
  fseek(f, 0, SEEK_END);
  const long fsize = ftell(f);
  if (fsize < 0 || (uint64_t)fsize != sizeof hd + hd.n_rows * row) {
    fclose(f);
    return fail(CQS_B200_ERR_INVALID, "%s: size does not match the header (truncated?)", path);
  }
  fseek(f, sizeof hd, SEEK_SET);
  cqs_b200_index* ix = nullptr;
  int rc = cqs_b200_create(device_ids, n_dev, hd.dim, (int)hd.metric, (int)hd.storage, &ix);
  if (rc) { fclose(f); return rc; }
  rc = cqs_b200_reserve(ix, hd.n_rows);
  const uint64_t chunk = std::max<uint64_t>(1, (64ull << 20) / row);
  std::vector<uint8_t> buf(chunk * row);
  std::vector<float> wide;
  uint64_t h = 0x243F6A8885A308D3ull;
  for (uint64_t r = 0; rc == 0 && r < hd.n_rows; r += chunk) {
    const uint64_t m = std::min(chunk, hd.n_rows - r);
    if (fread(buf.data(), 1, m * row, f) != m * row) { rc = fail(CQS_B200_ERR_INVALID, "%s: short read", path); break; }
    h = checksum_update(h, buf.data(), m * row);
    if (esz == 4) {
      rc = cqs_b200_append_rows_f32(ix, (const float*)buf.data(), m);
    } else {
      // bf16 -> f32 is exact, and the device's RNE rounding of an exact bf16 value is the identity
      wide.resize(m * hd.dim);
      const uint16_t* b16 = (const uint16_t*)buf.data();
      for (size_t i = 0; i < m * hd.dim; ++i) {
        uint32_t u = (uint32_t)b16[i] << 16;
        memcpy(&wide[i], &u, 4);
      }
      rc = cqs_b200_append_rows_f32(ix, wide.data(), m);
    }
  }
  fclose(f);
  if (rc == 0 && h != hd.checksum) rc = fail(CQS_B200_ERR_INVALID, "%s: checksum mismatch (corrupt file)", path);
  if (rc == 0) rc = cqs_b200_finalize(ix);
  if (rc) {
    std::string msg = t_last_error;
    cqs_b200_destroy(ix);
    t_last_error = msg;
    return rc;
  }
  *out = ix;
  return CQS_B200_OK;
} API_CATCH

// ---- introspection ---------------------------------------------------------

uint64_t cqs_b200_len(const cqs_b200_index* ix) { return ix ? ix->n_rows : 0; }
uint32_t cqs_b200_dim(const cqs_b200_index* ix) { return ix ? ix->dim : 0; }
uint32_t cqs_b200_max_k(const cqs_b200_index*) { return kMaxK; }
int cqs_b200_is_poisoned(const cqs_b200_index* ix) { return ix ? ix->poisoned.load() : 0; }
int cqs_b200_scores_are_cosine(const cqs_b200_index* ix) try {
  return ix && ix->storage != CQS_B200_STORAGE_BF16 && ix->metric == CQS_B200_METRIC_COSINE;
} API_CATCH
const char* cqs_b200_name(void) { return "B200"; }
const char* cqs_b200_last_error(void) { return t_last_error.c_str(); }
uint64_t cqs_b200_kernel_launches(void) { return g_kernel_launches.load(); }
float cqs_b200_last_kernel_ms(cqs_b200_index* ix) { return ix ? ix->last_kernel_ms : 0.f; }
int cqs_b200_set_timing(cqs_b200_index* ix, int enable) try {
  if (!ix) return fail(CQS_B200_ERR_INVALID, "index is NULL");
  std::lock_guard<std::mutex> g(ix->mu);
  ix->timing = enable != 0;
  return CQS_B200_OK;
} API_CATCH
// development aids, not declared in the public header
uint32_t cqs_b200_debug_batch_reruns(cqs_b200_index* ix) { return ix ? ix->batch_reruns_total : 0; }
uint32_t cqs_b200_debug_last_batch_reruns(cqs_b200_index* ix) { return ix ? ix->last_batch_reruns : 0; }
uint64_t cqs_b200_debug_shadow_reruns(cqs_b200_index* ix) { return ix ? ix->shadow_reruns : 0; }
int cqs_b200_debug_batch_flags(cqs_b200_index* ix, uint32_t* out, uint32_t n) try {
  if (!ix || ix->shards.empty() || !ix->shards[0].d_bflags) return CQS_B200_ERR_INVALID;
  cudaSetDevice(ix->shards[0].device);
  return cudaMemcpy(out, ix->shards[0].d_bflags, 4 * n, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : CQS_B200_ERR_CUDA;
} API_CATCH
float cqs_b200_debug_last_batch_ms(cqs_b200_index* ix) { return ix ? ix->last_batch_ms : 0.f; }
float cqs_b200_debug_max_row_delta(cqs_b200_index* ix) { return ix && !ix->shards.empty() ? ix->shards[0].max_row_delta : -1.f; }
float cqs_b200_debug_max_row_norm(cqs_b200_index* ix) { return ix && !ix->shards.empty() ? ix->shards[0].max_row_norm : -1.f; }
// copies the trace stamps of shard 0
int cqs_b200_debug_trace(cqs_b200_index* ix, unsigned long long* out, uint32_t n_words) try {
  if (!ix || ix->shards.empty() || !ix->shards[0].d_trace) return CQS_B200_ERR_INVALID;
  cudaSetDevice(ix->shards[0].device);
  return cudaMemcpy(out, ix->shards[0].d_trace, sizeof(unsigned long long) * n_words,
                    cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : CQS_B200_ERR_CUDA;
} API_CATCH

}  // extern "C"
