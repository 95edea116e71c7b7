// internal.h — host-side declarations shared by the translation units of
// libcqs_b200.so.  Nothing here is part of the C ABI (include/cqs_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace cqs {

typedef unsigned long long ckey_t;
struct PeerCtx;  // peer.cuh

constexpr uint32_t kMaxK = 1024;        // CQS_B200_MAX_K
constexpr uint32_t kMaxGrid = 1024;     // upper bound on scan CTAs (partial-list slots)
// Launch lanes: single-query scans issued round-robin over this many streams / scratch sets,
// so the tail of launch i (list merge, cross-shard exchange — slow while the next scans keep
// HBM saturated) is hidden behind launches i+1 .. i+kLanes-1.  Launch i only waits for i-kLanes.
constexpr uint32_t kLanes = 4;
// Bit 31 of an out_n word: the answer came from the bf16 shadow scan of a STORAGE_BF16_F32 index
// and its re-scored candidate pool could NOT be proven complete — the host entry points re-run such
// a query on the f32 master rows before returning; device-resident entry points hand the bit to the
// caller (CQS_B200_UNPROVEN in include/cqs_b200.h).
constexpr uint32_t kUnprovenBit = 0x80000000u;

extern std::atomic<uint64_t> g_kernel_launches;

// Row layout in HBM: row-major, `ld` elements per row (dim zero-padded up to
// ld), 16-byte aligned rows.  mode: 0 = f32 (float4 per lane), 1 = bf16 with
// 16-byte lane vectors (8 elems), 2 = bf16 with 8-byte lane vectors (4 elems).
// nv = lane vectors per row per lane: ld == 32 * elems_per_vec * nv.
struct RowLayout {
  uint32_t ld;
  int mode;
  int nv;
};
// Returns false when dim is unsupported (0 or > 2048).
bool choose_layout(uint32_t dim, int storage, RowLayout* out);

// Optional structured filter + per-row score signals: the string-free part of
// Store::search_filtered (src/search/query.rs:348-510, candidate.rs:420-562) on device.
struct ScanSignals {
  const uint8_t* d_ctype = nullptr;   // per-row chunk-type code (nullable = no type filter)
  const uint8_t* d_lang = nullptr;    // per-row language code
  uint64_t type_mask[4] = {~0ull, ~0ull, ~0ull, ~0ull};  // bit c set = code c passes
  uint64_t lang_mask[4] = {~0ull, ~0ull, ~0ull, ~0ull};
  const float* d_note_boost = nullptr;  // per-row 1 + sentiment*0.15 (nullable = 1.0)
  const float* d_importance = nullptr;  // per-row 0.70 / 0.80 / 1.0 (nullable or demotion off = skipped)
  float max_note_boost = 1.f;           // upper bounds used to skip rows that cannot qualify
  float max_importance = 1.f;
  float threshold = 0.f;                // ThresholdGate: keep score >= threshold
  int pipeline = 0;                     // 1 = clamp / note boost / demotion / threshold fold
};

struct ScanArgs {
  const void* d_rows;       // [n_rows][ld] f32 or bf16
  uint64_t n_rows;          // <= 2^32 - 1
  RowLayout layout;
  const float* d_query;     // f32[ld], zero padded
  const uint32_t* d_bitset; // nullable, ceil(n_rows/32) words
  uint32_t k;               // 1..kMaxK
  uint64_t row_base;        // added to local rows in the output
  ckey_t* d_partial;         // [kMaxGrid][kMaxK] scratch
  uint32_t* d_partial_cnt;  // [kMaxGrid]
  uint32_t* d_done;         // [2]: CTA ticket + tile counter, zero between launches
  ckey_t* d_col;            // [2 * kMaxGrid]: shared-threshold scratch of the large-k path (zero between launches)
                            // + per-CTA exclusion bounds of the shadow scan
  float* d_out_scores;      // [k]
  uint64_t* d_out_rows;     // [k]
  uint32_t* d_out_n;        // [1]
  void* d_trace = nullptr;  // optional [kMaxGrid][8] u64 timestamps (CQS_B200_TRACE=1)
  const ScanSignals* signals = nullptr;
  uint32_t* d_host_flag = nullptr;  // device alias of a host-mapped completion word (nullable)
  uint32_t seq = 0;                 // value written to it
  const PeerCtx* peer = nullptr;    // row-sharded corpus: exchange + merge in the kernel tail; out_* = GLOBAL top-k
  // STORAGE_BF16_F32 fast path: d_rows is the bf16 shadow, k the over-fetched candidate count k';
  // the kernel tail re-scores the k' candidates on the f32 master rows (the f32 scan's arithmetic:
  // bit-identical scores), keeps the best k_out and proves the pool complete or raises kUnprovenBit.
  const void* d_exact_rows = nullptr;
  int exact_nv = 0;                 // lane vectors per row of the f32 master layout
  uint32_t k_out = 0;
  float max_row_delta = 0.f;        // measured max |f32 row - bf16 row|_2
  float max_row_norm = 0.f;         // bound on |bf16 row|_2
};
// Over-fetch of the shadow scan for a final k (0 = k too large: scan the f32 master instead).
uint32_t shadow_kprime(uint32_t k);
// Kernel 1+3: single-query streaming scan with the top-k select fused in.
cudaError_t launch_scan_single(const ScanArgs& a, int num_sms, cudaStream_t stream);

// f32 -> bf16 (RNE) / f32 row conversion into the padded layout.
cudaError_t launch_convert_rows(const float* d_src, uint32_t dim, uint64_t n_rows, void* d_dst,
                                uint64_t dst_row0, RowLayout layout, cudaStream_t stream);

struct MergeArgs {
  const float* d_scores;   // [n_lists][n_queries][k]
  const uint64_t* d_rows;  // same shape, global rows; UINT64_MAX = empty slot
  uint32_t n_lists, n_queries, k;
  float* d_out_scores;     // [n_queries][k]
  uint64_t* d_out_rows;
  uint32_t* d_out_n;       // [n_queries]
};
cudaError_t launch_merge_topk(const MergeArgs& a, cudaStream_t stream);

// ---- batched tensor-core scan (kernel 2) ----
constexpr uint32_t kBatchMaxQ = 1024;        // queries per launch_scan_batch call
constexpr uint32_t kBatchCap = 20480;        // candidate slots per query pool
constexpr uint32_t kBatchDenseRows = 18944;  // round 0 (74 chunks of 256 rows)
struct BatchArgs {
  const void* d_rows;        // bf16 [n_rows][ld]: what the tensor cores scan
  uint64_t n_rows;           // < 2^31
  RowLayout layout;          // mode 1 or 2 (bf16)
  const void* d_exact_rows;  // rows the candidates are re-scored on (== d_rows, or the f32 master)
  RowLayout exact_layout;    // same ld as `layout`
  float max_row_delta;       // max over rows of |exact row - scanned bf16 row|_2 (0 when the bf16 rows ARE the corpus)
  const float* d_queries;    // f32 [nq][ld], zero padded beyond dim
  uint32_t nq;               // <= kBatchMaxQ
  uint32_t k;
  const uint32_t* d_bitset;  // nullable
  uint64_t row_base;
  float max_row_norm;        // upper bound of the L2 norm of every scanned (bf16) row (exactness bound)
  void* d_scratch;           // batch_scratch_bytes(round_up(nq,128), ld)
  float* d_out_scores;       // [nq][k]
  uint64_t* d_out_rows;      // [nq][k]
  uint32_t* d_out_n;         // [nq]
  uint32_t* d_flags;         // [nq]: 1 = not provably exact, caller re-runs the exact scan
};
size_t batch_scratch_bytes(uint32_t nq_pad, uint32_t ld);
uint32_t batch_kprime(uint32_t k, bool shadow);
cudaError_t launch_scan_batch(const BatchArgs& a, int num_sms, cudaStream_t stream);
cudaError_t launch_max_row_norm(const void* d_rows, uint64_t n_rows, RowLayout layout, float* d_out,
                                cudaStream_t stream);
// max over rows of |f32 master row - bf16 shadow row|_2 (same ld): the measured rounding distance
// that makes the batch path's exactness bound rigorous AND tight.
cudaError_t launch_max_row_delta(const void* d_rows_f32, const void* d_rows_bf16, uint64_t n_rows, uint32_t ld,
                                 float* d_out, cudaStream_t stream);

// ---- sparse (SPLADE) ----
struct SparseDev {
  // token-major postings (CSC): for token t, entries [tptr[t], tptr[t+1]) sorted by doc asc
  uint64_t* d_tptr = nullptr;   // [vocab+1]
  uint32_t* d_doc = nullptr;    // [nnz] doc ids alone (bounds pass)
  void* d_post = nullptr;       // [nnz] (doc, weight bits) pairs (accumulate pass)
  uint32_t vocab = 0;
  uint64_t nnz = 0;
  // Static block index (built once at attach / load): for the longest posting lists, where every
  // 64-doc block (kSparseDocsPerBlock) starts inside the list — what the per-query bounds pass would otherwise find by
  // streaming the list's doc ids.  slot_of[token] = row of the table or -1.
  int32_t* d_slot_of = nullptr;     // [vocab]
  uint32_t* d_block_index = nullptr;  // [n_slots][index_stride]
  uint32_t n_slots = 0;
  uint32_t index_stride = 0;        // n_blocks + 1 at build time
};
void free_sparse(SparseDev& sp);
// Builds the static block index for `n_tok` tokens (`d_tokens`, <= 1024 per call) into rows
// [row0, row0 + n_tok) of sp.d_block_index.
cudaError_t launch_sparse_block_index(const SparseDev& sp, const uint32_t* d_tokens, uint32_t n_tok,
                                      uint32_t row0, uint64_t n_docs, cudaStream_t stream);
struct SparseArgs {
  SparseDev sp;
  uint64_t n_docs;
  const uint32_t* d_q_tok;  // [q_nnz] (query order)
  const float* d_q_w;
  uint32_t q_nnz;
  uint32_t* d_bounds;       // scratch, sparse_bounds_bytes(n_docs, q_nnz)
  void* d_block_scratch = nullptr;   // sparse_block_scratch_bytes(n_docs): per-doc scores + touched bits
  uint32_t* d_claim = nullptr;       // [1] block counter of the accumulate kernel, zero between queries
  void* d_trace = nullptr;  // optional [kMaxGrid][8] u64 stamps (CQS_B200_TRACE=1)
  const uint32_t* d_bitset;
  uint32_t k;
  uint64_t row_base;
  ckey_t* d_partial;         // [n_blocks][kMaxK]
  uint32_t* d_partial_cnt;
  uint32_t* d_done;
  float* d_out_scores;
  uint64_t* d_out_rows;
  uint32_t* d_out_n;
};
constexpr uint32_t kSparseDocsPerBlock = 64;    // docs owned by one warp at a time
size_t sparse_bounds_bytes(uint64_t n_docs, uint32_t q_nnz);
size_t sparse_block_scratch_bytes(uint64_t n_docs);
cudaError_t launch_sparse_search(const SparseArgs& a, cudaStream_t stream);
// the leg's launches one by one (sparse_fuse.cu): the hybrid call places them around its dense scan
enum : uint32_t { kSparseBounds = 1, kSparseAccum = 2, kSparseSelect = 8 };
cudaError_t launch_sparse_stage(const SparseArgs& a, uint32_t stage, cudaStream_t stream);

// Inverted-index build on the device (sparse_build.cu): doc-major CSR -> token-major postings.
struct SparseBuildArgs {
  const uint64_t* d_indptr;  // [n_docs+1]
  const uint32_t* d_tok;     // [nnz]
  const float* d_w;          // [nnz]
  uint64_t n_docs, nnz;
  uint32_t vocab;            // <= 56,320 (the histogram lives in shared memory)
  uint64_t* d_tptr;          // out [vocab+1]
  uint32_t* d_doc;           // out [nnz]
  void* d_post;              // out [nnz] (doc, weight bits)
  void* d_scratch;           // sparse_build_scratch_bytes(vocab, num_sms)
  uint32_t* d_err;           // out: 0 ok, 1 token id >= vocab, 2 a doc lists a token twice
};
size_t sparse_build_scratch_bytes(uint32_t vocab, int num_sms);
cudaError_t launch_sparse_build(const SparseBuildArgs& a, int num_sms, cudaStream_t stream);

struct FuseArgs {
  const uint64_t* d_dense_rows; const float* d_dense_scores; const uint32_t* d_n_dense;
  const uint64_t* d_sparse_rows; const float* d_sparse_scores; const uint32_t* d_n_sparse;
  float alpha; uint32_t pool_k;
  uint64_t* d_out_rows; float* d_out_fused; float* d_out_dense; float* d_out_sparse_raw;
  uint8_t* d_out_present; uint32_t* d_out_n;
  uint32_t* d_host_flag = nullptr;  // optional: the outputs are host-mapped; publish `seq` here when they are written
  uint32_t seq = 0;
};
cudaError_t launch_fuse_pools(const FuseArgs& a, cudaStream_t stream);

constexpr uint32_t kRrfMaxEntries = 2048;  // total ids over all lists of one rrf_fuse call
cudaError_t launch_rrf_fuse(const uint64_t* d_ids, const uint32_t* d_list_off, uint32_t n_lists, float k,
                            uint32_t limit, uint64_t* d_out_ids, float* d_out_scores, uint32_t* d_out_n,
                            cudaStream_t stream);

cudaError_t launch_route_centroids(const float* d_centroids, uint32_t n_c, uint32_t dim,
                                   const float* d_queries, uint32_t nq, float threshold,
                                   int32_t* d_out_cat, float* d_out_margin, cudaStream_t stream);

// ---- peer-memory exchange (row-sharded corpora) ----
struct PeerGatherArgs {
  const float* d_scores;   // [nq][k] this rank's sorted lists
  const uint64_t* d_rows;  // [nq][k] GLOBAL rows
  const uint32_t* d_n;     // [nq]
  uint32_t nq, k;
  float* d_out_scores;     // [nq][k] GLOBAL top-k on every rank
  uint64_t* d_out_rows;
  uint32_t* d_out_n;
  uint32_t* d_ticket;      // [2] zero between launches
};
cudaError_t launch_peer_gather_merge(const PeerCtx& c, const PeerGatherArgs& a, int num_sms,
                                     cudaStream_t stream);

}  // namespace cqs
