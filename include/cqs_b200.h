/* cqs_b200.h — C ABI of the B200-native exact retrieval backend for cqs.
 *
 * This is the drop-in boundary: the only surface the reference's Rust side
 * binds (INTEGRATION.md shows the `extern "C"` block and the
 * `impl VectorIndex for B200Index` shim).  Every entry point cites the
 * reference interface it replaces (paths relative to jamie8johnson/cqs
 * v1.51.0).
 *
 * Conventions (mirroring the existing FFI backend, src/cagra.rs):
 *   - return 0 on success, a negative CQS_B200_ERR_* otherwise; never throws,
 *     never aborts.  `cqs_b200_last_error()` gives a thread-local message.
 *   - inputs are borrowed for the duration of the call; outputs are
 *     caller-allocated (capacity k, or nq*k for batches).
 *   - the device side speaks ROW INDICES only; chunk-id strings stay with the
 *     caller as `id_map[row]` (src/cagra.rs:268).  Rows must be appended in
 *     ascending chunk-id order so that the tie-break "(score desc, row asc)"
 *     equals the reference's "(score desc, id asc)"
 *     (src/search/scoring/candidate.rs:303-329).
 *   - any CUDA failure sets a sticky poison bit (`cqs_b200_is_poisoned`),
 *     the analogue of src/cagra.rs:276,472-489 / src/index.rs:203-205.  A failed
 *     cross-shard exchange (a peer rank that never answered) poisons the index that
 *     ran the search as well, so the daemon's is_poisoned() -> rebuild hook
 *     (src/cli/batch/view.rs:737-767) covers the sharded path too.
 *   - an index handle is internally serialised by one mutex
 *     (src/cagra.rs:263) and may be shared across threads.
 *   - score of a row = f32 dot(query, row) (== cosine for the unit-norm
 *     embeddings cqs stores; this is exactly math::cosine_similarity,
 *     src/math.rs:11-28).  Rows whose score is NaN/Inf are dropped
 *     (src/math.rs:23-27 -> None; candidate.rs:275).
 */
#ifndef CQS_B200_H
#define CQS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CQS_B200_OK 0
#define CQS_B200_ERR_INVALID (-1)   /* bad argument / wrong state            */
#define CQS_B200_ERR_CUDA (-2)      /* CUDA runtime failure (poisons index)   */
#define CQS_B200_ERR_POISONED (-3)  /* index already poisoned; rebuild it     */
#define CQS_B200_ERR_OOM (-4)       /* device or host allocation failed       */
#define CQS_B200_ERR_UNSUPPORTED (-5)

#define CQS_B200_METRIC_COSINE 0    /* DistanceMetric::Cosine  src/index.rs:45 */
#define CQS_B200_METRIC_DOT 1       /* DistanceMetric::DotProduct             */
#define CQS_B200_STORAGE_F32 0      /* rows kept as f32 (BLOB layout, src/store/helpers/embeddings.rs:14-41) */
#define CQS_B200_STORAGE_BF16 1     /* rows rounded (RNE) to bf16; the rounded matrix IS the corpus */
#define CQS_B200_STORAGE_BF16_F32 2 /* f32 master rows (every result is exact f32, as STORAGE_F32) plus a
                                       bf16 shadow copy that feeds the candidate scans: single queries
                                       stream the shadow (2 B/elem) and over-fetch k' candidates, batches
                                       run the tensor-core scan over it; the candidates are re-scored on
                                       the f32 rows and the pool is PROVEN complete from measured rounding
                                       distances — or the query is re-run on the f32 rows.  6 B/elem. */
/* Bit 31 of an out_n word written by the DEVICE-RESIDENT entry points (cqs_b200_search_device,
 * _search_sharded_device, _search_many_device) on a STORAGE_BF16_F32 index: the candidate pool of
 * that query could not be proven complete; the caller must repeat it through a host entry point
 * (which does this re-run itself and never returns the bit).  n = out_n & 0x7FFFFFFF. */
#define CQS_B200_UNPROVEN 0x80000000u

#define CQS_B200_MAX_K 1024u        /* VectorIndex::max_k()  src/index.rs:217  */

typedef struct cqs_b200_index cqs_b200_index; /* opaque */

/* ---- lifecycle ----------------------------------------------------------
 * Replaces CagraIndex::build_from_store_with_metric + CagraBackend::try_open
 * (src/cagra.rs:842-919, :1664-1803): the "build" of an exact-scan backend is
 * streaming the f32 BLOBs into HBM.
 * device_ids/n_dev: shards the rows in contiguous blocks over n_dev GPUs of
 * this process (n_dev > 1 requires cqs_b200_reserve before the first append). */
int cqs_b200_create(const int* device_ids, int n_dev, uint32_t dim, int metric, int storage,
                    cqs_b200_index** out);
/* Pre-size for n_rows total rows.  n_dev == 1: a capacity hint only (avoids regrowth; the
 * index still grows past it, so cqs_b200_reopen + append works on built and loaded indexes).
 * n_dev > 1: mandatory, and it fixes the shard boundaries — appends beyond
 * n_dev * ceil32(n_rows / n_dev) rows fail with CQS_B200_ERR_INVALID (rebuild with a larger
 * reserve to re-shard). */
int cqs_b200_reserve(cqs_b200_index* ix, uint64_t n_rows);
/* Global row number of local row 0.  Used when one process holds one shard of a
 * corpus that is row-sharded across processes; reported rows
 * are row_base + local row. */
int cqs_b200_set_row_base(cqs_b200_index* ix, uint64_t row_base);
/* Host rows, row-major f32[n_rows][dim] — the store.embedding_batches feed
 * (src/cagra.rs:888-916).  Copied; caller keeps ownership. */
int cqs_b200_append_rows_f32(cqs_b200_index* ix, const float* rows, uint64_t n_rows);
/* Same, but `d_rows` is a device pointer on the index's (single) device. */
int cqs_b200_append_rows_f32_device(cqs_b200_index* ix, const float* d_rows, uint64_t n_rows);
/* Seal the index: after this, searches are allowed and appends are rejected
 * until cqs_b200_reopen (the TieredIndex::extend analogue, src/tiered.rs:317-360).
 * cqs_b200_reopen drops everything aligned row-by-row with the matrix (sparse postings,
 * row meta, row signals): attach them again after the next finalize. */
int cqs_b200_finalize(cqs_b200_index* ix);
int cqs_b200_reopen(cqs_b200_index* ix);
void cqs_b200_destroy(cqs_b200_index* ix); /* syncs all streams first (src/cagra.rs:289-301) */

/* ---- persistence: the flat-matrix analogue of CagraIndex::save / load ---------------------
 * (src/cagra.rs:963-1652: blob + sidecar with magic, dim, chunk_count, checksum; atomic
 * temp + rename).  The file holds a 64-byte header (magic "CQSB2001", version, dim, storage,
 * metric, n_rows, payload checksum) and the master rows un-padded in their storage dtype; the
 * chunk-id sidecar stays with the caller exactly like CagraMeta's id_map.  Load verifies magic,
 * sizes and checksum and returns CQS_B200_ERR_INVALID on any mismatch (the caller then
 * rebuilds from the store, as src/cagra.rs:1676-1802 does). */
int cqs_b200_save(cqs_b200_index* ix, const char* path);
int cqs_b200_load(const char* path, const int* device_ids, int n_dev, cqs_b200_index** out);

/* ---- VectorIndex::search / search_with_filter  (src/index.rs:146,167) ------
 * query: f32[dim].  bitset: NULL or ceil(len/32) words, bit (i%32) of word
 * (i/32) set = row i passes (src/cagra.rs:747-757).
 * Output sorted (score desc by f32 total order, row asc); *out_n <= k.
 * k == 0, empty index or a non-finite query -> *out_n = 0, returns 0
 * (src/cagra.rs:445-470).  k > CQS_B200_MAX_K -> CQS_B200_ERR_INVALID. */
int cqs_b200_search(cqs_b200_index* ix, const float* query, uint32_t k, const uint32_t* bitset,
                    uint64_t* out_rows, float* out_scores, uint32_t* out_n);

/* ---- Store::search_filtered on the device (src/search/query.rs:316-510) ----------------
 * The brute-force callers (gather.rs:619, scout.rs:244, where_to_add.rs:167, onboard.rs:175,
 * the worktree overlay) pass no index; their semantics are: SQL type/language filter, then
 * per row  base = clamp(cos,0,1) -> max(base,0)*note_boost -> *importance (if demotion) ->
 * keep iff score >= threshold  (apply_scoring_pipeline, candidate.rs:420-562, without the
 * string-matching signals NameBlend/GlobGate, which SearchFilter::default() leaves off),
 * all BEFORE the bounded heap (query.rs:469-481).
 * Per-row inputs are uploaded once: chunk_type/language codes (u8, the caller's own
 * enumeration) and the two multipliers (note boost = 1 + sentiment*0.15 from the note
 * index, importance = chunk_importance(): 0.70 test / 0.80 private / 1.0).  NULL arrays
 * mean "no filter" / "multiplier 1.0". */
int cqs_b200_set_row_meta(cqs_b200_index* ix, const uint8_t* chunk_type, const uint8_t* lang,
                          uint64_t n_rows);
int cqs_b200_set_row_signals(cqs_b200_index* ix, const float* note_boost, const float* importance,
                             uint64_t n_rows);
/* type_mask / lang_mask: 256-bit sets (4 x u64), bit c set = code c passes; NULL = all pass.
 * Output: top-`limit` (score desc, row asc) of the rows that pass filter and threshold, with
 * the FOLDED score (what search_filtered puts in its heap). */
int cqs_b200_search_filtered(cqs_b200_index* ix, const float* query, uint32_t limit,
                             float threshold, const uint64_t* type_mask, const uint64_t* lang_mask,
                             int enable_demotion, uint64_t* out_rows, float* out_scores,
                             uint32_t* out_n);

/* VectorIndex::search_with_filter for the predicate search_hybrid_inner / the CLI build from
 * include_types / languages (src/search/query.rs:867-878): the same exact scan with the
 * type/language test done on the device from the uploaded row meta — no per-query host bitset
 * (12.5 MB at 100M rows).  Raw cosine scores.  NULL masks = all pass. */
int cqs_b200_search_typed(cqs_b200_index* ix, const float* query, uint32_t k,
                          const uint64_t* type_mask, const uint64_t* lang_mask, uint64_t* out_rows,
                          float* out_scores, uint32_t* out_n);

/* ---- rrf_fuse_n  (src/search/scoring/fusion.rs:36-68) -------------------------------------
 * ids: the ranked lists concatenated (row ids), list_len[n_lists]; k = the RRF constant (60,
 * CQS_RRF_K).  score[id] = sum over lists in list order of 1/(k + rank + 1) with rank the
 * 0-based first occurrence in that list; output top-`limit` (score desc, id asc). */
int cqs_b200_rrf_fuse(int device, const uint64_t* ids, const uint32_t* list_len, uint32_t n_lists,
                      float k, uint32_t limit, uint64_t* out_ids, float* out_scores, uint32_t* out_n);

/* nq independent queries (no trait counterpart — exposed as an inherent
 * B200Index::search_batch).  queries: f32[nq][dim]; outputs
 * [nq][k] with out_n[nq].  With STORAGE_BF16 / STORAGE_BF16_F32 and nq >= 8 this
 * runs the tcgen05 tile scan over bf16-rounded queries, over-fetches k' >= 64
 * candidates per query, re-scores them with the f32 queries (on the f32 master rows
 * when present) and falls back to the exact single-query scan for any query whose
 * candidate pool cannot be proven complete; results are identical to nq calls of
 * cqs_b200_search.  Other cases loop over cqs_b200_search. */
int cqs_b200_search_batch(cqs_b200_index* ix, const float* queries, uint32_t nq, uint32_t k,
                          const uint32_t* bitset, uint64_t* out_rows, float* out_scores,
                          uint32_t* out_n);

/* ---- SPLADE leg: SpladeIndex::build / search_with_filter -------------------
 * (src/splade/index.rs:191-211, :223-291).  Doc-major CSR over the SAME rows
 * (row i of the CSR is chunk i): indptr[len+1], tok/w[indptr[len]]. Copied. */
int cqs_b200_sparse_attach(cqs_b200_index* ix, const uint64_t* indptr, const uint32_t* tok,
                           const float* w, uint32_t vocab);
/* Same with the doc-major CSR already on the device (indptr[n_rows] must equal nnz). */
int cqs_b200_sparse_attach_device(cqs_b200_index* ix, const uint64_t* d_indptr, const uint32_t* d_tok,
                                  const float* d_w, uint64_t nnz, uint32_t vocab);
/* SpladeIndex::save / ::load (src/splade/index.rs:308-560): persist the built posting lists
 * with the store's `splade_generation` counter (src/splade/index.rs:13-16) in the header.
 * load fails (-> caller rebuilds from `sparse_vectors`, src/store/sparse.rs:342) when the file
 * is missing or damaged, its generation differs from `expected_generation`, or its chunk
 * count differs from cqs_b200_len(ix).  Atomic write (temp file + rename). */
int cqs_b200_sparse_save(cqs_b200_index* ix, const char* path, uint64_t generation);
int cqs_b200_sparse_load(cqs_b200_index* ix, const char* path, uint64_t expected_generation);
/* Sparse-only search: score[c] = sum over query tokens IN QUERY ORDER of
 * qw*dw (separate f32 mul and add), only chunks touched by >= 1 posting. */
int cqs_b200_search_sparse(cqs_b200_index* ix, const uint32_t* q_tok, const float* q_w,
                           uint32_t q_nnz, uint32_t k, const uint32_t* bitset, uint64_t* out_rows,
                           float* out_scores, uint32_t* out_n);

/* ---- hybrid: the two leg calls + alpha fusion of search_hybrid_inner -------
 * (src/search/query.rs:880-1005).  pool_k = cap_k_to_backend(candidate_count).
 * Output: the fused pool, sorted (fused desc, row asc), truncated to pool_k,
 * with the per-leg values the SearchLegs inspector reports
 * (src/search/query.rs:39-208): out_dense = raw cosine (0.0 if absent),
 * out_sparse_raw = raw SPLADE dot (0.0 if absent), out_present bit0 = in dense
 * pool, bit1 = in sparse pool.  Output capacity: pool_k each. */
int cqs_b200_search_hybrid(cqs_b200_index* ix, const float* query, const uint32_t* q_tok,
                           const float* q_w, uint32_t q_nnz, float alpha, uint32_t pool_k,
                           const uint32_t* bitset, uint64_t* out_rows, float* out_fused,
                           float* out_dense, float* out_sparse_raw, uint8_t* out_present,
                           uint32_t* out_n);

/* The fusion step alone on caller-supplied pools (what search_hybrid_inner does
 * after the two leg calls, src/search/query.rs:914-1005).  Pools are in leg
 * order (dense first in the union).  Runs on `device`.  Output capacity:
 * min(pool_k, n_dense + n_sparse). */
int cqs_b200_fuse_pools(int device, const uint64_t* dense_rows, const float* dense_scores,
                        uint32_t n_dense, const uint64_t* sparse_rows, const float* sparse_scores,
                        uint32_t n_sparse, float alpha, uint32_t pool_k, uint64_t* out_rows,
                        float* out_fused, float* out_dense, float* out_sparse_raw,
                        uint8_t* out_present, uint32_t* out_n);

/* ---- CentroidClassifier::classify  (src/search/router.rs:1415-1444) --------
 * centroids f32[n_c][dim], queries f32[nq][dim] (host).  out_cat[q] = index of
 * the best centroid if best-second >= threshold else -1; out_margin[q]. */
int cqs_b200_route_centroids(int device, const float* centroids, uint32_t n_c, uint32_t dim,
                             const float* queries, uint32_t nq, float threshold, int32_t* out_cat,
                             float* out_margin);

/* ---- row-sharded corpora across processes -------------------
 * Local top-k left ON THE DEVICE so a collective can follow without a host
 * round trip.  d_query f32[dim] (device), d_bitset nullable (device, local
 * rows).  d_out_scores f32[k], d_out_rows u64[k] (GLOBAL rows), d_out_n u32[1];
 * unused slots are filled with (-inf, UINT64_MAX).  `stream` is a
 * cudaStream_t (0 = the index's own stream).  Asynchronous. */
int cqs_b200_search_device(cqs_b200_index* ix, const float* d_query, uint32_t k,
                           const uint32_t* d_bitset, float* d_out_scores, uint64_t* d_out_rows,
                           uint32_t* d_out_n, void* stream);
/* Merge n_lists gathered candidate lists of k slots each (the all-gather
 * result) into the global top-k, same ordering rule.  All pointers device.
 * Batched: n_queries independent merges laid out [list][query][k]. */
int cqs_b200_merge_topk_device(int device, const float* d_scores, const uint64_t* d_rows,
                               uint32_t n_lists, uint32_t n_queries, uint32_t k,
                               float* d_out_scores, uint64_t* d_out_rows, uint32_t* d_out_n,
                               void* stream);

/* ---- row-sharded corpora WITHOUT a collective call: NVLink peer memory --------
 * (the reference has no multi-GPU path — src/index.rs:139-239 is one
 * index per process.)  A peer group is one mailbox per rank in that rank's HBM,
 * mapped into every other rank.  The scan kernel's last CTA stores the shard's
 * sorted top-k straight into every peer's mailbox over NVLink, raises one
 * release flag per peer, waits for the peers' lists and merges — all-gather and
 * merge ride in the tail of the kernel that produced the list.  Every rank gets the
 * GLOBAL top-k, identical to the unsharded result.
 *
 * Rules: every rank issues the same sequence of sharded searches on its group (the
 * exchanges are numbered); a rank that does not show up within the timeout (default
 * 5 s) makes the others return an error and marks the group failed (rebuild it).
 *
 * One process per GPU:   peer_create -> peer_handle -> exchange the 64-byte handles
 *                        through any side channel (the tests use torch.distributed
 *                        all_gather_object) -> peer_connect(all handles, rank order).
 * One process, n GPUs:   peer_create x n -> peer_connect_local(array, n). */
typedef struct cqs_b200_peer cqs_b200_peer; /* opaque */
#define CQS_B200_PEER_HANDLE_BYTES 64u
#define CQS_B200_PEER_MAX_WORLD 8u
/* max_elems: capacity of one exchange in (query, slot) pairs, nq*k <= max_elems
 * (0 = 65536: 1024 queries x top-64). */
int cqs_b200_peer_create(int device, uint32_t world, uint32_t rank, uint32_t max_elems,
                         cqs_b200_peer** out);
int cqs_b200_peer_handle(cqs_b200_peer* p, uint8_t* out_handle /* [64] */);
int cqs_b200_peer_connect(cqs_b200_peer* p, const uint8_t* handles /* [world][64], rank order */);
int cqs_b200_peer_connect_local(cqs_b200_peer** peers, uint32_t world);
int cqs_b200_peer_set_timeout_ms(cqs_b200_peer* p, uint32_t ms);
/* 0 = healthy, 1 = an exchange timed out (sticky), < 0 = error.  Synchronises the device. */
int cqs_b200_peer_status(cqs_b200_peer* p);
void cqs_b200_peer_destroy(cqs_b200_peer* p);
/* VectorIndex::search over the whole row-sharded corpus, host query in, GLOBAL host
 * top-k out on every rank (ONE kernel launch per rank: scan + exchange + merge; the
 * result lands in host-mapped memory).  `bitset` covers this rank's rows. */
int cqs_b200_search_sharded(cqs_b200_index* ix, cqs_b200_peer* peer, const float* query, uint32_t k,
                            const uint32_t* bitset, uint64_t* out_rows, float* out_scores,
                            uint32_t* out_n);
/* Same with device buffers, asynchronous on `stream` (0 = the index's own stream). */
int cqs_b200_search_sharded_device(cqs_b200_index* ix, cqs_b200_peer* peer, const float* d_query,
                                   uint32_t k, const uint32_t* d_bitset, float* d_out_scores,
                                   uint64_t* d_out_rows, uint32_t* d_out_n, void* stream);
/* nq exact single-query scans with everything on the device: one launch per query,
 * issued round-robin on four launch lanes (`stream` and three internal streams, each with its
 * own scratch set) so the tail of one launch (list merge, cross-shard exchange) overlaps the
 * streaming phase of the next three; `stream` is joined with the internal lanes before the call returns
 * (asynchronous).  d_queries f32 [nq][dim]; outputs [nq][k] / [nq].  peer == NULL: this
 * index alone; otherwise the GLOBAL top-k over the sharded corpus on every rank. */
int cqs_b200_search_many_device(cqs_b200_index* ix, cqs_b200_peer* peer, const float* d_queries,
                                uint32_t nq, uint32_t k, const uint32_t* d_bitset,
                                float* d_out_scores, uint64_t* d_out_rows, uint32_t* d_out_n,
                                void* stream);
/* cqs_b200_search_batch over the row-sharded corpus: per-shard tensor-core scan +
 * exact rescoring, then ONE gather+merge kernel over peer memory per <= 1024 queries. */
int cqs_b200_search_batch_sharded(cqs_b200_index* ix, cqs_b200_peer* peer, const float* queries,
                                  uint32_t nq, uint32_t k, const uint32_t* bitset,
                                  uint64_t* out_rows, float* out_scores, uint32_t* out_n);
/* cqs_b200_search_hybrid over the row-sharded corpus (each rank attached the SPLADE rows of
 * ITS chunks): GLOBAL dense pool from the scan's own exchange, GLOBAL sparse pool from one
 * gather+merge kernel, the same fusion on every rank (max_sparse = the merged pool's top-1).
 * Identical to the unsharded cqs_b200_search_hybrid. */
int cqs_b200_search_hybrid_sharded(cqs_b200_index* ix, cqs_b200_peer* peer, const float* query,
                                   const uint32_t* q_tok, const float* q_w, uint32_t q_nnz,
                                   float alpha, uint32_t pool_k, const uint32_t* bitset,
                                   uint64_t* out_rows, float* out_fused, float* out_dense,
                                   float* out_sparse_raw, uint8_t* out_present, uint32_t* out_n);
/* The exchange alone: this rank's nq sorted lists (device, [nq][k], GLOBAL rows) ->
 * GLOBAL top-k of every query on every rank.  nq <= 1024, nq*k <= max_elems. */
int cqs_b200_peer_gather_merge(cqs_b200_peer* p, const float* d_scores, const uint64_t* d_rows,
                               const uint32_t* d_n, uint32_t nq, uint32_t k, float* d_out_scores,
                               uint64_t* d_out_rows, uint32_t* d_out_n, void* stream);

/* ---- introspection: VectorIndex::{len,dim,max_k,is_poisoned,name} ----------
 * (src/index.rs:149-217) */
uint64_t cqs_b200_len(const cqs_b200_index* ix);
uint32_t cqs_b200_dim(const cqs_b200_index* ix);
uint32_t cqs_b200_max_k(const cqs_b200_index* ix);
int cqs_b200_is_poisoned(const cqs_b200_index* ix);
/* 1 iff returned scores are the brute-force cosine itself (f32 storage),
 * VectorIndex::index_scores_are_cosine, src/index.rs:236. */
int cqs_b200_scores_are_cosine(const cqs_b200_index* ix);
const char* cqs_b200_name(void);        /* "B200" */
const char* cqs_b200_last_error(void);  /* thread-local */
/* Number of CUDA kernels this library has launched in this process. */
uint64_t cqs_b200_kernel_launches(void);
/* Device time (ms) of the dominant kernel of the most recent search call on
 * this index, from CUDA events recorded on the launching stream.  The events are only
 * recorded after cqs_b200_set_timing(ix, 1) (off by default: they cost a few microseconds
 * on the single-query latency path). */
float cqs_b200_last_kernel_ms(cqs_b200_index* ix);
int cqs_b200_set_timing(cqs_b200_index* ix, int enable);

#ifdef __cplusplus
}
#endif
#endif /* CQS_B200_H */
