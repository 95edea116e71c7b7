/* cqs_oracle.c — CPU restatement of the cqs retrieval hot path in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker and the timed CPU baseline
 * ("port"): nothing under cqs_b200/ links or loads it.  The reference
 * (jamie8johnson/cqs v1.51.0, Rust) cannot be compiled in this image (no
 * cargo/rustc) and its f32 dot lives in the un-vendored crate simsimd 6.5.16
 * (Cargo.lock:4049-4052), so bit-level dense-score parity is UNPINNED; the
 * checker value is the reference's documented scalar fallback, the
 * f64-accumulated dot rounded to f32 (src/math.rs:17-22).  The timed baseline
 * uses an f32 SIMD dot (what simsimd's AVX2/AVX-512 kernels do) so that the CPU
 * number is as favourable as the reference's real path.
 *
 * Each function cites the reference file:line it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define CLONES __attribute__((target_clones("avx512f", "avx2,fma", "default")))

/* ---- synthetic generator: examples/exp_level_scale.rs:200-224 ------------- */
void oracle_synth_vectors(uint64_t n, uint32_t dim, uint64_t* state_io, float* out) {
  uint64_t s = *state_io;
  for (uint64_t r = 0; r < n; ++r) {
    float* v = out + r * dim;
    for (uint32_t i = 0; i < dim; ++i) {
      s ^= s << 13;
      s ^= s >> 7;
      s ^= s << 17;
      float u = (float)(s >> 11) / (float)(1ull << 53); /* (state >> 11) as f32 / 2^53 as f32 */
      v[i] = u * 2.0f - 1.0f;
    }
    float acc = 0.0f; /* sequential f32 sum of x*x */
    for (uint32_t i = 0; i < dim; ++i) {
      float sq = v[i] * v[i];
      acc = acc + sq;
    }
    float norm = sqrtf(acc);
    if (norm > 0.0f)
      for (uint32_t i = 0; i < dim; ++i) v[i] = v[i] / norm;
  }
  *state_io = s;
}

/* ---- a1: math::cosine_similarity, src/math.rs:11-28 ----------------------- */
/* checker: f64 accumulation (the reference's scalar fallback) */
CLONES float oracle_dot_f64(const float* a, const float* b, uint32_t n) {
  double acc = 0.0;
  for (uint32_t i = 0; i < n; ++i) acc += (double)a[i] * (double)b[i];
  return (float)acc;
}
/* timed baseline: f32 SIMD accumulation, 4 x 16-lane partial sums kept in
 * registers (GCC vector extension, lowered per target clone) */
typedef float v16f __attribute__((vector_size(64), aligned(4), may_alias));
CLONES float oracle_dot_f32_simd(const float* a, const float* b, uint32_t n) {
  v16f a0 = {0}, a1 = {0}, a2 = {0}, a3 = {0};
  uint32_t i = 0;
  for (; i + 64 <= n; i += 64) {
    a0 += *(const v16f*)(a + i) * *(const v16f*)(b + i);
    a1 += *(const v16f*)(a + i + 16) * *(const v16f*)(b + i + 16);
    a2 += *(const v16f*)(a + i + 32) * *(const v16f*)(b + i + 32);
    a3 += *(const v16f*)(a + i + 48) * *(const v16f*)(b + i + 48);
  }
  v16f t = (a0 + a1) + (a2 + a3);
  float s = 0.0f;
  for (int j = 0; j < 16; ++j) s += t[j];
  for (; i < n; ++i) s += a[i] * b[i];
  return s;
}
void oracle_dense_scores(const float* rows, uint64_t n, uint32_t dim, const float* q, float* out,
                         int use_f64) {
  for (uint64_t r = 0; r < n; ++r)
    out[r] = use_f64 ? oracle_dot_f64(rows + r * dim, q, dim) : oracle_dot_f32_simd(rows + r * dim, q, dim);
}

/* ---- a4: BoundedScoreHeap, src/search/scoring/candidate.rs:162-330 -------- */
/* ids are row numbers in ascending chunk-id order, so id order == row order. */
typedef struct {
  float score;
  uint64_t row;
} heap_item;

static inline int32_t total_key(float f) { /* f32::total_cmp as a signed compare */
  int32_t b;
  memcpy(&b, &f, 4);
  return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1);
}
/* 1 if a ranks WORSE than b under (score desc, id asc): lower score, or equal score and larger id */
static inline int worse(const heap_item* a, const heap_item* b) {
  int32_t ka = total_key(a->score), kb = total_key(b->score);
  return ka < kb || (ka == kb && a->row > b->row);
}
typedef struct {
  heap_item* items; /* binary heap, root = worst */
  uint32_t len, cap;
} bheap;
static void sift_down(bheap* h, uint32_t i) {
  for (;;) {
    uint32_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < h->len && worse(&h->items[l], &h->items[m])) m = l;
    if (r < h->len && worse(&h->items[r], &h->items[m])) m = r;
    if (m == i) return;
    heap_item t = h->items[i];
    h->items[i] = h->items[m];
    h->items[m] = t;
    i = m;
  }
}
static void sift_up(bheap* h, uint32_t i) {
  while (i) {
    uint32_t p = (i - 1) / 2;
    if (!worse(&h->items[i], &h->items[p])) return;
    heap_item t = h->items[i];
    h->items[i] = h->items[p];
    h->items[p] = t;
    i = p;
  }
}
static void heap_push(bheap* h, uint64_t row, float score) { /* candidate.rs:274-318 */
  if (!isfinite(score)) return;
  heap_item it = {score, row};
  if (h->len < h->cap) {
    h->items[h->len] = it;
    sift_up(h, h->len++);
    return;
  }
  if (h->cap == 0) return;
  if (worse(&h->items[0], &it)) { /* score > worst, or equal and id < worst_id */
    h->items[0] = it;
    sift_down(h, 0);
  }
}
static int cmp_final(const void* a, const void* b) { /* into_sorted_vec, candidate.rs:321-329 */
  const heap_item *x = a, *y = b;
  if (worse(y, x)) return -1;
  if (worse(x, y)) return 1;
  return 0;
}

/* ---- a3: the brute-force row loop, src/search/query.rs:453-484 ------------ */
int oracle_brute_force(const float* rows, uint64_t n, uint32_t dim, const float* query, uint32_t k,
                       const uint32_t* bitset, int use_f64, uint64_t* out_rows, float* out_scores,
                       uint32_t* out_n) {
  *out_n = 0;
  if (k == 0 || n == 0) return 0;
  for (uint32_t i = 0; i < dim; ++i)
    if (!isfinite(query[i])) return 0; /* src/cagra.rs:458-470 */
  bheap h = {malloc(sizeof(heap_item) * k), 0, k};
  if (!h.items) return -1;
  for (uint64_t r = 0; r < n; ++r) {
    if (bitset && !((bitset[r >> 5] >> (r & 31)) & 1u)) continue;
    float s = use_f64 ? oracle_dot_f64(rows + r * dim, query, dim)
                      : oracle_dot_f32_simd(rows + r * dim, query, dim);
    heap_push(&h, r, s);
  }
  qsort(h.items, h.len, sizeof(heap_item), cmp_final);
  for (uint32_t i = 0; i < h.len; ++i) {
    out_rows[i] = h.items[i].row;
    out_scores[i] = h.items[i].score;
  }
  *out_n = h.len;
  free(h.items);
  return 0;
}

/* nq queries; threads > 1 spreads QUERIES over cores with pthreads (each query
 * stays on one thread, as in the reference: src/search/query.rs:469). */
typedef struct {
  const float* rows; uint64_t n; uint32_t dim; const float* queries; uint32_t nq, k; int use_f64;
  uint64_t* out_rows; float* out_scores; uint32_t* out_n;
  volatile uint32_t* next; int rc;
} bf_job;
static void* bf_worker(void* arg) {
  bf_job* j = (bf_job*)arg;
  for (;;) {
    uint32_t q = __atomic_fetch_add(j->next, 1u, __ATOMIC_RELAXED);
    if (q >= j->nq) break;
    int r = oracle_brute_force(j->rows, j->n, j->dim, j->queries + (uint64_t)q * j->dim, j->k, NULL,
                               j->use_f64, j->out_rows + (uint64_t)q * j->k,
                               j->out_scores + (uint64_t)q * j->k, j->out_n + q);
    if (r) j->rc = r;
  }
  return NULL;
}
int oracle_brute_force_batch(const float* rows, uint64_t n, uint32_t dim, const float* queries,
                             uint32_t nq, uint32_t k, int use_f64, int threads, uint64_t* out_rows,
                             float* out_scores, uint32_t* out_n) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  volatile uint32_t next = 0;
  bf_job jobs[256];
  pthread_t tid[256];
  for (int t = 0; t < threads; ++t) {
    bf_job j = {rows, n, dim, queries, nq, k, use_f64, out_rows, out_scores, out_n, &next, 0};
    jobs[t] = j;
  }
  for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, bf_worker, &jobs[t]);
  bf_worker(&jobs[0]);
  int rc = jobs[0].rc;
  for (int t = 1; t < threads; ++t) {
    pthread_join(tid[t], NULL);
    if (jobs[t].rc) rc = jobs[t].rc;
  }
  return rc;
}

int oracle_num_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n < 1 ? 1 : (int)n;
}

/* ---- a10: SpladeIndex::search_with_filter, src/splade/index.rs:223-291 ----
 * postings: token-major (tptr[vocab+1], doc[], w[]) in build (doc) order.
 * Dense accumulator + touched flags replace the HashMap (same arithmetic:
 * per chunk, in query-token order, score += qw * dw with separate f32 ops). */
int oracle_sparse_search(const uint64_t* tptr, const uint32_t* doc, const float* w, uint32_t vocab,
                         uint64_t n_docs, const uint32_t* q_tok, const float* q_w, uint32_t q_nnz,
                         uint32_t k, const uint32_t* bitset, uint64_t* out_rows, float* out_scores,
                         uint32_t* out_n) {
  *out_n = 0;
  if (q_nnz == 0 || n_docs == 0) return 0;
  float* acc = calloc(n_docs, sizeof(float));
  uint8_t* touched = calloc(n_docs, 1);
  bheap h = {malloc(sizeof(heap_item) * (k ? k : 1)), 0, k};
  if (!acc || !touched || !h.items) return -1;
  for (uint32_t i = 0; i < q_nnz; ++i) {
    uint32_t t = q_tok[i];
    if (t >= vocab) continue;
    volatile float qw = q_w[i];
    for (uint64_t e = tptr[t]; e < tptr[t + 1]; ++e) {
      uint32_t d = doc[e];
      if (d >= n_docs) continue;
      if (bitset && !((bitset[d >> 5] >> (d & 31)) & 1u)) continue;
      volatile float prod = qw * w[e]; /* volatile: forbid FMA contraction */
      acc[d] = acc[d] + prod;
      touched[d] = 1;
    }
  }
  for (uint64_t d = 0; d < n_docs; ++d)
    if (touched[d]) heap_push(&h, d, acc[d]);
  qsort(h.items, h.len, sizeof(heap_item), cmp_final);
  for (uint32_t i = 0; i < h.len; ++i) {
    out_rows[i] = h.items[i].row;
    out_scores[i] = h.items[i].score;
  }
  *out_n = h.len;
  free(acc);
  free(touched);
  free(h.items);
  return 0;
}
