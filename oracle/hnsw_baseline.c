/* hnsw_baseline.c — CPU restatement of the reference's HNSW search path, for TIMING ONLY.
 *
 * TEST / BENCHMARK INFRASTRUCTURE.  Nothing under cqs_b200/ links or loads it.
 *
 * The reference's default CPU index is `hnsw_rs 0.3.4` + `anndists 0.1.4`
 * (Cargo.lock:1839-1842, :50-53), neither vendored under /root/reference, and there
 * is no Rust toolchain here: this file restates the PUBLISHED algorithm (Malkov &
 * Yashunin, "Efficient and robust approximate nearest neighbor search using
 * Hierarchical Navigable Small World graphs") with the reference's parameters:
 *   - tiers (M, efConstruction, efSearch): <5k (16,100,50), <100k (24,200,100),
 *     >=100k (32,400,200)                               src/hnsw/mod.rs:104-112
 *   - level scale factor 0.5: level = floor(-ln(u) * 0.5 / ln(M)), max 16 layers
 *                                                       src/hnsw/mod.rs:85
 *   - DistCosine on unit vectors = 1 - dot; score = 1 - dist      src/hnsw/search.rs:128
 *   - ef = min(max(efSearch, 2k), n)                               src/hnsw/search.rs:106
 *   - one thread per query; parallel build (hnsw_rs parallel_insert, src/hnsw/build.rs:234)
 * It is approximate and its results are NOT a parity oracle ("parity unpinned");
 * bench.py reports its p50 / queries/s / recall@20 beside the exact scan.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CLONES __attribute__((target_clones("avx512f", "avx2,fma", "default")))
typedef float v16f __attribute__((vector_size(64), aligned(4), may_alias));

CLONES static float dotf(const float* a, const float* b, uint32_t n) {
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

typedef struct {
  float d;
  uint32_t id;
} cand_t;

typedef struct hnsw {
  const float* rows;
  uint32_t n, dim, M, M0, efC;
  double level_mult;
  uint8_t* level;        /* [n] */
  uint32_t* link0;       /* [n][M0 + 1]: count, then neighbours (layer 0) */
  uint32_t** linku;      /* [n] -> [level][M + 1] (upper layers) */
  volatile int* lock;    /* [n] spin locks */
  volatile int glock;
  volatile uint32_t entry;
  volatile int max_level;
} hnsw_t;

static inline float dist(const hnsw_t* h, const float* q, uint32_t id) {
  return 1.0f - dotf(q, h->rows + (size_t)id * h->dim, h->dim);
}
static inline void lock(volatile int* l) {
  while (__atomic_exchange_n(l, 1, __ATOMIC_ACQUIRE)) {
    while (*l) __builtin_ia32_pause();
  }
}
static inline void unlock(volatile int* l) { __atomic_store_n(l, 0, __ATOMIC_RELEASE); }
static inline uint32_t* links(const hnsw_t* h, uint32_t id, int lev) {
  return lev == 0 ? h->link0 + (size_t)id * (h->M0 + 1) : h->linku[id] + (size_t)(lev - 1) * (h->M + 1);
}

/* binary heaps on cand_t: max-heap (worst on top) for results, min-heap for the frontier */
static void heap_push(cand_t* a, uint32_t* n, cand_t x, int maxheap) {
  uint32_t i = (*n)++;
  a[i] = x;
  while (i) {
    uint32_t p = (i - 1) / 2;
    int swap = maxheap ? (a[i].d > a[p].d) : (a[i].d < a[p].d);
    if (!swap) break;
    cand_t t = a[i]; a[i] = a[p]; a[p] = t;
    i = p;
  }
}
static cand_t heap_pop(cand_t* a, uint32_t* n, int maxheap) {
  cand_t top = a[0];
  a[0] = a[--(*n)];
  uint32_t i = 0;
  for (;;) {
    uint32_t l = 2 * i + 1, r = l + 1, m = i;
    if (l < *n && (maxheap ? a[l].d > a[m].d : a[l].d < a[m].d)) m = l;
    if (r < *n && (maxheap ? a[r].d > a[m].d : a[r].d < a[m].d)) m = r;
    if (m == i) break;
    cand_t t = a[i]; a[i] = a[m]; a[m] = t;
    i = m;
  }
  return top;
}

typedef struct {
  uint32_t* visited;  /* [n] epoch marks */
  uint32_t epoch;
  cand_t* frontier;   /* capacity: generous */
  cand_t* result;     /* capacity ef + 1 */
  uint32_t cap;
} scratch_t;

/* search one layer; returns count in s->result (max-heap, unsorted) */
static uint32_t search_layer(const hnsw_t* h, const float* q, uint32_t ep, float ep_d, uint32_t ef,
                             int lev, scratch_t* s, int locked) {
  uint32_t nf = 0, nr = 0;
  if (++s->epoch == 0) { memset(s->visited, 0, sizeof(uint32_t) * h->n); s->epoch = 1; }
  s->visited[ep] = s->epoch;
  cand_t c0 = {ep_d, ep};
  heap_push(s->frontier, &nf, c0, 0);
  heap_push(s->result, &nr, c0, 1);
  uint32_t nb[256];
  while (nf) {
    cand_t c = heap_pop(s->frontier, &nf, 0);
    if (nr >= ef && c.d > s->result[0].d) break;
    uint32_t* l = links(h, c.id, lev);
    uint32_t cnt;
    if (locked) { lock(&h->lock[c.id]); cnt = l[0]; memcpy(nb, l + 1, cnt * 4); unlock(&h->lock[c.id]); }
    else { cnt = l[0]; memcpy(nb, l + 1, cnt * 4); }
    for (uint32_t j = 0; j < cnt; ++j) {
      uint32_t e = nb[j];
      if (s->visited[e] == s->epoch) continue;
      s->visited[e] = s->epoch;
      float d = dist(h, q, e);
      if (nr < ef || d < s->result[0].d) {
        cand_t x = {d, e};
        if (nf < s->cap) heap_push(s->frontier, &nf, x, 0);
        heap_push(s->result, &nr, x, 1);
        if (nr > ef) heap_pop(s->result, &nr, 1);
      }
    }
  }
  return nr;
}

static int cmp_cand(const void* a, const void* b) {
  float x = ((const cand_t*)a)->d, y = ((const cand_t*)b)->d;
  return (x > y) - (x < y);
}

/* heuristic neighbour selection (Algorithm 4): keep a candidate only if it is closer to
 * the base point than to every neighbour kept so far */
static uint32_t select_neighbors(const hnsw_t* h, cand_t* c, uint32_t nc, uint32_t M, uint32_t* out) {
  qsort(c, nc, sizeof(cand_t), cmp_cand);
  uint32_t k = 0;
  for (uint32_t i = 0; i < nc && k < M; ++i) {
    int good = 1;
    const float* ci = h->rows + (size_t)c[i].id * h->dim;
    for (uint32_t j = 0; j < k; ++j)
      if (1.0f - dotf(ci, h->rows + (size_t)out[j] * h->dim, h->dim) < c[i].d) { good = 0; break; }
    if (good) out[k++] = c[i].id;
  }
  return k;
}

static void insert(hnsw_t* h, uint32_t id, scratch_t* s) {
  const float* q = h->rows + (size_t)id * h->dim;
  const int lev = h->level[id];
  lock(&h->glock);
  int maxl = h->max_level;
  uint32_t ep = h->entry;
  if (maxl < 0) {  /* first point */
    h->entry = id;
    h->max_level = lev;
    unlock(&h->glock);
    return;
  }
  const int hold = lev > maxl;  /* keep the global lock while raising the entry point */
  if (!hold) unlock(&h->glock);
  float ep_d = dist(h, q, ep);
  for (int l = maxl; l > lev; --l) {  /* greedy descent */
    int changed = 1;
    while (changed) {
      changed = 0;
      uint32_t nb[256];
      uint32_t* lk = links(h, ep, l);
      lock(&h->lock[ep]);
      uint32_t cnt = lk[0];
      memcpy(nb, lk + 1, cnt * 4);
      unlock(&h->lock[ep]);
      for (uint32_t j = 0; j < cnt; ++j) {
        float d = dist(h, q, nb[j]);
        if (d < ep_d) { ep_d = d; ep = nb[j]; changed = 1; }
      }
    }
  }
  cand_t* tmp = malloc(sizeof(cand_t) * (h->efC + 2 + 256));
  for (int l = lev < maxl ? lev : maxl; l >= 0; --l) {
    uint32_t nr = search_layer(h, q, ep, ep_d, h->efC, l, s, 1);
    memcpy(tmp, s->result, sizeof(cand_t) * nr);
    const uint32_t Ml = l == 0 ? h->M0 : h->M;
    uint32_t sel[256];
    uint32_t ns = select_neighbors(h, tmp, nr, h->M, sel);
    ep = tmp[0].id;  /* closest (tmp is sorted by select_neighbors) */
    ep_d = tmp[0].d;
    lock(&h->lock[id]);
    uint32_t* mine = links(h, id, l);
    mine[0] = ns;
    memcpy(mine + 1, sel, ns * 4);
    unlock(&h->lock[id]);
    for (uint32_t j = 0; j < ns; ++j) {  /* back links, shrink with the heuristic when full */
      uint32_t e = sel[j];
      lock(&h->lock[e]);
      uint32_t* le = links(h, e, l);
      if (le[0] < Ml) {
        le[1 + le[0]++] = id;
      } else {
        cand_t cc[260];
        const float* pe = h->rows + (size_t)e * h->dim;
        uint32_t n2 = 0;
        for (uint32_t t = 0; t < le[0]; ++t) { cc[n2].id = le[1 + t]; cc[n2].d = dist(h, pe, le[1 + t]); ++n2; }
        cc[n2].id = id; cc[n2].d = dist(h, pe, id); ++n2;
        uint32_t keep[256];
        uint32_t nk = select_neighbors(h, cc, n2, Ml, keep);
        le[0] = nk;
        memcpy(le + 1, keep, nk * 4);
      }
      unlock(&h->lock[e]);
    }
  }
  free(tmp);
  if (hold) {
    h->entry = id;
    h->max_level = lev;
    unlock(&h->glock);
  }
}

typedef struct {
  hnsw_t* h;
  volatile uint32_t* next;
} build_job;

static scratch_t* scratch_new(const hnsw_t* h, uint32_t ef) {
  scratch_t* s = calloc(1, sizeof(scratch_t));
  s->visited = calloc(h->n, sizeof(uint32_t));
  s->cap = 16 * ef + 4096;
  s->frontier = malloc(sizeof(cand_t) * (s->cap + 1));
  s->result = malloc(sizeof(cand_t) * (ef + 2));
  return s;
}
static void scratch_free(scratch_t* s) { free(s->visited); free(s->frontier); free(s->result); free(s); }

static void* build_worker(void* arg) {
  build_job* j = arg;
  scratch_t* s = scratch_new(j->h, j->h->efC);
  for (;;) {
    uint32_t id = __atomic_fetch_add(j->next, 1u, __ATOMIC_RELAXED);
    if (id >= j->h->n) break;
    insert(j->h, id, s);
  }
  scratch_free(s);
  return NULL;
}

void hnsw_tier(uint32_t n, uint32_t* M, uint32_t* efC, uint32_t* efS) { /* src/hnsw/mod.rs:104-112 */
  if (n < 5000) { *M = 16; *efC = 100; *efS = 50; }
  else if (n < 100000) { *M = 24; *efC = 200; *efS = 100; }
  else { *M = 32; *efC = 400; *efS = 200; }
}

hnsw_t* hnsw_build(const float* rows, uint32_t n, uint32_t dim, uint32_t M, uint32_t efC, int threads,
                   uint64_t seed) {
  hnsw_t* h = calloc(1, sizeof(hnsw_t));
  h->rows = rows; h->n = n; h->dim = dim; h->M = M; h->M0 = 2 * M; h->efC = efC;
  h->level_mult = 0.5 / log((double)M);  /* LEVEL_SCALE_FACTOR = 0.5 */
  h->level = malloc(n);
  h->link0 = calloc((size_t)n * (h->M0 + 1), 4);
  h->linku = calloc(n, sizeof(uint32_t*));
  h->lock = calloc(n, sizeof(int));
  h->max_level = -1;
  uint64_t st = seed ? seed : 0x9E3779B97F4A7C15ull;
  for (uint32_t i = 0; i < n; ++i) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    double u = ((double)(st >> 11) + 1.0) / 9007199254740993.0;
    int l = (int)floor(-log(u) * h->level_mult);
    if (l > 15) l = 15;
    h->level[i] = (uint8_t)l;
    if (l > 0) h->linku[i] = calloc((size_t)l * (M + 1), 4);
  }
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  volatile uint32_t next = 0;
  build_job job = {h, &next};
  pthread_t tid[256];
  for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, build_worker, &job);
  build_worker(&job);
  for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
  return h;
}

void hnsw_free(hnsw_t* h) {
  if (!h) return;
  for (uint32_t i = 0; i < h->n; ++i) free(h->linku[i]);
  free(h->linku); free(h->link0); free(h->level); free((void*)h->lock); free(h);
}

/* HnswIndex::search_impl (src/hnsw/search.rs:46-150): ef = min(max(efS, 2k), n); score = 1 - dist */
static uint32_t search_one(const hnsw_t* h, const float* q, uint32_t k, uint32_t efS, scratch_t* s,
                           uint32_t* out_ids, float* out_scores) {
  if (h->max_level < 0 || k == 0) return 0;
  uint32_t ef = efS > 2 * k ? efS : 2 * k;
  if (ef > h->n) ef = h->n;
  uint32_t ep = h->entry;
  float ep_d = dist(h, q, ep);
  for (int l = h->max_level; l > 0; --l) {
    int changed = 1;
    while (changed) {
      changed = 0;
      uint32_t* lk = links(h, ep, l);
      for (uint32_t j = 0; j < lk[0]; ++j) {
        float d = dist(h, q, lk[1 + j]);
        if (d < ep_d) { ep_d = d; ep = lk[1 + j]; changed = 1; }
      }
    }
  }
  uint32_t nr = search_layer(h, q, ep, ep_d, ef, 0, s, 0);
  qsort(s->result, nr, sizeof(cand_t), cmp_cand);
  uint32_t m = nr < k ? nr : k;
  for (uint32_t i = 0; i < m; ++i) { out_ids[i] = s->result[i].id; out_scores[i] = 1.0f - s->result[i].d; }
  return m;
}

typedef struct {
  const hnsw_t* h; const float* queries; uint32_t nq, k, efS;
  uint32_t* out_ids; float* out_scores; uint32_t* out_n; double* lat_s;
  volatile uint32_t* next;
} search_job;

#include <time.h>
static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void* search_worker(void* arg) {
  search_job* j = arg;
  uint32_t ef = j->efS > 2 * j->k ? j->efS : 2 * j->k;
  scratch_t* s = scratch_new(j->h, ef);
  for (;;) {
    uint32_t q = __atomic_fetch_add(j->next, 1u, __ATOMIC_RELAXED);
    if (q >= j->nq) break;
    double t0 = now_s();
    j->out_n[q] = search_one(j->h, j->queries + (size_t)q * j->h->dim, j->k, j->efS, s,
                             j->out_ids + (size_t)q * j->k, j->out_scores + (size_t)q * j->k);
    if (j->lat_s) j->lat_s[q] = now_s() - t0;
  }
  scratch_free(s);
  return NULL;
}

/* nq queries, one thread per query (the reference is single-threaded per query) */
void hnsw_search_batch(const hnsw_t* h, const float* queries, uint32_t nq, uint32_t k, uint32_t efS,
                       int threads, uint32_t* out_ids, float* out_scores, uint32_t* out_n,
                       double* lat_s) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  volatile uint32_t next = 0;
  search_job job = {h, queries, nq, k, efS, out_ids, out_scores, out_n, lat_s, &next};
  pthread_t tid[256];
  for (int t = 1; t < threads; ++t) pthread_create(&tid[t], NULL, search_worker, &job);
  search_worker(&job);
  for (int t = 1; t < threads; ++t) pthread_join(tid[t], NULL);
}
