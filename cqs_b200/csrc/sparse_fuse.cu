// sparse_fuse.cu — kernel 4 of the north star: SPLADE sparse·query scoring with
// a fused top-k pool, the dense+sparse alpha fusion, and centroid routing.
//
//   sparse_search_kernel   <- SpladeIndex::search_with_filter
//                             (src/splade/index.rs:223-291)
//   fuse_pools_kernel      <- the fusion block of search_hybrid_inner
//                             (src/search/query.rs:914-1005)
//   centroid_*_kernel      <- CentroidClassifier::classify
//                             (src/search/router.rs:1415-1444)
//
// Determinism / parity: all three reproduce the reference's f32 operation
// order exactly (separate multiply and add, __fmul_rn/__fadd_rn/__fdiv_rn so
// the compiler cannot contract to FMA), so results are bit-identical to a plain
// f32 restatement of the reference — no atomics on floating-point data anywhere.
//
// Sparse layout in HBM: token-major postings (CSC).  For token t the entries
// [tptr[t], tptr[t+1]) hold (doc, weight) sorted by doc ascending — the order
// SpladeIndex::build produces (index.rs:197-203).  A query touches only
// sum_t |postings(t)| * 8 bytes, versus the whole 8*nnz bytes of a doc-major
// scan.  Docs are processed in blocks of 256 owned by one warp (accumulators in
// shared memory); a first pass streams the doc ids of the touched lists once to
// find where every block starts in every query token's list, the second applies
// the tokens IN QUERY ORDER (the reference's accumulation order,
// index.rs:251-259).  Candidates go through the same shared-memory top-k
// accumulator as the dense scan; the last CTA merges.
#include "common.cuh"
#include "internal.h"

namespace cqs {

constexpr int kSpThreads = 256;          // 8 warps; light enough to share an SM with the dense scan CTA
constexpr int kSpWarps = kSpThreads / 32;
constexpr uint32_t kSpCap = 4096;        // top-k accumulator slots
constexpr uint32_t kSpMaxQ = 1024;       // max query nnz
constexpr uint32_t kSpBlock = kSparseDocsPerBlock;  // docs owned by one warp at a time (256)

// ---- pass 1: where does every 256-doc block start inside every query token's list? ----
// bounds[i][j] = number of postings of query token i with doc < j*256 (j = 0..n_blocks).
// Found by streaming the doc ids of the touched lists once (no dependent binary-search
// chains): the thread that sees the first posting of a block writes the offsets of that
// block and of the empty blocks before it.  bounds is zero-filled beforehand (empty lists).
struct BoundsParams {
  const uint64_t* tptr;
  const uint32_t* doc;
  uint32_t vocab;
  const uint32_t* q_tok;
  uint32_t q_nnz;
  uint32_t n_blocks;
  uint32_t* bounds;  // [q_nnz][n_blocks + 1]
};
constexpr uint32_t kBoundsChunk = 2048;  // postings per CTA work unit
__global__ void __launch_bounds__(256) sparse_bounds_kernel(const BoundsParams p) {
  __shared__ uint64_t s_base[kSpMaxQ];
  __shared__ uint64_t s_prefix[kSpMaxQ + 1];  // chunks before token i
  const uint32_t tid = threadIdx.x;
  for (uint32_t i = tid; i < p.q_nnz; i += blockDim.x) {
    const uint32_t t = __ldg(p.q_tok + i);
    uint64_t b0 = 0, b1 = 0;
    if (t < p.vocab) { b0 = __ldg(p.tptr + t); b1 = __ldg(p.tptr + t + 1); }
    s_base[i] = b0;
    s_prefix[i + 1] = (b1 - b0 + kBoundsChunk - 1) / kBoundsChunk;  // chunk count, scanned below
  }
  __syncthreads();
  if (tid == 0) {
    s_prefix[0] = 0;
    for (uint32_t i = 0; i < p.q_nnz; ++i) s_prefix[i + 1] += s_prefix[i];
  }
  __syncthreads();
  const uint64_t n_chunks = s_prefix[p.q_nnz];
  for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    uint32_t lo = 0, hi = p.q_nnz;  // token owning chunk c: last i with prefix[i] <= c
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (s_prefix[mid] <= c) lo = mid; else hi = mid;
    }
    const uint32_t i = lo, t = __ldg(p.q_tok + i);
    const uint64_t base = s_base[i];
    const uint32_t len = (uint32_t)(__ldg(p.tptr + t + 1) - base);
    const uint32_t e0 = (uint32_t)(c - s_prefix[i]) * kBoundsChunk;
    uint32_t* row = p.bounds + (size_t)i * (p.n_blocks + 1);
    for (uint32_t e = e0 + tid; e < min(len, e0 + kBoundsChunk); e += blockDim.x) {
      const uint32_t blk = __ldg(p.doc + base + e) / kSpBlock;
      const int prev = (e == 0) ? -1 : (int)(__ldg(p.doc + base + e - 1) / kSpBlock);
      for (int j = prev + 1; j <= (int)blk; ++j) row[j] = e;       // first posting at or after block j
      if (e == len - 1)
        for (uint32_t j = blk + 1; j <= p.n_blocks; ++j) row[j] = len;  // blocks after the last posting
    }
  }
}

struct SparseParams {
  const uint64_t* tptr;
  const uint32_t* doc;
  const float* w;
  uint32_t vocab;
  uint64_t n_docs;
  const uint32_t* q_tok;
  const float* q_w;
  uint32_t q_nnz;
  const uint32_t* bounds;
  uint32_t n_blocks;
  const uint32_t* bitset;
  uint32_t k;
  uint64_t row_base;
  ckey_t* partial;
  uint32_t* partial_cnt;
  uint32_t* done;
  float* out_scores;
  uint64_t* out_rows;
  uint32_t* out_n;
};

constexpr uint32_t kSpStage = 1024;      // postings staged per warp between apply phases

struct SpSmem {
  ckey_t buf[kSpCap];                    // 32 KB
  uint2 stage[kSpWarps][kSpStage];       // 64 KB  (local doc, weight bits)
  float acc[kSpWarps][kSpBlock];         //  8 KB
  uint8_t touched[kSpWarps][kSpBlock];   //  2 KB
  uint32_t off[kSpWarps][36];            //  staging offsets of the tokens of the current group
  uint64_t base[kSpMaxQ];                //  8 KB  start of query token i's posting list
  float qw[kSpMaxQ];                     //  4 KB
  uint32_t pos[kMaxGrid];                //  4 KB
  ckey_t thr;
  uint32_t cnt;
  uint32_t last;
};

// ---- pass 2: accumulate + select -------------------------------------------------------
// A warp owns a 256-doc block: its accumulators sit in shared memory.  Query tokens are
// taken in groups whose slices (the part of each token's posting list that falls in the
// block) fit a 1024-entry staging buffer: the group's postings are fetched with many
// independent loads in flight, then applied IN QUERY ORDER (the reference's accumulation
// order, index.rs:251-259): within one token the docs are distinct, so the lanes update
//   acc[doc] = acc[doc] + qw*dw   (separate f32 multiply and add)
// without conflicts, and only a __syncwarp separates tokens.  Touched docs that pass the
// filter go to the CTA's top-k accumulator.
__global__ void __launch_bounds__(kSpThreads) sparse_search_kernel(const SparseParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SpSmem& s = *reinterpret_cast<SpSmem*>(smem_raw);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TopK tk{s.buf, &s.cnt, &s.thr, kSpCap, Group{tid, kSpThreads, 0}};
  tk.init();
  for (uint32_t i = tid; i < p.q_nnz; i += kSpThreads) {
    const uint32_t t = __ldg(p.q_tok + i);
    s.base[i] = (t < p.vocab) ? __ldg(p.tptr + t) : 0;
    s.qw[i] = __ldg(p.q_w + i);
  }
  __syncthreads();
  const uint32_t k = p.k;
  const uint32_t stride = p.n_blocks + 1;
  const uint32_t n_steps = (p.n_blocks + kSpWarps - 1) / kSpWarps;
  float* acc = s.acc[warp];
  uint8_t* touched = s.touched[warp];
  uint2* stage = s.stage[warp];
  uint32_t* off = s.off[warp];
  for (uint32_t step = blockIdx.x; step < n_steps; step += gridDim.x) {
    const uint32_t blk = step * kSpWarps + warp;
    const ckey_t thr = s.thr;
    if (blk < p.n_blocks) {
      const uint32_t d0 = blk * kSpBlock;
#pragma unroll
      for (uint32_t i = lane; i < kSpBlock; i += 32) {
        acc[i] = 0.f;
        touched[i] = 0;
      }
      for (uint32_t i0 = 0; i0 < p.q_nnz; i0 += 32) {
        // slice [lo, hi) of 32 query tokens at once (one token per lane)
        uint32_t lo = 0, hi = 0;
        if (i0 + lane < p.q_nnz) {
          const uint32_t* row = p.bounds + (size_t)(i0 + lane) * stride + blk;
          lo = __ldg(row);
          hi = __ldg(row + 1);
        }
        const uint32_t len = hi - lo;           // <= 256 (a doc lists a token once)
        uint32_t incl = len;                    // inclusive warp scan of the slice lengths
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= (uint32_t)o) incl += v;
        }
        uint32_t done_tok = 0;                  // tokens of this 32-group already applied
        uint32_t consumed = 0;                  // postings of this 32-group already applied
        const uint32_t n_tok = min(32u, p.q_nnz - i0);
        while (done_tok < n_tok) {
          // group = maximal run of tokens from done_tok whose postings fit the staging buffer
          const uint32_t fits = __ballot_sync(0xffffffffu, lane >= done_tok && lane < n_tok &&
                                                               incl - consumed <= kSpStage);
          const uint32_t g1 = done_tok + __popc(fits);        // tokens [done_tok, g1)
          __syncwarp();
          if (lane >= done_tok && lane < g1) off[lane - done_tok + 1] = incl - consumed;
          if (lane == 0) off[0] = 0;
          __syncwarp();
          const uint32_t g_n = g1 - done_tok;
          const uint32_t total = off[g_n];
          // fetch: flat posting index -> (token, offset); 8 postings per lane in flight
          for (uint32_t f0 = 0; f0 < total; f0 += 32 * 8) {
            uint32_t dd[8];
            float ww[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint32_t f = f0 + u * 32 + lane;
              dd[u] = 0xFFFFFFFFu;
              ww[u] = 0.f;
              if (f < total) {
                uint32_t a = 0, b = g_n;        // token t with off[t] <= f < off[t+1]
                while (b - a > 1) {
                  const uint32_t m = (a + b) >> 1;
                  if (off[m] <= f) a = m; else b = m;
                }
                dd[u] = a;                      // token index inside the group, resolved after the loop
                ww[u] = __uint_as_float(f - off[a]);
              }
            }
            // resolve addresses (needs the slice start of the owning token, held by another lane)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint32_t f = f0 + u * 32 + lane;
              const uint32_t tsel = (f < total) ? dd[u] : 0;
              const uint32_t tok_lo = __shfl_sync(0xffffffffu, lo, done_tok + tsel);
              if (f < total) {
                const uint64_t e = s.base[i0 + done_tok + tsel] + tok_lo + __float_as_uint(ww[u]);
                dd[u] = __ldg(p.doc + e) - d0;
                ww[u] = __ldg(p.w + e);
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint32_t f = f0 + u * 32 + lane;
              if (f < total) stage[f] = make_uint2(dd[u], __float_as_uint(ww[u]));
            }
          }
          __syncwarp();
          // apply the group's tokens in query order
          for (uint32_t t = 0; t < g_n; ++t) {
            const float qw = s.qw[i0 + done_tok + t];
            const uint32_t e1 = off[t + 1];
            for (uint32_t e = off[t] + lane; e < e1; e += 32) {
              const uint2 ent = stage[e];
              // *scores.entry(idx).or_insert(0.0) += query_weight * doc_weight   (index.rs:259)
              acc[ent.x] = __fadd_rn(acc[ent.x], __fmul_rn(qw, __uint_as_float(ent.y)));
              touched[ent.x] = 1;
            }
            __syncwarp();
          }
          consumed += total;
          done_tok = g1;
        }
      }
      // candidates: touched docs that pass the filter, finite score (candidate.rs:275)
      for (uint32_t d = lane; d < kSpBlock; d += 32) {
        const uint64_t r = (uint64_t)d0 + d;
        if (!touched[d] || r >= p.n_docs) continue;
        if (p.bitset && !((__ldg(p.bitset + (r >> 5)) >> (r & 31)) & 1u)) continue;
        const float sc = acc[d];
        if (!finite_bits(__float_as_uint(sc))) continue;
        const ckey_t key = make_key(sc, (uint32_t)r);
        if (key > thr) tk.push(key);
      }
    }
    __syncthreads();
    // at most kSpWarps*kSpBlock = 2048 pushes per step
    if (s.cnt + kSpWarps * kSpBlock > kSpCap || (s.thr == 0 && s.cnt >= k)) tk.compact(k);
    __syncthreads();
  }
  tk.compact(k);
  const uint32_t mycnt = s.cnt;
  for (uint32_t i = tid; i < mycnt; i += kSpThreads)
    p.partial[(size_t)blockIdx.x * kMaxK + i] = s.buf[i];
  if (tid == 0) p.partial_cnt[blockIdx.x] = mycnt;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    uint32_t ticket = atomicAdd(p.done, 1u);
    s.last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s.last) return;
  __threadfence();
  merge_partials_and_emit(tk, s.pos, k, p.partial, p.partial_cnt, gridDim.x, p.row_base,
                          p.out_scores, p.out_rows, p.out_n);
  if (tid == 0) *p.done = 0;
}

size_t sparse_bounds_bytes(uint64_t n_docs, uint32_t q_nnz) {
  const uint64_t n_blocks = (n_docs + kSpBlock - 1) / kSpBlock;
  return (size_t)q_nnz * (n_blocks + 1) * sizeof(uint32_t);
}

cudaError_t launch_sparse_search(const SparseArgs& a, cudaStream_t st) {
  if (a.n_docs == 0 || a.k == 0 || a.k > kMaxK || a.q_nnz == 0 || a.q_nnz > kSpMaxQ ||
      a.n_docs > 0xFFFFFFFFull || !a.d_bounds)
    return cudaErrorInvalidValue;
  const uint32_t n_blocks = (uint32_t)((a.n_docs + kSpBlock - 1) / kSpBlock);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaMemsetAsync(a.d_bounds, 0, sparse_bounds_bytes(a.n_docs, a.q_nnz), st);
  if (e != cudaSuccess) return e;
  BoundsParams bp{a.sp.d_tptr, a.sp.d_doc, a.sp.vocab, a.d_q_tok, a.q_nnz, n_blocks, a.d_bounds};
  sparse_bounds_kernel<<<sms * 4, 256, 0, st>>>(bp);
  SparseParams p;
  p.tptr = a.sp.d_tptr; p.doc = a.sp.d_doc; p.w = a.sp.d_w; p.vocab = a.sp.vocab;
  p.n_docs = a.n_docs; p.q_tok = a.d_q_tok; p.q_w = a.d_q_w; p.q_nnz = a.q_nnz;
  p.bounds = a.d_bounds; p.n_blocks = n_blocks;
  p.bitset = a.d_bitset; p.k = a.k; p.row_base = a.row_base;
  p.partial = a.d_partial; p.partial_cnt = a.d_partial_cnt; p.done = a.d_done;
  p.out_scores = a.d_out_scores; p.out_rows = a.d_out_rows; p.out_n = a.d_out_n;
  const uint32_t n_steps = (n_blocks + kSpWarps - 1) / kSpWarps;
  int grid = (int)(n_steps < (uint32_t)(2 * sms) ? n_steps : (uint32_t)(2 * sms));
  if (grid > (int)kMaxGrid) grid = kMaxGrid;
  e = cudaFuncSetAttribute(sparse_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)sizeof(SpSmem));
  if (e != cudaSuccess) return e;
  sparse_search_kernel<<<grid, kSpThreads, sizeof(SpSmem), st>>>(p);
  g_kernel_launches.fetch_add(2, std::memory_order_relaxed);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// alpha fusion (src/search/query.rs:914-1005), one CTA, pools <= 1024 each.
// ---------------------------------------------------------------------------
constexpr uint32_t kFuseMax = kMaxK;       // per pool
constexpr uint32_t kFuseN = 2 * kFuseMax;  // union upper bound (power of two)

struct FuseSmem {
  uint64_t drow[kFuseMax];
  uint64_t srow[kFuseMax];
  float dsc[kFuseMax];
  float ssc[kFuseMax];
  uint64_t key[kFuseN];   // (1<<32 | ordered fused) or 0 when empty
  uint64_t row[kFuseN];
  uint32_t src[kFuseN];   // payload: dense idx | sparse idx<<12 | flags<<24
  float red[32];
  float max_sparse;
};

__device__ __forceinline__ float f32_max_rust(float a, float b) {
  // f32::max: if one operand is NaN the other is returned
  return fmaxf(a, b);
}

__global__ void __launch_bounds__(1024, 1) fuse_pools_kernel(const FuseArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FuseSmem& s = *reinterpret_cast<FuseSmem*>(smem_raw);
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  const uint32_t nd = min(*a.d_n_dense, kFuseMax), ns = min(*a.d_n_sparse, kFuseMax);
  for (uint32_t i = tid; i < nd; i += T) { s.drow[i] = a.d_dense_rows[i]; s.dsc[i] = a.d_dense_scores[i]; }
  for (uint32_t i = tid; i < ns; i += T) { s.srow[i] = a.d_sparse_rows[i]; s.ssc[i] = a.d_sparse_scores[i]; }
  for (uint32_t i = tid; i < kFuseN; i += T) { s.key[i] = 0; s.row[i] = ~0ull; s.src[i] = 0; }
  __syncthreads();
  // max_sparse = iter().map(score).reduce(f32::max).unwrap_or(0.0)   (query.rs:914-919)
  // (max is associative/commutative up to the sign of zero; NaN-ignoring)
  if (tid < 32) {
    float m = __uint_as_float(0x7FC00000u);  // NaN = identity of a NaN-ignoring max
    for (uint32_t i = tid; i < ns; i += 32) m = f32_max_rust(m, s.ssc[i]);
    for (int off = 16; off > 0; off >>= 1) m = f32_max_rust(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (tid == 0) s.max_sparse = (ns == 0) ? 0.f : m;
  }
  __syncthreads();
  const float max_sparse = s.max_sparse;
  const float alpha = a.alpha;
  // one thread per union slot: slots [0,nd) dense entries, [nd, nd+ns) sparse-only entries
  for (uint32_t i = tid; i < nd + ns; i += T) {
    float d = 0.f, sraw = 0.f, sn = 0.f;
    uint64_t row;
    uint32_t flags = 0;
    bool emit = true;
    if (i < nd) {
      row = s.drow[i];
      // HashMap::insert: a later duplicate overwrites; emit only the last occurrence
      for (uint32_t j = i + 1; j < nd; ++j) if (s.drow[j] == row) { emit = false; break; }
      d = s.dsc[i];
      flags = 1;
      for (uint32_t j = 0; j < ns; ++j) if (s.srow[j] == row) { sraw = s.ssc[j]; flags = 3; }
    } else {
      uint32_t j0 = i - nd;
      row = s.srow[j0];
      for (uint32_t j = j0 + 1; j < ns; ++j) if (s.srow[j] == row) { emit = false; break; }
      for (uint32_t j = 0; j < nd && emit; ++j) if (s.drow[j] == row) emit = false;
      sraw = s.ssc[j0];
      flags = 2;
    }
    if (!emit) continue;
    if (flags & 2) sn = (max_sparse > 0.f) ? __fdiv_rn(sraw, max_sparse) : 0.f;
    float fused;
    if (alpha <= 0.f)
      fused = __fadd_rn(d, __fmul_rn(sn, 0.1f));
    else
      fused = __fadd_rn(__fmul_rn(alpha, d), __fmul_rn(__fsub_rn(1.0f, alpha), sn));
    s.key[i] = (1ull << 32) | (uint64_t)ordered_u32(__float_as_uint(fused));
    s.row[i] = row;
    s.src[i] = i;
    // stash per-leg values in the (now consumed) score arrays' slots via registers:
    // we re-derive them at output time from src, so keep flags in the top bits.
    s.src[i] |= flags << 24;
  }
  __syncthreads();
  // sort (fused desc by total order, row asc)   (query.rs:1004)
  const uint32_t P = next_pow2(max(nd + ns, 1u));
  for (uint32_t size = 2; size <= P; size <<= 1)
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = tid; i < (P >> 1); i += T) {
        uint32_t pos = 2 * i - (i & (stride - 1));
        uint64_t ka = s.key[pos], kb = s.key[pos + stride];
        uint64_t ra = s.row[pos], rb = s.row[pos + stride];
        bool b_first = kb > ka || (kb == ka && rb < ra);
        bool dir = ((pos & size) == 0);
        if (b_first == dir) {
          uint32_t sa = s.src[pos], sb = s.src[pos + stride];
          s.key[pos] = kb; s.row[pos] = rb; s.src[pos] = sb;
          s.key[pos + stride] = ka; s.row[pos + stride] = ra; s.src[pos + stride] = sa;
        }
      }
      __syncthreads();
    }
  __shared__ uint32_t s_n;
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (uint32_t i = tid; i < a.pool_k && i < P; i += T) {
    if (s.key[i] == 0) continue;
    uint32_t src = s.src[i] & 0xFFFFFFu, flags = s.src[i] >> 24;
    uint64_t row = s.row[i];
    float d = 0.f, sraw = 0.f;
    if (src < nd) {
      d = s.dsc[src];
      if (flags & 2) for (uint32_t j = 0; j < ns; ++j) if (s.srow[j] == row) sraw = s.ssc[j];
    } else {
      sraw = s.ssc[src - nd];
    }
    a.d_out_rows[i] = row;
    a.d_out_fused[i] = __uint_as_float(unordered_u32((uint32_t)s.key[i]));
    a.d_out_dense[i] = d;
    a.d_out_sparse_raw[i] = sraw;
    a.d_out_present[i] = (uint8_t)flags;
    atomicAdd(&s_n, 1u);
  }
  __syncthreads();
  if (tid == 0) *a.d_out_n = s_n;
}

cudaError_t launch_fuse_pools(const FuseArgs& a, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(fuse_pools_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(FuseSmem));
  if (e != cudaSuccess) return e;
  fuse_pools_kernel<<<1, 1024, sizeof(FuseSmem), st>>>(a);
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// centroid routing (src/search/router.rs:1415-1444)
// ---------------------------------------------------------------------------
// score[q][c] = sequential f32 sum_i e_i * c_i (Iterator::sum over a.zip(b).map(a*b)).
__global__ void centroid_scores_kernel(const float* __restrict__ cen, uint32_t n_c, uint32_t dim,
                                       const float* __restrict__ qs, uint32_t nq,
                                       float* __restrict__ scores) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq * n_c) return;
  uint32_t q = i / n_c, c = i - q * n_c;
  const float* e = qs + (size_t)q * dim;
  const float* ce = cen + (size_t)c * dim;
  float acc = 0.f;
  for (uint32_t j = 0; j < dim; ++j) acc = __fadd_rn(acc, __fmul_rn(__ldg(e + j), __ldg(ce + j)));
  scores[i] = acc;
}
__global__ void centroid_pick_kernel(const float* __restrict__ scores, uint32_t n_c, uint32_t nq,
                                     float threshold, int32_t* __restrict__ out_cat,
                                     float* __restrict__ out_margin) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  float best = __uint_as_float(0xFF800000u), second = best;
  int32_t bc = -1;
  for (uint32_t c = 0; c < n_c; ++c) {
    float sc = scores[(size_t)q * n_c + c];
    if (sc > best) { second = best; best = sc; bc = (int32_t)c; }
    else if (sc > second) { second = sc; }
  }
  float margin = __fsub_rn(best, second);
  out_margin[q] = margin;
  out_cat[q] = (margin >= threshold) ? bc : -1;
}

cudaError_t launch_route_centroids(const float* d_centroids, uint32_t n_c, uint32_t dim,
                                   const float* d_queries, uint32_t nq, float threshold,
                                   int32_t* d_out_cat, float* d_out_margin, cudaStream_t st) {
  if (nq == 0) return cudaSuccess;
  float* d_scores = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&d_scores, sizeof(float) * (size_t)nq * n_c, st);
  if (e != cudaSuccess) return e;
  uint32_t total = nq * n_c;
  centroid_scores_kernel<<<(total + 127) / 128, 128, 0, st>>>(d_centroids, n_c, dim, d_queries, nq,
                                                              d_scores);
  centroid_pick_kernel<<<(nq + 127) / 128, 128, 0, st>>>(d_scores, n_c, nq, threshold, d_out_cat,
                                                         d_out_margin);
  g_kernel_launches.fetch_add(2, std::memory_order_relaxed);
  e = cudaGetLastError();
  cudaFreeAsync(d_scores, st);
  return e;
}

}  // namespace cqs
