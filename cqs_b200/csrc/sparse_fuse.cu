// sparse_fuse.cu — kernel 4 of the north star: SPLADE sparse·query scoring with
// a fused top-k pool, the dense+sparse alpha fusion, and centroid routing.
//
//   sparse_bounds_kernel, sparse_accum_kernel, sparse_select_kernel
//                          <- SpladeIndex::search_with_filter
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
// [tptr[t], tptr[t+1]) hold (doc, weight) pairs (8-byte AoS, plus a doc-only copy
// for the bounds pass) sorted by doc ascending — the order
// SpladeIndex::build produces (index.rs:197-203).  A query touches only
// sum_t |postings(t)| * 8 bytes, versus the whole 8*nnz bytes of a doc-major
// scan.  Docs are processed in blocks of 256 (four 64-doc index blocks) owned by one warp, accumulators
// in shared memory; a bounds pass (once per index for the long lists, per query for the short ones)
// finds where every 64-doc block starts in every token's list, the accumulate kernel applies
// the tokens IN QUERY ORDER (the reference's accumulation order,
// index.rs:251-259) and writes the blocks' scores to a scratch; the select kernel streams them
// through the same shared-memory top-k accumulator as the dense scan; its last CTA merges.
#include "common.cuh"
#include "internal.h"

namespace cqs {

constexpr int kSpThreads = 512;          // 16 warps, 2 CTAs per SM
constexpr int kSpWarps = kSpThreads / 32;
constexpr uint32_t kSpCap = 4096;        // top-k accumulator slots
constexpr uint32_t kSpMaxQ = 1024;       // max query nnz
constexpr uint32_t kSpBlock = kSparseDocsPerBlock;  // docs owned by one warp at a time (64)

// ---- pass 1: where does every 64-doc block start inside every query token's list? ----
// bounds[i][j] = number of postings of query token i with doc < j*64 (j = 0..n_blocks).
// Found by streaming the doc ids of the touched lists once (no dependent binary-search
// chains): the thread that sees the first posting of a block writes the offsets of that
// block and of the empty blocks before it.  Rows of empty lists (a token with no postings, or out
// of the vocabulary) are written as zeros by the same kernel — there is no per-query memset; rows
// of tokens served by the static block index are never read and are left alone.
struct BoundsParams {
  const uint64_t* tptr;
  const uint32_t* doc;
  uint32_t vocab;
  const uint32_t* q_tok;
  uint32_t q_nnz;
  uint32_t n_blocks;
  uint32_t* bounds;  // [q_nnz][n_blocks + 1]
  const int32_t* slot_of;  // nullable: tokens with slot_of[t] >= 0 have a static row and are skipped
  unsigned long long* trace = nullptr;  // optional stamps, rows 300.. of the sparse trace (development aid)
};
constexpr uint32_t kBoundsChunk = 8192;  // postings per streaming work unit
constexpr uint32_t kBoundsShort = 256;   // lists up to this long are binary-searched per block instead
// T threads per CTA (256).
template <int T>
__global__ void __launch_bounds__(T) sparse_bounds_kernel(const BoundsParams p) {
  __shared__ uint64_t s_base[kSpMaxQ];
  __shared__ uint32_t s_len[kSpMaxQ];
  __shared__ uint64_t s_prefix[kSpMaxQ + 1];  // work units before token i
  const uint32_t tid = threadIdx.x;
  if (p.trace && tid == 0 && blockIdx.x < 148) {
    unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    p.trace[(300 + blockIdx.x) * 8] = t_;
  }
  const uint32_t search_units = (p.n_blocks + 1 + 4 * T - 1) / (4 * T);  // 4 T block boundaries per unit
  for (uint32_t i = tid; i < p.q_nnz; i += blockDim.x) {
    const uint32_t t = __ldg(p.q_tok + i);
    uint64_t b0 = 0, b1 = 0;
    const bool covered = t < p.vocab && p.slot_of && __ldg(p.slot_of + t) >= 0;   // static index row exists
    if (t < p.vocab && !covered) {
      b0 = __ldg(p.tptr + t);
      b1 = __ldg(p.tptr + t + 1);
    }
    const uint64_t len = b1 - b0;
    s_base[i] = b0;
    s_len[i] = (uint32_t)len;
    // long lists are streamed (a thread that sees the first posting of a block writes the
    // offsets of that block and of the empty blocks before it); short lists would leave
    // long serial fills to a few threads, so every block boundary is binary-searched — which
    // for an empty list writes the all-zero row the search kernel expects.
    s_prefix[i + 1] = covered ? 0 : (len <= kBoundsShort ? search_units
                                                         : (len + kBoundsChunk - 1) / kBoundsChunk);
  }
  __syncthreads();
  if (tid == 0) {
    s_prefix[0] = 0;
    for (uint32_t i = 0; i < p.q_nnz; ++i) s_prefix[i + 1] += s_prefix[i];
  }
  __syncthreads();
  const uint64_t n_units = s_prefix[p.q_nnz];
  for (uint64_t c = blockIdx.x; c < n_units; c += gridDim.x) {
    uint32_t lo = 0, hi = p.q_nnz;  // token owning unit c: last i with prefix[i] <= c
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (s_prefix[mid] <= c) lo = mid; else hi = mid;
    }
    const uint32_t i = lo;
    const uint64_t base = s_base[i];
    const uint32_t len = s_len[i];
    const uint32_t unit = (uint32_t)(c - s_prefix[i]);
    uint32_t* row = p.bounds + (size_t)i * (p.n_blocks + 1);
    if (len <= kBoundsShort) {
      // four independent binary searches per thread (<= 8 steps each)
      uint32_t a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = 0; b[u] = len; }
      for (int step = 0; step < 9; ++step) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint64_t target = (uint64_t)(unit * 4 * T + u * T + tid) * kSpBlock;
          if (a[u] < b[u]) {
            const uint32_t m = (a[u] + b[u]) >> 1;
            if ((uint64_t)__ldg(p.doc + base + m) < target) a[u] = m + 1; else b[u] = m;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t j = unit * 4 * T + u * T + tid;  // block boundary
        if (j <= p.n_blocks) row[j] = a[u];              // first posting with doc >= j*64
      }
      continue;
    }
    const uint32_t e0 = unit * kBoundsChunk;
    const uint32_t e_end = min(len, e0 + kBoundsChunk);
    const uint32_t lane = tid & 31;
    constexpr int kU = 8;  // independent loads in flight per thread
    for (uint32_t eb = e0; eb < e_end; eb += T * kU) {
      uint32_t cur[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const uint32_t e = eb + u * T + tid;
        cur[u] = (e < e_end) ? __ldg(p.doc + base + e) : 0;
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const uint32_t e = eb + u * T + tid;
        // block of the previous posting: the neighbouring lane holds it, lane 0 re-reads it
        uint32_t prv = __shfl_up_sync(0xffffffffu, cur[u], 1);
        if (lane == 0 && e > 0 && e < e_end) prv = __ldg(p.doc + base + e - 1);
        if (e >= e_end) continue;
        const uint32_t blk = cur[u] / kSpBlock;
        const int prev = (e == 0) ? -1 : (int)(prv / kSpBlock);
        if (prev != (int)blk)
          for (int j = prev + 1; j <= (int)blk; ++j) row[j] = e;     // first posting at or after block j
        if (e == len - 1)
          for (uint32_t j = blk + 1; j <= p.n_blocks; ++j) row[j] = len;  // blocks after the last posting
      }
    }
  }
  if (p.trace && tid == 0 && blockIdx.x < 148) {
    unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    p.trace[(300 + blockIdx.x) * 8 + 1] = t_;
  }
}

struct SparseParams {
  const uint64_t* tptr;
  const uint2* post;      // (doc, weight bits) pairs, token-major
  uint32_t vocab;
  uint64_t n_docs;
  const uint32_t* q_tok;
  const float* q_w;
  uint32_t q_nnz;
  const uint32_t* bounds;
  uint32_t n_blocks;
  const int32_t* slot_of;        // nullable: static block index (see SparseDev)
  const uint32_t* block_index;
  const uint32_t* bitset;
  uint32_t k;
  uint64_t row_base;
  ckey_t* partial;
  uint32_t* partial_cnt;
  uint32_t* done;
  uint32_t* claim;            // accumulate kernel: next unclaimed 256-doc block (zero between queries)
  float* blk_scores;          // [n_docs rounded up to 256] scores of the accumulated blocks
  uint32_t* blk_mask;         // [n_wblocks][8] "touched" bits
  float* out_scores;
  uint64_t* out_rows;
  uint32_t* out_n;
  unsigned long long* trace;  // optional [grid][8] stamps (development aid)
};
#define SPTRACE(slot) do { if (p.trace && tid == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.trace[blockIdx.x * 8 + (slot)] = t_; } } while (0)

constexpr uint32_t kSpR = 4;                       // 64-doc index blocks per warp block
constexpr uint32_t kSpWDocs = kSpBlock * kSpR;     // docs owned by one warp at a time (256)

// The leg is two kernels.  ACCUMULATE (sparse_accum_kernel): warps claim 256-doc blocks from a global
// counter, build the block's scores in shared memory and write them — 256 f32 + a 256-bit "touched"
// mask — to a per-index scratch.  SELECT (sparse_select_kernel): streams that scratch through the top-k
// accumulator, last-CTA merge.  Split, the two cost 148 us at k = 500 where the single kernel cost 162 us
// (the accumulate loop runs without the accumulator's registers and shared memory, 2 CTAs per SM).
// (A 3-warp "slim" accumulate kernel that runs on the SMs BESIDE the hybrid call's dense scan was built
// and measured — DESIGN.md §4.3c: it fits, but its 444 warps do a third of the work in the scan's 440 us
// and slow the scan by 18 us; dropped.)
struct SpAccSmemBase {
  uint64_t base[kSpMaxQ];                //  8 KB  start of query token i's posting list
  const uint32_t* brow[kSpMaxQ];         //  8 KB  its block-boundary row (static index or this query's bounds)
  float qw[kSpMaxQ];                     //  4 KB
};
template <int WARPS>
struct SpAccSmem : SpAccSmemBase {
  float acc[WARPS][kSpWDocs];            //  1 KB per warp
  uint8_t touched[WARPS][kSpWDocs];      //  256 B per warp
};

// ---- pass 2a: accumulate --------------------------------------------------------------------
// A warp owns a block of 256 docs (four 64-doc index blocks): its accumulators sit in shared
// memory.  The block's slice of every query token's posting list is contiguous; the slices of 32
// tokens at a time are cut into 32-posting chunks and the (token, chunk) items are walked in query
// order, eight per batch — all of a batch's loads in flight together, every load a full 256-byte
// line, then applied IN QUERY ORDER (the reference's accumulation order, index.rs:251-259):
// within one token the docs are distinct, so the lanes update
//   acc[doc] = acc[doc] + qw*dw   (separate f32 multiply and add)
// without conflicts, and only a __syncwarp separates items.
// (Round 1 used 64-doc blocks: a heavy token has ~26 postings there, so a third of the lanes idled,
// every (token, block) visit paid ~40 instructions of bookkeeping, and the kernel was issue-bound at
// 62 % of the slots for 17 % of the DRAM bandwidth — profiles/r02_sparse_search_*.)
constexpr int kSpBatch = 8;
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2) sparse_accum_kernel(const SparseParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SpAccSmem<WARPS>& s = *reinterpret_cast<SpAccSmem<WARPS>*>(smem_raw);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t stride = p.n_blocks + 1;
  for (uint32_t i = tid; i < p.q_nnz; i += WARPS * 32) {
    const uint32_t t = __ldg(p.q_tok + i);
    s.base[i] = (t < p.vocab) ? __ldg(p.tptr + t) : 0;
    s.qw[i] = __ldg(p.q_w + i);
    const int32_t slot = (p.slot_of && t < p.vocab) ? __ldg(p.slot_of + t) : -1;
    s.brow[i] = slot >= 0 ? p.block_index + (size_t)slot * stride : p.bounds + (size_t)i * stride;
  }
  __syncthreads();
  SPTRACE(6);
  const uint32_t n_wblocks = (p.n_blocks + kSpR - 1) / kSpR;
  float* acc = s.acc[warp];
  uint8_t* touched = s.touched[warp];
  for (;;) {
    // warps claim blocks one at a time: no warp waits for the slowest block of a static slice
    uint32_t wb = 0;
    if (lane == 0) wb = atomicAdd(p.claim, 1u);
    wb = __shfl_sync(0xffffffffu, wb, 0);
    if (wb >= n_wblocks) break;
    const uint32_t d0 = wb * kSpWDocs;
    const uint32_t b_lo = wb * kSpR, b_hi = min(b_lo + kSpR, p.n_blocks);
#pragma unroll
    for (uint32_t i = lane; i < kSpWDocs; i += 32) {
      acc[i] = 0.f;
      touched[i] = 0;
    }
    for (uint32_t i0 = 0; i0 < p.q_nnz; i0 += 32) {
      // slice [lo, hi) of 32 query tokens at once (one token per lane), and its chunk count
      uint32_t lo = 0, hi = 0;
      if (i0 + lane < p.q_nnz) {
        const uint32_t* row = s.brow[i0 + lane];
        lo = __ldg(row + b_lo);
        hi = __ldg(row + b_hi);
      }
      const uint32_t nch = (hi - lo + 31) >> 5;
      uint32_t incl = nch;                                  // inclusive prefix of the chunk counts
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
      }
      const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
      for (uint32_t j0 = 0; j0 < total; j0 += kSpBatch) {
        uint2 pp[kSpBatch];
        uint32_t cl[kSpBatch], tix[kSpBatch];
#pragma unroll
        for (int u = 0; u < kSpBatch; ++u) {
          const uint32_t j = j0 + u;
          const bool have = j < total;
          // the token that owns item j = the number of tokens whose chunks all come before it
          tix[u] = min((uint32_t)__popc(__ballot_sync(0xffffffffu, incl <= j)), 31u);
          const uint32_t first = __shfl_sync(0xffffffffu, incl - nch, tix[u]);
          const uint32_t tl = __shfl_sync(0xffffffffu, lo, tix[u]);
          const uint32_t th = __shfl_sync(0xffffffffu, hi, tix[u]);
          const uint32_t start = tl + (j - first) * 32;     // chunk (j - first) of that token's slice
          cl[u] = have ? min(32u, th - start) : 0;
          pp[u] = make_uint2(0u, 0u);
          if (lane < cl[u]) pp[u] = __ldg(p.post + s.base[i0 + tix[u]] + start + lane);
        }
#pragma unroll
        for (int u = 0; u < kSpBatch; ++u) {
          if (cl[u] == 0) continue;  // warp-uniform
          const float qw = s.qw[i0 + tix[u]];
          if (lane < cl[u]) {
            const uint32_t d = pp[u].x - d0;
            // *scores.entry(idx).or_insert(0.0) += query_weight * doc_weight   (index.rs:259)
            acc[d] = __fadd_rn(acc[d], __fmul_rn(qw, __uint_as_float(pp[u].y)));
            touched[d] = 1;
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
    // hand the block to the select kernel: 256 scores (coalesced) + 8 words of "touched" bits
#pragma unroll
    for (uint32_t j = 0; j < kSpWDocs / 32; ++j) {
      const uint32_t d = j * 32 + lane;
      p.blk_scores[(size_t)d0 + d] = acc[d];
      const uint32_t bits = __ballot_sync(0xffffffffu, touched[d] != 0);
      if (lane == 0) p.blk_mask[(size_t)wb * (kSpWDocs / 32) + j] = bits;
    }
    __syncwarp();
  }
  SPTRACE(7);
}

struct SpSmem {
  ckey_t buf[kSpCap];                    // 32 KB
  uint32_t pos[kMaxGrid];                //  4 KB
  uint32_t hist[kSelBuckets + 96];       //  8 KB  select() histogram
  ckey_t thr;
  uint32_t cnt;
  uint32_t last;
};

// ---- pass 2b: select ------------------------------------------------------------------------
// Streams the accumulated blocks (1 KB of scores + 32 B of mask each) through the CTA's top-k
// accumulator: touched docs that pass the filter, finite score (candidate.rs:275).
__global__ void __launch_bounds__(kSpThreads, 2) sparse_select_kernel(const SparseParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SpSmem& s = *reinterpret_cast<SpSmem*>(smem_raw);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TopK tk{s.buf, &s.cnt, &s.thr, kSpCap, Group{tid, kSpThreads, 0}, s.hist};
  tk.init();
  __syncthreads();
  SPTRACE(0);
  const uint32_t k = p.k;
  const uint32_t n_wblocks = (p.n_blocks + kSpR - 1) / kSpR;
  const uint32_t n_steps = (n_wblocks + kSpWarps - 1) / kSpWarps;
  for (uint32_t step = blockIdx.x; step < n_steps; step += gridDim.x) {
    const uint32_t wb = step * kSpWarps + warp;
    const uint32_t d0 = wb * kSpWDocs;
    // pushed in two halves of <= kSpWarps * 128 = 2048 keys with a re-selection in between: a selection
    // leaves at most kSpCap / 2 keys, so the accumulator (4096 slots) cannot overflow whatever the scores are
#pragma unroll 1
    for (uint32_t half = 0; half < 2; ++half) {
      const ckey_t thr = s.thr;
      if (wb < n_wblocks) {
#pragma unroll
        for (uint32_t j = half * (kSpWDocs / 64); j < (half + 1) * (kSpWDocs / 64); ++j) {
          const uint32_t d = j * 32 + lane;
          const uint64_t r = (uint64_t)d0 + d;
          const uint32_t bits = __ldcg(p.blk_mask + (size_t)wb * (kSpWDocs / 32) + j);
          if (!((bits >> lane) & 1u) || r >= p.n_docs) continue;
          if (p.bitset && !((__ldg(p.bitset + (r >> 5)) >> (r & 31)) & 1u)) continue;
          const float sc = __ldcg(p.blk_scores + r);
          if (!finite_bits(__float_as_uint(sc))) continue;
          const ckey_t key = make_key(sc, (uint32_t)r);
          if (key > thr) tk.push(key);
        }
      }
      __syncthreads();
      if (step == blockIdx.x && half == 0) SPTRACE(1);
      if (s.cnt + kSpWarps * (kSpWDocs / 2) > kSpCap || (s.thr == 0 && s.cnt >= k))
        tk.template select<kSpCap / kSpThreads>(k);
      __syncthreads();
    }
    if (step == blockIdx.x) SPTRACE(2);
  }
  SPTRACE(3);
  topk_finish<kSpCap / kSpThreads>(tk, k);
  SPTRACE(4);
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
  merge_partials_and_emit<kSpCap / kSpThreads>(tk, s.pos, k, p.partial, p.partial_cnt, gridDim.x,
                                               p.row_base, p.out_scores, p.out_rows, p.out_n);
  SPTRACE(5);
  if (tid == 0) {
    *p.done = 0;
    *p.claim = 0;     // the accumulate kernel's block counter, for the next query
  }
}

size_t sparse_bounds_bytes(uint64_t n_docs, uint32_t q_nnz) {
  const uint64_t n_blocks = (n_docs + kSpBlock - 1) / kSpBlock;
  return (size_t)q_nnz * (n_blocks + 1) * sizeof(uint32_t);
}

size_t sparse_block_scratch_bytes(uint64_t n_docs) {
  const uint64_t n_blocks = (n_docs + kSpBlock - 1) / kSpBlock;
  const uint64_t n_wblocks = (n_blocks + kSpR - 1) / kSpR;
  return (size_t)n_wblocks * kSpWDocs * sizeof(float) + (size_t)n_wblocks * (kSpWDocs / 32) * sizeof(uint32_t);
}

// The leg's three launches, separately, so that the hybrid call can place them around its dense scan
// (index.cu, search_hybrid_impl).  stage: kSparseBounds | kSparseAccum | kSparseSelect.
cudaError_t launch_sparse_stage(const SparseArgs& a, uint32_t stage, cudaStream_t st) {
  if (a.n_docs == 0 || a.k == 0 || a.k > kMaxK || a.q_nnz == 0 || a.q_nnz > kSpMaxQ ||
      a.n_docs > 0xFFFFFFFFull || !a.d_bounds || !a.d_block_scratch || !a.d_claim)
    return cudaErrorInvalidValue;
  const uint32_t n_blocks = (uint32_t)((a.n_docs + kSpBlock - 1) / kSpBlock);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaError_t e = cudaSuccess;
  // the static block index only applies while the corpus still has the block count it was built for
  const bool use_index = a.sp.d_slot_of && a.sp.d_block_index && a.sp.index_stride == n_blocks + 1;
  if (stage & kSparseBounds) {
    BoundsParams bp{a.sp.d_tptr, a.sp.d_doc, a.sp.vocab, a.d_q_tok, a.q_nnz, n_blocks, a.d_bounds,
                    use_index ? a.sp.d_slot_of : nullptr};
    bp.trace = (unsigned long long*)a.d_trace;
    sparse_bounds_kernel<256><<<sms * 6, 256, 0, st>>>(bp);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  }
  SparseParams p;
  p.tptr = a.sp.d_tptr; p.post = (const uint2*)a.sp.d_post; p.vocab = a.sp.vocab;
  p.n_docs = a.n_docs; p.q_tok = a.d_q_tok; p.q_w = a.d_q_w; p.q_nnz = a.q_nnz;
  p.bounds = a.d_bounds; p.n_blocks = n_blocks;
  p.slot_of = use_index ? a.sp.d_slot_of : nullptr;
  p.block_index = use_index ? a.sp.d_block_index : nullptr;
  p.bitset = a.d_bitset; p.k = a.k; p.row_base = a.row_base;
  p.partial = a.d_partial; p.partial_cnt = a.d_partial_cnt; p.done = a.d_done;
  p.out_scores = a.d_out_scores; p.out_rows = a.d_out_rows; p.out_n = a.d_out_n;
  p.trace = (unsigned long long*)a.d_trace;
  const uint32_t n_wblocks = (n_blocks + kSpR - 1) / kSpR;
  p.claim = a.d_claim;
  p.blk_scores = (float*)a.d_block_scratch;
  p.blk_mask = (uint32_t*)((uint8_t*)a.d_block_scratch + (size_t)n_wblocks * kSpWDocs * sizeof(float));
  if (stage & kSparseAccum) {
    e = cudaFuncSetAttribute(sparse_accum_kernel<kSpWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(SpAccSmem<kSpWarps>));
    if (e != cudaSuccess) return e;
    const uint32_t want = (n_wblocks + kSpWarps - 1) / kSpWarps;
    const int grid = (int)(want < (uint32_t)(2 * sms) ? want : (uint32_t)(2 * sms));
    if (p.trace) p.trace += 150 * 8;   // development aid: the select kernel's stamps use rows 0..
    sparse_accum_kernel<kSpWarps><<<grid, kSpThreads, sizeof(SpAccSmem<kSpWarps>), st>>>(p);
    p.trace = (unsigned long long*)a.d_trace;
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  }
  if (stage & kSparseSelect) {
    const uint32_t n_steps = (n_wblocks + kSpWarps - 1) / kSpWarps;
    int grid = (int)(n_steps < (uint32_t)(2 * sms) ? n_steps : (uint32_t)(2 * sms));
    if (grid > (int)kMaxGrid) grid = kMaxGrid;
    e = cudaFuncSetAttribute(sparse_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(SpSmem));
    if (e != cudaSuccess) return e;
    sparse_select_kernel<<<grid, kSpThreads, sizeof(SpSmem), st>>>(p);
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  }
  return cudaGetLastError();
}

cudaError_t launch_sparse_search(const SparseArgs& a, cudaStream_t st) {
  return launch_sparse_stage(a, kSparseBounds | kSparseAccum | kSparseSelect, st);
}

void free_sparse(SparseDev& sp) {
  cudaFree(sp.d_tptr); cudaFree(sp.d_doc); cudaFree(sp.d_post);
  cudaFree(sp.d_slot_of); cudaFree(sp.d_block_index);
  sp = SparseDev();
}

// The static block index is the bounds pass run once, at build time, over the longest lists.
cudaError_t launch_sparse_block_index(const SparseDev& sp, const uint32_t* d_tokens, uint32_t n_tok,
                                      uint32_t row0, uint64_t n_docs, cudaStream_t st) {
  if (n_tok == 0) return cudaSuccess;
  if (n_tok > kSpMaxQ || !sp.d_block_index) return cudaErrorInvalidValue;
  const uint32_t n_blocks = (uint32_t)((n_docs + kSpBlock - 1) / kSpBlock);
  if (sp.index_stride != n_blocks + 1) return cudaErrorInvalidValue;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint32_t* rows = sp.d_block_index + (size_t)row0 * sp.index_stride;
  cudaError_t e = cudaMemsetAsync(rows, 0, (size_t)n_tok * sp.index_stride * sizeof(uint32_t), st);
  if (e != cudaSuccess) return e;
  BoundsParams bp{sp.d_tptr, sp.d_doc, sp.vocab, d_tokens, n_tok, n_blocks, rows, nullptr};
  sparse_bounds_kernel<256><<<sms * 6, 256, 0, st>>>(bp);
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// alpha fusion (src/search/query.rs:914-1005), one CTA, pools <= 1024 each.
// ---------------------------------------------------------------------------
// The reference builds two HashMaps (id -> dense score, id -> normalised sparse score; a
// later duplicate overwrites), walks the union in insertion order and sorts by
// (fused desc, id asc).  Here: one open-addressing table in shared memory keyed by row
// ("last index wins" via atomicMax on (index, score bits)), one pass over the table to
// form (fused, row) keys, the register/shuffle block sort, and a table lookup per output
// row for the per-leg values.  Every f32 step keeps the reference's operation order.
constexpr uint32_t kFuseMax = kMaxK;           // per pool
constexpr uint32_t kFuseSlots = 4096;          // >= 2 * (2 * kFuseMax) entries' worth of slots
constexpr uint32_t kFuseThreads = 1024;
constexpr uint64_t kFuseEmpty = ~0ull;

struct FuseSmem {
  unsigned long long hrow[kFuseSlots];   // 32 KB
  unsigned long long hd[kFuseSlots];     // (index + 1) << 32 | dense score bits, 0 = not in the dense pool
  unsigned long long hs[kFuseSlots];     // same for the raw sparse score
  ckey_t keys[2 * kFuseMax];             // 16 KB
  unsigned long long row_min;
  uint32_t n_union;
  float max_sparse;
};

__device__ __forceinline__ uint32_t fuse_hash(uint64_t row) {
  return (uint32_t)((row * 0x9E3779B97F4A7C15ull) >> 52);  // 12 bits
}
__device__ __forceinline__ uint32_t fuse_find(const FuseSmem& s, uint64_t row) {
  uint32_t slot = fuse_hash(row);
  while (s.hrow[slot] != row) slot = (slot + 1) & (kFuseSlots - 1);
  return slot;
}
__device__ __forceinline__ uint32_t fuse_insert(FuseSmem& s, uint64_t row) {
  uint32_t slot = fuse_hash(row);
  while (true) {
    const unsigned long long prev = atomicCAS(&s.hrow[slot], kFuseEmpty, row);
    if (prev == kFuseEmpty || prev == row) return slot;
    slot = (slot + 1) & (kFuseSlots - 1);
  }
}

__global__ void __launch_bounds__(kFuseThreads, 1) fuse_pools_kernel(const FuseArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FuseSmem& s = *reinterpret_cast<FuseSmem*>(smem_raw);
  const uint32_t tid = threadIdx.x, T = kFuseThreads;
  const uint32_t nd = min(*a.d_n_dense, kFuseMax), ns = min(*a.d_n_sparse, kFuseMax);
  for (uint32_t i = tid; i < kFuseSlots; i += T) {
    s.hrow[i] = kFuseEmpty;
    s.hd[i] = 0;
    s.hs[i] = 0;
  }
  for (uint32_t i = tid; i < 2 * kFuseMax; i += T) s.keys[i] = 0;
  if (tid == 0) {
    s.row_min = ~0ull;
    s.n_union = 0;
  }
  // max_sparse = iter().map(score).reduce(f32::max).unwrap_or(0.0)   (query.rs:914-919);
  // f32::max ignores NaN, and max is order independent up to the sign of zero
  if (tid < 32) {
    float m = __uint_as_float(0x7FC00000u);  // NaN = identity of a NaN-ignoring max
    for (uint32_t i = tid; i < ns; i += 32) m = fmaxf(m, a.d_sparse_scores[i]);
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (tid == 0) s.max_sparse = (ns == 0) ? 0.f : m;
  }
  __syncthreads();
  // insert both pools: HashMap::insert semantics, a later duplicate overwrites
  for (uint32_t i = tid; i < nd + ns; i += T) {
    const bool dense = i < nd;
    const uint32_t j = dense ? i : i - nd;
    const uint64_t row = dense ? a.d_dense_rows[j] : a.d_sparse_rows[j];
    if (row == kFuseEmpty) continue;
    const uint32_t bits = __float_as_uint(dense ? a.d_dense_scores[j] : a.d_sparse_scores[j]);
    const uint32_t slot = fuse_insert(s, row);
    atomicMax(dense ? &s.hd[slot] : &s.hs[slot], ((unsigned long long)(j + 1) << 32) | bits);
    atomicMin(&s.row_min, (unsigned long long)row);
  }
  __syncthreads();
  const float max_sparse = s.max_sparse;
  const float alpha = a.alpha;
  const uint64_t row_min = s.row_min;
  for (uint32_t slot = tid; slot < kFuseSlots; slot += T) {
    const uint64_t row = s.hrow[slot];
    if (row == kFuseEmpty) continue;
    const unsigned long long pd = s.hd[slot], ps = s.hs[slot];
    const float d = pd ? __uint_as_float((uint32_t)pd) : 0.f;          // unwrap_or(0.0)
    float sn = 0.f;
    if (ps) sn = (max_sparse > 0.f) ? __fdiv_rn(__uint_as_float((uint32_t)ps), max_sparse) : 0.f;
    float fused;
    if (alpha <= 0.f)
      fused = __fadd_rn(d, __fmul_rn(sn, 0.1f));
    else
      fused = __fadd_rn(__fmul_rn(alpha, d), __fmul_rn(__fsub_rn(1.0f, alpha), sn));
    ckey_t key = ((ckey_t)ordered_u32(__float_as_uint(fused)) << 32) | (ckey_t)(~(uint32_t)(row - row_min));
    if (key == 0) key = 1;
    s.keys[atomicAdd(&s.n_union, 1u)] = key;
  }
  __syncthreads();
  // sort (fused desc by total order, row asc)   (query.rs:1004)
  const uint32_t n_union = s.n_union;
  const uint32_t P = max(next_pow2(n_union), kSortChunk);
  block_sort_desc(Group{tid, T, 0}, s.keys, P);
  const uint32_t n_out = min(n_union, a.pool_k);
  for (uint32_t i = tid; i < n_out; i += T) {
    const ckey_t key = s.keys[i];
    const uint64_t row = row_min + (uint64_t)(~(uint32_t)key);
    const uint32_t slot = fuse_find(s, row);
    const unsigned long long pd = s.hd[slot], ps = s.hs[slot];
    a.d_out_rows[i] = row;
    a.d_out_fused[i] = __uint_as_float(unordered_u32((uint32_t)(key >> 32)));
    a.d_out_dense[i] = pd ? __uint_as_float((uint32_t)pd) : 0.f;
    a.d_out_sparse_raw[i] = ps ? __uint_as_float((uint32_t)ps) : 0.f;
    a.d_out_present[i] = (uint8_t)((pd ? 1 : 0) | (ps ? 2 : 0));
  }
  if (tid == 0) *a.d_out_n = n_out;
  if (a.d_host_flag) {
    // the pool went straight into host-mapped memory: make it visible system-wide, then publish the
    // completion word the host polls (saves six D2H copies and the stream-sync wake-up per query)
    __threadfence_system();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile uint32_t*>(a.d_host_flag) = a.seq;
  }
}

cudaError_t launch_fuse_pools(const FuseArgs& a, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(fuse_pools_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(FuseSmem));
  if (e != cudaSuccess) return e;
  fuse_pools_kernel<<<1, kFuseThreads, sizeof(FuseSmem), st>>>(a);
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// rrf_fuse_n (src/search/scoring/fusion.rs:36-68), one CTA.
// ---------------------------------------------------------------------------
// score[id] = sum over lists IN LIST ORDER of 1/(K + rank + 1) (f32; rank = 0-based position
// of the FIRST occurrence of id in that list), then the bounded heap: top-`limit` by
// (score desc, id asc).  Lists are processed one after the other (two barriers each) so
// the f32 additions happen in the reference's order; within a list every id is owned by
// the thread holding its first occurrence.
constexpr uint32_t kRrfMaxTotal = 2048;   // total entries over all lists
constexpr uint32_t kRrfSlots = 4096;
struct RrfSmem {
  unsigned long long hrow[kRrfSlots];   // 32 KB
  uint32_t hfirst[kRrfSlots];           // first rank of the id in the CURRENT list
  float hscore[kRrfSlots];
  ckey_t keys[kRrfMaxTotal];            // 16 KB
  unsigned long long row_min;
  uint32_t n_union;
};
struct RrfParams {
  const uint64_t* ids;        // concatenated lists
  const uint32_t* list_off;   // [n_lists + 1]
  uint32_t n_lists;
  float k;
  uint32_t limit;
  uint64_t* out_ids;
  float* out_scores;
  uint32_t* out_n;
};
__global__ void __launch_bounds__(1024, 1) rrf_fuse_kernel(const RrfParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  RrfSmem& s = *reinterpret_cast<RrfSmem*>(smem_raw);
  const uint32_t tid = threadIdx.x, T = 1024;
  for (uint32_t i = tid; i < kRrfSlots; i += T) {
    s.hrow[i] = kFuseEmpty;
    s.hfirst[i] = 0xFFFFFFFFu;
    s.hscore[i] = 0.f;
  }
  for (uint32_t i = tid; i < kRrfMaxTotal; i += T) s.keys[i] = 0;
  if (tid == 0) {
    s.row_min = ~0ull;
    s.n_union = 0;
  }
  __syncthreads();
  auto insert = [&](uint64_t row) {
    uint32_t slot = fuse_hash(row);
    while (true) {
      const unsigned long long prev = atomicCAS(&s.hrow[slot], kFuseEmpty, row);
      if (prev == kFuseEmpty || prev == row) return slot;
      slot = (slot + 1) & (kRrfSlots - 1);
    }
  };
  for (uint32_t l = 0; l < p.n_lists; ++l) {
    const uint32_t b = p.list_off[l], e = p.list_off[l + 1];
    for (uint32_t r = tid; r < e - b; r += T) {          // first occurrence per id
      const uint64_t row = p.ids[b + r];
      if (row == kFuseEmpty) continue;
      const uint32_t slot = insert(row);
      atomicMin(&s.hfirst[slot], r);
      atomicMin(&s.row_min, (unsigned long long)row);
    }
    __syncthreads();
    for (uint32_t r = tid; r < e - b; r += T) {
      const uint64_t row = p.ids[b + r];
      if (row == kFuseEmpty) continue;
      uint32_t slot = fuse_hash(row);
      while (s.hrow[slot] != row) slot = (slot + 1) & (kRrfSlots - 1);
      if (s.hfirst[slot] == r) {
        // contribution = 1.0 / (k + rank as f32 + 1.0);  *entry += contribution
        const float c = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(p.k, (float)r), 1.0f));
        s.hscore[slot] = __fadd_rn(s.hscore[slot], c);
      }
    }
    __syncthreads();
    for (uint32_t r = tid; r < e - b; r += T) {          // reset the per-list first-rank marks
      const uint64_t row = p.ids[b + r];
      if (row == kFuseEmpty) continue;
      uint32_t slot = fuse_hash(row);
      while (s.hrow[slot] != row) slot = (slot + 1) & (kRrfSlots - 1);
      s.hfirst[slot] = 0xFFFFFFFFu;
    }
    __syncthreads();
  }
  const uint64_t row_min = s.row_min;
  for (uint32_t slot = tid; slot < kRrfSlots; slot += T) {
    const uint64_t row = s.hrow[slot];
    if (row == kFuseEmpty) continue;
    const float sc = s.hscore[slot];
    if (!finite_bits(__float_as_uint(sc))) continue;     // BoundedScoreHeap drops non-finite
    ckey_t key = ((ckey_t)ordered_u32(__float_as_uint(sc)) << 32) | (ckey_t)(~(uint32_t)(row - row_min));
    s.keys[atomicAdd(&s.n_union, 1u)] = key;
  }
  __syncthreads();
  const uint32_t n_union = s.n_union;
  block_sort_desc(Group{tid, T, 0}, s.keys, max(next_pow2(n_union), kSortChunk));
  const uint32_t n_out = min(n_union, p.limit);
  for (uint32_t i = tid; i < n_out; i += T) {
    p.out_ids[i] = row_min + (uint64_t)(~(uint32_t)s.keys[i]);
    p.out_scores[i] = __uint_as_float(unordered_u32((uint32_t)(s.keys[i] >> 32)));
  }
  if (tid == 0) *p.out_n = n_out;
}

cudaError_t launch_rrf_fuse(const uint64_t* d_ids, const uint32_t* d_list_off, uint32_t n_lists, float k,
                            uint32_t limit, uint64_t* d_out_ids, float* d_out_scores, uint32_t* d_out_n,
                            cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(rrf_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(RrfSmem));
  if (e != cudaSuccess) return e;
  RrfParams p{d_ids, d_list_off, n_lists, k, limit, d_out_ids, d_out_scores, d_out_n};
  rrf_fuse_kernel<<<1, 1024, sizeof(RrfSmem), st>>>(p);
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
