// scan_single_impl.cuh — the single-query scan kernel template and its launcher, included by
// the six scan_v_*.cu translation units (one per row mode x top-k variant) so that the 48
// instantiations compile in parallel.  See scan_single.cu for the description of the kernel.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "internal.h"
#include "peer.cuh"
#include "scan_params.cuh"

namespace cqs {

constexpr int kWarps = 8;                       // consumer warps
constexpr int kConsumers = kWarps * 32;
constexpr int kThreads = kConsumers + 32;       // + producer warp
constexpr uint32_t kCap = 4096;                 // candidate slots per CTA (32 KB)
constexpr uint32_t kStageTarget = 24576;        // bytes per pipeline stage (target)
constexpr uint32_t kInFlight = 131072;          // bytes in flight per SM (target)

// ---- mbarrier / TMA bulk-copy primitives ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem)),
      "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}


__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(slot) do { if (p.trace && ctid == 0) p.trace[blockIdx.x * 8 + (slot)] = gtimer(); } while (0)

template <int MODE>
struct LaneVec;
template <>
struct LaneVec<0> {  // f32, 4 elems / 16 B
  static constexpr int E = 4;
  static constexpr int BYTES = 16;
  uint32_t r[4];
  __device__ __forceinline__ void load(const uint8_t* p) {  // p: shared memory
    uint4 v = *reinterpret_cast<const uint4*>(p);
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  }
  __device__ __forceinline__ void zero() { r[0] = r[1] = r[2] = r[3] = 0; }
  __device__ __forceinline__ void fma(const float* q, float* acc) const {
    acc[0] = fmaf(__uint_as_float(r[0]), q[0], acc[0]);
    acc[1] = fmaf(__uint_as_float(r[1]), q[1], acc[1]);
    acc[2] = fmaf(__uint_as_float(r[2]), q[2], acc[2]);
    acc[3] = fmaf(__uint_as_float(r[3]), q[3], acc[3]);
  }
};
template <>
struct LaneVec<1> {  // bf16, 8 elems / 16 B
  static constexpr int E = 8;
  static constexpr int BYTES = 16;
  uint32_t r[4];
  __device__ __forceinline__ void load(const uint8_t* p) {  // p: shared memory
    uint4 v = *reinterpret_cast<const uint4*>(p);
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  }
  __device__ __forceinline__ void zero() { r[0] = r[1] = r[2] = r[3] = 0; }
  __device__ __forceinline__ void fma(const float* q, float* acc) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // bf16 -> f32 is a 16-bit shift (exact)
      acc[(2 * i) & 3] = fmaf(__uint_as_float(r[i] << 16), q[2 * i], acc[(2 * i) & 3]);
      acc[(2 * i + 1) & 3] =
          fmaf(__uint_as_float(r[i] & 0xFFFF0000u), q[2 * i + 1], acc[(2 * i + 1) & 3]);
    }
  }
};
template <>
struct LaneVec<2> {  // bf16, 4 elems / 8 B
  static constexpr int E = 4;
  static constexpr int BYTES = 8;
  uint32_t r[2];
  __device__ __forceinline__ void load(const uint8_t* p) {  // p: shared memory
    uint2 v = *reinterpret_cast<const uint2*>(p);
    r[0] = v.x; r[1] = v.y;
  }
  __device__ __forceinline__ void zero() { r[0] = r[1] = 0; }
  __device__ __forceinline__ void fma(const float* q, float* acc) const {
    acc[0] = fmaf(__uint_as_float(r[0] << 16), q[0], acc[0]);
    acc[1] = fmaf(__uint_as_float(r[0] & 0xFFFF0000u), q[1], acc[1]);
    acc[2] = fmaf(__uint_as_float(r[1] << 16), q[2], acc[2]);
    acc[3] = fmaf(__uint_as_float(r[1] & 0xFFFF0000u), q[3], acc[3]);
  }
};

template <int MODE, int NV>
struct ScanCfg {
  using V = LaneVec<MODE>;
  static constexpr uint32_t ROWB = NV * 32 * V::BYTES;  // bytes per (padded) row
  static constexpr uint32_t U0 = kStageTarget / (kWarps * ROWB);
  static constexpr uint32_t U = U0 < 1 ? 1 : (U0 > 8 ? 8 : U0);  // rows per consumer warp per stage
  static constexpr uint32_t RPS = kWarps * U;                   // rows per stage
  static constexpr uint32_t STAGE_BYTES = RPS * ROWB;
  static constexpr uint32_t S0 = kInFlight / STAGE_BYTES;
  static constexpr uint32_t STAGES = S0 < 2 ? 2 : (S0 > 8 ? 8 : S0);
  static constexpr uint32_t SMEM = STAGES * STAGE_BYTES + kCap * sizeof(ckey_t);
};

// Collective over tk.g: raise the CTA's threshold to the j-th largest of the values the CTAs have
// published in col[0..G) (0 = nothing published yet), minus one ("key > thr" keeps keys >= it).
// A published value is usable iff at least j published values are >= it (those j CTAs hold >= m
// keys >= it each); the largest usable one is the j-th largest.  colv: shared scratch for G keys.
// Not inlined: one copy per translation unit instead of one per instantiation.
static __device__ __noinline__ void refresh_global_thr(TopK tk, ckey_t* colv, const ckey_t* col,
                                                       uint32_t G, uint32_t j) {
  tk.assume_shared();
  CQS_ASSUME_SHARED(colv);
  const uint32_t tid = tk.g.tid, T = tk.g.nthr;
  for (uint32_t l = tid; l < G; l += T) colv[l] = __ldcg(col + l);
  tk.g.sync();
  for (uint32_t t = tid; t < G; t += T) {
    const ckey_t mine = colv[t];
    if (mine == 0) continue;
    uint32_t ge = 0;
    for (uint32_t i = 0; i < G; ++i) ge += (colv[i] >= mine) ? 1u : 0u;
    if (ge >= j) atomicMax(tk.thr, mine - 1);
  }
  tk.g.sync();
}

// STORAGE_BF16_F32 fast path (the scan streams the bf16 shadow rows, p.k = k' candidates per CTA).
// EVERY CTA re-scores its own candidates (tk.buf[0..n), keys of the shadow scan) on their f32 master
// rows before it emits its list — in parallel on all SMs, ~k' random 3 KB rows each — with the f32
// scan's own arithmetic (LaneVec<0>: lane l owns elements (v*32+l)*4.., four accumulators, the same
// butterfly), so the scores are bit-identical to a STORAGE_F32 scan.  The list is then sorted by
// EXACT score and cut to k_out: the last CTA merges G short exact lists, exactly as for a k_out scan.
// Four rows are in flight per warp.  Collective over tk.g; on return buf[0..*cnt) holds the valid
// exact keys, descending.
static __device__ __noinline__ void rescore_local(TopK tk, const ScanParams& p) {
  tk.assume_shared();
  const uint32_t tid = tk.g.tid, T = tk.g.nthr, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  tk.g.sync();
  const uint32_t n = min(*tk.cnt, tk.cap);
  const uint32_t ld4 = p.exact_nv * 32;   // float4 per row
  const float4* q4 = reinterpret_cast<const float4*>(p.query);
  for (uint32_t j0 = warp * 4; j0 < n; j0 += nwarps * 4) {
    const float4* row[4];
    uint32_t r[4];
    float acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t j = min(j0 + u, n - 1);
      r[u] = key_row(tk.buf[j]);
      row[u] = reinterpret_cast<const float4*>(p.exact_rows) + (size_t)r[u] * ld4 + lane;
      acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f;
    }
    for (uint32_t v = 0; v < p.exact_nv; ++v) {
      const float4 q = __ldg(q4 + v * 32 + lane);
      float4 x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = __ldg(row[u] + v * 32);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[u][0] = fmaf(x[u].x, q.x, acc[u][0]);
        acc[u][1] = fmaf(x[u].y, q.y, acc[u][1]);
        acc[u][2] = fmaf(x[u].z, q.z, acc[u][2]);
        acc[u][3] = fmaf(x[u].w, q.w, acc[u][3]);
      }
    }
    float sc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) sc[u] = (acc[u][0] + acc[u][1]) + (acc[u][2] + acc[u][3]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int u = 0; u < 4; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], off);
    __syncwarp();
    // a non-finite exact score drops the row (candidate.rs:275): its key becomes (0, ~row) — below every
    // valid key (whose high word is >= 0x00800000) yet unique, as the exact sort requires
    if (lane < 4 && j0 + lane < n) {
      float mine = sc[0];
#pragma unroll
      for (int u = 1; u < 4; ++u)
        if (lane == u) mine = sc[u];
      uint32_t rr = r[0];
#pragma unroll
      for (int u = 1; u < 4; ++u)
        if (lane == u) rr = r[u];
      tk.buf[j0 + lane] = finite_bits(__float_as_uint(mine)) ? make_key(mine, rr) : (ckey_t)(~rr);
    }
  }
  tk.g.sync();
  tk.compact(p.k_out);                       // exact sort; dropped candidates sink to the end
  const uint32_t m = *tk.cnt;
  uint32_t valid = 0;
  for (uint32_t i = 0; i < m; ++i) valid += (tk.buf[i] >> 32) != 0 ? 1u : 0u;   // m <= k_out, broadcast reads
  tk.g.sync();
  if (tid == 0) *tk.cnt = valid;
  tk.g.sync();
}

// Completeness proof of the shadow scan, by the last CTA (collective over tk.g).  excl[c] is an upper
// bound of the SHADOW key of every row CTA c scanned but did not re-score (its streaming threshold /
// the cut of its candidate list; 0 = it re-scored every eligible row).  With cut = max_c excl[c]:
//   a row that was not re-scored has shadow score <= cut, and |q.r - q.r16| = |q.(r - r16)| <= |q| * D
//   (D = max_row_delta, measured at finalize) plus the fp32 accumulation error of both dots
//   (<= 2^-18 |q| R for ld <= 2048), so its exact score is <= cut + E.  If the k_out-th exact score of
//   the merged answer is > cut + E, no such row can belong to the answer.
// Returns kUnprovenBit when the proof fails (the host then re-runs the query on the f32 master rows).
static __device__ __noinline__ uint32_t prove_shadow(TopK tk, const ScanParams& p, const ckey_t* excl, uint32_t G,
                                                     float* s_red, ckey_t* s_cut) {
  tk.assume_shared();
  CQS_ASSUME_SHARED(s_red);
  CQS_ASSUME_SHARED(s_cut);
  const uint32_t tid = tk.g.tid, T = tk.g.nthr, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  if (tid == 0) *s_cut = 0;
  const uint32_t ld = p.exact_nv * 128;
  float qq = 0.f;
  for (uint32_t i = tid; i < ld; i += T) {
    const float v = __ldg(p.query + i);
    qq = fmaf(v, v, qq);
  }
  for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
  if (lane == 0) s_red[warp] = qq;
  tk.g.sync();
  ckey_t mx = 0;
  for (uint32_t l = tid; l < G; l += T) {
    const ckey_t e = __ldcg(excl + l);
    mx = e > mx ? e : mx;
  }
  if (mx) atomicMax(s_cut, mx);
  tk.g.sync();
  const ckey_t cut_key = *s_cut;
  if (cut_key == 0) return 0;                 // every eligible row of the corpus was re-scored
  float qn = 0.f;
  for (uint32_t w = 0; w < nwarps; ++w) qn += s_red[w];
  qn = sqrtf(qn);
  const float E = 1.001f * qn * (p.max_row_delta + 3.814697265625e-6f * p.max_row_norm);
  const uint32_t n = *tk.cnt;                 // merged exact keys, descending, in tk.buf
  if (n < p.k_out) return kUnprovenBit;
  return key_score(tk.buf[p.k_out - 1]) > key_score(cut_key) + E ? 0u : kUnprovenBit;
}

// SMALLK (k <= 32): every consumer warp keeps its own sorted top-32 in registers
// (lane i holds the i-th best key); a row that beats the warp's k-th key is
// inserted with one ballot and one shuffle.  No shared-memory traffic, no
// barriers and no re-selection stalls while streaming.  Larger k use the
// CTA-level shared-memory accumulator (TopK).
template <int MODE, int NV, bool SMALLK>
__global__ void __launch_bounds__(kThreads, 1) scan_topk_kernel(const ScanParams p) {
  using Cfg = ScanCfg<MODE, NV>;
  using V = LaneVec<MODE>;
  constexpr int E = V::E;
  constexpr int U = Cfg::U;
  constexpr uint32_t kBurst = 1024;                       // max pushes per check interval
  constexpr uint32_t kCheckEvery = kBurst / Cfg::RPS;     // tiles between checks
  static_assert(kCheckEvery >= 1, "stage too large");
  static_assert(kCap >= kMaxK + 2 * kBurst, "candidate buffer too small");

  extern __shared__ __align__(128) uint8_t smem[];
  ckey_t* s_buf = reinterpret_cast<ckey_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  __shared__ __align__(8) uint64_t s_full[Cfg::STAGES];
  __shared__ __align__(8) uint64_t s_empty[Cfg::STAGES];
  __shared__ uint64_t s_tile[Cfg::STAGES];  // which tile a stage holds (~0 = no more tiles)
  __shared__ uint32_t s_cnt;
  __shared__ ckey_t s_thr;
  __shared__ uint32_t s_last;
  __shared__ __align__(8) uint32_t s_pos[kMaxGrid];
  __shared__ __align__(8) uint32_t s_hist[kSelBuckets + 96];

  const uint32_t lane = threadIdx.x & 31;
  const uint64_t n = p.n_rows;
  const uint64_t tiles = (n + Cfg::RPS - 1) / Cfg::RPS;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (threadIdx.x < 32) {
    // ===== producer warp: one lane streams row tiles into the ring =====
    // Tiles are handed out dynamically (one global atomic per 24-32 KB tile, fetched
    // one tile ahead so its latency hides behind the empty-slot wait): SMs that
    // stream faster simply take more tiles, which removes the ~5% finish-time
    // spread of a static round-robin split.
    if (lane == 0) {
      // Work units: the first p.n_big units are runs of p.chunk consecutive tiles (an SM then
      // stays inside one or two 2 MB pages for a while instead of touching every page of the
      // corpus — at 19 GB per GPU the interleaved hand-out streamed ~12 % slower), the rest
      // are single tiles so that the SMs still finish together.
      uint32_t s = 0, ph = 0;
      const uint64_t big_tiles = p.n_big * p.chunk;
      const uint64_t units = p.n_big + (tiles - big_tiles);
      uint64_t unit = blockIdx.x;
      uint64_t nxt = (uint64_t)atomicAdd(p.done + 1, 1u) + gridDim.x;
      while (unit < units) {
        uint64_t tile = unit < p.n_big ? unit * p.chunk : big_tiles + (unit - p.n_big);
        const uint64_t tile_end = unit < p.n_big ? tile + p.chunk : tile + 1;
        for (; tile < tile_end; ++tile) {
          mbar_wait(&s_empty[s], ph ^ 1);
          const uint64_t row0 = tile * Cfg::RPS;
          const uint64_t rows = (n - row0 < Cfg::RPS) ? (n - row0) : Cfg::RPS;
          const uint32_t bytes = (uint32_t)(rows * Cfg::ROWB);
          s_tile[s] = tile;
          mbar_expect_tx(&s_full[s], bytes);
          bulk_g2s(smem + (size_t)s * Cfg::STAGE_BYTES, p.rows + row0 * Cfg::ROWB, bytes, &s_full[s]);
          if (++s == Cfg::STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        unit = nxt;
        nxt = (uint64_t)atomicAdd(p.done + 1, 1u) + gridDim.x;
      }
      mbar_wait(&s_empty[s], ph ^ 1);
      s_tile[s] = ~0ull;       // end-of-stream marker
      mbar_arrive(&s_full[s]);
    }
    return;
  }

  // ===== consumer warps =====
  const uint32_t ctid = threadIdx.x - 32, warp = ctid >> 5;
  TopK tk{s_buf, &s_cnt, &s_thr, kCap, Group{ctid, kConsumers, 1}, s_hist};
  tk.init();
  float q[NV * E];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < E; ++e) q[v * E + e] = __ldg(p.query + (v * 32 + lane) * E + e);
  tk.g.sync();
  TRACE(0);

  const uint32_t k = p.k;
  ckey_t thr = 0;
  ckey_t slot = 0, wthr = 0;  // SMALLK: this lane's entry of the warp's sorted list / its k-th key
  // Large k: a CTA sees only n/G rows, so its own k-th best is a weak filter (k = 500 of 6,757 rows
  // lets 7 % of the rows through) and every CTA would carry ~k candidates into its final sort and
  // into the merge.  The CTAs therefore share a GLOBAL threshold: whenever a CTA re-selects its
  // buffer it also publishes v = a lower bound of its m-th best key (p.col[blockIdx.x], monotone);
  // the j-th largest published value, j = ceil(k/m), is a lower bound of the global k-th best
  // (those j CTAs hold >= m keys >= it each, j*m >= k keys in all, and keys of different CTAs are
  // different rows) — the column bound of the merge, applied while streaming.  m = ceil(2k/G) makes
  // j ~ G/2: the bound sits near the median of the CTAs' m-th keys.  Always valid, never too tight:
  // no fallback pass.  Refreshed at every check interval (one 1.2 KB L2 read per CTA).
  uint32_t m_pub = 0, j_need = 0;
  if (!SMALLK && gridDim.x >= 8) {
    m_pub = (2 * k + gridDim.x - 1) / gridDim.x;
    j_need = (k + m_pub - 1) / m_pub;
    if (j_need > gridDim.x) m_pub = 0;
  }
  float thr_s = __uint_as_float(0xFF800000u);   // score of the current threshold key (-inf: none yet)
  ckey_t* s_vm = reinterpret_cast<ckey_t*>(s_pos);   // [0]: select()'s m-th-key bound (s_pos is free until the merge)
  uint32_t it = 0, s = 0, ph = 0;
  for (;; ++it) {
    mbar_wait(&s_full[s], ph);
    const uint64_t tile = s_tile[s];
    if (tile == ~0ull) break;
    const uint8_t* st = smem + (size_t)s * Cfg::STAGE_BYTES + (size_t)(warp * U) * Cfg::ROWB +
                        (size_t)lane * V::BYTES;
    V d[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v) d[u][v].load(st + (size_t)u * Cfg::ROWB + (size_t)v * 32 * V::BYTES);
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[s]);  // rows are in registers: hand the stage back
    if (++s == Cfg::STAGES) {
      s = 0;
      ph ^= 1;
    }
    const uint64_t row0 = tile * Cfg::RPS + (uint64_t)warp * U;
    float sc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int v = 0; v < NV; ++v) d[u][v].fma(q + v * E, acc);
      sc[u] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
      for (int u = 0; u < U; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], off);
    float mine = sc[0];
#pragma unroll
    for (int u = 1; u < U; ++u)
      if (lane == u) mine = sc[u];
    // Warp-uniform fast reject: in steady state none of a tile's rows beats the current
    // threshold, and the key / filter / push code below (a third of the loop's instructions)
    // is skipped.  thr_s is the threshold's SCORE (-inf while there is none): a key can only
    // beat the threshold key if its score is >= that (ties go on to the exact key compare);
    // NaN never qualifies and is dropped anyway.  Not applied with the score fold (rare).
    const bool maybe = p.sig.pipeline || __any_sync(0xffffffffu, lane < U && mine >= thr_s);
    ckey_t mykey = 0;
    if (maybe && lane < U && row0 + lane < n) {
      const uint64_t r = row0 + lane;
      const uint32_t bits = __float_as_uint(mine);
      bool ok = finite_bits(bits);
      if (ok && p.bitset) ok = (__ldg(p.bitset + (r >> 5)) >> (r & 31)) & 1u;
      float sc = mine;
      if (ok && (p.sig.pipeline || p.sig.d_ctype)) {
        // Store::search_filtered on device: SQL-style type/language filter, then the
        // score fold base = clamp(cos,0,1); max(base,0)*note_boost; *importance; >= threshold
        // (candidate.rs:420-562) applied BEFORE the top-k, in the reference's op order.
        const float base = p.sig.pipeline ? fminf(fmaxf(mine, 0.f), 1.f) : mine;
        // cheap upper bound first: a row that cannot beat the current k-th key needs no loads
        const float ub = p.sig.pipeline
                             ? __fmul_rn(__fmul_rn(base, p.sig.max_note_boost), p.sig.max_importance)
                             : mine;
        ok = make_key(ub, 0u) > (SMALLK ? wthr : thr);
        if (ok && p.sig.d_ctype) {
          const uint32_t ct = __ldg(p.sig.d_ctype + r), lg = __ldg(p.sig.d_lang + r);
          ok = ((p.sig.type_mask[ct >> 6] >> (ct & 63)) & 1ull) && ((p.sig.lang_mask[lg >> 6] >> (lg & 63)) & 1ull);
        }
        if (ok && p.sig.pipeline) {
          sc = base;
          if (p.sig.d_note_boost) sc = __fmul_rn(fmaxf(sc, 0.f), __ldg(p.sig.d_note_boost + r));
          else sc = fmaxf(sc, 0.f);
          if (p.sig.d_importance) sc = __fmul_rn(sc, __ldg(p.sig.d_importance + r));
          ok = sc >= p.sig.threshold;
        }
      }
      if (ok) mykey = make_key(sc, (uint32_t)r);
    }
    if (SMALLK) {
      if (maybe) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const ckey_t key = __shfl_sync(0xffffffffu, mykey, u);
          if (key > wthr) {  // warp-uniform
            const uint32_t pos = __popc(__ballot_sync(0xffffffffu, slot > key));
            const ckey_t up = __shfl_up_sync(0xffffffffu, slot, 1);
            slot = lane < pos ? slot : (lane == pos ? key : up);
            wthr = __shfl_sync(0xffffffffu, slot, k - 1);
            if (wthr) thr_s = key_score(wthr);
          }
        }
      }
      continue;
    }
    if (mykey > thr) tk.push(mykey);
    if ((it % kCheckEvery) == kCheckEvery - 1) {
      // consumer 0's view of the count may miss pushes of this interval that
      // are still in flight in other warps (< kBurst), hence the 2*kBurst.
      bool need = false;
      if (ctid == 0) {
        uint32_t c = s_cnt;
        need = (c + 2 * kBurst > kCap) || (s_thr == 0 && c >= k);
      }
      const bool selected = tk.g.any(need);
      if (selected) {
        if (ctid == 0) s_vm[0] = 0;
        tk.template select<kCap / kConsumers>(k, m_pub, s_vm);
        if (m_pub && ctid == 0 && s_vm[0]) atomicMax(p.col + blockIdx.x, s_vm[0]);
      }
      // the shared bound is re-read after a re-selection (this CTA just published), at the first two
      // checks (the other CTAs publish their first values around then) and every 8th check after
      // that: on a 12.5M-row shard a CTA passes ~80 checks, and a refresh per check cost 3-6 %
      const uint32_t chk = it / kCheckEvery;
      if (m_pub && (selected || chk < 2 || (chk & 7) == 7))
        refresh_global_thr(tk, reinterpret_cast<ckey_t*>(s_hist), p.col, gridDim.x, j_need);
      thr = s_thr;
      if (thr) thr_s = key_score(thr);
    }
  }
  TRACE(1);
  ckey_t gthr = 0;
  const bool rescore = p.exact_rows != nullptr;
  if (SMALLK) {
    // combine the 8 warp lists: rank sort of 256 keys straight into the partial list (or, when the
    // candidates are re-scored first, into a second shared-memory list)
    s_buf[ctid] = (lane < k) ? slot : 0;
    tk.g.sync();
    const ckey_t mine = s_buf[ctid];
    uint32_t rank = 0;
    if (mine != 0) {
      for (uint32_t j = 0; j < kConsumers; ++j) rank += (s_buf[j] > mine) ? 1u : 0u;
      if (rank < k) {
        if (rescore) s_buf[kConsumers + rank] = mine;
        else p.partial[(size_t)blockIdx.x * kMaxK + rank] = mine;
        atomicAdd(&s_cnt, 1u);
      }
    }
    if (rescore) {
      // anything this CTA dropped — a warp's full list rejecting a row, the cut of the 256-key union —
      // is <= the union's k-th key (it contains every warp's k best)
      const bool dropped = tk.g.any(wthr != 0 || (mine != 0 && rank >= k));
      const uint32_t n_c = s_cnt;
      const ckey_t excl = (dropped && n_c >= k) ? s_buf[kConsumers + k - 1] : 0;
      tk.g.sync();
      if (ctid < n_c) s_buf[ctid] = s_buf[kConsumers + ctid];   // k <= 32 < kConsumers
      tk.g.sync();
      rescore_local(tk, p);
      const uint32_t mycnt = s_cnt;
      if (ctid < mycnt) p.partial[(size_t)blockIdx.x * kMaxK + ctid] = s_buf[ctid];
      if (ctid == 0) {
        p.partial_cnt[blockIdx.x] = mycnt;
        p.excl[blockIdx.x] = excl;
      }
    } else {
      tk.g.sync();
      if (ctid == 0) p.partial_cnt[blockIdx.x] = s_cnt;
    }
    TRACE(2);
  } else {
    if (m_pub) {
      // whatever the other CTAs have published by now bounds the global k-th best: drop the
      // keys below it before the exact sort (this CTA then sorts and emits tens of keys, not k)
      refresh_global_thr(tk, reinterpret_cast<ckey_t*>(s_hist), p.col, gridDim.x, j_need);
      gthr = s_thr;   // max(local k-th bound, shared bound): a lower bound of the global k-th best
      tk.template prune<kCap / kConsumers>(gthr);
    }
    topk_finish<kCap / kConsumers>(tk, k);
    if (rescore) {
      // rows this CTA scanned but does not re-score were filtered by its (monotone) streaming
      // threshold or cut from its list: their shadow keys are <= max(shared bound, list cut)
      const ckey_t cut = s_thr;
      rescore_local(tk, p);
      if (ctid == 0) p.excl[blockIdx.x] = gthr > cut ? gthr : cut;
    }
    TRACE(2);
    const uint32_t mycnt = s_cnt;
    for (uint32_t i = ctid; i < mycnt; i += kConsumers)
      p.partial[(size_t)blockIdx.x * kMaxK + i] = s_buf[i];
    if (ctid == 0) p.partial_cnt[blockIdx.x] = mycnt;
  }
  __threadfence();
  tk.g.sync();
  if (ctid == 0) {
    uint32_t ticket = atomicAdd(p.done, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  tk.g.sync();
  TRACE(3);
  if (!s_last) return;
  __threadfence();
  const ckey_t thr0 = gthr;   // the shared threshold is a valid start for the merge
  if (m_pub)
    for (uint32_t l = ctid; l < gridDim.x; l += kConsumers) p.col[l] = 0;  // every CTA is past its last read
  // with re-scoring the partial lists hold EXACT keys cut to k_out: a plain k_out merge (the shared
  // threshold of the shadow scan does not apply to exact keys)
  const uint32_t k_emit = rescore ? p.k_out : k;   // what leaves this kernel / goes to the peers
  const ckey_t thr_m = rescore ? 0 : thr0;
  if (p.peer.world == 0) {
    merge_partials_and_emit<kCap / kConsumers>(tk, s_pos, k_emit, p.partial, p.partial_cnt, gridDim.x,
                                               p.row_base, p.out_scores, p.out_rows, p.out_n,
                                               p.trace ? p.trace + blockIdx.x * 8 : nullptr, thr_m);
    if (rescore) {
      const uint32_t flag = prove_shadow(tk, p, p.excl, gridDim.x, reinterpret_cast<float*>(s_hist), &s_thr);
      if (ctid == 0 && flag) *p.out_n |= flag;
    }
  } else {
    // Row-sharded corpus: the shard's list goes into this rank's own mailbox
    // block and, by plain stores over NVLink, into every peer's; one release flag per peer
    // publishes it; then wait for the peers' lists and merge — the all-gather and the merge
    // ride in the tail of the scan, no collective call and no extra launch.
    const PeerBlock own = peer_block(p.peer, p.peer.rank, p.peer.rank);
    merge_partials_and_emit<kCap / kConsumers>(tk, s_pos, k_emit, p.partial, p.partial_cnt, gridDim.x,
                                               p.row_base, own.scores, own.rows, own.n,
                                               p.trace ? p.trace + blockIdx.x * 8 : nullptr, thr_m);
    const uint32_t n = *tk.cnt;  // sorted keys are still in tk.buf[0..n)
    uint32_t n_word = n;         // carries kUnprovenBit to every rank
    if (rescore) {
      n_word |= prove_shadow(tk, p, p.excl, gridDim.x, reinterpret_cast<float*>(s_hist), &s_thr);
      if (ctid == 0) *own.n = n_word;
    }
    for (uint32_t e = ctid; e < p.peer.world * k_emit; e += kConsumers) {
      const uint32_t g = e / k_emit, i = e - g * k_emit;
      if (g == p.peer.rank || i >= n) continue;
      const PeerBlock pb = peer_block(p.peer, g, p.peer.rank);
      const ckey_t key = s_buf[i];
      pb.scores[i] = key_score(key);
      pb.rows[i] = p.row_base + key_row(key);
    }
    if (ctid < p.peer.world && ctid != p.peer.rank) peer_block(p.peer, ctid, p.peer.rank).n[0] = n_word;
    peer_signal(p.peer, tk.g);
    if (peer_wait(p.peer, tk.g)) {
      // the tile ring is idle by now (the producer has retired, every bulk copy has landed): stage the
      // world lists there so that the merge's binary searches probe shared memory, not the mailbox
      if ((size_t)p.peer.world * k_emit * 12 + 64 <= (size_t)Cfg::STAGES * Cfg::STAGE_BYTES)
        peer_merge_query_staged(p.peer, tk.g, 0, k_emit, smem, p.out_scores, p.out_rows, p.out_n);
      else
        peer_merge_query(p.peer, tk.g, 0, k_emit, p.out_scores, p.out_rows, p.out_n);
    } else {
      peer_emit_empty(tk.g, k_emit, p.out_scores, p.out_rows, p.out_n);
    }
  }
  TRACE(4);
  if (p.host_flag) {
    // the result went straight into host-mapped memory: make it visible system-wide, then
    // publish the completion word the host is polling (saves the D2H copies and the
    // driver's stream-sync wake-up on the latency path)
    __threadfence_system();
    tk.g.sync();
    if (ctid == 0) *reinterpret_cast<volatile uint32_t*>(p.host_flag) = p.seq;
  }
  if (ctid == 0) {
    p.done[0] = 0;
    p.done[1] = 0;
  }
}

template <int MODE, int NV, bool SMALLK>
static cudaError_t launch_nv(const ScanParams& p, int num_sms, cudaStream_t st) {
  using Cfg = ScanCfg<MODE, NV>;
  uint64_t tiles = (p.n_rows + Cfg::RPS - 1) / Cfg::RPS;
  int grid = (int)(tiles < (uint64_t)num_sms ? tiles : (uint64_t)num_sms);
  if (grid > (int)kMaxGrid) grid = kMaxGrid;
  // hand-out granularity: ~8 runs per SM over the first 7/8 of the corpus, at most 64 tiles
  // (1.5 MB) each; single tiles for the rest (see the producer loop)
  ScanParams q = p;
  uint64_t chunk = tiles / ((uint64_t)grid * 8);
  if (p.chunk_override) chunk = p.chunk_override;
  chunk = chunk < 1 ? 1 : (chunk > 64 ? 64 : chunk);
  q.chunk = (uint32_t)chunk;
  q.n_big = chunk > 1 ? (tiles - tiles / 8) / chunk : 0;
  auto kern = scan_topk_kernel<MODE, NV, SMALLK>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Cfg::SMEM);
  if (e != cudaSuccess) return e;
  // CQS_B200_MAX_CARVEOUT (development aid, tools/probe_coresidency.py): with the largest shared-memory
  // carveout the ~57 KB this kernel leaves unused become available to a small CTA of another kernel on
  // the same SM.  Off by default — DESIGN.md §4.3c records why nothing is launched beside the scan.
  static const bool max_carveout = getenv("CQS_B200_MAX_CARVEOUT") != nullptr;
  if (max_carveout) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  kern<<<grid, kThreads, Cfg::SMEM, st>>>(q);
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

template <int MODE, bool SMALLK>
static cudaError_t launch_variant(const ScanParams& p, int nv, int num_sms, cudaStream_t st) {
  switch (nv) {
    case 1: return launch_nv<MODE, 1, SMALLK>(p, num_sms, st);
    case 2: return launch_nv<MODE, 2, SMALLK>(p, num_sms, st);
    case 3: return launch_nv<MODE, 3, SMALLK>(p, num_sms, st);
    case 4: return launch_nv<MODE, 4, SMALLK>(p, num_sms, st);
    case 6: return launch_nv<MODE, 6, SMALLK>(p, num_sms, st);
    case 8: return launch_nv<MODE, 8, SMALLK>(p, num_sms, st);
    case 12: return launch_nv<MODE, 12, SMALLK>(p, num_sms, st);
    case 16: return launch_nv<MODE, 16, SMALLK>(p, num_sms, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace cqs
