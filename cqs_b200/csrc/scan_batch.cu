// scan_batch.cu — kernel 2 of the north star: the batched scan.  A batch of
// queries against the corpus really is a dense contraction, so it runs on the
// 5th-generation tensor cores: bf16 x bf16 -> fp32 with tcgen05.mma, operands
// staged by TMA (128-byte swizzle), accumulators in TMEM, and the top-k
// selection fused into the epilogue straight out of TMEM — the nq x N score
// matrix is never written.
//
// Replaces nq calls of VectorIndex::search (src/index.rs:146) /
// Store::search_filtered_with_notes's row loop (src/search/query.rs:453-482),
// e.g. the 218 sequential searches of the eval runner
// (src/cli/commands/eval/runner.rs:279-367).
//
// Orientation: QUERIES are the M side (A operand, 128 per tile = the 128 TMEM
// lanes), CORPUS ROWS are the N side (B operand, 256 per MMA).  An epilogue
// thread therefore owns one query: its threshold lives in a register and a
// whole 32-column TMEM load is rejected with one max-tree.
//
// Exactness: tensor-core scores use bf16-rounded queries, so they only
// GENERATE candidates (k' >= 64 per query, kept in per-query pools in HBM under
// per-query thresholds).  Thresholds are tightened between rounds of
// geometrically growing row ranges (round 0 is dense: every score of the first
// 18,944 rows is kept and selected from).  The k' candidates are then re-scored
// with the f32 query by the SAME per-row arithmetic as the single-query scan
// (scan_single.cu) and re-ranked; a query whose k-th exact score is not
// separated from the pool's cut-off by the rigorous rounding bound is flagged
// and re-run through the exact single-query kernel by the caller.  Results are
// identical to nq calls of cqs_b200_search.
#define CQS_FORCE_NETWORK 1   // the batch kernels keep the bitonic network: see DESIGN.md §4.6 (chunk sort, update_thr_kernel)
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "internal.h"

namespace cqs {

namespace {

constexpr uint32_t kBM = 128;           // queries per tile (UMMA M)
constexpr uint32_t kBN = 256;           // corpus rows per tile (UMMA N)
constexpr uint32_t kBK = 64;            // K elements per pipeline stage (one 128-byte swizzle row)
constexpr uint32_t kUmmaK = 16;
constexpr uint32_t kStages = 4;
constexpr uint32_t kABytes = kBM * kBK * 2;   // 16 KB
constexpr uint32_t kBBytes = kBN * kBK * 2;   // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr uint32_t kEpiParts = 4;          // column parts per TMEM lane quarter
constexpr uint32_t kBatchThreads = 128 + 128 * kEpiParts;  // 4 control warps + 4*kEpiParts epilogue warps
constexpr uint32_t kAccStages = 2;            // 2 x 256 TMEM columns
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kBatchSmem = kStages * kStageBytes + 1024;  // + alignment slack

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
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* tm, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// K-major operand tile, 128-byte swizzle: rows at 128 B pitch, 8-row atoms 1024 B apart.
// SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14) | LBO [16,30) |
// SBO [32,46) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// InstrDescriptor: c=F32 [4,6)=1 | a=BF16 [7,10)=1 | b=BF16 [10,13)=1 | K-major A,B |
// N>>3 [17,23) | M>>4 [24,29)
constexpr uint32_t kIdesc =
    (1u << 4) | (1u << 7) | (1u << 10) | ((kBN >> 3) << 17) | ((kBM >> 4) << 24);

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct BatchParams {
  uint64_t n_rows;        // rows of the shard
  uint64_t row_begin;     // this launch covers rows [row_begin, row_end)
  uint64_t row_end;
  uint32_t n_qt;          // query tiles (nq_pad / 128)
  uint32_t nq;            // real queries
  uint32_t k_blocks;      // ld / 64
  int dense;              // round 0: keep every score at slot (row - row_begin)
  const uint32_t* bitset; // nullable
  const ckey_t* thr;      // [nq_pad] per-query threshold key (0 = none)
  ckey_t* cand;           // [nq_pad][cap]
  uint32_t* cnt;          // [nq_pad]
  uint32_t* overflow;     // [nq_pad]
  uint32_t cap;
  // sparse rounds: candidates go to per-(query, CTA) sub-pools so that no global
  // atomic sits on the epilogue's path (a query is owned by exactly one thread of a CTA)
  ckey_t* sub;            // [nq_pad][grid][kSub]
  uint8_t* subcnt;        // [grid][nq_pad]
  uint32_t nq_pad;
};
constexpr uint32_t kSub = 64;        // slots per (query, CTA) sub-pool and round
constexpr uint32_t kMaxGridB = 160;  // sub-pool stride bound (SMs)

// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = idle,
// 4..11 = epilogue: warp % 4 selects the TMEM lane quarter (32 queries), (warp-4)/4 the
// column half (128 corpus rows): two threads share a query and split its columns.
// A work item is (row chunk of 256, query tile of 128); items are dealt
// round-robin so that neighbouring CTAs share the same row chunk through L2.
__global__ void __launch_bounds__(kBatchThreads, 1)
scan_batch_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmR,
                  const BatchParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];
  __shared__ __align__(8) uint64_t s_tfull[kAccStages], s_tempty[kAccStages];
  __shared__ uint32_t s_tmem_base;
  __shared__ uint32_t s_subcnt[kBatchMaxQ];

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < kBatchMaxQ; i += kBatchThreads) s_subcnt[i] = 0;
  const uint64_t n_chunks = (p.row_end - p.row_begin + kBN - 1) / kBN;
  const uint64_t n_items = n_chunks * p.n_qt;

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < kStages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    for (uint32_t a = 0; a < kAccStages; ++a) {
      mbar_init(&s_tfull[a], 1);
      mbar_init(&s_tempty[a], 4 * kEpiParts);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&s_tmem_base)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint64_t chunk = item / p.n_qt;
        const uint32_t qt = (uint32_t)(item - chunk * p.n_qt);
        const int row0 = (int)(p.row_begin + chunk * kBN);
        for (uint32_t kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&s_empty[s], ph ^ 1);
          mbar_expect_tx(&s_full[s], kStageBytes);
          uint8_t* st = smem + (size_t)s * kStageBytes;
          tma_load_2d(st, &tmQ, (int)(kb * kBK), (int)(qt * kBM), &s_full[s]);
          tma_load_2d(st + kABytes, &tmR, (int)(kb * kBK), row0, &s_full[s]);
          if (++s == kStages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t s = 0, ph = 0, acc = 0, aph = 0;
      for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait(&s_tempty[acc], aph ^ 1);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBN;
        for (uint32_t kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&s_full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)s * kStageBytes);
          const uint64_t a_desc = make_sw128_desc(a_addr);
          const uint64_t b_desc = make_sw128_desc(a_addr + kABytes);
#pragma unroll
          for (uint32_t kk = 0; kk < kBK / kUmmaK; ++kk) {
            // advance 16 K-elements = 32 bytes inside the swizzle row: +2 in the encoded address
            umma_f16(d_tmem, a_desc + 2 * kk, b_desc + 2 * kk, kIdesc, (kb | kk) != 0);
          }
          umma_commit(&s_empty[s]);            // smem slot reusable once these MMAs retire
          if (++s == kStages) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&s_tfull[acc]);            // accumulator complete -> epilogue
        if (++acc == kAccStages) {
          acc = 0;
          aph ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread = query, columns = corpus rows =====
    const uint32_t quarter = warp & 3, half = (warp - 4) >> 2;  // half = column part
    uint32_t acc = 0, aph = 0;
    for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint64_t chunk = item / p.n_qt;
      const uint32_t qt = (uint32_t)(item - chunk * p.n_qt);
      const uint64_t row0 = p.row_begin + chunk * kBN + half * (kBN / kEpiParts);
      const uint32_t q = qt * kBM + quarter * 32 + lane;
      const bool active = q < p.nq;
      ckey_t thr_key = 0;
      float thr_score = __uint_as_float(0xFF800000u);  // -inf
      if (active && !p.dense) {
        thr_key = __ldcg(p.thr + q);
        if (thr_key) thr_score = key_score(thr_key);
      }
      mbar_wait(&s_tfull[acc], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kBN + half * (kBN / kEpiParts);
      ckey_t* my_cand = p.cand + (size_t)q * p.cap;
      ckey_t* my_sub = p.sub + ((size_t)q * gridDim.x + blockIdx.x) * kSub;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < kBN / kEpiParts; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        if (!active) continue;
        if (p.dense) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const uint64_t r = row0 + c0 + c;
            if (r < p.row_end) {
              bool ok = finite_bits(v[c]);
              if (ok && p.bitset) ok = (__ldg(p.bitset + (r >> 5)) >> (r & 31)) & 1u;
              my_cand[r - p.row_begin] = ok ? make_key(__uint_as_float(v[c]), (uint32_t)r) : 0;
            }
          }
          continue;
        }
        float m = __uint_as_float(v[0]);
#pragma unroll
        for (int c = 1; c < 32; ++c) m = fmaxf(m, __uint_as_float(v[c]));
        if (!(m >= thr_score)) continue;  // the common case: nothing in these 32 rows qualifies
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float sc = __uint_as_float(v[c]);
          if (sc >= thr_score && finite_bits(v[c])) {
            const uint64_t r = row0 + c0 + c;
            if (r < p.row_end) {
              const ckey_t key = make_key(sc, (uint32_t)r);
              bool ok = key > thr_key;
              if (ok && p.bitset) ok = (__ldg(p.bitset + (r >> 5)) >> (r & 31)) & 1u;
              if (ok) {
                const uint32_t slot = atomicAdd(&s_subcnt[q], 1u);  // shared with the other half's thread
                if (slot < kSub) my_sub[slot] = key; else p.overflow[q] = 1;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_tempty[acc]);
      if (++acc == kAccStages) {
        acc = 0;
        aph ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (!p.dense) {
    // publish this CTA's sub-pool fill levels
    for (uint32_t q = threadIdx.x; q < p.nq_pad; q += kBatchThreads)
      p.subcnt[(size_t)blockIdx.x * p.nq_pad + q] = (uint8_t)min(s_subcnt[q], kSub);
  }
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(kTmemCols)
                 : "memory");
  }
}

// ---- query preparation: f32 -> bf16 (RNE), zero padding rows, norms, state reset -------
__global__ void prep_queries_kernel(const float* __restrict__ q32, uint32_t nq, uint32_t nq_pad,
                                    uint32_t ld, __nv_bfloat16* __restrict__ q16,
                                    float* __restrict__ qnorm, float* __restrict__ dqnorm, ckey_t* thr,
                                    uint32_t* cnt, uint32_t* overflow, uint32_t dense_count) {
  const uint32_t q = blockIdx.x;
  float acc = 0.f, dacc = 0.f;
  for (uint32_t i = threadIdx.x; i < ld; i += blockDim.x) {
    float v = q < nq ? q32[(size_t)q * ld + i] : 0.f;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    q16[(size_t)q * ld + i] = b;
    const float d = v - __bfloat162float(b);   // exact (Sterbenz): the rounding error of this element
    acc += v * v;
    dacc += d * d;
  }
  __shared__ float red[32], dred[32];
  for (int off = 16; off > 0; off >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, off);
    dacc += __shfl_xor_sync(0xffffffffu, dacc, off);
  }
  if ((threadIdx.x & 31) == 0) {
    red[threadIdx.x >> 5] = acc;
    dred[threadIdx.x >> 5] = dacc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f, dt = 0.f;
    for (uint32_t w = 0; w < (blockDim.x + 31) / 32; ++w) {
      t += red[w];
      dt += dred[w];
    }
    qnorm[q] = sqrtf(t);
    dqnorm[q] = sqrtf(dt);      // |q - bf16(q)|_2, measured: what the tensor-core scan did not see of the query
    thr[q] = 0;
    cnt[q] = q < nq ? dense_count : 0;
    overflow[q] = 0;
  }
  (void)nq_pad;
}

// ---- threshold update: keep the k' best keys of every pool, publish the k'-th ------------
constexpr uint32_t kUpdThreads = 256;
constexpr uint32_t kUpdCap = 4096;
__global__ void __launch_bounds__(kUpdThreads) update_thr_kernel(ckey_t* cand, uint32_t* cnt,
                                                                 ckey_t* thr, uint32_t cap,
                                                                 uint32_t kprime, const ckey_t* sub,
                                                                 const uint8_t* subcnt,
                                                                 uint32_t sub_grid, uint32_t nq_pad) {
  __shared__ __align__(16) ckey_t s_buf[kUpdCap];
  __shared__ __align__(8) uint32_t s_hist[kSelBuckets + 96];
  __shared__ uint32_t s_cnt;
  __shared__ ckey_t s_thr;
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  TopK tk{s_buf, &s_cnt, &s_thr, kUpdCap, Group{tid, kUpdThreads, 0}, s_hist};
  tk.init();
  __syncthreads();
  ckey_t* pool = cand + (size_t)q * cap;
  const uint32_t n = min(cnt[q], cap);
  // the pool itself (k' keys, or the 18,944 keys of the dense round): chunks of cap/2 pushes,
  // shrunk by the histogram select in between (a superset of the k' best survives)
  constexpr uint32_t kStep = kUpdCap / 2;
  for (uint32_t base = 0; base < n; base += kStep) {
    if (s_cnt + kStep > kUpdCap) tk.template select<kUpdCap / kUpdThreads>(kprime);
    const ckey_t t = s_thr;
    const uint32_t end = min(n, base + kStep);
    for (uint32_t i = base + tid; i < end; i += kUpdThreads) {
      const ckey_t key = pool[i];
      if (key > t) tk.push(key);
    }
    __syncthreads();
  }
  // this round's sub-pools: one thread per (CTA) sub-pool walks its few keys; at most
  // kSub * 32 = cap/2 pushes between two selects
  for (uint32_t c0 = 0; c0 < sub_grid; c0 += 32) {
    if (s_cnt + 32 * kSub > kUpdCap) tk.template select<kUpdCap / kUpdThreads>(kprime);
    const ckey_t t = s_thr;
    const uint32_t c1 = min(sub_grid, c0 + 32);
    // 8 threads per sub-pool
    for (uint32_t i = tid; i < (c1 - c0) * 8; i += kUpdThreads) {
      const uint32_t c = c0 + (i >> 3);
      const uint32_t m = subcnt[(size_t)c * nq_pad + q];
      const ckey_t* sp = sub + ((size_t)q * sub_grid + c) * kSub;
      for (uint32_t j = i & 7; j < m; j += 8) {
        const ckey_t key = sp[j];
        if (key > t) tk.push(key);
      }
    }
    __syncthreads();
  }
  tk.compact(kprime);
  const uint32_t m = s_cnt;
  for (uint32_t i = tid; i < m; i += kUpdThreads) pool[i] = s_buf[i];
  if (tid == 0) {
    cnt[q] = m;
    thr[q] = (m >= kprime) ? s_buf[kprime - 1] : 0;
  }
}

// ---- exact re-scoring: f32 query x bf16 row, the single-query kernel's arithmetic ---------
// One warp per (query, candidate).  The per-lane element assignment, the four
// accumulators and the butterfly are those of scan_topk_kernel (LaneVec<1>/<2>),
// so the score is bit-identical to what cqs_b200_search returns for that row.
template <int MODE>
__device__ __forceinline__ float exact_row_dot(const uint8_t* row, const float* q, uint32_t nv,
                                               uint32_t lane) {
  constexpr int E = (MODE == 1) ? 8 : 4;
  constexpr int BYTES = (MODE == 2) ? 8 : 16;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (uint32_t v = 0; v < nv; ++v) {
    const uint8_t* rp = row + ((size_t)v * 32 + lane) * BYTES;
    const float* qp = q + ((size_t)v * 32 + lane) * E;
    if (MODE == 0) {
      const float4 x = *reinterpret_cast<const float4*>(rp);
      acc[0] = fmaf(x.x, qp[0], acc[0]);
      acc[1] = fmaf(x.y, qp[1], acc[1]);
      acc[2] = fmaf(x.z, qp[2], acc[2]);
      acc[3] = fmaf(x.w, qp[3], acc[3]);
    } else if (MODE == 1) {
      const uint4 x = *reinterpret_cast<const uint4*>(rp);
      const uint32_t r[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[(2 * i) & 3] = fmaf(__uint_as_float(r[i] << 16), qp[2 * i], acc[(2 * i) & 3]);
        acc[(2 * i + 1) & 3] =
            fmaf(__uint_as_float(r[i] & 0xFFFF0000u), qp[2 * i + 1], acc[(2 * i + 1) & 3]);
      }
    } else {
      const uint2 x = *reinterpret_cast<const uint2*>(rp);
      acc[0] = fmaf(__uint_as_float(x.x << 16), qp[0], acc[0]);
      acc[1] = fmaf(__uint_as_float(x.x & 0xFFFF0000u), qp[1], acc[1]);
      acc[2] = fmaf(__uint_as_float(x.y << 16), qp[2], acc[2]);
      acc[3] = fmaf(__uint_as_float(x.y & 0xFFFF0000u), qp[3], acc[3]);
    }
  }
  float s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  return s;
}

__global__ void __launch_bounds__(256) rescore_kernel(const uint8_t* __restrict__ rows,
                                                      uint32_t ld, int mode, uint32_t nv,
                                                      const float* __restrict__ q32,
                                                      ckey_t* cand, const uint32_t* cnt,
                                                      uint32_t cap, uint32_t kprime) {
  const uint32_t q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n = min(cnt[q], kprime);
  ckey_t* pool = cand + (size_t)q * cap;
  const float* qp = q32 + (size_t)q * ld;
  for (uint32_t j = warp; j < n; j += 8) {
    const ckey_t key = pool[j];
    const uint32_t r = key_row(key);
    const uint8_t* row = rows + (size_t)r * ld * (mode == 0 ? 4 : 2);
    const float s = (mode == 0)   ? exact_row_dot<0>(row, qp, nv, lane)
                    : (mode == 1) ? exact_row_dot<1>(row, qp, nv, lane)
                                  : exact_row_dot<2>(row, qp, nv, lane);
    // exact keys go to the second half of the pool ([cap/2, cap/2 + n)); a non-finite
    // exact score is dropped like in the single-query kernel
    if (lane == 0) pool[cap / 2 + j] = finite_bits(__float_as_uint(s)) ? make_key(s, r) : 0;
  }
}

// ---- final selection + exactness check ---------------------------------------------------
__global__ void __launch_bounds__(kUpdThreads) final_select_kernel(
    const ckey_t* cand, const uint32_t* cnt, const ckey_t* thr, const uint32_t* overflow,
    const float* qnorm, const float* dqnorm, float max_row_norm, float max_row_delta, uint32_t cap, uint32_t kprime, uint32_t k,
    uint64_t row_base, float* out_scores, uint64_t* out_rows, uint32_t* out_n, uint32_t* flags) {
  __shared__ __align__(16) ckey_t s_buf[kUpdCap];
  __shared__ uint32_t s_cnt;
  __shared__ ckey_t s_thr;
  const uint32_t q = blockIdx.x, tid = threadIdx.x;
  TopK tk{s_buf, &s_cnt, &s_thr, kUpdCap, Group{tid, kUpdThreads, 0}};
  tk.init();
  __syncthreads();
  const uint32_t n = min(cnt[q], kprime);
  const ckey_t* exact = cand + (size_t)q * cap + cap / 2;
  for (uint32_t i = tid; i < n; i += kUpdThreads) {
    const ckey_t key = exact[i];
    if (key) tk.push(key);
  }
  tk.compact(k);
  const uint32_t m = s_cnt;
  for (uint32_t i = tid; i < k; i += kUpdThreads) {
    const bool valid = i < m;
    out_scores[(size_t)q * k + i] = valid ? key_score(s_buf[i]) : __uint_as_float(0xFF800000u);
    out_rows[(size_t)q * k + i] = valid ? row_base + key_row(s_buf[i]) : ~0ull;
  }
  if (tid == 0) {
    out_n[q] = m;
    // Rows outside the pool have approx score <= the pool's cut-off.  With q16 = bf16(q),
    // r16 = the scanned bf16 row, r = the exact row:  q.r - q16.r16 = q.(r - r16) + (q - q16).r16,
    // so |exact - approx| <= |q| * D + |q - q16| * R + fp32 accumulation, where D = max_row_delta
    // (MEASURED max |r - r16|_2 over the corpus; 0 when the bf16 rows are the corpus),
    // |q - q16| is measured per query, R bounds |r16|, and 2^-13 |q| R covers 768-term fp32
    // accumulation with truncation (768 * 2^-23 < 2^-13).  Cauchy-Schwarz on measured norms:
    // rigorous, and ~2.5x tighter than the worst-case unit roundoff 2^-8.  If the k-th exact
    // score clears cut-off + E nothing was missed.
    // flag codes: 1 = pool overflowed, 2 = too few exact survivors, 3 = margin not proven
    uint32_t flag = overflow[q] ? 1u : 0u;
    const ckey_t t = thr[q];
    if (t != 0 && !flag) {
      const float cut = key_score(t);
      const float E = 1.001f * (qnorm[q] * max_row_delta + dqnorm[q] * max_row_norm +
                                0.0001220703125f * qnorm[q] * max_row_norm);
      if (m < k) flag = 2;  // the pool was full yet fewer than k exact survivors: be safe
      else if (!(key_score(s_buf[k - 1]) > cut + E)) flag = 3;
    }
    flags[q] = flag;
  }
}

__global__ void max_row_norm_kernel(const uint8_t* __restrict__ rows, uint64_t n_rows, uint32_t ld,
                                    int is_bf16, float* out) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  float best = 0.f;
  for (uint64_t r = warp; r < n_rows; r += nwarps) {
    float acc = 0.f;
    for (uint32_t i = lane; i < ld; i += 32) {
      float v = is_bf16 ? __bfloat162float(((const __nv_bfloat16*)rows)[r * ld + i])
                        : ((const float*)rows)[r * ld + i];
      acc += v * v;
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (acc == acc && acc < 3.0e38f) best = fmaxf(best, acc);
  }
  if (lane == 0 && best > 0.f) atomicMax((int*)out, __float_as_int(sqrtf(best)));
}

// max over rows of |f32 master row - bf16 shadow row|_2; one warp per row
__global__ void max_row_delta_kernel(const float* __restrict__ r32, const __nv_bfloat16* __restrict__ r16,
                                     uint64_t n_rows, uint32_t ld, float* out) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  float best = 0.f;
  for (uint64_t r = warp; r < n_rows; r += nwarps) {
    float acc = 0.f;
    for (uint32_t i = lane; i < ld; i += 32) {
      const float d = r32[r * ld + i] - __bfloat162float(r16[r * ld + i]);
      acc += d * d;
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (acc == acc && acc < 3.0e38f) best = fmaxf(best, acc);
  }
  if (lane == 0 && best > 0.f) atomicMax((int*)out, __float_as_int(sqrtf(best)));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// 2-D bf16 tensor [rows][ld], box {64, box_rows}, 128-byte swizzle, OOB rows read as zero
static bool make_tmap(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t ld,
                      uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {ld, rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {kBK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

cudaError_t launch_max_row_norm(const void* d_rows, uint64_t n_rows, RowLayout layout, float* d_out,
                                cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(float), st);
  if (e != cudaSuccess || n_rows == 0) return e;
  max_row_norm_kernel<<<148 * 4, 256, 0, st>>>((const uint8_t*)d_rows, n_rows, layout.ld,
                                               layout.mode != 0, d_out);
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t launch_max_row_delta(const void* d_rows_f32, const void* d_rows_bf16, uint64_t n_rows, uint32_t ld,
                                 float* d_out, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(float), st);
  if (e != cudaSuccess || n_rows == 0) return e;
  max_row_delta_kernel<<<148 * 4, 256, 0, st>>>((const float*)d_rows_f32, (const __nv_bfloat16*)d_rows_bf16, n_rows,
                                               ld, d_out);
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

// Candidates kept per query.  The pool must reach down to (k-th exact score - E) for the proof to
// succeed; with an f32 master behind bf16 rows E carries the row rounding distance as well as the
// query's (~2x), so the pool is deeper there (measured on clustered rows, 2M..10M: the exact-score gap
// rank 20 -> rank 64 is 0.0036..0.0043 against E = 0.0037, rank 20 -> rank 128 is 0.0072).
uint32_t batch_kprime(uint32_t k, bool shadow) {
  const uint32_t floor_extra = shadow ? 108 : 44, frac = shadow ? k / 2 : k / 4;
  uint32_t extra = frac < floor_extra ? floor_extra : frac;
  uint32_t kp = (k + extra + 31) / 32 * 32;
  return kp > 1536 ? 1536 : kp;
}

size_t batch_scratch_bytes(uint32_t nq_pad, uint32_t ld) {
  size_t b = 0;
  b += (size_t)nq_pad * ld * 2;                 // bf16 queries
  b += (size_t)nq_pad * kBatchCap * 8;          // candidate pools
  b += (size_t)nq_pad * (8 + 4 + 4 + 4 + 4);    // thr, cnt, overflow, qnorm, flags
  b += (size_t)nq_pad * kMaxGridB * kSub * 8;   // sub-pools
  b += (size_t)nq_pad * kMaxGridB;              // sub-pool fill levels
  return b + 4096;
}

cudaError_t launch_scan_batch(const BatchArgs& a, int num_sms, cudaStream_t st) {
  if (a.nq == 0 || a.k == 0 || a.k > kMaxK || a.n_rows == 0 || a.n_rows >= (1ull << 31) ||
      a.layout.mode == 0 || a.nq > kBatchMaxQ || a.exact_layout.ld != a.layout.ld)
    return cudaErrorInvalidValue;
  const uint32_t nq_pad = (a.nq + kBM - 1) / kBM * kBM;
  const uint32_t ld = a.layout.ld;
  const uint32_t kprime = batch_kprime(a.k, a.max_row_delta > 0.f);
  // carve the scratch
  uint8_t* base = (uint8_t*)a.d_scratch;
  __nv_bfloat16* q16 = (__nv_bfloat16*)base;
  base += ((size_t)nq_pad * ld * 2 + 255) / 256 * 256;
  ckey_t* cand = (ckey_t*)base;
  base += (size_t)nq_pad * kBatchCap * 8;
  ckey_t* thr = (ckey_t*)base;
  base += (size_t)nq_pad * 8;
  uint32_t* cnt = (uint32_t*)base;
  base += (size_t)nq_pad * 4;
  uint32_t* overflow = (uint32_t*)base;
  base += (size_t)nq_pad * 4;
  float* qnorm = (float*)base;
  base += (size_t)nq_pad * 4;
  float* dqnorm = (float*)base;
  base += (size_t)nq_pad * 4;
  base = (uint8_t*)(((uintptr_t)base + 255) & ~(uintptr_t)255);
  ckey_t* sub = (ckey_t*)base;
  base += (size_t)nq_pad * kMaxGridB * kSub * 8;
  uint8_t* subcnt = base;
  if (num_sms > (int)kMaxGridB) num_sms = kMaxGridB;

  CUtensorMap tmQ, tmR;
  if (!make_tmap(&tmQ, q16, nq_pad, ld, kBM) || !make_tmap(&tmR, a.d_rows, a.n_rows, ld, kBN))
    return cudaErrorNotSupported;

  // CQS_B200_DEBUG_SYNC (development aid): synchronise after every launch and name the one that failed
  static const bool dbg = getenv("CQS_B200_DEBUG_SYNC") != nullptr;
  auto check = [&](const char* what) {
    if (!dbg) return cudaSuccess;
    const cudaError_t ee = cudaStreamSynchronize(st);
    if (ee != cudaSuccess) fprintf(stderr, "[cqs_b200] %s failed: %s\n", what, cudaGetErrorString(ee));
    return ee;
  };
  const uint64_t r0 = a.n_rows < kBatchDenseRows ? a.n_rows : kBatchDenseRows;
  prep_queries_kernel<<<nq_pad, 128, 0, st>>>(a.d_queries, a.nq, nq_pad, ld, q16, qnorm, dqnorm, thr, cnt,
                                              overflow, (uint32_t)r0);
  cudaError_t e = cudaFuncSetAttribute(scan_batch_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kBatchSmem);
  if (e != cudaSuccess) return e;
  uint64_t launches = 1;
  BatchParams p;
  p.n_rows = a.n_rows;
  p.n_qt = nq_pad / kBM;
  p.nq = a.nq;
  p.k_blocks = ld / kBK;
  p.bitset = a.d_bitset;
  p.thr = thr;
  p.cand = cand;
  p.cnt = cnt;
  p.overflow = overflow;
  p.cap = kBatchCap;
  p.sub = sub;
  p.subcnt = subcnt;
  p.nq_pad = nq_pad;
  uint64_t begin = 0, end = r0;
  int round = 0;
  while (begin < a.n_rows) {
    p.row_begin = begin;
    p.row_end = end;
    p.dense = (round == 0);
    const uint64_t items = ((end - begin + kBN - 1) / kBN) * p.n_qt;
    const int grid = (int)(items < (uint64_t)num_sms ? items : (uint64_t)num_sms);
    scan_batch_kernel<<<grid, kBatchThreads, kBatchSmem, st>>>(tmQ, tmR, p);
    if ((e = check("scan_batch_kernel")) != cudaSuccess) return e;
    update_thr_kernel<<<nq_pad, kUpdThreads, 0, st>>>(cand, cnt, thr, kBatchCap, kprime, sub, subcnt,
                                                      p.dense ? 0u : (uint32_t)grid, nq_pad);
    if ((e = check("update_thr_kernel")) != cudaSuccess) return e;
    launches += 2;
    begin = end;
    end = (end * 9 < a.n_rows) ? end * 9 : a.n_rows;
    ++round;
  }
  rescore_kernel<<<a.nq, 256, 0, st>>>((const uint8_t*)a.d_exact_rows, ld, a.exact_layout.mode,
                                       a.exact_layout.nv, a.d_queries, cand, cnt, kBatchCap, kprime);
  if ((e = check("rescore_kernel")) != cudaSuccess) return e;
  final_select_kernel<<<a.nq, kUpdThreads, 0, st>>>(cand, cnt, thr, overflow, qnorm, dqnorm,
                                                    a.max_row_norm, a.max_row_delta, kBatchCap, kprime, a.k,
                                                    a.row_base, a.d_out_scores, a.d_out_rows,
                                                    a.d_out_n, a.d_flags);
  if ((e = check("final_select_kernel")) != cudaSuccess) return e;
  launches += 2;
  g_kernel_launches.fetch_add(launches, std::memory_order_relaxed);
  return cudaGetLastError();
}

}  // namespace cqs
