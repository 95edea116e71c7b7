// scan_single.cu — kernels 1 and 3 of the north star: the single-query exact
// scan (GEMV-style streaming over f32 / bf16 rows, HBM-bound) with the top-k
// selection fused into its epilogue, so the score vector never touches HBM.
//
// Replaces, per query: the row loop of Store::search_filtered_with_notes
// (src/search/query.rs:453-482) = math::cosine_similarity per row
// (src/math.rs:11-28) + BoundedScoreHeap::push (src/search/scoring/
// candidate.rs:274-318) + into_sorted_vec (:321-329).
//
// Shape of the kernel (one persistent 288-thread CTA per SM)
//   * warp 0 is the PRODUCER: one elected lane streams the corpus with 1-D TMA
//     bulk copies (cp.async.bulk.shared.global, 24-32 KB per stage — rows are
//     contiguous in HBM so a tile of rows is one linear copy) into a 4-5 stage
//     shared-memory ring guarded by full/empty mbarriers.  ~128 KB in flight
//     per SM is the measured sweet spot on B200 (tools/bw_probe.cu: 7.4 TB/s
//     read-only versus 6.9 TB/s for the best register-staged LDG.128 loop).
//   * warps 1-8 are CONSUMERS: a warp owns whole rows of the stage; 32 lanes x
//     16-byte LDS cover 512 contiguous bytes (conflict-free), NV per row; the
//     stage is released as soon as its rows sit in registers.
//   * the query lives in registers (NV*E floats per lane) for the whole kernel.
//   * per lane 4 independent fp32 accumulators, then a 5-step shuffle
//     butterfly: the summation tree is fixed (independent of grid size), error
//     a few ulp, well inside the 1e-5 relative contract.
//   * tiles are handed out by a global atomic counter (dynamic scheduling):
//     SMs that stream faster take more tiles, all CTAs finish within ~1 us.
//   * top-k, k <= 32: every consumer warp keeps a sorted 32-entry list in
//     registers (ballot + shuffle insert; no shared memory, no barriers while
//     streaming).  k > 32: a row whose key beats the CTA's current threshold
//     is appended to a shared-memory candidate buffer that is re-selected
//     (histogram select, ~3 us) only when it might overflow.  Each CTA emits
//     <= k sorted keys; the last CTA to finish (atomic ticket) bounds the
//     global k-th key from the list heads, pulls the few survivors and writes
//     the final (score desc, row asc) result — straight into host-mapped
//     memory on the latency path.  One launch per query.
//   * optional structured filter (type / language codes) and the string-free
//     part of the reference's scoring fold run inside the same loop, before
//     the top-k (Store::search_filtered semantics).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"
#include "scan_params.cuh"

namespace cqs {

bool choose_layout(uint32_t dim, int storage, RowLayout* out) {
  if (dim == 0 || dim > 2048) return false;
  static const int kNv[] = {1, 2, 3, 4, 6, 8, 12, 16};
  auto pick = [&](uint32_t per_nv) -> int {
    uint32_t need = (dim + per_nv - 1) / per_nv;
    for (int nv : kNv)
      if ((uint32_t)nv >= need) return nv;
    return 0;
  };
  if (storage == 0) {
    int nv = pick(128);
    if (!nv) return false;
    *out = {(uint32_t)nv * 128u, 0, nv};
    return true;
  }
  // bf16: prefer 16-byte lane vectors when they do not add padding
  int nv8 = pick(256), nv4 = pick(128);
  if (nv8 && (!nv4 || (uint32_t)nv8 * 256u <= (uint32_t)nv4 * 128u)) {
    *out = {(uint32_t)nv8 * 256u, 1, nv8};
    return true;
  }
  if (!nv4) return false;
  *out = {(uint32_t)nv4 * 128u, 2, nv4};
  return true;
}

uint32_t shadow_kprime(uint32_t k) {
  static const uint32_t kp_env = getenv("CQS_B200_SHADOW_KP") ? (uint32_t)atoi(getenv("CQS_B200_SHADOW_KP")) : 0;
  if (kp_env >= k && kp_env <= kMaxK) return kp_env;   // development aid
  if (k <= 24) return 32;                      // per-warp register lists (SMALLK) still apply
  const uint32_t extra = k / 4 < 44 ? 44 : k / 4;
  const uint32_t kp = (k + extra + 31) / 32 * 32;
  return kp > kMaxK ? 0 : kp;
}

// one translation unit per (row mode, top-k variant): scan_v_*.cu
#define CQS_DECL_VARIANT(name) cudaError_t name(const ScanParams& p, int nv, int num_sms, cudaStream_t st)
CQS_DECL_VARIANT(launch_scan_m0_small); CQS_DECL_VARIANT(launch_scan_m0_large);
CQS_DECL_VARIANT(launch_scan_m1_small); CQS_DECL_VARIANT(launch_scan_m1_large);
CQS_DECL_VARIANT(launch_scan_m2_small); CQS_DECL_VARIANT(launch_scan_m2_large);

cudaError_t launch_scan_single(const ScanArgs& a, int num_sms, cudaStream_t st) {
  if (a.n_rows == 0 || a.k == 0 || a.k > kMaxK || a.n_rows > 0xFFFFFFFFull)
    return cudaErrorInvalidValue;
  ScanParams p;
  p.rows = (const uint8_t*)a.d_rows;
  p.n_rows = a.n_rows;
  p.row_bytes = (uint64_t)a.layout.ld * (a.layout.mode == 0 ? 4 : 2);
  p.query = a.d_query;
  p.bitset = a.d_bitset;
  p.k = a.k;
  p.row_base = a.row_base;
  p.partial = a.d_partial;
  p.partial_cnt = a.d_partial_cnt;
  p.done = a.d_done;
  p.col = a.d_col;
  p.excl = a.d_col + kMaxGrid;   // second half of the same scratch (written by every CTA before its ticket)
  p.out_scores = a.d_out_scores;
  p.out_rows = a.d_out_rows;
  p.out_n = a.d_out_n;
  p.trace = (unsigned long long*)a.d_trace;
  if (a.signals) p.sig = *a.signals;
  p.host_flag = a.d_host_flag;
  p.seq = a.seq;
  if (a.peer) p.peer = *a.peer;
  if (a.d_exact_rows) {
    if (a.k_out == 0 || a.k_out > a.k || a.layout.mode == 0 || a.exact_nv <= 0 || a.signals) return cudaErrorInvalidValue;
    p.exact_rows = (const uint8_t*)a.d_exact_rows;
    p.exact_nv = (uint32_t)a.exact_nv;
    p.k_out = a.k_out;
    p.max_row_delta = a.max_row_delta;
    p.max_row_norm = a.max_row_norm;
  }
  static const uint32_t chunk_env = getenv("CQS_B200_CHUNK") ? (uint32_t)atoi(getenv("CQS_B200_CHUNK")) : 0;
  p.chunk_override = chunk_env;
  const bool small = a.k <= 32;  // per-warp register lists; larger k: shared-memory accumulator
  switch (a.layout.mode) {
    case 0: return small ? launch_scan_m0_small(p, a.layout.nv, num_sms, st) : launch_scan_m0_large(p, a.layout.nv, num_sms, st);
    case 1: return small ? launch_scan_m1_small(p, a.layout.nv, num_sms, st) : launch_scan_m1_large(p, a.layout.nv, num_sms, st);
    case 2: return small ? launch_scan_m2_small(p, a.layout.nv, num_sms, st) : launch_scan_m2_large(p, a.layout.nv, num_sms, st);
  }
  return cudaErrorInvalidValue;
}

// ---- row ingest: f32 [n][dim] -> padded f32 / bf16(RNE) [n][ld] -------------
__global__ void convert_rows_kernel(const float* __restrict__ src, uint32_t dim, uint64_t n_rows,
                                    uint8_t* __restrict__ dst, uint64_t dst_row0, uint32_t ld,
                                    int is_bf16) {
  const uint64_t total = n_rows * ld;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t r = i / ld;
    uint32_t c = (uint32_t)(i - r * ld);
    float v = c < dim ? src[r * dim + c] : 0.f;
    uint64_t o = (dst_row0 + r) * ld + c;
    if (is_bf16)
      ((__nv_bfloat16*)dst)[o] = __float2bfloat16_rn(v);
    else
      ((float*)dst)[o] = v;
  }
}

cudaError_t launch_convert_rows(const float* d_src, uint32_t dim, uint64_t n_rows, void* d_dst,
                                uint64_t dst_row0, RowLayout layout, cudaStream_t st) {
  if (n_rows == 0) return cudaSuccess;
  uint64_t total = n_rows * layout.ld;
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  convert_rows_kernel<<<grid, 256, 0, st>>>(d_src, dim, n_rows, (uint8_t*)d_dst, dst_row0,
                                            layout.ld, layout.mode != 0);
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

// ---- cross-shard merge (after the all-gather) ------------------
// One CTA per query; n_lists*k <= 8192 candidates sorted by (score desc, row asc).
// Global rows need all 64 bits, so the sort runs on (ordered score, row) pairs.
struct Cand {
  uint32_t s;  // ordered score, 0 = empty
  uint64_t r;
};
__device__ __forceinline__ bool cand_before(const Cand& a, const Cand& b) {
  return a.s > b.s || (a.s == b.s && a.r < b.r);  // a ranks ahead of b
}
constexpr uint32_t kMergeCap = 8192;
__global__ void __launch_bounds__(512) merge_topk_kernel(const MergeArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint32_t* s_s = (uint32_t*)smem;                       // [P]
  uint64_t* s_r = (uint64_t*)(smem + sizeof(uint32_t) * kMergeCap);  // [P]
  const uint32_t qi = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const uint32_t total = a.n_lists * a.k;
  const uint32_t P = next_pow2(total);
  for (uint32_t i = tid; i < P; i += T) {
    uint32_t s = 0;
    uint64_t r = ~0ull;
    if (i < total) {
      uint32_t l = i / a.k, j = i - l * a.k;
      size_t off = ((size_t)l * a.n_queries + qi) * a.k + j;
      r = a.d_rows[off];
      uint32_t bits = __float_as_uint(a.d_scores[off]);
      if (r != ~0ull && finite_bits(bits)) s = ordered_u32(bits);
    }
    s_s[i] = s;
    s_r[i] = r;
  }
  __syncthreads();
  for (uint32_t size = 2; size <= P; size <<= 1)
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = tid; i < (P >> 1); i += T) {
        uint32_t pos = 2 * i - (i & (stride - 1));
        Cand x{s_s[pos], s_r[pos]}, y{s_s[pos + stride], s_r[pos + stride]};
        bool dir = ((pos & size) == 0);
        if (cand_before(y, x) == dir) {
          s_s[pos] = y.s; s_r[pos] = y.r;
          s_s[pos + stride] = x.s; s_r[pos + stride] = x.r;
        }
      }
      __syncthreads();
    }
  __shared__ uint32_t s_n;
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (uint32_t i = tid; i < a.k; i += T) {
    bool valid = i < P && s_s[i] != 0;
    a.d_out_scores[(size_t)qi * a.k + i] =
        valid ? __uint_as_float(unordered_u32(s_s[i])) : __uint_as_float(0xFF800000u);
    a.d_out_rows[(size_t)qi * a.k + i] = valid ? s_r[i] : ~0ull;
    if (valid) atomicAdd(&s_n, 1u);
  }
  __syncthreads();
  if (tid == 0) a.d_out_n[qi] = s_n;
}

cudaError_t launch_merge_topk(const MergeArgs& a, cudaStream_t st) {
  if (a.n_queries == 0 || a.k == 0) return cudaSuccess;
  if ((uint64_t)a.n_lists * a.k > kMergeCap) return cudaErrorInvalidValue;
  size_t smem = (sizeof(uint32_t) + sizeof(uint64_t)) * kMergeCap;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    attr_set = true;
  }
  merge_topk_kernel<<<a.n_queries, 512, smem, st>>>(a);
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

}  // namespace cqs
