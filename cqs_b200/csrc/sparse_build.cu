// sparse_build.cu — SPLADE inverted-index build on the device.
//
// Replaces SpladeIndex::build (src/splade/index.rs:177-221): the reference walks
// every chunk's sparse vector and pushes (chunk_idx, weight) onto the posting
// list of each token, i.e. it transposes the doc-major `sparse_vectors` rows
// (src/store/sparse.rs:342, ORDER BY chunk_id, token_id) into token-major lists
// whose entries are in ascending chunk order.  Here that is a stable counting
// sort by token id over the CSR entries, entirely on the GPU:
//
//   K1 count    P chunks of consecutive entries, one CTA each; a shared-memory
//               histogram over the whole vocabulary (30,522 tokens = 119 KB)
//               -> counts[chunk][token]
//   K2 colscan  per token: exclusive prefix over the chunks (coalesced across
//               tokens) -> every chunk's first slot inside every token's list
//   K3 tptr     exclusive scan of the token totals -> tptr[vocab+1]
//   K4 scatter  every CTA replays its chunk IN ORDER, 1024 entries per tile; the
//               32 warps of a tile take turns (32 barriers per tile) and, inside
//               a warp, lanes holding the same token are ranked with
//               match.any — so equal tokens keep their entry (= doc) order.
//               Writes doc[pos] and (doc, weight) post[pos].
//   K5 check    a doc that lists a token twice (two adjacent equal docs inside
//               one list) is rejected: the search kernel's lanes would collide.
//
// HBM-bound integer work: reads 12 B per entry twice, writes 12 B per entry
// (scattered over the vocabulary's lists).
#include <cuda_runtime.h>
#include <stdint.h>

#include "internal.h"

namespace cqs {

constexpr int kBuildThreads = 1024;

__device__ __forceinline__ uint64_t chunk_begin(uint64_t nnz, uint32_t P, uint32_t c) {
  // chunk boundaries on multiples of 1024 entries so tiles never straddle chunks
  const uint64_t tiles = (nnz + kBuildThreads - 1) / kBuildThreads;
  const uint64_t per = (tiles + P - 1) / P;
  const uint64_t b = (uint64_t)c * per * kBuildThreads;
  return b < nnz ? b : nnz;
}

__global__ void __launch_bounds__(kBuildThreads) sparse_count_kernel(const uint32_t* __restrict__ tok,
                                                                     uint64_t nnz, uint32_t vocab,
                                                                     uint32_t* __restrict__ counts,
                                                                     uint32_t* __restrict__ err) {
  extern __shared__ uint32_t s_hist[];
  for (uint32_t t = threadIdx.x; t < vocab; t += blockDim.x) s_hist[t] = 0;
  __syncthreads();
  const uint64_t e0 = chunk_begin(nnz, gridDim.x, blockIdx.x), e1 = chunk_begin(nnz, gridDim.x, blockIdx.x + 1);
  for (uint64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const uint32_t t = __ldg(tok + e);
    if (t >= vocab) atomicExch(err, 1u);
    else atomicAdd(&s_hist[t], 1u);
  }
  __syncthreads();
  uint32_t* out = counts + (size_t)blockIdx.x * vocab;
  for (uint32_t t = threadIdx.x; t < vocab; t += blockDim.x) out[t] = s_hist[t];
}

__global__ void sparse_colscan_kernel(uint32_t* __restrict__ counts, uint32_t P, uint32_t vocab,
                                      uint64_t* __restrict__ totals) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= vocab) return;
  uint64_t run = 0;
  for (uint32_t c = 0; c < P; ++c) {
    const uint32_t v = counts[(size_t)c * vocab + t];
    counts[(size_t)c * vocab + t] = (uint32_t)run;  // a posting list holds < 2^32 entries (n_docs < 2^32)
    run += v;
  }
  totals[t] = run;
}

// tptr[t] = sum of totals[0..t); tptr[vocab] = nnz.  One CTA.
__global__ void __launch_bounds__(1024) sparse_tptr_kernel(const uint64_t* __restrict__ totals, uint32_t vocab,
                                                           uint64_t* __restrict__ tptr) {
  __shared__ uint64_t s_warp[32];
  __shared__ uint64_t s_carry;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < vocab; base += 1024) {
    const uint32_t t = base + threadIdx.x;
    const uint64_t v = t < vocab ? totals[t] : 0;
    uint64_t incl = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += u;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t before = s_carry;
    for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
    if (t < vocab) tptr[t] = before + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) tptr[vocab] = s_carry;
}

__global__ void __launch_bounds__(kBuildThreads) sparse_scatter_kernel(
    const uint64_t* __restrict__ indptr, const uint32_t* __restrict__ tok, const float* __restrict__ w,
    uint64_t n_docs, uint64_t nnz, uint32_t vocab, const uint32_t* __restrict__ counts,
    const uint64_t* __restrict__ tptr, uint32_t* __restrict__ out_doc, uint2* __restrict__ out_post) {
  extern __shared__ uint32_t s_cur[];  // next free slot of this chunk inside every token's list
  const uint32_t* base = counts + (size_t)blockIdx.x * vocab;
  for (uint32_t t = threadIdx.x; t < vocab; t += blockDim.x) s_cur[t] = base[t];
  __syncthreads();
  const uint64_t e0 = chunk_begin(nnz, gridDim.x, blockIdx.x), e1 = chunk_begin(nnz, gridDim.x, blockIdx.x + 1);
  if (e0 >= e1) return;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // docs this chunk can touch: [dlo, dhi] with indptr[dlo] <= e0 < ..., found once per CTA
  __shared__ uint64_t s_dlo, s_dhi;
  if (threadIdx.x < 2) {
    const uint64_t target = threadIdx.x == 0 ? e0 : e1 - 1;
    uint64_t lo = 0, hi = n_docs;  // last d with indptr[d] <= target
    while (hi - lo > 1) {
      const uint64_t mid = (lo + hi) >> 1;
      if (__ldg(indptr + mid) <= target) lo = mid;
      else hi = mid;
    }
    if (threadIdx.x == 0) s_dlo = lo;
    else s_dhi = lo;
  }
  __syncthreads();
  const uint64_t dlo = s_dlo, dhi = s_dhi;
  for (uint64_t tile = e0; tile < e1; tile += kBuildThreads) {
    const uint64_t e = tile + threadIdx.x;
    const bool valid = e < e1;
    uint32_t t = 0xFFFFFFFFu - lane;  // distinct per lane: never matches a neighbour
    uint32_t doc = 0, wbits = 0;
    if (valid) {
      t = __ldg(tok + e);
      wbits = __float_as_uint(__ldg(w + e));
      uint64_t lo = dlo, hi = dhi + 1;  // last d in [dlo, dhi] with indptr[d] <= e
      while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (__ldg(indptr + mid) <= e) lo = mid;
        else hi = mid;
      }
      doc = (uint32_t)lo;
      if (t >= vocab) t = 0xFFFFFFFFu - lane;  // reported by the count pass; skipped here
    }
    const bool live = valid && t < vocab;
    uint32_t slot = 0;
    for (uint32_t turn = 0; turn < kBuildThreads / 32; ++turn) {
      if (warp == turn) {
        const uint32_t peers = __match_any_sync(0xffffffffu, t);
        const uint32_t leader = __ffs(peers) - 1;
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t first = 0;
        if (live && lane == leader) {
          first = s_cur[t];
          s_cur[t] = first + __popc(peers);
        }
        first = __shfl_sync(0xffffffffu, first, leader);
        slot = first + rank;
      }
      __syncthreads();
    }
    if (live) {
      const uint64_t pos = __ldg(tptr + t) + slot;
      out_doc[pos] = doc;
      out_post[pos] = make_uint2(doc, wbits);
    }
  }
}

__global__ void sparse_dupcheck_kernel(const uint32_t* __restrict__ doc, uint64_t nnz,
                                       const uint64_t* __restrict__ tptr, uint32_t vocab,
                                       uint32_t* __restrict__ err) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < nnz;
       i += (uint64_t)gridDim.x * blockDim.x) {
    if (doc[i] != doc[i - 1]) continue;
    // equal neighbours: legal only across a list boundary (i is the first entry of its list)
    uint32_t lo = 0, hi = vocab;  // last t with tptr[t] <= i
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (tptr[mid] <= i) lo = mid;
      else hi = mid;
    }
    // skip empty lists that start at the same offset
    if (tptr[lo] != i) atomicExch(err, 2u);
  }
}

size_t sparse_build_scratch_bytes(uint32_t vocab, int num_sms) {
  const uint32_t P = (uint32_t)num_sms * 4;
  return (size_t)P * vocab * sizeof(uint32_t) + (size_t)vocab * sizeof(uint64_t) + 256;
}

cudaError_t launch_sparse_build(const SparseBuildArgs& a, int num_sms, cudaStream_t st) {
  const uint32_t P = (uint32_t)num_sms * 4;
  const size_t smem = (size_t)a.vocab * sizeof(uint32_t);
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  uint32_t* counts = (uint32_t*)a.d_scratch;
  uint64_t* totals = (uint64_t*)((uint8_t*)a.d_scratch + ((size_t)P * a.vocab * sizeof(uint32_t) + 255) / 256 * 256);
  cudaError_t e = cudaFuncSetAttribute(sparse_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(sparse_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  sparse_count_kernel<<<P, kBuildThreads, smem, st>>>(a.d_tok, a.nnz, a.vocab, counts, a.d_err);
  sparse_colscan_kernel<<<(a.vocab + 255) / 256, 256, 0, st>>>(counts, P, a.vocab, totals);
  sparse_tptr_kernel<<<1, 1024, 0, st>>>(totals, a.vocab, a.d_tptr);
  g_kernel_launches.fetch_add(3, std::memory_order_relaxed);
  if (a.nnz) {
    sparse_scatter_kernel<<<P, kBuildThreads, smem, st>>>(a.d_indptr, a.d_tok, a.d_w, a.n_docs, a.nnz, a.vocab,
                                                          counts, a.d_tptr, a.d_doc, (uint2*)a.d_post);
    sparse_dupcheck_kernel<<<num_sms * 8, 256, 0, st>>>(a.d_doc, a.nnz, a.d_tptr, a.vocab, a.d_err);
    g_kernel_launches.fetch_add(2, std::memory_order_relaxed);
  }
  return cudaGetLastError();
}

}  // namespace cqs
