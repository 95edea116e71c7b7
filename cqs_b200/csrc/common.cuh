// common.cuh — device-side building blocks shared by every kernel:
//   * the 64-bit candidate key that encodes the reference's result order
//   * a CTA-level streaming top-k accumulator in shared memory
//
// Ordering contract (src/search/scoring/candidate.rs:303-329): results are
// ordered by (score desc under f32::total_cmp, id asc).  Rows are stored in
// ascending chunk-id order, so id asc == row asc.  We fold both into ONE
// unsigned 64-bit key whose natural descending order is the result order:
//     key = ordered_u32(score) << 32 | ~row
// ordered_u32 is the usual monotone float->uint map (it reproduces total_cmp,
// including -0.0 < +0.0).  Non-finite scores never become keys
// (candidate.rs:275).  key == 0 is the "empty slot" sentinel: every valid key
// has a high word >= ordered(-FLT_MAX) = 0x00800000.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cqs {

typedef unsigned long long ckey_t;

__host__ __device__ __forceinline__ uint32_t ordered_u32(uint32_t bits) {
  return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t unordered_u32(uint32_t o) {
  return (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
}
__device__ __forceinline__ bool finite_bits(uint32_t bits) {
  return (bits & 0x7F800000u) != 0x7F800000u;
}
__device__ __forceinline__ ckey_t make_key(float score, uint32_t row) {
  return ((ckey_t)ordered_u32(__float_as_uint(score)) << 32) | (ckey_t)(~row);
}
__device__ __forceinline__ float key_score(ckey_t k) {
  return __uint_as_float(unordered_u32((uint32_t)(k >> 32)));
}
__device__ __forceinline__ uint32_t key_row(ckey_t k) { return ~(uint32_t)k; }

__device__ __forceinline__ uint32_t next_pow2(uint32_t v) {
  return v <= 1 ? 1u : 1u << (32 - __clz(v - 1));
}

// In-place descending bitonic sort of buf[0..P), P a power of two, by the whole
// CTA.  Ends with a __syncthreads().
__device__ __forceinline__ void bitonic_sort_desc(ckey_t* buf, uint32_t P) {
  const uint32_t T = blockDim.x, tid = threadIdx.x;
  for (uint32_t size = 2; size <= P; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t i = tid; i < (P >> 1); i += T) {
        uint32_t pos = 2 * i - (i & (stride - 1));
        ckey_t a = buf[pos], b = buf[pos + stride];
        bool dir = ((pos & size) == 0);
        if ((a < b) == dir) {
          buf[pos] = b;
          buf[pos + stride] = a;
        }
      }
      __syncthreads();
    }
  }
}

// CTA-level streaming top-k accumulator.  `buf` has CAP slots (power of two).
// Any thread may push() a key that beats the current threshold; the CTA calls
// compact() (collectively) often enough that the buffer cannot overflow: the
// caller guarantees at most (CAP - k) pushes between two compactions.
struct TopK {
  ckey_t* buf;      // shared, CAP entries
  uint32_t* cnt;   // shared
  ckey_t* thr;      // shared: k-th best key so far (0 while fewer than k)
  uint32_t cap;

  __device__ __forceinline__ void init() {
    if (threadIdx.x == 0) {
      *cnt = 0;
      *thr = 0;
    }
  }
  __device__ __forceinline__ void push(ckey_t key) {
    uint32_t slot = atomicAdd(cnt, 1u);
    if (slot < cap) buf[slot] = key;  // cannot fail when the caller honours the bound
  }
  // Collective.  Afterwards buf[0..min(n,k)) holds the best keys, descending.
  __device__ __forceinline__ void compact(uint32_t k) {
    __syncthreads();
    uint32_t n = min(*cnt, cap);
    uint32_t P = next_pow2(n);
    for (uint32_t i = n + threadIdx.x; i < P; i += blockDim.x) buf[i] = 0;
    __syncthreads();
    if (n > 1) bitonic_sort_desc(buf, P);
    if (threadIdx.x == 0) {
      *cnt = min(n, k);
      *thr = (n >= k && k > 0) ? buf[k - 1] : 0;
    }
    __syncthreads();
  }
};

constexpr uint32_t kPartialStride = 1024;  // == kMaxK: slots per CTA in the partial-list scratch

// Last-CTA merge of the per-CTA sorted candidate lists (shared by the dense and
// the sparse kernels).  On entry tk holds this CTA's own sorted list and
// threshold.  Every list is sorted descending, so a list is abandoned at its
// first key <= threshold; lists advance by `budget` keys per round so the
// accumulator cannot overflow.  Collective.  Writes the final result:
// (score desc, row asc), unused slots = (-inf, UINT64_MAX).
__device__ __forceinline__ void merge_partials_and_emit(TopK& tk, uint32_t* s_pos, uint32_t k,
                                                        const ckey_t* partial,
                                                        const uint32_t* partial_cnt, uint32_t G,
                                                        uint32_t self, uint64_t row_base,
                                                        float* out_scores, uint64_t* out_rows,
                                                        uint32_t* out_n) {
  const uint32_t tid = threadIdx.x, T = blockDim.x;
  for (uint32_t l = tid; l < G; l += T) s_pos[l] = 0;
  __syncthreads();
  uint32_t budget = (tk.cap - k) / G;
  if (budget == 0) budget = 1;
  uint32_t group = (tk.cap - k) / budget;  // lists per compaction: group * budget <= cap - k
  if (group == 0) group = 1;
  while (true) {
    int more = 0;
    for (uint32_t g0 = 0; g0 < G; g0 += group) {
      ckey_t thr = *tk.thr;
      uint32_t g1 = min(G, g0 + group);
      for (uint32_t l = g0 + tid; l < g1; l += T) {
        if (l == self) continue;
        uint32_t len = __ldcg(partial_cnt + l);
        uint32_t pos = s_pos[l];
        uint32_t pushed = 0;
        while (pos < len && pushed < budget) {
          ckey_t key = __ldcg(partial + (size_t)l * kPartialStride + pos);
          if (key <= thr) {  // sorted list: nothing further can qualify
            pos = len;
            break;
          }
          tk.push(key);
          ++pos;
          ++pushed;
        }
        s_pos[l] = pos;
        if (pos < len) more = 1;
      }
      tk.compact(k);
    }
    if (!__syncthreads_or(more)) break;
  }
  uint32_t n = *tk.cnt;
  for (uint32_t i = tid; i < k; i += T) {
    if (i < n) {
      ckey_t key = tk.buf[i];
      out_scores[i] = key_score(key);
      out_rows[i] = row_base + key_row(key);
    } else {
      out_scores[i] = __uint_as_float(0xFF800000u);  // -inf
      out_rows[i] = ~0ull;
    }
  }
  if (tid == 0) *out_n = n;
}

}  // namespace cqs
