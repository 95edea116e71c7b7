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

// "this generic pointer is in the shared window" hint (see TopK::assume_shared)
#ifdef CQS_NO_ASSUME_SHARED   // development aid: build without the hints
#define CQS_ASSUME_SHARED(p) ((void)0)
#else
#define CQS_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#endif

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

// A set of threads of one CTA that synchronise on a named barrier: the whole CTA
// (barrier 0) or, in the warp-specialised scan, just the consumer warps.
struct Group {
  uint32_t tid, nthr, bar;
  __device__ __forceinline__ void sync() const {
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory");
  }
  __device__ __forceinline__ bool any(bool p) const {  // barrier + OR-reduction
    uint32_t r;
    asm volatile(
        "{\n.reg .pred q, o;\nsetp.ne.u32 q, %3, 0;\n"
        "bar.red.or.pred o, %1, %2, q;\nselp.u32 %0, 1, 0, o;\n}\n"
        : "=r"(r)
        : "r"(bar), "r"(nthr), "r"((uint32_t)p)
        : "memory");
    return r != 0;
  }
};

// ---- block sort ---------------------------------------------------------------
// In-place descending bitonic sort of buf[0..P) by the group; P is a power of two
// and a multiple of kSortChunk.  A warp owns a 256-key chunk in registers (8 keys
// per lane, element index = base + r*32 + lane): network strides < 32 are lane
// shuffles, strides 32..128 are register-to-register, only strides >= 256 go
// through shared memory with a barrier.  Sorting 2048 keys costs 10 barriers
// instead of the 66 of the plain shared-memory network.
constexpr uint32_t kSortR = 8;
constexpr uint32_t kSortChunk = 32 * kSortR;  // 256

// Apply the strides stride_start, stride_start/2, .., 1 (stride_start <= 128) of
// the bitonic merge of width `size` to one chunk held in registers.
__device__ __forceinline__ void sort_chunk_pass(ckey_t (&reg)[kSortR], uint32_t lane,
                                                uint32_t base, uint32_t size,
                                                uint32_t stride_start) {
#pragma unroll
  for (int ls = 7; ls >= 0; --ls) {
    const uint32_t st = 1u << ls;
    if (st > stride_start) continue;
    if (st >= 32) {
      const uint32_t rs = st >> 5;
#pragma unroll
      for (uint32_t r = 0; r < kSortR; ++r) {
        if (r & rs) continue;
        const uint32_t i = base + r * 32 + lane;  // lower index of the pair
        const bool dir = (i & size) == 0;
        const ckey_t a = reg[r], b2 = reg[r | rs];
        if ((a < b2) == dir) {
          reg[r] = b2;
          reg[r | rs] = a;
        }
      }
    } else {
      const bool lower = (lane & st) == 0;
#pragma unroll
      for (uint32_t r = 0; r < kSortR; ++r) {
        const uint32_t il = (base + r * 32 + lane) & ~st;
        const bool dir = (il & size) == 0;
        const ckey_t mine = reg[r];
        const ckey_t other = __shfl_xor_sync(0xffffffffu, mine, st);
        const bool take_max = (lower == dir);
        reg[r] = take_max ? (mine > other ? mine : other) : (mine < other ? mine : other);
      }
    }
  }
}

__device__ __forceinline__ void block_sort_desc(const Group& g, ckey_t* buf, uint32_t P) {
  const uint32_t lane = g.tid & 31, warp = g.tid >> 5, nwarps = g.nthr >> 5;
  const uint32_t nchunks = P / kSortChunk;
  ckey_t reg[kSortR];
  // phase 1: every chunk fully sorted (direction alternates with bit 8 of the index)
  for (uint32_t c = warp; c < nchunks; c += nwarps) {
    const uint32_t base = c * kSortChunk;
#pragma unroll
    for (uint32_t r = 0; r < kSortR; ++r) reg[r] = buf[base + r * 32 + lane];
    for (uint32_t size = 2; size <= kSortChunk; size <<= 1)
      sort_chunk_pass(reg, lane, base, size, size >> 1);
#pragma unroll
    for (uint32_t r = 0; r < kSortR; ++r) buf[base + r * 32 + lane] = reg[r];
  }
  g.sync();
  for (uint32_t size = 2 * kSortChunk; size <= P; size <<= 1) {
    for (uint32_t stride = size >> 1; stride >= kSortChunk; stride >>= 1) {
      for (uint32_t i = g.tid; i < (P >> 1); i += g.nthr) {
        const uint32_t pos = 2 * i - (i & (stride - 1));
        const ckey_t a = buf[pos], b2 = buf[pos + stride];
        const bool dir = ((pos & size) == 0);
        if ((a < b2) == dir) {
          buf[pos] = b2;
          buf[pos + stride] = a;
        }
      }
      g.sync();
    }
    for (uint32_t c = warp; c < nchunks; c += nwarps) {
      const uint32_t base = c * kSortChunk;
#pragma unroll
      for (uint32_t r = 0; r < kSortR; ++r) reg[r] = buf[base + r * 32 + lane];
      sort_chunk_pass(reg, lane, base, size, kSortChunk >> 1);
#pragma unroll
      for (uint32_t r = 0; r < kSortR; ++r) buf[base + r * 32 + lane] = reg[r];
    }
    g.sync();
  }
}

// ---- chunk sort + cross rank ----------------------------------------------------
// Descending sort of buf[0..n), result in buf[0..n); `scratch` holds n rounded up to C keys.  The warps
// sort chunks of C = 32 R keys in registers (bitonic network: strides < 32 are lane
// shuffles, the others register to register) and writes it to scratch; then every key's final position
// is its index in its own chunk plus, for every other chunk, the number of keys there that beat it — a
// branch-free binary search (log2 C + 1 probes, four chunks' searches interleaved).  Keys are unique;
// padding keys are 0 and land at positions >= n, where they are dropped.  Two barriers, no quadratic loop:
// ~860 keys (a k = 500 merge) cost 5.6 us on 8 warps inside the scan kernel where selection + rank sort took
// 15 us and selection + the full bitonic network 22 us (tools/trace_scan.py); 500 keys 2.5 us in isolation.
// NOT used by the batch kernels: scan_batch.cu defines CQS_FORCE_NETWORK because this sort, executed inside
// update_thr_kernel, faulted about once per 100-500 batch calls — cause not found; it is
// clean in isolation (tools/stress_sort.py) and in the scan and sparse kernels (tools/stress_single.py).
template <int R>
__device__ __forceinline__ void chunk_network(ckey_t (&reg)[R], uint32_t lane) {
  constexpr uint32_t C = 32 * R;
#pragma unroll
  for (uint32_t size = 2; size <= C; size <<= 1) {
#pragma unroll
    for (uint32_t st = size >> 1; st > 0; st >>= 1) {
      if (st >= 32) {
        const uint32_t rs = st >> 5;
#pragma unroll
        for (uint32_t r = 0; r < (uint32_t)R; ++r) {
          if (r & rs) continue;
          const bool dir = ((r * 32 + lane) & size) == 0;   // true: larger key first
          const ckey_t a = reg[r], b2 = reg[r | rs];
          if ((a < b2) == dir) {
            reg[r] = b2;
            reg[r | rs] = a;
          }
        }
      } else {
        const bool lower = (lane & st) == 0;
#pragma unroll
        for (uint32_t r = 0; r < (uint32_t)R; ++r) {
          const bool dir = (((r * 32 + lane) & ~st) & size) == 0;
          const ckey_t mine = reg[r];
          const ckey_t other = __shfl_xor_sync(0xffffffffu, mine, st);
          reg[r] = (lower == dir) ? (mine > other ? mine : other) : (mine < other ? mine : other);
        }
      }
    }
  }
}

template <int R>
__device__ __forceinline__ void chunk_rank_sort(const Group& g, ckey_t* buf, ckey_t* scratch, uint32_t n) {
  constexpr uint32_t C = 32 * R;
  const uint32_t lane = g.tid & 31, warp = g.tid >> 5, nwarps = g.nthr >> 5;
  const uint32_t nchunks = (n + C - 1) / C;   // a warp takes chunks warp, warp + nwarps, ..
  ckey_t reg[R];
  for (uint32_t c = warp; c < nchunks; c += nwarps) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint32_t i = c * C + r * 32 + lane;
      reg[r] = i < n ? buf[i] : 0;
    }
    __syncwarp();   // the network's shuffles run on the converged-warp fast path
    chunk_network<R>(reg, lane);
#pragma unroll
    for (int r = 0; r < R; ++r) scratch[c * C + r * 32 + lane] = reg[r];
  }
  g.sync();
  for (uint32_t c = warp; c < nchunks; c += nwarps) {
    uint32_t pos[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      pos[r] = r * 32 + lane;
      reg[r] = scratch[c * C + pos[r]];
    }
    // other chunks, four at a time: 4 R independent probe chains
    for (uint32_t c0 = 0; c0 < nchunks; c0 += 4) {
      uint32_t lo[4][R];
      const ckey_t* arr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        // the own chunk and chunks past the end are searched too (keeps the code uniform); their counts
        // are discarded below
        arr[u] = scratch + min(c0 + u, nchunks - 1) * C;
#pragma unroll
        for (int r = 0; r < R; ++r) lo[u][r] = 0;
      }
#pragma unroll
      for (uint32_t st = C >> 1; st > 0; st >>= 1)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (arr[u][lo[u][r] + st - 1] > reg[r]) lo[u][r] += st;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (arr[u][lo[u][r]] > reg[r]) lo[u][r] += 1;
          if (c0 + u < nchunks && c0 + u != c) pos[r] += lo[u][r];
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (pos[r] < n) buf[pos[r]] = reg[r];
  }
  g.sync();
}

// Up to this many keys the finishing code sorts directly (chunk sort); above it a histogram selection
// (3 us) first cuts the buffer down to ~k keys.
constexpr uint32_t kDirectSortMax = 1024;
constexpr uint32_t kSelBuckets = 2048;  // histogram resolution of select()

// Group-level streaming top-k accumulator.  `buf` has CAP slots (power of two).
// Any thread may push() a key that beats the current threshold; the group calls
// compact() (collectively) often enough that the buffer cannot overflow: the
// caller guarantees at most (CAP - k) pushes between two compactions.
struct TopK {
  ckey_t* buf;     // shared, CAP entries
  uint32_t* cnt;   // shared
  ckey_t* thr;     // shared: k-th best key so far (0 while fewer than k)
  uint32_t cap;
  Group g;
  uint32_t* hist = nullptr;  // shared [kSelBuckets + 96] words, needed by select()

  __device__ __forceinline__ void init() {
    if (g.tid == 0) {
      *cnt = 0;
      *thr = 0;
    }
  }
  // The out-of-line functions below receive this struct by value and would otherwise address the
  // buffer with GENERIC loads and stores (LD.E/ST.E through the global pipe: a final sort of 500 keys
  // took 15 us inside the scan kernel and 5.5 us in a test kernel that could see the __shared__
  // declaration).  Tells the compiler which window the pointers are in.
  __device__ __forceinline__ void assume_shared() const {
    CQS_ASSUME_SHARED(buf);
    CQS_ASSUME_SHARED(cnt);
    CQS_ASSUME_SHARED(thr);
  }
  __device__ __forceinline__ void push(ckey_t key) {
    uint32_t slot = atomicAdd(cnt, 1u);
    if (slot < cap) buf[slot] = key;  // cannot fail when the caller honours the bound
  }
  // Collective.  Afterwards buf[0..min(n,k)) holds the best keys, descending.
  __device__ __forceinline__ void compact(uint32_t k, bool force_network = false);
  __device__ __forceinline__ void compact_impl(uint32_t k, bool force_network) {
    g.sync();
    uint32_t n = min(*cnt, cap);
    const uint32_t per_r = g.nthr;   // keys per register of chunk_rank_sort when every warp takes one chunk
    ckey_t* scratch = buf + (cap >> 1);   // the sorted chunks (n <= cap / 2 on that path)
#ifdef CQS_FORCE_NETWORK   // development aid: always the bitonic network
    force_network = true;
#endif
    if (force_network || n > (cap >> 1)) {
      uint32_t P = max(next_pow2(n), kSortChunk);  // cap >= kSortChunk
      for (uint32_t i = n + g.tid; i < P; i += g.nthr) buf[i] = 0;
      g.sync();
      block_sort_desc(g, buf, P);
    } else if (n <= per_r) {
      chunk_rank_sort<1>(g, buf, scratch, n);
    } else if (n <= 2 * per_r) {
      chunk_rank_sort<2>(g, buf, scratch, n);
    } else {
      // Above 4 keys per thread the warps take several chunks each — except for a FEW keys more than one
      // chunk per warp (a selection for k = 4 nthr leaves k plus the ties of the k-th key's bucket), where a
      // second round of chunks would double the cost: sort the first 4 nthr keys, then merge the e <= 32
      // others in (a sorted key moves down by the number of extras that beat it; an extra goes to its rank
      // among the sorted keys plus its rank among the extras).
      const uint32_t n0 = (n > 4 * per_r && n <= 4 * per_r + 32) ? 4 * per_r : n;
      chunk_rank_sort<4>(g, buf, scratch, n0);          // leaves buf[n0..n) alone; scratch is free again
      if (n0 < n) {
        const uint32_t e = n - n0;
        for (uint32_t i = g.tid; i < n0; i += g.nthr) {
          const ckey_t key = buf[i];
          uint32_t pos = i;
          for (uint32_t x = 0; x < e; ++x) pos += (buf[n0 + x] > key) ? 1u : 0u;
          scratch[pos] = key;
        }
        if (g.tid < e) {
          const ckey_t key = buf[n0 + g.tid];
          uint32_t lo = 0, hi = n0;                     // first sorted key <= key
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (buf[mid] > key) lo = mid + 1; else hi = mid;
          }
          uint32_t pos = lo;
          for (uint32_t x = 0; x < e; ++x) pos += (buf[n0 + x] > key) ? 1u : 0u;
          scratch[pos] = key;
        }
        g.sync();
        for (uint32_t i = g.tid; i < n; i += g.nthr) buf[i] = scratch[i];
        g.sync();
      }
    }
    if (g.tid == 0) {
      *cnt = min(n, k);
      *thr = (n >= k && k > 0) ? buf[k - 1] : 0;
    }
    g.sync();
  }

  // Collective partial selection for the streaming phase: shrinks the buffer to a SUPERSET
  // of the k best keys (unsorted, at most cap/2 of them) and raises the threshold to a
  // valid lower bound of the k-th best — with one histogram pass instead of a sort.
  // Keys live in registers (ITEMS per thread, ITEMS * nthr >= cap) while the buffer is
  // rewritten.  Buckets are linear in key space between the current min and max.
  //
  // m / vm_out (optional): also reports a lower bound of the m-th best key held BEFORE the
  // shrink (m <= k): at least m keys of the buffer are >= *vm_out (0 when fewer than m keys).
  // The scan kernel publishes it so that the CTAs can share a global threshold.
  template <int ITEMS>
  __device__ __forceinline__ void select(uint32_t k, uint32_t m = 0, ckey_t* vm_out = nullptr);
  template <int ITEMS>
  __device__ __forceinline__ void select_impl(uint32_t k, uint32_t m, ckey_t* vm_out) {
    g.sync();
    const uint32_t n = min(*cnt, cap);
    if (n <= kDirectSortMax || n <= k || hist == nullptr) {
      compact(k);
      if (vm_out && g.tid == 0) *vm_out = (m > 0 && m <= min(n, k)) ? buf[m - 1] : 0;
      g.sync();
      return;
    }
    CQS_ASSUME_SHARED(hist);   // (non-null past the early return)
    if (vm_out) CQS_ASSUME_SHARED(vm_out);
    const uint32_t lane = g.tid & 31, warp = g.tid >> 5, nwarps = g.nthr >> 5;
    ckey_t r[ITEMS];
    ckey_t lo = ~0ull, hi = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const uint32_t i = j * g.nthr + g.tid;
      r[j] = (i < n) ? buf[i] : 0;
      if (i < n) {
        lo = r[j] < lo ? r[j] : lo;
        hi = r[j] > hi ? r[j] : hi;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const ckey_t a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
      lo = a < lo ? a : lo;
      hi = b > hi ? b : hi;
    }
    ckey_t* aux64 = reinterpret_cast<ckey_t*>(hist + kSelBuckets);  // [2 * 16] keys
    uint32_t* aux32 = hist + kSelBuckets + 64;                      // [16] words
    if (lane == 0) {
      aux64[warp] = lo;
      aux64[16 + warp] = hi;
    }
    for (uint32_t i = g.tid; i < kSelBuckets; i += g.nthr) hist[i] = 0;
    g.sync();
    for (uint32_t w = 0; w < nwarps; ++w) {
      lo = aux64[w] < lo ? aux64[w] : lo;
      hi = aux64[16 + w] > hi ? aux64[16 + w] : hi;
    }
    const int bits = 64 - __clzll((long long)((hi - lo) | 1ull));
    const int sh = bits > 11 ? bits - 11 : 0;  // (key - lo) >> sh < 2048
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
      if ((uint32_t)(j * g.nthr + g.tid) < n) atomicAdd(&hist[(uint32_t)((r[j] - lo) >> sh)], 1u);
    g.sync();
    // cumulative counts from the TOP bucket down: thread t owns buckets
    // [top - per*(t+1) + 1, top - per*t], per = kSelBuckets / nthr
    const uint32_t per = kSelBuckets / g.nthr;
    const uint32_t b_hi = kSelBuckets - 1 - per * g.tid;  // highest bucket of this thread
    uint32_t local = 0;
    for (uint32_t j = 0; j < per; ++j) local += hist[b_hi - j];
    uint32_t incl = local;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += v;
    }
    if (lane == 31) aux32[warp] = incl;
    g.sync();
    uint32_t before = incl - local;  // keys in buckets above this thread's segment
    for (uint32_t w = 0; w < warp; ++w) before += aux32[w];
    g.sync();
    if (before < k && k <= before + local) {  // exactly one thread: the k-th best falls in its segment
      uint32_t acc = before;
      for (uint32_t j = 0; j < per; ++j) {
        const uint32_t h = hist[b_hi - j];
        if (acc + h >= k) {
          aux32[0] = b_hi - j;   // b*: bucket holding the k-th best
          aux32[1] = acc + h;    // keys in buckets >= b*
          break;
        }
        acc += h;
      }
    }
    if (vm_out && m > 0 && before < m && m <= before + local) {  // likewise for the m-th best (m <= k <= n)
      uint32_t acc = before;
      for (uint32_t j = 0; j < per; ++j) {
        acc += hist[b_hi - j];
        if (acc >= m) {
          *vm_out = lo + ((ckey_t)(b_hi - j) << sh);   // lower edge of its bucket: >= m keys are >= this
          break;
        }
      }
    }
    g.sync();
    const uint32_t bstar = aux32[0], n_keep = aux32[1];
    if (n_keep > (cap >> 1)) {  // pathological ties: fall back to the exact sort (buffer untouched so far)
      compact(k);               // (*vm_out, if any, stays: it was derived from the untouched buffer)
      return;
    }
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
      if ((uint32_t)(j * g.nthr + g.tid) < n && (uint32_t)((r[j] - lo) >> sh) >= bstar) ++mine;
    uint32_t inc2 = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc2, o);
      if (lane >= (uint32_t)o) inc2 += v;
    }
    g.sync();
    if (lane == 31) aux32[2 + warp] = inc2;
    g.sync();
    uint32_t off = inc2 - mine;
    for (uint32_t w = 0; w < warp; ++w) off += aux32[2 + w];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
      if ((uint32_t)(j * g.nthr + g.tid) < n && (uint32_t)((r[j] - lo) >> sh) >= bstar) buf[off++] = r[j];
    if (g.tid == 0) {
      *cnt = n_keep;
      *thr = lo + ((ckey_t)bstar << sh) - 1;  // every kept key is > thr; thr >= the previous threshold
    }
    g.sync();
  }
  // Collective: drop every key that is not > t (a valid lower bound of the final k-th best key
  // obtained elsewhere, e.g. the global threshold shared by the scan CTAs); raises thr to t.
  template <int ITEMS>
  __device__ __forceinline__ void prune(ckey_t t) {
    g.sync();
    const uint32_t n = min(*cnt, cap);
    const uint32_t lane = g.tid & 31, warp = g.tid >> 5, nwarps = g.nthr >> 5;
    uint32_t* aux32 = hist + kSelBuckets + 64;
    ckey_t r[ITEMS];
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const uint32_t i = j * g.nthr + g.tid;
      r[j] = (i < n) ? buf[i] : 0;
      if (r[j] > t) ++mine;
    }
    uint32_t inc = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (uint32_t)o) inc += v;
    }
    if (lane == 31) aux32[2 + warp] = inc;
    g.sync();
    uint32_t off = inc - mine, total = 0;
    for (uint32_t w = 0; w < nwarps; ++w) {
      if (w < warp) off += aux32[2 + w];
      total += aux32[2 + w];
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
      if (r[j] > t) buf[off++] = r[j];
    if (g.tid == 0) {
      *cnt = total;
      if (t > *thr) *thr = t;
    }
    g.sync();
  }
};

// ONE out-of-line copy of the selection per translation unit, shared by the streaming loop, the
// per-CTA finish and the last CTA's merge: the merge runs once per kernel, and its private inlined
// copies of this code were instruction-cache misses from L2 (2.6 us warm, 11 us measured there); the
// streaming loop has executed the shared copy on the same SM moments earlier.
static __device__ __noinline__ void topk_compact(TopK tk, uint32_t k, bool force_network) {
  tk.assume_shared();
  tk.compact_impl(k, force_network);
}
__device__ __forceinline__ void TopK::compact(uint32_t k, bool force_network) {
  topk_compact(*this, k, force_network);
}
template <int ITEMS>
__device__ __noinline__ void topk_select(TopK tk, uint32_t k, uint32_t m, ckey_t* vm_out) {
  tk.assume_shared();
  tk.template select_impl<ITEMS>(k, m, vm_out);
}
template <int ITEMS>
__device__ __forceinline__ void TopK::select(uint32_t k, uint32_t m, ckey_t* vm_out) {
  topk_select<ITEMS>(*this, k, m, vm_out);
}

// Collective: exact top-k of whatever the accumulator holds, descending in buf[0..min(n,k)).
// Two histogram selections (3 us each) first cut the buffer down to the k best plus the few
// keys sharing the k-th key's bucket, so the exact sort runs on ~k keys (the chunk sort above)
// instead of on the whole buffer (a 2048-4096 key bitonic sort costs 20-40 us,
// paid by every CTA at the end of its scan and again by the last CTA's merge).
// Not inlined (TopK by value: a handful of shared-memory pointers): the scan file instantiates
// 48 kernels and this tail code would otherwise be compiled into each of them.
#define CQS_STAMP(tr, i)                                                        \
  do {                                                                          \
    if ((tr) && tk.g.tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"((tr)[i])); \
  } while (0)
template <int ITEMS>
__device__ __noinline__ void topk_finish(TopK tk, uint32_t k, unsigned long long* stamps = nullptr) {
  tk.assume_shared();
  tk.g.sync();
  if (stamps && tk.g.tid == 0) stamps[7] = *tk.cnt;
  // a selection (3 us) pays off only above the range the chunk sort handles directly
  if (min(*tk.cnt, tk.cap) > kDirectSortMax && *tk.cnt > k + 64) {
    const ckey_t before = *tk.thr;
    tk.template select<ITEMS>(k);
    if (tk.g.tid == 0 && before > *tk.thr) *tk.thr = before;
    tk.g.sync();
    CQS_STAMP(stamps, 4);
    if (min(*tk.cnt, tk.cap) > kDirectSortMax && *tk.cnt > k + 64) tk.template select<ITEMS>(k);
    CQS_STAMP(stamps, 5);
  }
  if (stamps && tk.g.tid == 0) stamps[6] = *tk.cnt;
  long long c0 = 0;
  if (stamps && tk.g.tid == 0) {   // development aid: SM cycles beside the wall-clock stamps (row below)
    c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"((stamps - 8)[0]));
  }
  tk.compact(k);
  if (stamps && tk.g.tid == 0) {
    (stamps - 8)[2] = (unsigned long long)(clock64() - c0);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"((stamps - 8)[1]));
  }
}

constexpr uint32_t kPartialStride = 1024;  // == kMaxK: slots per CTA in the partial-list scratch

// Last-CTA merge of the per-CTA sorted candidate lists (shared by the dense and
// the sparse kernels).  All G lists (the caller's own included) are read back
// from the partial-list scratch in L2.  Collective over tk.g.
//
//  1. bound: for a column m (the m-th key of every list) the ceil(k/m)-th largest
//     column value v has at least ceil(k/m) lists with >= m keys >= v, i.e. >= k keys
//     >= v overall: v is a lower bound of the global k-th best.  Two columns are
//     evaluated (the first power of two m with ceil(k/m) <= G, and 2m; rank counting
//     over G values each) and the better bound kept.  For k <= G column 1 alone
//     leaves ~k survivors out of G*k candidates.
//  2. rounds: in round r every still-active list exposes its window
//     [r*b, (r+1)*b) and all window slots are examined in parallel (independent L2
//     loads); keys above the threshold are pushed; a list retires at its first key
//     <= threshold (lists are sorted) or when exhausted.  The accumulator is only
//     re-selected when the next round could overflow it.  s_pos[l] = keys list l
//     still offers.
//  3. final exact sort, emit (score desc, row asc); unused slots = (-inf, UINT64_MAX).
template <int ITEMS>
__device__ __noinline__ void merge_partials_and_emit(TopK tk, uint32_t* s_pos, uint32_t k,
                                                        const ckey_t* partial,
                                                        const uint32_t* partial_cnt, uint32_t G,
                                                        uint64_t row_base, float* out_scores,
                                                        uint64_t* out_rows, uint32_t* out_n,
                                                        unsigned long long* trace = nullptr,
                                                        ckey_t thr0 = 0) {
  // thr0: a lower bound of the global k-th best key the caller already knows (0 = none)
  tk.assume_shared();
  CQS_ASSUME_SHARED(s_pos);
  const uint32_t tid = tk.g.tid, T = tk.g.nthr;
  // development aid: extra stamps of the merge go to the last row of the trace buffer
  unsigned long long* mstamps = trace ? trace - (size_t)blockIdx.x * 8 + (size_t)(kPartialStride - 1) * 8 : nullptr;
  CQS_STAMP(mstamps, 0);
  for (uint32_t l = tid; l < G; l += T) s_pos[l] = __ldcg(partial_cnt + l);
  if (tid == 0) {
    *tk.cnt = 0;
    *tk.thr = thr0;
  }
  tk.g.sync();
  // ---- 1. column bounds (two columns: the first m with ceil(k/m) <= G, and 2m) ----
  {
    ckey_t* col = tk.buf + (tk.cap >> 1);  // [2][G] scratch (2G <= 2*kMaxGrid <= cap/2)
    uint32_t m0 = 1;
    while ((k + m0 - 1) / m0 > G) m0 <<= 1;
    const uint32_t ms[2] = {m0, 2 * m0};
    for (uint32_t l = tid; l < G; l += T) {
      const uint32_t len = s_pos[l];
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2)
        col[c2 * G + l] = (ms[c2] <= len && ms[c2] <= k)
                              ? __ldcg(partial + (size_t)l * kPartialStride + (ms[c2] - 1)) : 0;
    }
    tk.g.sync();
    for (uint32_t t = tid; t < 2 * G; t += T) {
      const uint32_t c2 = t / G;
      const ckey_t mine = col[t];
      if (mine == 0) continue;
      const uint32_t j = (k + ms[c2] - 1) / ms[c2];  // need the j-th largest of this column
      const ckey_t* cc = col + c2 * G;
      uint32_t rank = 0, i = 0;  // keys are unique
      for (; i + 8 <= G; i += 8) {
        ckey_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = cc[i + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) rank += (v[u] > mine) ? 1u : 0u;
      }
      for (; i < G; ++i) rank += (cc[i] > mine) ? 1u : 0u;
      if (rank == j - 1) atomicMax(tk.thr, mine - 1);  // keys >= v pass "key > thr"
    }
    tk.g.sync();
  }
  CQS_STAMP(mstamps, 1);
  // ---- 2. rounds ----
  uint32_t b = (tk.cap / 2) / G;  // at most cap/2 pushes per round
  if (b == 0) b = 1;
  if (b > 64) b = 64;
  constexpr int kMlp = 8;  // independent L2 loads in flight per thread
  for (uint32_t r = 0;; ++r) {
    if (*tk.cnt + G * b > tk.cap) {  // uniform: cnt only changes between barriers
      const ckey_t before = *tk.thr;
      tk.template select<ITEMS>(k);
      if (tid == 0 && before > *tk.thr) *tk.thr = before;
      tk.g.sync();
    }
    const ckey_t thr = *tk.thr;
    const uint32_t lo = r * b, total = G * b;
    bool more = false;
    for (uint32_t base = 0; base < total; base += T * kMlp) {
      ckey_t keys[kMlp];
      bool tail[kMlp];
#pragma unroll
      for (int j = 0; j < kMlp; ++j) {
        const uint32_t t = base + j * T + tid;
        keys[j] = 0;
        tail[j] = false;
        if (t < total) {
          const uint32_t l = t / b, pos = lo + (t - l * b), len = s_pos[l];
          if (pos < len) {
            keys[j] = __ldcg(partial + (size_t)l * kPartialStride + pos);
            tail[j] = (pos == lo + b - 1) && (pos + 1 < len);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kMlp; ++j)
        if (keys[j] > thr) {
          tk.push(keys[j]);
          if (tail[j]) more = true;  // window full of winners and the list goes on
        }
    }
    const bool any_more = tk.g.any(more);
    if (trace && tid == 0 && r == 0) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[5]));
      trace[7] = *tk.cnt;
    }
    if (!any_more) break;
    // retire lists whose window ended at or below the threshold, or that are exhausted
    for (uint32_t l = tid; l < G; l += T) {
      const uint32_t len = s_pos[l], last = lo + b - 1;
      if (len == 0) continue;
      if (last + 1 >= len || __ldcg(partial + (size_t)l * kPartialStride + last) <= thr) s_pos[l] = 0;
    }
    tk.g.sync();
  }
  CQS_STAMP(mstamps, 2);
  if (mstamps && tid == 0) mstamps[3] = *tk.thr;
  topk_finish<ITEMS>(tk, k, mstamps);
  if (trace && tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace[6]));
  if (out_scores == nullptr) return;   // caller post-processes the sorted keys in tk.buf[0..*tk.cnt)
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
