// peer.cuh — device side of the cross-shard exchange over NVLink peer memory
// (SURVEY.md §8e): every rank PUSHES its local sorted top-k lists straight into
// a mailbox in each peer's HBM (plain stores through the NVSwitch fabric), then
// publishes one release-flag per peer; the receiver spins on its own flags and
// merges the G sorted lists by rank counting.  No NCCL call, no host round trip:
// the exchange rides in the tail of the kernel that produced the lists.
//
// The reference has no multi-GPU path; ordering is the reference's result order
// (score desc by f32 total order, id asc — src/search/scoring/candidate.rs:321-329),
// so the merged result is identical to the unsharded one.
//
// Mailbox of rank r (lives in r's HBM, mapped into every peer):
//   [slot 0..kPeerSlots)[source rank 0..world)  block
//   block = | flag u32, pad to 128 B | n[maxq] u32 | rows[cap] u64 | scores[cap] f32 |
// Exchange number `seq` (same on every rank, 1, 2, 3, ...) uses slot seq % kPeerSlots
// and writes `seq` into the flag.  A rank orders exchange seq after its own exchange
// seq-L (L = kPeerLanes may be in flight: the tail of one scan overlaps the next scans),
// and it can only FINISH exchange seq-L after every peer has pushed seq-L, which a peer
// does after finishing (reading) its seq-2L.  So when a push of `seq` lands in a peer's
// slot, that peer is done with seq-2L, the previous user of the slot: 2L slots suffice.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace cqs {

constexpr uint32_t kPeerMaxWorld = 8;
constexpr uint32_t kPeerLanes = 4;   // == kLanes (internal.h)
constexpr uint32_t kPeerSlots = 2 * kPeerLanes;
constexpr uint32_t kPeerMaxQ = 1024;   // queries per exchange
constexpr uint32_t kPeerFlagBytes = 128;

struct PeerCtx {
  uint32_t world = 0;       // 0 = no exchange
  uint32_t rank = 0;
  uint32_t seq = 0;
  uint32_t cap = 0;         // (query, slot) elements per block; nq * k <= cap
  uint64_t block_bytes = 0;
  uint64_t timeout_ns = 0;  // give up waiting for a peer after this long
  uint32_t* status = nullptr;  // local sticky word: 1 = a wait timed out
  uint8_t* mbox[kPeerMaxWorld] = {};  // base of every rank's mailbox in this process' address space
};

__host__ __device__ __forceinline__ uint64_t peer_block_bytes(uint32_t cap) {
  uint64_t b = kPeerFlagBytes + 4ull * kPeerMaxQ + 12ull * cap;
  return (b + 255) / 256 * 256;
}

struct PeerBlock {
  uint32_t* flag;
  uint32_t* n;
  uint64_t* rows;
  float* scores;
};
// block in `owner`'s mailbox that receives the lists of `src`
__device__ __forceinline__ PeerBlock peer_block(const PeerCtx& c, uint32_t owner, uint32_t src) {
  uint8_t* b = c.mbox[owner] + ((uint64_t)(c.seq % kPeerSlots) * c.world + src) * c.block_bytes;
  PeerBlock p;
  p.flag = reinterpret_cast<uint32_t*>(b);
  p.n = reinterpret_cast<uint32_t*>(b + kPeerFlagBytes);
  p.rows = reinterpret_cast<uint64_t*>(b + kPeerFlagBytes + 4ull * kPeerMaxQ);
  p.scores = reinterpret_cast<float*>(b + kPeerFlagBytes + 4ull * kPeerMaxQ + 8ull * c.cap);
  return p;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Collective over g.  Tell every rank (this one included) that this rank's lists of
// exchange c.seq are in place.  The caller's pushes must precede this call; the
// system-scope fence + barrier makes all of the group's stores visible before the flags.
__device__ __forceinline__ void peer_signal(const PeerCtx& c, const Group& g) {
  __threadfence_system();
  g.sync();
  if (g.tid < c.world) st_release_sys(peer_block(c, g.tid, c.rank).flag, c.seq);
}

// Collective over g.  Wait until the lists of every rank for exchange c.seq have landed
// in this rank's mailbox.  Returns false (and raises the sticky status word) on timeout.
__device__ __forceinline__ bool peer_wait(const PeerCtx& c, const Group& g) {
  bool late = false;
  if (g.tid < c.world) {
    const uint32_t* f = peer_block(c, c.rank, g.tid).flag;
    const unsigned long long t0 = peer_now_ns();
    uint32_t spins = 0;
    while (ld_acquire_sys(f) != c.seq) {
      if ((++spins & 0x3FF) == 0 && peer_now_ns() - t0 > c.timeout_ns) {
        late = true;
        break;
      }
    }
    if (late) atomicExch(c.status, 1u);
  }
  return !g.any(late);
}

// Collective over g.  Merge the `world` sorted lists of query qi (k slots each, n valid)
// that sit in this rank's mailbox into the global top-k.  Every list is sorted in result
// order and rows are unique across lists, so the final position of an element is its own
// position plus, for every other list, the number of that list's elements ahead of it
// (one binary search each): no sort, no shared memory.
__device__ __forceinline__ void peer_merge_query(const PeerCtx& c, const Group& g, uint32_t qi,
                                                 uint32_t k, float* out_scores, uint64_t* out_rows,
                                                 uint32_t* out_n) {
  uint32_t nl[kPeerMaxWorld];
  const float* sl[kPeerMaxWorld];
  const uint64_t* rl[kPeerMaxWorld];
  uint32_t total = 0, unproven = 0;   // bit 31 of a list length: that shard could not prove its list (kUnprovenBit)
#pragma unroll
  for (uint32_t l = 0; l < kPeerMaxWorld; ++l) {
    nl[l] = 0;
    sl[l] = nullptr;
    rl[l] = nullptr;
    if (l < c.world) {
      const PeerBlock b = peer_block(c, c.rank, l);
      const uint32_t raw = __ldcg(b.n + qi);
      unproven |= raw & 0x80000000u;
      nl[l] = min(raw & 0x7FFFFFFFu, k);
      sl[l] = b.scores + (size_t)qi * k;
      rl[l] = b.rows + (size_t)qi * k;
      total += nl[l];
    }
  }
  for (uint32_t e = g.tid; e < c.world * k; e += g.nthr) {
    const uint32_t l = e / k, j = e - l * k;
    uint32_t mine_n = 0;
    const float* ms = nullptr;
    const uint64_t* mr = nullptr;
#pragma unroll
    for (uint32_t x = 0; x < kPeerMaxWorld; ++x)
      if (x == l) {
        mine_n = nl[x];
        ms = sl[x];
        mr = rl[x];
      }
    if (j >= mine_n) continue;
    const uint32_t sbits = __float_as_uint(__ldcg(ms + j));
    const uint32_t s = ordered_u32(sbits);
    const uint64_t r = __ldcg(mr + j);
    uint32_t rank = j;
#pragma unroll
    for (uint32_t x = 0; x < kPeerMaxWorld; ++x) {
      if (x == l || nl[x] == 0) continue;
      uint32_t lo = 0, hi = nl[x];  // first position of list x that is NOT ahead of (s, r)
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint32_t s2 = ordered_u32(__float_as_uint(__ldcg(sl[x] + mid)));
        const bool ahead = s2 > s || (s2 == s && __ldcg(rl[x] + mid) < r);
        if (ahead) lo = mid + 1;
        else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      out_scores[rank] = __uint_as_float(sbits);
      out_rows[rank] = r;
    }
  }
  const uint32_t n = min(total, k);
  for (uint32_t i = n + g.tid; i < k; i += g.nthr) {
    out_scores[i] = __uint_as_float(0xFF800000u);  // -inf
    out_rows[i] = ~0ull;
  }
  if (g.tid == 0) *out_n = n | unproven;
}

// Same merge with the lists staged in SHARED memory first (`stage`: world * k * 12 + 64 bytes).  The
// rank counting does world-1 binary searches per element; out of the mailbox (global memory, L2
// latency per probe) that is ~35 dependent loads per thread at k = 20 but ~1000 at k = 500 — 0.3 ms,
// measured at 8 GPUs — while from shared memory it is a few microseconds at any k.  The scan kernel
// passes its (by then idle) tile ring.
__device__ __forceinline__ void peer_merge_query_staged(const PeerCtx& c, const Group& g, uint32_t qi,
                                                        uint32_t k, uint8_t* stage, float* out_scores,
                                                        uint64_t* out_rows, uint32_t* out_n) {
  uint64_t* srow = reinterpret_cast<uint64_t*>(stage);             // [world][k]
  uint32_t* ssc = reinterpret_cast<uint32_t*>(srow + (size_t)c.world * k);   // [world][k] ordered scores
  uint32_t* sn = ssc + (size_t)c.world * k;                        // [world] lengths, [world] = flags
  if (g.tid == 0) sn[c.world] = 0;
  g.sync();
  if (g.tid < c.world) {
    const uint32_t raw = __ldcg(peer_block(c, c.rank, g.tid).n + qi);
    sn[g.tid] = min(raw & 0x7FFFFFFFu, k);
    if (raw & 0x80000000u) atomicOr(&sn[c.world], 0x80000000u);
  }
  g.sync();
  for (uint32_t e = g.tid; e < c.world * k; e += g.nthr) {
    const uint32_t l = e / k, j = e - l * k;
    if (j >= sn[l]) continue;
    const PeerBlock b = peer_block(c, c.rank, l);
    srow[e] = __ldcg(b.rows + (size_t)qi * k + j);
    ssc[e] = ordered_u32(__float_as_uint(__ldcg(b.scores + (size_t)qi * k + j)));
  }
  g.sync();
  uint32_t total = 0;
  for (uint32_t l = 0; l < c.world; ++l) total += sn[l];
  for (uint32_t e = g.tid; e < c.world * k; e += g.nthr) {
    const uint32_t l = e / k, j = e - l * k;
    if (j >= sn[l]) continue;
    const uint32_t s = ssc[e];
    const uint64_t r = srow[e];
    uint32_t rank = j;
    for (uint32_t x = 0; x < c.world; ++x) {
      if (x == l || sn[x] == 0) continue;
      const uint32_t* xs = ssc + (size_t)x * k;
      const uint64_t* xr = srow + (size_t)x * k;
      uint32_t lo = 0, hi = sn[x];  // first position of list x that is NOT ahead of (s, r)
      while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const bool ahead = xs[mid] > s || (xs[mid] == s && xr[mid] < r);
        if (ahead) lo = mid + 1;
        else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      out_scores[rank] = __uint_as_float(unordered_u32(s));
      out_rows[rank] = r;
    }
  }
  const uint32_t n = min(total, k);
  for (uint32_t i = n + g.tid; i < k; i += g.nthr) {
    out_scores[i] = __uint_as_float(0xFF800000u);  // -inf
    out_rows[i] = ~0ull;
  }
  if (g.tid == 0) *out_n = n | sn[c.world];
}

// Collective over g: empty result (exchange failed).
__device__ __forceinline__ void peer_emit_empty(const Group& g, uint32_t k, float* out_scores,
                                                uint64_t* out_rows, uint32_t* out_n) {
  for (uint32_t i = g.tid; i < k; i += g.nthr) {
    out_scores[i] = __uint_as_float(0xFF800000u);
    out_rows[i] = ~0ull;
  }
  if (g.tid == 0) *out_n = 0;
}

}  // namespace cqs
