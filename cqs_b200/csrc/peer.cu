// peer.cu — the cross-shard exchange of a row-sharded corpus over
// NVLink peer memory: lifecycle of a peer group (mailboxes exported over CUDA IPC
// between the one-process-per-GPU ranks, or plain peer access inside one process)
// and the stand-alone gather+merge kernel used after batched searches.  The
// single-query scan carries the same exchange in its own tail (scan_single.cu).
//
// The reference has no multi-GPU path (its VectorIndex is one index per process,
// src/index.rs:139-239); this is the B200-side extension that keeps the answer
// identical to the unsharded one.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <string>

#include "../../include/cqs_b200.h"
#include "internal.h"
#include "peer_host.h"

namespace cqs {

static_assert(kPeerLanes == kLanes, "peer lanes must match the scan launch lanes");

// One launch = one exchange of nq lists per rank.  Grid <= number of SMs so every CTA is
// resident (phase 2 spins on flags that the PEERS' phase 1 raises; nothing in phase 1 waits).
//   phase 1: copy this rank's lists into every rank's mailbox (own HBM + NVLink stores);
//            the last CTA to finish (ticket) publishes one flag per rank.
//   phase 2: wait for every rank's flag in the own mailbox, then merge per query.
__global__ void __launch_bounds__(256) peer_gather_merge_kernel(const PeerCtx c, const PeerGatherArgs a) {
  __shared__ uint32_t s_last;
  const Group g{threadIdx.x, blockDim.x, 0};
  const uint32_t k = a.k;
  for (uint32_t q = blockIdx.x; q < a.nq; q += gridDim.x) {
    const uint32_t n = min(__ldg(a.d_n + q), k);
    for (uint32_t e = g.tid; e < c.world * k; e += g.nthr) {
      const uint32_t r = e / k, i = e - r * k;
      if (i >= n) continue;
      const PeerBlock pb = peer_block(c, r, c.rank);
      pb.scores[(size_t)q * k + i] = a.d_scores[(size_t)q * k + i];
      pb.rows[(size_t)q * k + i] = a.d_rows[(size_t)q * k + i];
    }
    if (g.tid < c.world) peer_block(c, g.tid, c.rank).n[q] = n;
  }
  __threadfence_system();
  __syncthreads();
  if (g.tid == 0) s_last = (atomicAdd(a.d_ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) peer_signal(c, g);  // fence.sys again: the other CTAs' stores were observed through the ticket
  const bool ok = peer_wait(c, g);
  for (uint32_t q = blockIdx.x; q < a.nq; q += gridDim.x) {
    float* os = a.d_out_scores + (size_t)q * k;
    uint64_t* orow = a.d_out_rows + (size_t)q * k;
    if (ok) peer_merge_query(c, g, q, k, os, orow, a.d_out_n + q);
    else peer_emit_empty(g, k, os, orow, a.d_out_n + q);
  }
  __syncthreads();
  if (g.tid == 0 && atomicAdd(a.d_ticket + 1, 1u) == gridDim.x - 1) {
    a.d_ticket[0] = 0;
    a.d_ticket[1] = 0;
  }
}

cudaError_t launch_peer_gather_merge(const PeerCtx& c, const PeerGatherArgs& a, int num_sms,
                                     cudaStream_t st) {
  if (a.nq == 0 || a.k == 0 || a.nq > kPeerMaxQ || (uint64_t)a.nq * a.k > c.cap)
    return cudaErrorInvalidValue;
  const int grid = (int)(a.nq < (uint32_t)num_sms ? a.nq : (uint32_t)num_sms);
  peer_gather_merge_kernel<<<grid, 256, 0, st>>>(c, a);
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError();
}

cudaError_t peer_begin(cqs_b200_peer* p, cudaStream_t st, PeerCtx* c, bool exclusive) {
  if (++p->seq == 0) p->seq = kPeerSlots;  // 0 is the mailbox's initial flag value
  for (uint32_t i = 0; i < kPeerLanes; ++i) {
    if (!exclusive && i != p->seq % kPeerLanes) continue;
    if (p->ev_used[i] && p->ev_stream[i] != st) {
      cudaError_t e = cudaStreamWaitEvent(st, p->ev[i], 0);
      if (e != cudaSuccess) return e;
    }
  }
  c->world = p->world;
  c->rank = p->rank;
  c->seq = p->seq;
  c->cap = p->cap;
  c->block_bytes = p->block_bytes;
  c->timeout_ns = p->timeout_ns;
  c->status = p->d_status;
  for (uint32_t r = 0; r < kPeerMaxWorld; ++r) c->mbox[r] = p->mbox[r];
  return cudaSuccess;
}
cudaError_t peer_mark(cqs_b200_peer* p, cudaStream_t st, bool exclusive) {
  for (uint32_t i = 0; i < kPeerLanes; ++i) {
    if (!exclusive && i != p->seq % kPeerLanes) continue;
    cudaError_t e = cudaEventRecord(p->ev[i], st);
    if (e != cudaSuccess) return e;
    p->ev_stream[i] = st;
    p->ev_used[i] = true;
  }
  return cudaSuccess;
}

}  // namespace cqs

using namespace cqs;

// error text goes through the library's thread-local slot (index.cu)
int cqs_b200_internal_fail(int code, const char* msg);
static int pfail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  return cqs_b200_internal_fail(code, buf);
}

#define PCK(p, expr)                                                                          \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      if (p) (p)->failed.store(1);                                                            \
      cudaGetLastError();                                                                     \
      return pfail(_e == cudaErrorMemoryAllocation ? CQS_B200_ERR_OOM : CQS_B200_ERR_CUDA,    \
                   "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                         \
  } while (0)

extern "C" {

int cqs_b200_peer_create(int device, uint32_t world, uint32_t rank, uint32_t max_elems,
                         cqs_b200_peer** out) {
  if (!out) return pfail(CQS_B200_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (world < 1 || world > kPeerMaxWorld || rank >= world)
    return pfail(CQS_B200_ERR_INVALID, "world must be 1..%u and rank < world", kPeerMaxWorld);
  if (max_elems == 0) max_elems = 65536;
  if (max_elems < kMaxK) max_elems = kMaxK;
  if (max_elems > (1u << 22)) return pfail(CQS_B200_ERR_INVALID, "max_elems too large");
  cqs_b200_peer* p = new (std::nothrow) cqs_b200_peer();
  if (!p) return pfail(CQS_B200_ERR_OOM, "host allocation failed");
  p->device = device;
  p->world = world;
  p->rank = rank;
  p->cap = max_elems;
  p->block_bytes = peer_block_bytes(max_elems);
  p->bytes = (size_t)kPeerSlots * world * p->block_bytes;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_mbox, p->bytes);
  if (e == cudaSuccess) e = cudaMemset(p->d_mbox, 0, p->bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_status, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(p->d_status, 0, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_ticket, 2 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(p->d_ticket, 0, 2 * sizeof(uint32_t));
  for (uint32_t i = 0; i < kPeerLanes; ++i)
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaGetLastError();
    cqs_b200_peer_destroy(p);
    return pfail(e == cudaErrorMemoryAllocation ? CQS_B200_ERR_OOM : CQS_B200_ERR_CUDA,
                 "peer_create: %s", cudaGetErrorString(e));
  }
  p->mbox[rank] = p->d_mbox;
  if (world == 1) p->connected = true;
  *out = p;
  return CQS_B200_OK;
}

int cqs_b200_peer_handle(cqs_b200_peer* p, uint8_t* out_handle) {
  if (!p || !out_handle) return pfail(CQS_B200_ERR_INVALID, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == CQS_B200_PEER_HANDLE_BYTES, "handle size");
  std::lock_guard<std::mutex> g(p->mu);
  PCK(p, cudaSetDevice(p->device));
  cudaIpcMemHandle_t h;
  PCK(p, cudaIpcGetMemHandle(&h, p->d_mbox));
  memcpy(out_handle, &h, sizeof h);
  return CQS_B200_OK;
}

int cqs_b200_peer_connect(cqs_b200_peer* p, const uint8_t* handles) {
  if (!p || !handles) return pfail(CQS_B200_ERR_INVALID, "NULL argument");
  std::lock_guard<std::mutex> g(p->mu);
  if (p->connected) return pfail(CQS_B200_ERR_INVALID, "peer group already connected");
  PCK(p, cudaSetDevice(p->device));
  for (uint32_t r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * sizeof h, sizeof h);
    void* ptr = nullptr;
    PCK(p, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->mbox[r] = (uint8_t*)ptr;
    p->ipc_opened[r] = true;
  }
  p->connected = true;
  return CQS_B200_OK;
}

int cqs_b200_peer_connect_local(cqs_b200_peer** peers, uint32_t world) {
  if (!peers || world < 1 || world > kPeerMaxWorld) return pfail(CQS_B200_ERR_INVALID, "bad peer list");
  for (uint32_t r = 0; r < world; ++r)
    if (!peers[r] || peers[r]->world != world || peers[r]->rank != r || peers[r]->connected ||
        peers[r]->cap != peers[0]->cap)
      return pfail(CQS_B200_ERR_INVALID, "peers[%u] does not belong to this group", r);
  for (uint32_t r = 0; r < world; ++r) {
    cqs_b200_peer* p = peers[r];
    PCK(p, cudaSetDevice(p->device));
    for (uint32_t o = 0; o < world; ++o) {
      if (o == r) continue;
      if (peers[o]->device != p->device) {
        int can = 0;
        PCK(p, cudaDeviceCanAccessPeer(&can, p->device, peers[o]->device));
        if (!can) return pfail(CQS_B200_ERR_UNSUPPORTED, "device %d cannot access device %d", p->device, peers[o]->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(peers[o]->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else PCK(p, e);
      }
      p->mbox[o] = peers[o]->d_mbox;
    }
  }
  for (uint32_t r = 0; r < world; ++r) peers[r]->connected = true;
  return CQS_B200_OK;
}

int cqs_b200_peer_gather_merge(cqs_b200_peer* p, const float* d_scores, const uint64_t* d_rows,
                               const uint32_t* d_n, uint32_t nq, uint32_t k, float* d_out_scores,
                               uint64_t* d_out_rows, uint32_t* d_out_n, void* stream) {
  if (!p || !d_scores || !d_rows || !d_n || !d_out_scores || !d_out_rows || !d_out_n)
    return pfail(CQS_B200_ERR_INVALID, "NULL argument");
  if (nq == 0 || nq > kPeerMaxQ || k == 0 || k > kMaxK || (uint64_t)nq * k > p->cap)
    return pfail(CQS_B200_ERR_INVALID, "nq=%u k=%u exceed the mailbox (nq <= %u, nq*k <= %u)", nq, k,
                 kPeerMaxQ, p->cap);
  std::lock_guard<std::mutex> g(p->mu);
  if (!p->connected) return pfail(CQS_B200_ERR_INVALID, "peer group is not connected");
  if (p->failed.load()) return pfail(CQS_B200_ERR_POISONED, "peer group failed earlier; rebuild it");
  PCK(p, cudaSetDevice(p->device));
  cudaStream_t st = (cudaStream_t)stream;
  PeerCtx c;
  PCK(p, peer_begin(p, st, &c, /*exclusive=*/true));
  PeerGatherArgs a{d_scores, d_rows, d_n, nq, k, d_out_scores, d_out_rows, d_out_n, p->d_ticket};
  PCK(p, launch_peer_gather_merge(c, a, p->num_sms, st));
  PCK(p, peer_mark(p, st, /*exclusive=*/true));
  return CQS_B200_OK;
}

int cqs_b200_peer_status(cqs_b200_peer* p) {
  if (!p) return pfail(CQS_B200_ERR_INVALID, "peer is NULL");
  std::lock_guard<std::mutex> g(p->mu);
  if (p->failed.load()) return 1;
  PCK(p, cudaSetDevice(p->device));
  uint32_t s = 0;
  PCK(p, cudaMemcpy(&s, p->d_status, sizeof s, cudaMemcpyDeviceToHost));  // synchronises the device
  if (s) p->failed.store(1);
  return s ? 1 : 0;
}

int cqs_b200_peer_set_timeout_ms(cqs_b200_peer* p, uint32_t ms) {
  if (!p || ms == 0) return pfail(CQS_B200_ERR_INVALID, "bad argument");
  std::lock_guard<std::mutex> g(p->mu);
  p->timeout_ns = (uint64_t)ms * 1000000ull;
  return CQS_B200_OK;
}

void cqs_b200_peer_destroy(cqs_b200_peer* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  for (uint32_t r = 0; r < kPeerMaxWorld; ++r)
    if (p->ipc_opened[r]) cudaIpcCloseMemHandle(p->mbox[r]);
  cudaFree(p->d_mbox);
  cudaFree(p->d_status);
  cudaFree(p->d_ticket);
  for (uint32_t i = 0; i < kPeerLanes; ++i)
    if (p->ev[i]) cudaEventDestroy(p->ev[i]);
  cudaGetLastError();
  delete p;
}

}  // extern "C"
