// peer_host.h — host-side state of a peer group (cqs_b200_peer, include/cqs_b200.h).
// Shared by peer.cu (lifecycle, the stand-alone gather+merge) and index.cu (the
// scan kernels that carry the exchange in their tail).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>

#include "peer.cuh"

struct cqs_b200_peer {
  std::mutex mu;
  int device = 0;
  int num_sms = 148;
  uint32_t world = 0, rank = 0, cap = 0;
  uint64_t block_bytes = 0;
  size_t bytes = 0;                 // size of one mailbox
  uint8_t* d_mbox = nullptr;        // this rank's mailbox (cudaMalloc: exportable over CUDA IPC)
  uint8_t* mbox[cqs::kPeerMaxWorld] = {};
  bool ipc_opened[cqs::kPeerMaxWorld] = {};
  bool connected = false;
  uint32_t* d_status = nullptr;     // sticky: 1 = an exchange timed out
  uint32_t* d_ticket = nullptr;     // [2] CTA tickets of the stand-alone kernel
  uint32_t seq = 0;                 // exchanges issued so far (identical on every rank)
  uint64_t timeout_ns = 5ull * 1000 * 1000 * 1000;
  // Mailbox slots are reused every kPeerSlots exchanges and a fused scan may overlap the
  // kPeerLanes-1 before it (the caller rotates over that many streams): exchange seq waits for
  // exchange seq-kPeerLanes (events by seq % kPeerLanes).  The stand-alone kernel shares
  // d_ticket, so it is `exclusive`: ordered after every earlier exchange, and every later one after it.
  cudaEvent_t ev[cqs::kPeerLanes] = {};
  cudaStream_t ev_stream[cqs::kPeerLanes] = {};
  bool ev_used[cqs::kPeerLanes] = {};
  std::atomic<int> failed{0};
};

namespace cqs {
// Next exchange of the group: fills `c` and orders stream `st` after the previous exchange.
// Call with p->mu held; follow the launch with peer_mark(p, st).
cudaError_t peer_begin(cqs_b200_peer* p, cudaStream_t st, PeerCtx* c, bool exclusive);
cudaError_t peer_mark(cqs_b200_peer* p, cudaStream_t st, bool exclusive);
}  // namespace cqs
