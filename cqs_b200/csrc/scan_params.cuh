// scan_params.cuh — kernel parameter block of scan_topk_kernel (shared by the dispatcher in
// scan_single.cu and the per-variant translation units).
#pragma once
#include <stdint.h>

#include "internal.h"
#include "peer.cuh"

namespace cqs {

struct ScanParams {
  const uint8_t* rows;
  uint64_t n_rows;
  uint64_t row_bytes;
  const float* query;
  const uint32_t* bitset;
  uint32_t k;
  uint64_t row_base;
  ckey_t* partial;
  uint32_t* partial_cnt;
  uint32_t* done;        // [0] finished-CTA ticket, [1] dynamic tile counter; zero between launches
  ckey_t* col;           // [kMaxGrid] large k: per-CTA published m-th-key bounds; zero between launches
  ckey_t* excl;          // [kMaxGrid] shadow scan: per-CTA bound of the shadow keys it did not re-score
  float* out_scores;
  uint64_t* out_rows;
  uint32_t* out_n;
  unsigned long long* trace;  // optional [grid][8] globaltimer stamps (development aid)
  ScanSignals sig;            // structured filter + per-row signals (sig.pipeline / sig.d_ctype gate them)
  uint32_t* host_flag;        // optional: host-mapped word that receives `seq` once the result is written
  uint32_t seq;
  uint32_t chunk = 1;         // tiles per big work unit (filled in by the launcher)
  uint64_t n_big = 0;         // number of big work units; the remaining tiles are handed out one by one
  uint32_t chunk_override = 0;  // development aid: CQS_B200_CHUNK
  PeerCtx peer;               // peer.world != 0: exchange the local list with the other shards and emit the GLOBAL top-k
  const uint8_t* exact_rows = nullptr;  // f32 master rows: re-score the k candidates, keep k_out (ScanArgs)
  uint32_t exact_nv = 0;
  uint32_t k_out = 0;
  float max_row_delta = 0.f;
  float max_row_norm = 0.f;
};

}  // namespace cqs
