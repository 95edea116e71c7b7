// b200_index.hpp — C++ host-side mirror of the reference's Rust interface for the
// retrieval hot path, written above the C ABI (include/cqs_b200.h).  The reference
// is Rust and there is no Rust toolchain in this image, so this header plays the
// role of the shim in INTEGRATION.md: same names, same argument meaning, same
// error behaviour as
//   trait VectorIndex                       src/index.rs:139-239
//   CagraIndex::{search,search_with_filter} src/cagra.rs:443-500, :727-820 (conventions)
//   SpladeIndex::search_with_filter         src/splade/index.rs:223-291
//   search_hybrid_inner (legs + fusion)     src/search/query.rs:880-1005
// Header-only; link with libcqs_b200.so.  No arithmetic happens here.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <numeric>
#include <optional>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "../../include/cqs_b200.h"

namespace cqs {

struct IndexResult {  // src/index.rs:129
  std::string id;
  float score;
};

enum class DistanceMetric { Cosine = CQS_B200_METRIC_COSINE, DotProduct = CQS_B200_METRIC_DOT };
enum class Storage { F32 = CQS_B200_STORAGE_F32, BF16 = CQS_B200_STORAGE_BF16, BF16_F32 = CQS_B200_STORAGE_BF16_F32 };
using Embedding = std::vector<float>;
using SparseVector = std::vector<std::pair<uint32_t, float>>;  // src/splade/mod.rs SparseVector
using Filter = std::function<bool(std::string_view)>;

class VectorIndex {  // trait VectorIndex: Send + Sync
 public:
  virtual ~VectorIndex() = default;
  virtual std::vector<IndexResult> search(const Embedding& query, size_t k) const = 0;
  virtual size_t len() const = 0;
  virtual bool is_empty() const { return len() == 0; }
  virtual const char* name() const = 0;
  virtual size_t dim() const = 0;
  // default impl: over-fetch 3k, post-filter (src/index.rs:167-193)
  virtual std::vector<IndexResult> search_with_filter(const Embedding& query, size_t k, const Filter& filter) const {
    std::vector<IndexResult> out;
    for (auto& r : search(query, k > SIZE_MAX / 3 ? SIZE_MAX : k * 3)) {
      if (out.size() >= k) break;
      if (filter(r.id)) out.push_back(std::move(r));
    }
    return out;
  }
  virtual bool is_poisoned() const { return false; }
  virtual std::optional<size_t> max_k() const { return std::nullopt; }
  virtual bool index_scores_are_cosine() const { return false; }
};

struct FusedCandidate {  // one row of the fused pool + the SearchLegs values (src/search/query.rs:39-208)
  std::string id;
  float fused, dense, sparse_raw;
  bool in_dense, in_sparse;
};

inline size_t candidate_count_for(size_t limit, size_t floor = 500) {  // src/limits.rs:315-320
  size_t five = limit > SIZE_MAX / 5 ? SIZE_MAX : limit * 5;
  return std::max(five, floor);
}
inline size_t cap_k_to_backend(const VectorIndex& idx, size_t k) {  // src/search/query.rs:232-245
  auto cap = idx.max_k();
  return (cap && k > *cap) ? *cap : k;
}

// A rank's handle on the peer-memory exchange of a row-sharded corpus (include/cqs_b200.h, "row-sharded
// corpora WITHOUT a collective call").  One process per GPU: create(), send handle() to the other ranks over
// any channel, connect() with all handles in rank order.
class PeerGroup {
 public:
  static std::unique_ptr<PeerGroup> create(int device, uint32_t world, uint32_t rank, uint32_t max_elems = 0) {
    std::unique_ptr<PeerGroup> g(new PeerGroup());
    if (cqs_b200_peer_create(device, world, rank, max_elems, &g->h_) != CQS_B200_OK) return nullptr;
    g->world_ = world;
    return g;
  }
  ~PeerGroup() { cqs_b200_peer_destroy(h_); }
  PeerGroup(const PeerGroup&) = delete;
  PeerGroup& operator=(const PeerGroup&) = delete;
  std::array<uint8_t, CQS_B200_PEER_HANDLE_BYTES> handle() const {
    std::array<uint8_t, CQS_B200_PEER_HANDLE_BYTES> h{};
    cqs_b200_peer_handle(h_, h.data());
    return h;
  }
  bool connect(const std::vector<std::array<uint8_t, CQS_B200_PEER_HANDLE_BYTES>>& all) {
    if (all.size() != world_) return false;
    std::vector<uint8_t> flat;
    for (auto& h : all) flat.insert(flat.end(), h.begin(), h.end());
    return cqs_b200_peer_connect(h_, flat.data()) == CQS_B200_OK;
  }
  bool healthy() const { return cqs_b200_peer_status(h_) == 0; }
  cqs_b200_peer* raw() const { return h_; }

 private:
  PeerGroup() = default;
  cqs_b200_peer* h_ = nullptr;
  uint32_t world_ = 0;
};

class B200Index final : public VectorIndex {
 public:
  // build_from_store analogue (src/cagra.rs:842-919): takes the (chunk_id, embedding) feed in
  // store (rowid) order, drops zero / non-finite vectors (src/hnsw/mod.rs:716-736), sorts by
  // chunk id so that device row order == id order, uploads.  Returns nullptr on failure
  // (try_open -> Ok(None): fall through to the next backend, src/index.rs:281-290).
  static std::unique_ptr<B200Index> build(const std::vector<std::pair<std::string, Embedding>>& feed, size_t dim,
                                          DistanceMetric metric = DistanceMetric::Cosine,
                                          Storage storage = Storage::F32, int device = 0) {
    std::vector<size_t> order;
    for (size_t i = 0; i < feed.size(); ++i) {
      const Embedding& e = feed[i].second;
      if (e.size() != dim) continue;
      bool finite = true, nonzero = false;
      for (float x : e) { finite &= std::isfinite(x); nonzero |= (x != 0.0f); }
      if (finite && nonzero) order.push_back(i);
    }
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return feed[a].first < feed[b].first; });
    std::unique_ptr<B200Index> ix(new B200Index());
    ix->dim_ = dim;
    if (cqs_b200_create(&device, 1, (uint32_t)dim, (int)metric, (int)storage, &ix->h_) != CQS_B200_OK) return nullptr;
    if (cqs_b200_reserve(ix->h_, order.size()) != CQS_B200_OK) return nullptr;
    std::vector<float> block;
    const size_t kBlock = 4096;
    for (size_t b = 0; b < order.size(); b += kBlock) {
      size_t m = std::min(kBlock, order.size() - b);
      block.resize(m * dim);
      for (size_t i = 0; i < m; ++i) {
        const auto& f = feed[order[b + i]];
        std::copy(f.second.begin(), f.second.end(), block.begin() + i * dim);
        ix->id_map_.push_back(f.first);
      }
      if (cqs_b200_append_rows_f32(ix->h_, block.data(), m) != CQS_B200_OK) return nullptr;
    }
    if (cqs_b200_finalize(ix->h_) != CQS_B200_OK) return nullptr;
    return ix;
  }
  ~B200Index() override { cqs_b200_destroy(h_); }
  B200Index(const B200Index&) = delete;
  B200Index& operator=(const B200Index&) = delete;

  std::vector<IndexResult> search(const Embedding& query, size_t k) const override {
    return search_bits(query, k, nullptr);
  }
  size_t len() const override { return (size_t)cqs_b200_len(h_); }
  const char* name() const override { return cqs_b200_name(); }
  size_t dim() const override { return dim_; }
  bool is_poisoned() const override { return cqs_b200_is_poisoned(h_) != 0; }
  std::optional<size_t> max_k() const override { return (size_t)cqs_b200_max_k(h_); }
  bool index_scores_are_cosine() const override { return cqs_b200_scores_are_cosine(h_) != 0; }

  // src/cagra.rs:727-820: host bitset, all-pass -> unfiltered, none-pass -> empty, k = min(k, included)
  std::vector<IndexResult> search_with_filter(const Embedding& query, size_t k, const Filter& filter) const override {
    if (id_map_.empty() || k == 0) return {};
    size_t included = 0;
    std::vector<uint32_t> bits = bitset_for(filter, &included);
    if (included == id_map_.size()) return search(query, k);
    if (included == 0) return {};
    return search_bits(query, std::min(k, included), bits.data());
  }

  // inherent (no trait counterpart): nq queries in one call
  std::vector<std::vector<IndexResult>> search_batch(const std::vector<Embedding>& queries, size_t k) const {
    std::vector<std::vector<IndexResult>> out(queries.size());
    if (queries.empty() || k == 0 || k > CQS_B200_MAX_K) return out;
    std::vector<float> flat(queries.size() * dim_, 0.f);
    for (size_t i = 0; i < queries.size(); ++i) {
      if (queries[i].size() != dim_) { flat[i * dim_] = NAN; continue; }  // wrong dim -> empty
      std::copy(queries[i].begin(), queries[i].end(), flat.begin() + i * dim_);
    }
    std::vector<uint64_t> rows(queries.size() * k);
    std::vector<float> scores(queries.size() * k);
    std::vector<uint32_t> n(queries.size());
    if (cqs_b200_search_batch(h_, flat.data(), (uint32_t)queries.size(), (uint32_t)k, nullptr, rows.data(),
                              scores.data(), n.data()) != CQS_B200_OK) {
      log_error("search_batch");
      return out;
    }
    for (size_t i = 0; i < queries.size(); ++i) out[i] = to_results(rows.data() + i * k, scores.data() + i * k, n[i]);
    return out;
  }

  // This index holds ONE shard of a row-sharded corpus (cqs_b200_set_row_base): VectorIndex::search over
  // the whole corpus.  Every rank must make the same call; each gets the global top-k as (global row,
  // score) — the chunk-id strings of other shards live with their owners, so rows are returned as they are.
  std::vector<std::pair<uint64_t, float>> search_sharded_rows(const PeerGroup& peer, const Embedding& query, size_t k) const {
    std::vector<std::pair<uint64_t, float>> out;
    if (k == 0 || k > CQS_B200_MAX_K || query.size() != dim_ || is_poisoned()) return out;
    std::vector<uint64_t> rows(k); std::vector<float> scores(k); uint32_t n = 0;
    if (cqs_b200_search_sharded(h_, peer.raw(), query.data(), (uint32_t)k, nullptr, rows.data(), scores.data(), &n) != CQS_B200_OK) {
      log_error("search_sharded");
      return out;
    }
    for (uint32_t i = 0; i < n; ++i)
      if (std::isfinite(scores[i])) out.push_back({rows[i], scores[i]});
    return out;
  }

  // SpladeIndex::build over the same chunk ids (src/splade/index.rs:191-211)
  bool attach_splade(const std::vector<std::pair<std::string, SparseVector>>& chunks, uint32_t vocab = 30522) {
    std::vector<uint64_t> indptr(id_map_.size() + 1, 0);
    std::vector<const SparseVector*> by_row(id_map_.size(), nullptr);
    for (auto& c : chunks) {
      auto it = std::lower_bound(id_map_.begin(), id_map_.end(), c.first);
      if (it != id_map_.end() && *it == c.first) by_row[it - id_map_.begin()] = &c.second;
    }
    std::vector<uint32_t> tok;
    std::vector<float> w;
    for (size_t r = 0; r < id_map_.size(); ++r) {
      if (by_row[r]) for (auto& tw : *by_row[r]) { tok.push_back(tw.first); w.push_back(tw.second); }
      indptr[r + 1] = tok.size();
    }
    return cqs_b200_sparse_attach(h_, indptr.data(), tok.data(), w.data(), vocab) == CQS_B200_OK;
  }
  // SpladeIndex::search_with_filter (src/splade/index.rs:223-291)
  std::vector<IndexResult> splade_search_with_filter(const SparseVector& query, size_t k, const Filter* filter = nullptr) const {
    if (query.empty() || id_map_.empty() || k == 0) return {};
    std::vector<uint32_t> bits;
    if (filter) { size_t inc = 0; bits = bitset_for(*filter, &inc); if (inc == 0) return {}; }
    std::vector<uint32_t> qt; std::vector<float> qw;
    for (auto& tw : query) { qt.push_back(tw.first); qw.push_back(tw.second); }
    k = std::min<size_t>(k, CQS_B200_MAX_K);
    std::vector<uint64_t> rows(k); std::vector<float> scores(k); uint32_t n = 0;
    if (cqs_b200_search_sparse(h_, qt.data(), qw.data(), (uint32_t)qt.size(), (uint32_t)k, filter ? bits.data() : nullptr,
                               rows.data(), scores.data(), &n) != CQS_B200_OK) { log_error("splade_search"); return {}; }
    return to_results(rows.data(), scores.data(), n);
  }
  // the two leg calls + alpha fusion of search_hybrid_inner (src/search/query.rs:880-1005) in one call
  std::vector<FusedCandidate> search_hybrid(const Embedding& query, const SparseVector& sparse_query, float alpha,
                                            size_t limit, const Filter* filter = nullptr) const {
    size_t pool_k = cap_k_to_backend(*this, candidate_count_for(limit));
    std::vector<uint32_t> bits;
    if (filter) { size_t inc = 0; bits = bitset_for(*filter, &inc); if (inc == 0) return {}; }
    std::vector<float> q(dim_, NAN);  // wrong-dim dense query -> empty dense leg, sparse leg still runs
    if (query.size() == dim_) q = query;
    std::vector<uint32_t> qt; std::vector<float> qw;
    for (auto& tw : sparse_query) { qt.push_back(tw.first); qw.push_back(tw.second); }
    std::vector<uint64_t> rows(pool_k); std::vector<float> fused(pool_k), dense(pool_k), sraw(pool_k);
    std::vector<uint8_t> present(pool_k); uint32_t n = 0;
    if (cqs_b200_search_hybrid(h_, q.data(), qt.data(), qw.data(), (uint32_t)qt.size(), alpha, (uint32_t)pool_k,
                               filter ? bits.data() : nullptr, rows.data(), fused.data(), dense.data(), sraw.data(),
                               present.data(), &n) != CQS_B200_OK) { log_error("search_hybrid"); return {}; }
    std::vector<FusedCandidate> out;
    for (uint32_t i = 0; i < n; ++i)
      if (rows[i] < id_map_.size())
        out.push_back({id_map_[rows[i]], fused[i], dense[i], sraw[i], (present[i] & 1) != 0, (present[i] & 2) != 0});
    return out;
  }
  const std::vector<std::string>& id_map() const { return id_map_; }

 private:
  B200Index() = default;
  std::vector<uint32_t> bitset_for(const Filter& f, size_t* included) const {  // src/cagra.rs:747-757
    std::vector<uint32_t> bits((id_map_.size() + 31) / 32, 0u);
    size_t inc = 0;
    for (size_t i = 0; i < id_map_.size(); ++i)
      if (f(id_map_[i])) { bits[i / 32] |= 1u << (i % 32); ++inc; }
    *included = inc;
    return bits;
  }
  std::vector<IndexResult> to_results(const uint64_t* rows, const float* scores, uint32_t n) const {
    std::vector<IndexResult> out;
    for (uint32_t i = 0; i < n; ++i)
      if (rows[i] < id_map_.size() && std::isfinite(scores[i]))  // drop out-of-range slots (src/cagra.rs:642-669)
        out.push_back({id_map_[rows[i]], scores[i]});
    return out;
  }
  std::vector<IndexResult> search_bits(const Embedding& query, size_t k, const uint32_t* bits) const {
    if (id_map_.empty() || k == 0) return {};        // src/cagra.rs:445
    if (query.size() != dim_) return {};             // dimension mismatch -> warn + empty (:449-456)
    if (is_poisoned()) return {};                    // :488
    std::vector<uint64_t> rows(k); std::vector<float> scores(k); uint32_t n = 0;
    if (k > CQS_B200_MAX_K || cqs_b200_search(h_, query.data(), (uint32_t)k, bits, rows.data(), scores.data(), &n) != CQS_B200_OK) {
      log_error("search");                           // every device failure -> error log + empty Vec (:541-627)
      return {};
    }
    return to_results(rows.data(), scores.data(), n);
  }
  static void log_error(const char* what) { std::fprintf(stderr, "[b200] %s failed: %s\n", what, cqs_b200_last_error()); }
  cqs_b200_index* h_ = nullptr;
  size_t dim_ = 0;
  std::vector<std::string> id_map_;
};

}  // namespace cqs
