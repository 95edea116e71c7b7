// Conformance tests of the C++ host mirror, written after the reference's own backend tests:
//   k = 0 / k > n / duplicates            src/hnsw/build.rs:603-632, src/cagra.rs:2012-2044
//   all-reject filter, filtered ids only  src/hnsw/build.rs:637-689
//   wrong-dim / NaN / Inf query -> empty  tests/hnsw_test.rs:212, :460-552, src/cagra.rs:2564-2605
//   zero vector never yields non-finite   src/cagra.rs:2606-2640
//   sorted descending                     src/cagra.rs:1924-1940
//   SpladeIndex fixture                   src/splade/index.rs:1114-1241
//   legs corpus fusion                    tests/search_test.rs:549-791
#include <cassert>
#include <cmath>
#include <cstdio>
#include <memory>
#include <set>

#include "../../cqs_b200/host/b200_index.hpp"

using namespace cqs;
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static Embedding unit(size_t dim, uint64_t seed) {
  Embedding v(dim); double n = 0; uint64_t s = seed * 0x9E3779B97F4A7C15ull + 1;
  for (auto& x : v) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (float)((double)(s >> 11) / 9007199254740992.0) * 2.f - 1.f; n += (double)x * x; }
  for (auto& x : v) x = (float)(x / std::sqrt(n));
  return v;
}

int main() {
  const size_t dim = 768, n = 600;
  std::vector<std::pair<std::string, Embedding>> feed;
  for (size_t i = 0; i < n; ++i) { char id[32]; std::snprintf(id, sizeof id, "chunk_%04zu", (i * 7919) % 10000); feed.push_back({id, unit(dim, i + 1)}); }
  feed[5].second.assign(dim, 0.f);  // zero vector: dropped by the feed like prepare_index_data
  auto ix = B200Index::build(feed, dim);
  CHECK(ix != nullptr);
  CHECK(ix->len() == n - 1 && ix->dim() == dim && std::string(ix->name()) == "B200");
  CHECK(!ix->is_poisoned() && ix->max_k().value() == 1024 && ix->index_scores_are_cosine());
  CHECK(std::is_sorted(ix->id_map().begin(), ix->id_map().end()));
  const Embedding q = feed[42].second;
  CHECK(ix->search(q, 0).empty());                                         // k = 0
  auto r = ix->search(q, 5);
  CHECK(r.size() == 5 && r[0].id == feed[42].first && std::fabs(r[0].score - 1.f) < 1e-5f);
  for (size_t i = 1; i < r.size(); ++i) CHECK(r[i - 1].score >= r[i].score); // sorted descending
  auto all = ix->search(q, 1024);                                          // k > n: <= n unique real ids
  CHECK(all.size() == n - 1);
  std::set<std::string> uniq; for (auto& x : all) { uniq.insert(x.id); CHECK(std::isfinite(x.score)); }
  CHECK(uniq.size() == n - 1);
  CHECK(ix->search(q, 2000).empty());                                      // k > max_k -> refused -> empty Vec
  CHECK(ix->search(Embedding(100, 0.1f), 5).empty());                      // wrong dim
  Embedding bad = q; bad[3] = NAN; CHECK(ix->search(bad, 5).empty());      // NaN
  bad[3] = INFINITY; CHECK(ix->search(bad, 5).empty());                    // Inf
  Filter even = [](std::string_view id) { return (id.back() - '0') % 2 == 0; };
  auto fr = ix->search_with_filter(q, 10, even);
  CHECK(fr.size() == 10); for (auto& x : fr) CHECK(even(x.id));            // filtered ids only
  CHECK(ix->search_with_filter(q, 10, [](std::string_view) { return false; }).empty());  // all-reject
  auto ar = ix->search_with_filter(q, 5, [](std::string_view) { return true; });
  CHECK(ar.size() == 5 && ar[0].id == r[0].id);
  Filter one = [&](std::string_view id) { return id == feed[7].first; };
  auto o = ix->search_with_filter(q, 10, one);
  CHECK(o.size() == 1 && o[0].id == feed[7].first);                        // k capped at `included`
  std::vector<Embedding> qs; for (int i = 0; i < 12; ++i) qs.push_back(feed[100 + i].second);
  qs[3] = Embedding(5, 1.f);
  auto br = ix->search_batch(qs, 4);
  CHECK(br.size() == 12 && br[3].empty() && br[0].size() == 4 && br[0][0].id == feed[100].first);

  // --- SpladeIndex fixture (src/splade/index.rs:1114-1241) on a tiny index ---
  std::vector<std::pair<std::string, Embedding>> f3 = {{"chunk_a", unit(dim, 11)}, {"chunk_b", unit(dim, 12)}, {"chunk_c", unit(dim, 13)}};
  auto s3 = B200Index::build(f3, dim);
  CHECK(s3 && s3->attach_splade({{"chunk_a", {{1, .5f}, {2, .3f}, {3, .8f}}}, {"chunk_b", {{1, .7f}, {4, .6f}}}, {"chunk_c", {{2, .9f}, {3, .1f}, {5, .4f}}}}));
  auto sr = s3->splade_search_with_filter({{1, 1.f}, {2, 1.f}}, 10);
  CHECK(sr.size() == 3 && sr[0].id == "chunk_c" && sr[1].id == "chunk_a" && sr[2].id == "chunk_b");
  CHECK(std::fabs(sr[0].score - .9f) < 1e-5f && std::fabs(sr[1].score - .8f) < 1e-5f && std::fabs(sr[2].score - .7f) < 1e-5f);
  CHECK(s3->splade_search_with_filter({{999, 1.f}}, 10).empty());
  CHECK(s3->splade_search_with_filter({}, 10).empty());
  CHECK(s3->splade_search_with_filter({{1, 1.f}}, 0).empty());
  Filter onlya = [](std::string_view id) { return id == "chunk_a"; };
  auto sf = s3->splade_search_with_filter({{1, 1.f}}, 10, &onlya);
  CHECK(sf.size() == 1 && sf[0].id == "chunk_a");
  auto hy = s3->search_hybrid(f3[0].second, {{1, 1.f}}, 0.5f, 5);
  CHECK(hy.size() == 3 && hy[0].id == "chunk_a" && hy[0].in_dense);
  for (size_t i = 1; i < hy.size(); ++i) CHECK(hy[i - 1].fused >= hy[i].fused);
  CHECK(candidate_count_for(5) == 500 && candidate_count_for(300) == 1500 && cap_k_to_backend(*ix, 1500) == 1024);
  std::printf("host mirror OK\n");
  return 0;
}
