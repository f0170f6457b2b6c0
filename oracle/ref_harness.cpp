// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Harness that instantiates the UNMODIFIED reference headers
//   /root/reference/library/include/binary/algorithm/interval_tree.hpp
//   /root/reference/library/include/binary/algorithm/rb_tree.hpp
// (included from where they lie; no reference source is copied into this repo) and
// exposes the reference IntervalTree through a small C ABI so tests and the
// `bench.py --impl reference` arm can drive it from Python via ctypes.
//
// How ids are carried: the reference returns interval *copies* (interval_tree.hpp:306-328),
// so we use the same mechanism sv2nl uses to carry a VCF record (parser/vcf.hpp:598-639):
// a subclass of UIntInterval with a payload -- here a u32 target id.
//
// One tree per group, mirroring sv2nl's one-tree-per-chromosome (sv2nl mapper.hpp:147-162).
// Built into oracle/_ref/libbinary_ref.so by oracle/Makefile (only when /root/reference exists).

#include <spdlog/fmt/ostr.h>
#include <spdlog/spdlog.h>

#include <binary/algorithm/interval_tree.hpp>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <thread>
#include <vector>

namespace bt = binary::algorithm::tree;

// The image ships spdlog 1.14/fmt 10 (the reference pins 1.10/fmt 8, library/CMakeLists.txt:38-43);
// fmt >= 9 no longer formats operator<< types implicitly, so the trace calls in the reference
// rotations (interval_tree.hpp:208-215) need this opt-in. It lives here, outside the reference.
template <class I>
struct fmt::formatter<bt::IntervalNode<I>> : fmt::ostream_formatter {};

namespace {

struct IdInterval : bt::UIntInterval {
  std::uint32_t id{};
  IdInterval() = default;
  // field-set instead of the 2-arg base ctor: that ctor asserts low<=high in debug builds
  // (interval_tree.hpp:115-117) and TraMapper legitimately inserts inverted intervals.
  IdInterval(std::uint32_t l, std::uint32_t h, std::uint32_t i) : id(i) {
    low = l;
    high = h;
  }
};
static_assert(bt::IntervalConcept<IdInterval>);

using Tree = bt::IntervalTree<bt::IntervalNode<IdInterval>>;

struct Forest {
  std::map<std::uint32_t, std::unique_ptr<Tree>> trees;
  std::uint64_t n = 0;
};

int black_height(const bt::IntervalNode<IdInterval>* node, bool& ok) {
  if (node == nullptr) return 0;
  int l = black_height(node->leftr(), ok);
  int r = black_height(node->rightr(), ok);
  if (l != r) ok = false;
  return l + (node->is_black() ? 1 : 0);
}

// checks the augmentation invariant max == max(high, max(left), max(right)) everywhere
bool check_max(const bt::IntervalNode<IdInterval>* node, std::uint32_t& out_max) {
  if (node == nullptr) {
    out_max = 0;
    return true;
  }
  std::uint32_t lm = 0, rm = 0;
  bool ok = check_max(node->leftr(), lm) && check_max(node->rightr(), rm);
  std::uint32_t m = node->interval.high;
  if (node->leftr()) m = std::max(m, lm);
  if (node->rightr()) m = std::max(m, rm);
  out_max = m;
  return ok && node->max == m;
}

}  // namespace

extern "C" {

void* ref_build(std::uint64_t n, const std::uint32_t* group, const std::uint32_t* low,
                const std::uint32_t* high) {
  spdlog::set_level(spdlog::level::off);
  auto* f = new Forest();
  f->n = n;
  for (std::uint64_t i = 0; i < n; ++i) {
    std::uint32_t g = group ? group[i] : 0u;
    auto& t = f->trees[g];
    if (!t) t = std::make_unique<Tree>();
    // reference: RbTree::insert_node(Args&&...) rb_tree.hpp:145-149
    t->insert_node(IdInterval{low[i], high[i], static_cast<std::uint32_t>(i)});
  }
  return f;
}

void ref_free(void* forest) { delete static_cast<Forest*>(forest); }

std::uint64_t ref_size(const void* forest, std::uint32_t group) {
  auto* f = static_cast<const Forest*>(forest);
  auto it = f->trees.find(group);
  return it == f->trees.end() ? 0 : it->second->size();  // rb_tree.hpp:173-180
}

// returns 1 and fills key/low/high/id of the root node; 0 when the group has no tree
int ref_root(const void* forest, std::uint32_t group, std::uint32_t* key, std::uint32_t* low,
             std::uint32_t* high, std::uint32_t* id, std::uint32_t* max) {
  auto* f = static_cast<const Forest*>(forest);
  auto it = f->trees.find(group);
  if (it == f->trees.end() || it->second->root() == nullptr) return 0;
  auto* r = it->second->root();
  *key = r->key;
  *low = r->interval.low;
  *high = r->interval.high;
  *id = r->interval.id;
  *max = r->max;
  return 1;
}

// black height of the group's tree, or -1 when two root-to-leaf paths disagree
// (the reference test's check_black_height, test_interval_tree.cpp:18-29); also checks `max`.
int ref_check_invariants(const void* forest, std::uint32_t group) {
  auto* f = static_cast<const Forest*>(forest);
  auto it = f->trees.find(group);
  if (it == f->trees.end()) return 0;
  bool ok = true;
  int h = black_height(it->second->root(), ok);
  std::uint32_t m = 0;
  if (!check_max(it->second->root(), m)) return -2;
  return ok ? h : -1;
}

// reference find_overlap (first hit), interval_tree.hpp:290-304
int ref_find_overlap(const void* forest, std::uint32_t group, std::uint32_t qlow,
                     std::uint32_t qhigh, std::uint32_t* low, std::uint32_t* high,
                     std::uint32_t* id) {
  auto* f = static_cast<const Forest*>(forest);
  auto it = f->trees.find(group);
  if (it == f->trees.end()) return 0;
  auto res = it->second->find_overlap(IdInterval{qlow, qhigh, 0u});
  if (!res) return 0;
  *low = res->low;
  *high = res->high;
  *id = res->id;
  return 1;
}

// Batched reference find_overlaps (interval_tree.hpp:161-168,306-334) from `threads` std::threads over
// contiguous query ranges; trees are shared read-only as TraMapper does (sv2nl mapper.cpp:136-140).
// offsets: n_q+1 entries (CSR). *targets: malloc'd, `offsets[n_q]` u32 ids in the reference's
// NATIVE (preorder) order per query; caller frees with ref_free_buf. seconds: wall time of the
// query phase only (result materialisation included, as the reference returns vectors by value).
int ref_query(const void* forest, std::uint64_t n_q, const std::uint32_t* qgroup,
              const std::uint32_t* qlow, const std::uint32_t* qhigh, int threads,
              std::uint64_t* offsets, std::uint32_t** targets, double* seconds) {
  auto* f = static_cast<const Forest*>(forest);
  if (threads < 1) threads = 1;
  if (static_cast<std::uint64_t>(threads) > n_q && n_q > 0) threads = static_cast<int>(n_q);
  std::vector<std::vector<std::uint32_t>> hits(threads);
  std::vector<std::uint64_t> counts(n_q);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int t) {
    std::uint64_t b = n_q * t / threads, e = n_q * (t + 1) / threads;
    auto& out = hits[t];
    for (std::uint64_t i = b; i < e; ++i) {
      auto it = f->trees.find(qgroup ? qgroup[i] : 0u);
      if (it == f->trees.end()) {
        counts[i] = 0;
        continue;
      }
      auto res = it->second->find_overlaps(IdInterval{qlow[i], qhigh[i], 0u});
      counts[i] = res.size();
      for (auto const& r : res) out.push_back(r.id);
    }
  };
  if (threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  std::uint64_t acc = 0;
  for (std::uint64_t i = 0; i < n_q; ++i) {
    offsets[i] = acc;
    acc += counts[i];
  }
  offsets[n_q] = acc;
  if (targets) {
    auto* buf = static_cast<std::uint32_t*>(std::malloc(std::max<std::uint64_t>(acc, 1) * 4));
    if (!buf) return -1;
    std::uint64_t pos = 0;
    for (int t = 0; t < threads; ++t) {
      if (!hits[t].empty()) std::memcpy(buf + pos, hits[t].data(), hits[t].size() * 4);
      pos += hits[t].size();
    }
    *targets = buf;
  }
  return 0;
}

void ref_free_buf(void* p) { std::free(p); }

int ref_hardware_threads(void) {
  unsigned n = std::thread::hardware_concurrency();
  return n ? static_cast<int>(n) : 1;
}

}  // extern "C"
