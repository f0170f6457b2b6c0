// Drop-in C++20 front end for binary::algorithm::tree::IntervalTree on top of libbinary_cuda.
//
// Mirrors the reference interface (paths relative to the reference checkout,
// library/include/binary/algorithm/):
//   BaseInterval<K>, UIntInterval, IntInterval          interval_tree.hpp:105-138
//   IntervalNode<Interval>                              interval_tree.hpp:51-103   (value holder only)
//   IntervalTree<Node>::insert_node(R&&)                rb_tree.hpp:111-117
//   IntervalTree<Node>::insert_node(Args&&...)          rb_tree.hpp:145-149
//   IntervalTree<Node>::insert_node(unique_ptr<Node>)   rb_tree.hpp:143,372 (takes ownership)
//   IntervalTree<Node>::find_overlaps(...)              interval_tree.hpp:161-168  -> vector of COPIES
//   IntervalTree<Node>::find_overlap(...)               interval_tree.hpp:152-159  -> optional copy
//   size(), empty()                                     rb_tree.hpp:126-129
// plus the NEW batched entry point the sv2nl loop (standalone/sv2nl/include/mapper.hpp:207-218) needs:
//   find_overlaps_batch(span<const interval_type>) -> {offsets, target ids}
// and use_devices({0, 1, ...}): batches split over several GPUs (bcu_join_multi; the counterpart of the reference's
// one-task-per-chromosome pool over a shared tree, mapper.hpp:238-246, mapper.cpp:136-140).
//
// What differs, on purpose:
//   * there is no pointer tree: intervals (with whatever payload the Interval subclass carries) stay in a
//     host vector in insertion order; only (low, high) go to the GPU; hits come back as insertion
//     ordinals and are turned into copies of the stored intervals. root()/to_dot()/inorder_walk()/
//     delete_node() -- not on the sv2nl path -- are not provided.
//   * the overlap predicate is the reference's BaseInterval::is_overlap, fixed: low <= o.high &&
//     o.low <= high (interval_tree.hpp:119-121). A subclass overriding the virtual is NOT consulted on
//     the device.
//   * find_overlaps returns the hits of a query sorted by (low, insertion ordinal) instead of tree-shape
//     preorder, and find_overlap returns the FIRST of that order (the reference returns whichever overlap its
//     root descent meets first, interval_tree.hpp:291-304). The batched entry point returns target ids in the
//     device's order, which is unspecified inside one query (it depends on the index's length classes).
//   * cost model: every single find_overlaps / find_overlap call is one host->device->host round trip
//     (tens of microseconds), and an insert after a query rebuilds the whole device index. Callers that loop
//     over records should collect them and call find_overlaps_batch once; after kUnbatchedHint single
//     queries the header says so once on stderr (set BINARY_CUDA_QUIET to silence it).
//   * errors: the reference tree never throws; this one throws binary::cuda_error when the CUDA
//     library reports a failure (there is no CPU fallback).
//   * lvalue interval arguments are accepted as well (the reference only compiles for rvalues/args).
// Thread safety as in the reference: concurrent const queries are safe (TraMapper shares one tree across
// threads, sv2nl mapper.cpp:136-140); inserts must not race with anything.
#ifndef BINARY_B200_INCLUDE_BINARY_ALGORITHM_INTERVAL_TREE_HPP_
#define BINARY_B200_INCLUDE_BINARY_ALGORITHM_INTERVAL_TREE_HPP_

#include <binary_cuda.h>

#include <algorithm>
#include <atomic>
#include <cassert>
#include <concepts>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <optional>
#include <ostream>
#include <ranges>
#include <span>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace binary {

class cuda_error : public std::runtime_error {
public:
  cuda_error(int status, const char* what_arg)
      : std::runtime_error(std::string("libbinary_cuda: ") + what_arg), status_(status) {}
  [[nodiscard]] int status() const noexcept { return status_; }

private:
  int status_;
};

namespace concepts {
  template <typename T, typename... U>
  concept IsAnyOf = (std::same_as<T, U> || ...);
  // reference: library/include/binary/concepts.hpp:16-18
  template <typename T, typename... Args>
  concept ArgsConstructible
      = std::constructible_from<T, Args...> && (!IsAnyOf<T, std::remove_cvref_t<Args>...>);
}  // namespace concepts

namespace algorithm::tree {

  template <typename T>
  concept KeyConcept = std::totally_ordered<T> && std::default_initializable<T>;

  // Closed interval [low, high]. The 2-arg constructor asserts low <= high in debug builds only; sv2nl's
  // TraMapper fills the fields directly and may leave them inverted, which the join honours as-is.
  template <KeyConcept KeyType = std::uint32_t> class BaseInterval {
  public:
    using key_type = std::remove_cv_t<KeyType>;

    constexpr BaseInterval() = default;
    constexpr BaseInterval(key_type lo, key_type hi) : low{lo}, high{hi} { assert(low <= high); }
    virtual ~BaseInterval() = default;

    /// Host-side twin of the device predicate (join.cu `overlaps`): both ends inclusive.
    [[nodiscard]] virtual bool is_overlap(BaseInterval const& o) const {
      return low <= o.high && o.low <= high;
    }

    friend std::ostream& operator<<(std::ostream& os, BaseInterval const& v) {
      return os << "BaseInterval: " << v.low << "-" << v.high;
    }

    key_type low{};
    key_type high{};
  };

  using IntInterval = BaseInterval<std::int32_t>;
  using UIntInterval = BaseInterval<std::uint32_t>;

  template <typename Interval>
  concept IntervalConcept = requires(Interval const& interval) {
    requires std::semiregular<Interval>;
    requires std::movable<Interval>;
    requires std::same_as<decltype(interval.low), decltype(interval.high)>;
    interval.low <= interval.high;
    typename Interval::key_type;
  };

  // Kept so that `IntervalTree<IntervalNode<I>>` spells the same as in the reference. It only carries
  // the interval (+ max/key as the reference constructors set them); there are no links or colours.
  template <IntervalConcept Interval> class IntervalNode {
  public:
    using interval_type = Interval;
    using key_type = typename Interval::key_type;
    using pointer = std::unique_ptr<IntervalNode>;

    constexpr IntervalNode() = default;
    IntervalNode(IntervalNode&&) noexcept = default;
    IntervalNode& operator=(IntervalNode&&) noexcept = default;

    template <typename... Arg>
      requires std::constructible_from<Interval, Arg...>
    explicit constexpr IntervalNode(Arg&&... args)
        : interval{std::forward<Arg>(args)...}, max{interval.high}, key{interval.low} {}
    explicit constexpr IntervalNode(Interval const& v) : interval{v}, max{v.high}, key{v.low} {}
    explicit constexpr IntervalNode(Interval&& v)
        : interval{std::move(v)}, max{interval.high}, key{interval.low} {}

    Interval interval{};
    key_type max{};
    key_type key{};
  };

  using IntIntervalNode = IntervalNode<IntInterval>;
  using UIntIntervalNode = IntervalNode<UIntInterval>;

  namespace detail {
    // order-preserving map of the key onto the u32 coordinate the device index sorts by
    template <typename K> constexpr std::uint32_t to_device_key(K k) noexcept {
      static_assert(std::is_integral_v<K> && sizeof(K) <= 4,
                    "the CUDA index handles integral keys of at most 32 bits");
      if constexpr (std::is_signed_v<K>) {
        return static_cast<std::uint32_t>(static_cast<std::int32_t>(k)) ^ 0x80000000u;
      } else {
        return static_cast<std::uint32_t>(k);
      }
    }
    inline void check(int status) {
      if (status != BCU_OK) throw cuda_error(status, bcu_last_error());
    }
    struct IndexDeleter {
      void operator()(bcu_index* p) const noexcept { bcu_index_free(p); }
    };
  }  // namespace detail

  /// CSR result of a batched query: hits of query i are target_ids[offsets[i] .. offsets[i+1]),
  /// each a 0-based insertion ordinal into the tree.
  struct BatchOverlaps {
    std::vector<std::uint64_t> offsets;
    std::vector<std::uint32_t> target_ids;
    [[nodiscard]] std::span<const std::uint32_t> hits(std::size_t query) const {
      return {target_ids.data() + offsets[query], target_ids.data() + offsets[query + 1]};
    }
  };

  template <typename NodeType> class IntervalTree {
  public:
    using interval_type = typename NodeType::interval_type;
    using key_type = typename NodeType::key_type;
    using pointer = std::unique_ptr<NodeType>;

    explicit IntervalTree(int device = 0) : device_{device} {}
    // moves leave the source empty but usable (it gets a mutex of its own again)
    IntervalTree(IntervalTree&& o) noexcept
        : device_{o.device_}, devices_{std::move(o.devices_)}, items_{std::move(o.items_)},
          index_{std::move(o.index_)}, replicas_{std::move(o.replicas_)}, mutex_{std::move(o.mutex_)} {
      o.items_.clear();
      o.replicas_.clear();
      o.mutex_ = std::make_unique<std::mutex>();
    }
    IntervalTree& operator=(IntervalTree&& o) noexcept {
      if (this != &o) {
        device_ = o.device_;
        devices_ = std::move(o.devices_);
        items_ = std::move(o.items_);
        index_ = std::move(o.index_);
        replicas_ = std::move(o.replicas_);
        o.items_.clear();
        o.index_.reset();
        o.replicas_.clear();
      }
      return *this;
    }

    /// Several GPUs for find_overlaps_batch: the index is replicated on every listed device and each batch is cut
    /// into one contiguous query range per device (bcu_join_multi; the reference's counterpart is one thread-pool
    /// task per chromosome over a shared tree, sv2nl mapper.hpp:238-246, mapper.cpp:136-140). The single-query
    /// calls keep using the first device. An empty list returns to one device.
    void use_devices(std::vector<int> devices) {
      std::lock_guard lock{*mutex_};
      devices_ = std::move(devices);
      if (!devices_.empty()) device_ = devices_.front();
      index_.reset();
      replicas_.clear();
    }
    IntervalTree(IntervalTree const&) = delete;
    IntervalTree& operator=(IntervalTree const&) = delete;
    virtual ~IntervalTree() = default;

    // ---- build (reference: rb_tree.hpp:111-117, 143-149) ------------------------------------------
    template <std::ranges::input_range R>
      requires std::constructible_from<NodeType, std::ranges::range_value_t<R>>
    void insert_node(R&& range) {
      for (auto&& item : range) insert_node(std::forward<decltype(item)>(item));
    }

    void insert_node(pointer node) {  // ownership transfer, as in the reference
      items_.push_back(std::move(node->interval));
      index_.reset();
      replicas_.clear();
    }

    template <typename... Args>
      requires std::constructible_from<NodeType, Args...>
    void insert_node(Args&&... args) {
      items_.push_back(std::move(NodeType(std::forward<Args>(args)...).interval));
      index_.reset();
      replicas_.clear();
    }

    [[nodiscard]] auto size() const -> std::size_t { return items_.size(); }
    [[nodiscard]] auto empty() const -> bool { return items_.empty(); }
    /// the stored interval behind a target id returned by find_overlaps_batch
    [[nodiscard]] auto at(std::uint32_t target_id) const -> interval_type const& { return items_[target_id]; }

    // ---- single queries (reference: interval_tree.hpp:152-168) -------------------------------------
    [[nodiscard]] auto find_overlaps(interval_type const& interval) const -> std::vector<interval_type> {
      const std::uint32_t ql = detail::to_device_key(interval.low), qh = detail::to_device_key(interval.high);
      note_single_query();
      auto res = batch(1, &ql, &qh);
      sort_hits(res.target_ids);
      std::vector<interval_type> out;
      out.reserve(res.target_ids.size());
      for (auto id : res.target_ids) out.push_back(items_[id]);
      return out;
    }

    template <typename... Args>
      requires concepts::ArgsConstructible<interval_type, Args...>
    [[nodiscard]] auto find_overlaps(Args&&... args) const -> std::vector<interval_type> {
      return find_overlaps(make_query(std::forward<Args>(args)...));
    }

    [[nodiscard]] auto find_overlap(interval_type const& interval) const -> std::optional<interval_type> {
      // deterministic whatever the index layout: the hit with the smallest (low, insertion ordinal)
      const std::uint32_t ql = detail::to_device_key(interval.low), qh = detail::to_device_key(interval.high);
      note_single_query();
      auto res = batch(1, &ql, &qh);
      if (res.target_ids.empty()) return {};
      const auto best = *std::min_element(res.target_ids.begin(), res.target_ids.end(),
                                          [&](std::uint32_t a, std::uint32_t b) { return hit_less(a, b); });
      return items_[best];
    }

    template <typename... Args>
      requires concepts::ArgsConstructible<interval_type, Args...>
    [[nodiscard]] auto find_overlap(Args&&... args) const -> std::optional<interval_type> {
      return find_overlap(make_query(std::forward<Args>(args)...));
    }

    // ---- NEW: the batched join ------------------------------------------------------------------------
    [[nodiscard]] auto find_overlaps_batch(std::span<const interval_type> queries) const -> BatchOverlaps {
      std::vector<std::uint32_t> ql(queries.size()), qh(queries.size());
      for (std::size_t i = 0; i < queries.size(); ++i) {
        ql[i] = detail::to_device_key(queries[i].low);
        qh[i] = detail::to_device_key(queries[i].high);
      }
      return batch(queries.size(), ql.data(), qh.data());
    }

    /// SoA variant: keys already split into low/high columns (no per-query objects on the host).
    [[nodiscard]] auto find_overlaps_batch(std::span<const key_type> low, std::span<const key_type> high) const
        -> BatchOverlaps {
      if (low.size() != high.size()) throw std::invalid_argument("low/high differ in length");
      if constexpr (std::is_same_v<key_type, std::uint32_t>) {
        return batch(low.size(), low.data(), high.data());
      } else {
        std::vector<std::uint32_t> ql(low.size()), qh(low.size());
        for (std::size_t i = 0; i < low.size(); ++i) {
          ql[i] = detail::to_device_key(low[i]);
          qh[i] = detail::to_device_key(high[i]);
        }
        return batch(low.size(), ql.data(), qh.data());
      }
    }

    static constexpr std::uint64_t kUnbatchedHint = 4096;

  private:
    [[nodiscard]] bool hit_less(std::uint32_t a, std::uint32_t b) const {
      return items_[a].low != items_[b].low ? items_[a].low < items_[b].low : a < b;
    }
    void sort_hits(std::vector<std::uint32_t>& ids) const {
      std::sort(ids.begin(), ids.end(), [&](std::uint32_t a, std::uint32_t b) { return hit_less(a, b); });
    }
    // the drop-in's cost model differs from the reference's (see the header comment): say so once
    static void note_single_query() {
      static std::atomic<std::uint64_t> calls{0};
      if (calls.fetch_add(1, std::memory_order_relaxed) + 1 == kUnbatchedHint && !std::getenv("BINARY_CUDA_QUIET"))
        std::fprintf(stderr,
                     "binary::IntervalTree (CUDA): %llu single find_overlap(s) calls so far, each a full "
                     "host<->device round trip; collect the queries and call find_overlaps_batch once\n",
                     static_cast<unsigned long long>(kUnbatchedHint));
    }
    template <typename... Args> static auto make_query(Args&&... args) -> interval_type {
      return interval_type{std::forward<Args>(args)...};
    }

    // (re)build the device index on first use after an insert; const like the reference's queries
    auto index() const -> bcu_index* {
      std::lock_guard lock{*mutex_};
      if (!index_) {
        std::vector<std::uint32_t> low(items_.size()), high(items_.size());
        for (std::size_t i = 0; i < items_.size(); ++i) {
          low[i] = detail::to_device_key(items_[i].low);
          high[i] = detail::to_device_key(items_[i].high);
        }
        bcu_index* raw = nullptr;
        detail::check(bcu_index_build(device_, items_.size(), nullptr, low.data(), high.data(), &raw));
        index_.reset(raw);
        for (std::size_t d = 1; d < devices_.size(); ++d) {  // replicas for bcu_join_multi
          detail::check(bcu_index_build(devices_[d], items_.size(), nullptr, low.data(), high.data(), &raw));
          replicas_.emplace_back(raw);
        }
      }
      return index_.get();
    }

    auto batch(std::size_t n, const std::uint32_t* ql, const std::uint32_t* qh) const -> BatchOverlaps {
      BatchOverlaps res;
      res.offsets.assign(n + 1, 0);
      if (n == 0 || items_.empty()) return res;
      bcu_index* ix = index();
      std::uint64_t total = 0;
      // first try with room for 4 hits per query, grow once to the exact size if that is not enough
      std::uint64_t capacity = 4 * static_cast<std::uint64_t>(n) + 1024;
      for (int attempt = 0; attempt < 2; ++attempt) {
        res.target_ids.resize(capacity);
        int rc;
        if (devices_.empty()) {
          rc = bcu_join(ix, n, nullptr, ql, qh, res.offsets.data(), capacity, /*hit_query=*/nullptr,
                        res.target_ids.data(), &total);
        } else {
          std::vector<const bcu_index*> all{ix};
          for (auto const& r : replicas_) all.push_back(r.get());
          rc = bcu_join_multi(all.data(), static_cast<int>(all.size()), n, nullptr, ql, qh, res.offsets.data(),
                              /*counts=*/nullptr, capacity, /*hit_query=*/nullptr, res.target_ids.data(), &total);
        }
        if (rc == BCU_E_CAPACITY) {
          capacity = total;
          continue;
        }
        detail::check(rc);
        res.target_ids.resize(total);
        return res;
      }
      throw cuda_error(BCU_E_CAPACITY, "pair capacity still too small after growing");
    }

    int device_{0};
    std::vector<int> devices_{};  // use_devices(): non-empty = batches go through bcu_join_multi
    std::vector<interval_type> items_{};
    mutable std::unique_ptr<bcu_index, detail::IndexDeleter> index_{};
    mutable std::vector<std::unique_ptr<bcu_index, detail::IndexDeleter>> replicas_{};
    mutable std::unique_ptr<std::mutex> mutex_{std::make_unique<std::mutex>()};
  };

}  // namespace algorithm::tree
}  // namespace binary

#endif  // BINARY_B200_INCLUDE_BINARY_ALGORITHM_INTERVAL_TREE_HPP_
