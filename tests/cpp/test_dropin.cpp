// The reference's own interval-tree unit tests (test/source/test_algorithm/test_interval_tree.cpp),
// re-expressed with plain checks (doctest is not in the image) against the drop-in front end.
// `./test_dropin`        : construction / insert / size only (no GPU needed)
// `./test_dropin --gpu`  : everything, queries run on the GPU through libbinary_cuda
#include <binary/algorithm/interval_tree.hpp>

#include <array>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>

using namespace binary::algorithm::tree;

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)

// an interval with a payload, the way sv2nl's BaseVcfInterval carries a record (parser/vcf.hpp:598-639)
struct NamedInterval : UIntInterval {
  std::string name;
  NamedInterval() = default;
  NamedInterval(std::uint32_t l, std::uint32_t h, std::string n) : name(std::move(n)) { low = l; high = h; }
};

static std::array<UIntInterval, 10> clrs() {
  // test_interval_tree.cpp:88-92
  return {UIntInterval(16u, 21u), UIntInterval(8u, 9u),   UIntInterval(5u, 8u),   UIntInterval(0u, 3u),
          UIntInterval(6u, 10u),  UIntInterval(15u, 23u), UIntInterval(25u, 30u), UIntInterval(17u, 19u),
          UIntInterval(19u, 20u), UIntInterval(26u, 26u)};
}

int main(int argc, char** argv) {
  const bool gpu = argc > 1 && std::strcmp(argv[1], "--gpu") == 0;

  {  // "test construct interval" / "test construct interval node" (:31-72)
    UIntInterval a{};
    CHECK(a.low == 0 && a.high == 0);
    UIntInterval b{1, 2};
    CHECK(b.low == 1 && b.high == 2);
    UIntIntervalNode n1{1u, 10u};
    CHECK(n1.interval.low == 1 && n1.interval.high == 10 && n1.max == 10);
    UIntInterval c{100u, 2000u};
    UIntIntervalNode n2{c};
    CHECK(n2.interval.low == 100 && n2.interval.high == 2000 && n2.max == 2000);
    IntervalNode<IntInterval> n3{1, 10};
    CHECK(n3.interval.low == 1 && n3.max == 10);
  }
  {  // "test construct interval tree" (:74-85) and "test insert multiple nodes" (:87-99)
    IntervalTree<UIntIntervalNode> t{};
    CHECK(t.empty());
    t.insert_node(16u, 21u);
    CHECK(t.size() == 1 && !t.empty());
    IntervalTree<IntIntervalNode> ti{};
    ti.insert_node(16, 21);
    CHECK(ti.size() == 1);
    IntervalTree<UIntIntervalNode> many{};
    auto nodes = clrs();
    many.insert_node(nodes);
    CHECK(many.size() == nodes.size());
    many.insert_node(std::make_unique<UIntIntervalNode>(40u, 41u));  // ownership overload
    CHECK(many.size() == nodes.size() + 1);
  }
  {  // "test insert nodes fuzzy test" (:101-109)
    IntervalTree<IntIntervalNode> t{};
    for (int i = 0; i < 1000; i += 2) t.insert_node(i, i + 3);
    CHECK(t.size() == 500);
  }
  if (!gpu) {
    std::printf("%s (host-only part)\n", failures ? "FAILED" : "OK");
    return failures ? 1 : 0;
  }

  {  // "test find overlap" (:111-144)
    IntervalTree<UIntIntervalNode> t{};
    auto nodes = clrs();
    t.insert_node(nodes);
    auto one = t.find_overlap(22u, 25u);
    CHECK(one.has_value());
    // the reference returns [15,23] here; with index order the first hit is also [15,23] (low 15 < 25)
    CHECK(one && one->low == 15u && one->high == 23u);
    auto none = t.find_overlap(UIntInterval{100u, 111u});
    CHECK(!none.has_value());
    auto r1 = t.find_overlaps(UIntInterval{7u, 25u});
    CHECK(r1.size() == 8);
    auto r2 = t.find_overlaps(15u, 25u);
    CHECK(r2.size() == 5);
    std::set<std::pair<unsigned, unsigned>> got;
    for (auto const& v : r2) got.insert({v.low, v.high});
    CHECK((got == std::set<std::pair<unsigned, unsigned>>{{16, 21}, {15, 23}, {19, 20}, {17, 19}, {25, 30}}));
    UIntInterval lvalue{15u, 25u};
    CHECK(t.find_overlaps(lvalue).size() == 5);  // lvalues are accepted too
  }
  {  // "test same interval value" (:146-155)
    IntervalTree<UIntIntervalNode> t{};
    for (int i = 0; i < 4; ++i) t.insert_node(1u, 4u);
    CHECK(t.size() == 4);
    CHECK(t.find_overlaps(2u, 5u).size() == 4);
  }
  {  // signed keys keep their order on the device
    IntervalTree<IntIntervalNode> t{};
    t.insert_node(-10, -5);
    t.insert_node(-3, 4);
    t.insert_node(6, 9);
    CHECK(t.find_overlaps(-4, 5).size() == 1);
    CHECK(t.find_overlaps(-100, 100).size() == 3);
    CHECK(t.find_overlaps(-5, -5).size() == 1);
  }
  {  // payload-carrying intervals come back as copies, like sv2nl's records
    IntervalTree<IntervalNode<NamedInterval>> t{};
    t.insert_node(NamedInterval{10, 20, "a"});
    t.insert_node(NamedInterval{15, 30, "b"});
    t.insert_node(NamedInterval{40, 50, "c"});
    auto hits = t.find_overlaps(NamedInterval{18, 19, "q"});
    CHECK(hits.size() == 2);
    std::set<std::string> names;
    for (auto const& h : hits) names.insert(h.name);
    CHECK((names == std::set<std::string>{"a", "b"}));
  }
  {  // the new batched entry point
    IntervalTree<UIntIntervalNode> t{};
    auto nodes = clrs();
    t.insert_node(nodes);
    std::vector<UIntInterval> q{{7u, 25u}, {15u, 25u}, {100u, 111u}, {26u, 26u}};
    auto res = t.find_overlaps_batch(q);
    CHECK(res.offsets.size() == 5);
    CHECK(res.offsets[1] - res.offsets[0] == 8 && res.offsets[2] - res.offsets[1] == 5);
    CHECK(res.offsets[3] - res.offsets[2] == 0 && res.offsets[4] - res.offsets[3] == 2);
    for (auto id : res.hits(3)) CHECK(t.at(id).low <= 26u && 26u <= t.at(id).high);
    // a batch big enough to need the capacity retry: every query hits all 10 intervals
    std::vector<std::uint32_t> lo(3000, 0u), hi(3000, 100u);
    auto big = t.find_overlaps_batch(std::span<const std::uint32_t>(lo), std::span<const std::uint32_t>(hi));
    CHECK(big.target_ids.size() == 30000);
    // inserting after a query rebuilds the index
    t.insert_node(200u, 300u);
    CHECK(t.find_overlaps(250u, 250u).size() == 1);
  }
  {  // use_devices(): batches through bcu_join_multi (one device here when the box has one, every GPU otherwise)
    int n_dev = 0;
    CHECK(bcu_device_count(&n_dev) == BCU_OK && n_dev >= 1);
    IntervalTree<IntIntervalNode> one{}, many{};
    for (int i = 0; i < 20000; ++i) {
      const int lo = (i * 7919) % 100003 - 50000, len = (i * 31) % 400;
      one.insert_node(lo, lo + len);
      many.insert_node(lo, lo + len);
    }
    std::vector<int> lo(50000), hi(50000);
    for (int i = 0; i < 50000; ++i) { lo[i] = (int)(((long long)i * 104729) % 100003) - 50000; hi[i] = lo[i] + (i % 300); }
    std::vector<int> devices;
    for (int d = 0; d < n_dev; ++d) devices.push_back(d);
    many.use_devices(devices);
    auto a = one.find_overlaps_batch(std::span<const int>(lo), std::span<const int>(hi));
    auto b = many.find_overlaps_batch(std::span<const int>(lo), std::span<const int>(hi));
    CHECK(a.offsets == b.offsets && a.target_ids.size() == b.target_ids.size() && !a.target_ids.empty());
    bool same = true;
    for (std::size_t q = 0; q < lo.size() && same; ++q) {
      std::multiset<std::uint32_t> x(a.hits(q).begin(), a.hits(q).end()), y(b.hits(q).begin(), b.hits(q).end());
      same = x == y;
    }
    CHECK(same);
    many.insert_node(0, 1);  // replicas are rebuilt after an insert
    many.use_devices({0});
    CHECK(many.find_overlaps_batch(std::span<const int>(lo), std::span<const int>(hi)).target_ids.size() > a.target_ids.size());
    CHECK(many.find_overlaps(0, 0).size() == one.find_overlaps(0, 0).size() + 1);
  }
  {  // empty tree
    IntervalTree<UIntIntervalNode> t{};
    CHECK(t.find_overlaps(1u, 2u).empty());
    CHECK(!t.find_overlap(1u, 2u).has_value());
  }
  std::printf("%s\n", failures ? "FAILED" : "OK");
  return failures ? 1 : 0;
}
