// sv2nl's three mappers on top of the batched GPU join (libbinary_cuda, include/binary_cuda.h).
//
// Reference being restructured (standalone/sv2nl in the reference checkout):
//   Mapper::map_delegate / map_impl / build_tree   include/mapper.hpp:147-162, 194-246  (one tree per
//                                                  chromosome, one find_overlaps per NL record)
//   Dup/Inv/TraMapper::check_condition             source/mapper.cpp:50-79, 144-156
//   TraMapper: one tree over all BND records       source/mapper.cpp:86-170
//   validate_record, get_2chroms_with_pos,
//   format_map_key                                 include/helper.hpp:52-91
//   is_contained, distance_less                    include/helper.hpp:16-41 (on the device: csrc/join.cu accept<>)
//   SV2NL_USE_CACHE duplicate-key rule             include/mapper.hpp:212-234, options.hpp:8
// Here every mapper issues ONE call, bcu_sv2nl_join, over all chromosomes (chromosome = group): the join, the
// reference's post-filters and its duplicate-key rule run on the device (csrc/join.cu accept<>, csrc/sv2nl_rules.cu);
// the host formats the pairs that come back. Both VCFs are parsed once. Output lines equal the reference's as a
// multiset (its line order is thread-dependent).
#pragma once

#include <binary_cuda.h>

#include <algorithm>
#include <array>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "vcf_text.hpp"

namespace sv2nl {

constexpr const char* HEADER = "chrom\tpos\tend\tsvtype\tchrom\tpos\tend\tsvtype";  // mapper.hpp:30

// Records are handled as integers (chromosome ids, type codes); strings are only produced for the lines
// that are actually written. Chromosome ids are ranks in lexicographic name order, so the reference's
// string comparison `record.chrom > chr2` (helper.hpp:76-82) is an integer comparison here.
struct Rec {  // the fields of Sv2nlVcfRecord the mapping reads
  std::uint32_t chrom{}, chr2{};
  std::uint32_t pos{}, svend{};
  std::uint8_t type{};  // index into Names::types
  bool two_chrom{};     // SVTYPE is TRA or BND
  bool strand1{true}, strand2{true};
};

struct Names {  // id <-> string tables shared by both files
  std::vector<std::string> chroms;  // sorted: id order == lexicographic order
  std::vector<std::string> types;
  std::unordered_map<std::string, std::uint32_t> chrom_id, type_id;
  // per-file dictionary id (vcf_text.hpp interns while parsing) -> shared id; index 0 = NL file, 1 = SV file
  std::array<std::vector<std::uint32_t>, 2> chrom_of, type_of;
  std::uint32_t no_chrom = 0;  // what a record without CHR2 gets: the smallest id, never used
  void build(const VcfTable& nl, const VcfTable& sv) {
    const VcfTable* tables[2] = {&nl, &sv};
    chrom_id.emplace("", 0);  // the reference's default-constructed chr2 (empty string sorts first)
    for (const VcfTable* t : tables) {
      for (auto const& c : t->contigs) chrom_id.emplace(c, 0);
      for (auto const& c : t->chrom_names) chrom_id.emplace(c, 0);
      for (auto const& c : t->type_names) type_id.emplace(c, 0);
    }
    for (auto const& kv : chrom_id) chroms.push_back(kv.first);
    std::sort(chroms.begin(), chroms.end());
    for (std::uint32_t i = 0; i < chroms.size(); ++i) chrom_id[chroms[i]] = i;
    for (auto& kv : type_id) { kv.second = (std::uint32_t)types.size(); types.push_back(kv.first); }
    no_chrom = chrom_id.at("");
    for (int f = 0; f < 2; ++f) {
      for (auto const& c : tables[f]->chrom_names) chrom_of[f].push_back(chrom_id.at(c));
      for (auto const& c : tables[f]->type_names) type_of[f].push_back(type_id.at(c));
    }
  }
  [[nodiscard]] std::uint32_t type_or_none(const std::string& name) const {  // shared id of an SVTYPE, or ~0
    auto it = type_id.find(name);
    return it == type_id.end() ? 0xffffffffu : it->second;
  }
};

// file: 0 = the NL table, 1 = the SV table (selects the id translation)
inline Rec record_at(const VcfTable& t, const Names& names, int file, std::size_t i, std::uint32_t tra_type,
                     std::uint32_t bnd_type) {
  Rec r;
  r.chrom = names.chrom_of[file][t.chrom[i]];
  r.chr2 = t.chr2[i] == kNoChrom ? names.no_chrom : names.chrom_of[file][t.chr2[i]];
  r.pos = t.pos[i];
  r.svend = t.svend[i];
  const std::uint32_t type = names.type_of[file][t.svtype[i]];
  r.type = (std::uint8_t)type;
  r.two_chrom = type == tra_type || type == bnd_type;
  r.strand1 = t.strand1[i] != 0;
  r.strand2 = t.strand2[i] != 0;
  return r;
}

inline Rec validate_record(Rec r) {  // helper.hpp:52-63
  if (r.pos > r.svend) {
    std::swap(r.pos, r.svend);
    if (r.two_chrom) std::swap(r.chrom, r.chr2);
  }
  return r;
}
struct Breakpoints { std::uint32_t c1, c2, p1, p2; };
inline Breakpoints ordered_breakpoints(const Rec& r) {  // get_2chroms_with_pos, helper.hpp:76-82
  return r.chrom > r.chr2 ? Breakpoints{r.chr2, r.chrom, r.svend, r.pos} : Breakpoints{r.chrom, r.chr2, r.pos, r.svend};
}
// format_map_key (helper.hpp:84-91) as integers: two records get the same key iff the reference's
// formatted strings are equal
struct MapKey {
  std::uint32_t a, b, c, d;
  bool operator==(const MapKey& o) const { return a == o.a && b == o.b && c == o.c && d == o.d; }
};
inline MapKey map_key(const Rec& r) {
  if (r.two_chrom) {
    auto b = ordered_breakpoints(r);
    return {b.c1, b.c2, b.p1, b.p2};
  }
  return {r.chrom, 0xffffffffu, r.pos, r.svend};
}
// writer.cpp:21-27 (note pos + 1), appended to `out` without temporaries
inline void append_keys(std::string& out, const Rec& r, const Names& names) {
  out += names.chroms[r.chrom];
  if (r.two_chrom) { out += ','; out += names.chroms[r.chr2]; }
  char num[24];
  out += '\t';
  out.append(num, std::to_chars(num, num + sizeof(num), (std::uint64_t)r.pos + 1).ptr);
  out += '\t';
  out.append(num, std::to_chars(num, num + sizeof(num), r.svend).ptr);
  out += '\t';
  out += names.types[r.type];
}

struct JoinResult {
  std::vector<std::uint64_t> offsets;
  std::vector<std::uint32_t> targets;
};

// one mapper through the C ABI (join + post-filter + duplicate-key rule on the device); throws
// binary::VcfReaderError with the library's message on failure. Returns the per-RECORD CSR of the written pairs.
inline JoinResult gpu_join(int device, const std::vector<std::uint32_t>& tg, const std::vector<std::uint32_t>& tl,
                           const std::vector<std::uint32_t>& th, const std::vector<std::uint32_t>& qg,
                           const std::vector<std::uint32_t>& ql, const std::vector<std::uint32_t>& qh,
                           const bcu_sv2nl_rules& rules, const bcu_filter* filter = nullptr,
                           const std::vector<std::uint8_t>* qstrand = nullptr) {
  JoinResult r;
  const std::size_t n_rec = ql.size() / rules.probes_per_record;
  r.offsets.assign(n_rec + 1, 0);
  if (tl.empty() || ql.empty()) return r;
  auto check = [](int rc) {
    if (rc != BCU_OK) throw binary::VcfReaderError(std::string("libbinary_cuda: ") + bcu_last_error());
  };
  bcu_index* ix = nullptr;
  check(bcu_index_build(device, tl.size(), tg.data(), tl.data(), th.data(), &ix));
  std::uint64_t total = 0, cap = 4 * n_rec + 1024;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r.targets.resize(cap);
    int rc = bcu_sv2nl_join(ix, filter, &rules, n_rec, qg.data(), ql.data(), qh.data(),
                            qstrand ? qstrand->data() : nullptr, r.offsets.data(), cap, r.targets.data(), &total);
    if (rc == BCU_E_CAPACITY) { cap = total; continue; }
    if (rc != BCU_OK) { bcu_index_free(ix); check(rc); }
    break;
  }
  bcu_index_free(ix);
  r.targets.resize(total);
  return r;
}

struct Options {
  std::uint32_t diff = 1000000;  // --dis default, main.cpp:91
  bool use_strand = true;
  int device = 0;
  bool debug = false;  // per-mapper phase times on stderr
};

struct Lines {  // newline-terminated data lines of one output file, in one buffer
  std::string text;
  std::size_t count = 0;
  [[nodiscard]] std::size_t size() const { return count; }
};
struct Sv2nlOutput { Lines dup, inv, tra; };

namespace detail {
// the tail of the three map_impl loops: formatting (the pairs are the ones to write)
inline Lines emit_lines(const Names& names, const std::vector<Rec>& nl_orig, const std::vector<Rec>& sv_recs,
                        const JoinResult& jr) {
  Lines lines;
  std::string left;
  for (std::size_t q = 0; q < nl_orig.size(); ++q) {
    if (jr.offsets[q] == jr.offsets[q + 1]) continue;
    left.clear();
    append_keys(left, nl_orig[q], names);
    for (std::uint64_t k = jr.offsets[q]; k < jr.offsets[q + 1]; ++k) {
      lines.text += left;
      lines.text += '\t';
      append_keys(lines.text, sv_recs[jr.targets[k]], names);
      lines.text += '\n';
      ++lines.count;
    }
  }
  return lines;
}
inline std::vector<std::uint32_t> key_words(const std::vector<Rec>& nl_orig) {  // format_map_key, four words each
  std::vector<std::uint32_t> w;
  w.reserve(4 * nl_orig.size());
  for (auto const& r : nl_orig) {
    const MapKey k = map_key(r);
    w.insert(w.end(), {k.a, k.b, k.c, k.d});
  }
  return w;
}
}  // namespace detail

inline Sv2nlOutput map_sv2nl(const VcfTable& nl, const VcfTable& sv, const Options& opt) {
  Sv2nlOutput out;
  Names names;
  auto clock = [] { return std::chrono::steady_clock::now(); };
  auto t_last = clock();
  auto lap = [&](const char* what) {
    if (!opt.debug) return;
    const auto t = clock();
    std::fprintf(stderr, "  [map] %-22s %.3f s\n", what, std::chrono::duration<double>(t - t_last).count());
    t_last = t;
  };
  names.build(nl, sv);
  lap("name tables");
  std::vector<std::uint8_t> is_main(names.chroms.size(), 0);  // header contigs without '_' (mapper.hpp:239-244)
  for (auto const& c : nl.contigs)
    if (c.find('_') == std::string::npos) is_main[names.chrom_id.at(c)] = 1;
  const std::uint32_t tra_type = names.type_or_none("TRA"), bnd_type = names.type_or_none("BND");
  auto nl_rec = [&](std::size_t i) { return record_at(nl, names, 0, i, tra_type, bnd_type); };
  auto sv_rec = [&](std::size_t i) { return record_at(sv, names, 1, i, tra_type, bnd_type); };
  auto nl_is = [&](std::size_t i, std::uint32_t type) {  // NL record i has this type and sits on a main contig
    return names.type_of[0][nl.svtype[i]] == type && is_main[names.chrom_of[0][nl.chrom[i]]];
  };

  // ---- DupMapper / InvMapper: overlap join per chromosome ------------------------------------------
  struct Kind { const char* nl_type; const char* sv_type; bool inv; };
  for (Kind kind : {Kind{"TDUP", "DUP", false}, Kind{"INV", "INV", true}}) {
    std::vector<Rec> sv_recs, nl_orig, nl_valid;
    std::vector<std::uint32_t> tg, tl, th, qg, ql, qh;
    std::vector<std::uint8_t> qstrand;
    for (auto* v : {&tg, &tl, &th}) v->reserve(sv.size() / 2);
    for (auto* v : {&qg, &ql, &qh}) v->reserve(nl.size() / 2);
    sv_recs.reserve(sv.size() / 2); nl_orig.reserve(nl.size() / 2); nl_valid.reserve(nl.size() / 2);
    qstrand.reserve(nl.size() / 2);
    const std::uint32_t sv_type = names.type_or_none(kind.sv_type), nl_type = names.type_or_none(kind.nl_type);
    for (std::size_t i = 0; i < sv.size(); ++i)
      if (names.type_of[1][sv.svtype[i]] == sv_type) {
        sv_recs.push_back(validate_record(sv_rec(i)));  // build_tree validates (mapper.hpp:151)
        tg.push_back(sv_recs.back().chrom); tl.push_back(sv_recs.back().pos); th.push_back(sv_recs.back().svend);
      }
    for (std::size_t i = 0; i < nl.size(); ++i)
      if (nl_is(i, nl_type)) {
        nl_orig.push_back(nl_rec(i));
        nl_valid.push_back(validate_record(nl_orig.back()));
        qg.push_back(nl_valid.back().chrom); ql.push_back(nl_valid.back().pos); qh.push_back(nl_valid.back().svend);
        qstrand.push_back((std::uint8_t)((nl_valid.back().strand1 ? 1 : 0) | (nl_valid.back().strand2 ? 2 : 0)));
      }
    // check_condition (mapper.cpp:50-79) is fused into the join kernels, the duplicate-key rule follows on the device
    const bcu_filter filter{kind.inv ? (std::uint32_t)BCU_FILTER_SV2NL_INV : (std::uint32_t)BCU_FILTER_SV2NL_DUP,
                            opt.diff, opt.use_strand ? 1u : 0u, 0u};
    const std::vector<std::uint32_t> keys = detail::key_words(nl_orig);
    bcu_sv2nl_rules rules{};
    rules.probes_per_record = 1;
    rules.diff = opt.diff;
    rules.dedup = 1;
    rules.rec_key = keys.data();
    lap(kind.inv ? "inv: select records" : "dup: select records");
    JoinResult jr = gpu_join(opt.device, tg, tl, th, qg, ql, qh, rules, &filter, &qstrand);
    lap(kind.inv ? "inv: index + join + rules" : "dup: index + join + rules");
    (kind.inv ? out.inv : out.dup) = detail::emit_lines(names, nl_orig, sv_recs, jr);
    lap(kind.inv ? "inv: format" : "dup: format");
  }

  // ---- TraMapper ----------------------------------------------------------------------------------
  // The reference joins on the raw [pos, POS2] intervals of ALL BND records (not validated, chromosome
  // not part of the key) and filters afterwards (mapper.cpp:144-156). Same result with far fewer pairs:
  // join on the selective conditions -- group = (ordered chromosome pair, bucket of the SECOND breakpoint,
  // buckets `diff` wide), target = the point p1, query = [p1 - diff, p1 + diff] in each of the three buckets
  // a partner's p2 can fall into -- and apply the exact rules (both breakpoints within diff, and the
  // reference's raw-interval overlap, which can still reject a pair) to the pairs on the device. The three
  // probes of a record are consecutive queries, so its hits are one contiguous CSR range.
  {
    std::vector<Rec> sv_recs, nl_orig, nl_valid;
    std::vector<std::uint32_t> tg, tl, th, qg, ql, qh;
    const std::uint32_t width = std::max<std::uint32_t>(opt.diff, 1u);
    const std::uint32_t n_buckets = 0xffffffffu / width + 1u;
    const std::uint64_t n_chrom = names.chroms.size();
    std::vector<std::uint32_t> pair_index(n_chrom * n_chrom, 0xffffffffu);  // dense ids of the pairs in use
    std::uint32_t n_pairs = 0;
    bool overflow = false;
    auto group_of = [&](const Breakpoints& b, std::uint32_t bucket) -> std::uint32_t {
      std::uint32_t& pi = pair_index[(std::uint64_t)b.c1 * n_chrom + b.c2];
      if (pi == 0xffffffffu) pi = n_pairs++;
      const std::uint64_t g = (std::uint64_t)pi * n_buckets + bucket;
      if (g >= 0xffffffffull) overflow = true;  // (pairs x buckets) beyond 32 bits: see the fallback below
      return (std::uint32_t)g;
    };
    constexpr std::uint32_t kNoGroup = 0xffffffffu;  // no target carries it: such a probe finds nothing
    for (std::size_t i = 0; i < sv.size(); ++i)
      if (names.type_of[1][sv.svtype[i]] == bnd_type) {
        sv_recs.push_back(sv_rec(i));  // NOT validated (mapper.cpp:158-170)
        auto b = ordered_breakpoints(sv_recs.back());
        tg.push_back(group_of(b, b.p2 / width)); tl.push_back(b.p1); th.push_back(b.p1);
      }
    for (std::size_t i = 0; i < nl.size(); ++i)
      if (nl_is(i, tra_type)) {
        nl_orig.push_back(nl_rec(i));
        nl_valid.push_back(validate_record(nl_orig.back()));
        auto b = ordered_breakpoints(nl_valid.back());
        const std::uint32_t lo = b.p1 > opt.diff ? b.p1 - opt.diff : 0u;
        const std::uint32_t hi = b.p1 <= 0xffffffffu - opt.diff ? b.p1 + opt.diff : 0xffffffffu;
        const std::uint32_t mid = b.p2 / width;
        for (int k = -1; k <= 1; ++k) {
          const bool exists = !(k < 0 && mid == 0) && !(k > 0 && mid + 1 >= n_buckets);
          qg.push_back(exists ? group_of(b, (std::uint32_t)((std::int64_t)mid + k)) : kNoGroup);
          ql.push_back(lo);
          qh.push_back(hi);
        }
      }
    if (overflow) {  // absurdly many chromosome pairs for a tiny diff: fall back to one probe per record
      for (std::size_t t = 0; t < sv_recs.size(); ++t) {
        auto b = ordered_breakpoints(sv_recs[t]);
        tg[t] = pair_index[(std::uint64_t)b.c1 * n_chrom + b.c2];
      }
      for (std::size_t q = 0; q < nl_valid.size(); ++q) {
        auto b = ordered_breakpoints(nl_valid[q]);
        qg[3 * q] = pair_index[(std::uint64_t)b.c1 * n_chrom + b.c2];
        qg[3 * q + 1] = qg[3 * q + 2] = kNoGroup;
      }
    }
    // TraMapper::check_condition (mapper.cpp:144-156), the raw-interval overlap of the reference's tree and the
    // duplicate-key rule: on the device, over the three probes of each record (bcu_sv2nl_rules)
    std::vector<std::uint32_t> rec_p1, rec_p2, tgt_p1, tgt_p2, tgt_pos, tgt_end;
    for (auto const& n : nl_valid) { auto b = ordered_breakpoints(n); rec_p1.push_back(b.p1); rec_p2.push_back(b.p2); }
    for (auto const& t : sv_recs) {
      auto b = ordered_breakpoints(t);
      tgt_p1.push_back(b.p1); tgt_p2.push_back(b.p2); tgt_pos.push_back(t.pos); tgt_end.push_back(t.svend);
    }
    const std::vector<std::uint32_t> keys = detail::key_words(nl_orig);
    bcu_sv2nl_rules rules{};
    rules.probes_per_record = 3;
    rules.tra = 1;
    rules.diff = opt.diff;
    rules.dedup = 1;
    rules.rec_p1 = rec_p1.data(); rules.rec_p2 = rec_p2.data();
    rules.tgt_p1 = tgt_p1.data(); rules.tgt_p2 = tgt_p2.data();
    rules.tgt_pos = tgt_pos.data(); rules.tgt_end = tgt_end.data();
    rules.rec_key = keys.data();
    lap("tra: select records");
    JoinResult jr = gpu_join(opt.device, tg, tl, th, qg, ql, qh, rules);
    lap("tra: index + join + rules");
    out.tra = detail::emit_lines(names, nl_orig, sv_recs, jr);
    lap("tra: format");
  }
  return out;
}

}  // namespace sv2nl
