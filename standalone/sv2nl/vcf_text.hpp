// Minimal text-VCF reader (plain or gzip/bgzip via zlib) for sv2nl: parse once, straight into integer
// columns (SURVEY.md section 8f.2 "parse-once columnar ingest").
//
// Stands in for the reference's htslib-backed VcfRanges (library/include/binary/parser/vcf.hpp), which
// cannot be built in this image; it reads exactly the fields sv2nl uses:
//   chrom = column 1, pos = POS - 1 (0-based, vcf.hpp:305-310, test_vcf.cpp:100)
//   INFO SVTYPE (required), CHR2 (TRA/BND), STRAND1/STRAND2 == "+" (INV, missing keeps the default true),
//   end = POS2 if SVTYPE == BND, else SVEND for source "nls", else END   (sv2nl vcf_info.cpp:9-43)
//   contigs = ##contig=<ID=...> lines in header order                    (vcf.hpp:577-589)
// Chromosome names and SVTYPE values are interned while parsing (a file has a few hundred distinct names at
// most): the table holds ids, the strings exist once in the dictionaries. The file is read in blocks,
// each block is parsed by several threads (memchr splitting, std::from_chars); nothing is allocated per record.
#pragma once

#include <zlib.h>

#include <algorithm>
#include <charconv>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

namespace binary {
class VcfReaderError : public std::runtime_error {  // reference: library/include/binary/exception.hpp:13
public:
  using std::runtime_error::runtime_error;
};
}  // namespace binary

namespace sv2nl {

constexpr std::uint32_t kNoChrom = 0xffffffffu;  // chr2 of a record that has none

struct VcfTable {
  std::vector<std::string> contigs;                   // ##contig IDs, header order
  std::vector<std::string> chrom_names, type_names;   // dictionaries, first-seen order
  std::vector<std::uint32_t> chrom, chr2;             // ids into chrom_names (chr2: kNoChrom if absent)
  std::vector<std::uint8_t> svtype;                   // ids into type_names
  std::vector<std::uint32_t> pos, svend;
  std::vector<std::uint8_t> strand1, strand2;         // 1 = "+"
  [[nodiscard]] std::size_t size() const { return pos.size(); }
  [[nodiscard]] const std::string& chrom_name(std::size_t i) const { return chrom_names[chrom[i]]; }
  [[nodiscard]] const std::string& type_name(std::size_t i) const { return type_names[svtype[i]]; }
};

namespace detail {
struct SvHash {
  using is_transparent = void;
  std::size_t operator()(std::string_view s) const { return std::hash<std::string_view>{}(s); }
};
struct SvEq {
  using is_transparent = void;
  bool operator()(std::string_view a, std::string_view b) const { return a == b; }
};
// string -> dense id, remembering the last hit (records of one chromosome / type come in runs)
class Interner {
public:
  explicit Interner(std::vector<std::string>& names) : names_(names) {}
  std::uint32_t id(std::string_view s) {
    if (last_ != kNoChrom && names_[last_] == s) return last_;
    auto it = map_.find(s);
    if (it == map_.end()) {
      it = map_.emplace(std::string(s), (std::uint32_t)names_.size()).first;
      names_.emplace_back(s);
    }
    return last_ = it->second;
  }

private:
  std::vector<std::string>& names_;
  std::unordered_map<std::string, std::uint32_t, SvHash, SvEq> map_;
  std::uint32_t last_ = kNoChrom;
};

// INFO values sv2nl reads, found in ONE pass over the field
struct InfoFields {
  std::string_view svtype, end, svend, pos2, chr2, strand1, strand2;
};
inline void scan_info(std::string_view info, InfoFields& f) {
  std::size_t p = 0;
  while (p < info.size()) {
    const char* semi = static_cast<const char*>(std::memchr(info.data() + p, ';', info.size() - p));
    const std::size_t e = semi ? (std::size_t)(semi - info.data()) : info.size();
    const std::string_view kv = info.substr(p, e - p);
    const char* eqp = static_cast<const char*>(std::memchr(kv.data(), '=', kv.size()));
    if (eqp) {
      const std::string_view key = kv.substr(0, (std::size_t)(eqp - kv.data()));
      const std::string_view val = kv.substr(key.size() + 1);
      // first occurrence wins, as in a left-to-right key search
      auto set = [&](std::string_view& slot) { if (!slot.data()) slot = val; };
      switch (key.size()) {
        case 3: if (key == "END") set(f.end); break;
        case 4: if (key == "POS2") set(f.pos2); else if (key == "CHR2") set(f.chr2); break;
        case 5: if (key == "SVEND") set(f.svend); break;
        case 6: if (key == "SVTYPE") set(f.svtype); break;
        case 7: if (key == "STRAND1") set(f.strand1); else if (key == "STRAND2") set(f.strand2); break;
        default: break;
      }
    }
    p = e + 1;
  }
}
inline bool parse_i64(std::string_view s, long long& out) {
  auto r = std::from_chars(s.data(), s.data() + s.size(), out);
  return r.ec == std::errc() && r.ptr == s.data() + s.size();
}
}  // namespace detail

namespace detail {
constexpr std::uint8_t kInherit = 2;
// Sequential pass over the finished table: an INV record without STRAND1 keeps both strands of the previous
// INV record (initially '+','+'), one with STRAND1 but without STRAND2 keeps the previous strand2; records of
// other types show the carried values too, as the reference's in-place record does.
inline void carry_strands(VcfTable& t) {
  std::uint32_t inv = 0xffffffffu;
  for (std::size_t k = 0; k < t.type_names.size(); ++k)
    if (t.type_names[k] == "INV") inv = (std::uint32_t)k;
  std::uint8_t c1 = 1, c2 = 1;
  for (std::size_t i = 0; i < t.size(); ++i) {
    if (t.svtype[i] == inv && t.strand1[i] != kInherit) {
      c1 = t.strand1[i];
      if (t.strand2[i] != kInherit) c2 = t.strand2[i];
    }
    t.strand1[i] = c1;
    t.strand2[i] = c2;
  }
}
// Parses whole lines of VCF text into `t` through the given interners. n_lines: lines seen. Returns an
// empty string, or the error, with bad_line = 1-based index of the offending line within `text`.
inline std::string parse_lines(std::string_view text, bool nls, VcfTable& t, Interner& chroms, Interner& types,
                               std::size_t& n_lines, std::size_t& bad_line) {
  std::size_t line_no = 0, start = 0;
  std::string error;
  auto handle_line = [&](std::string_view line) -> bool {
    ++line_no;
    auto fail = [&](std::string what) { error = std::move(what); bad_line = line_no; return false; };
    if (!line.empty() && line.back() == '\r') line.remove_suffix(1);
    if (line.empty()) return true;
    if (line[0] == '#') {
      if (line.rfind("##contig=<", 0) == 0) {
        const std::size_t p = line.find("ID=");
        if (p != std::string_view::npos) {
          const std::size_t e = line.find_first_of(",>", p);
          t.contigs.emplace_back(line.substr(p + 3, (e == std::string_view::npos ? line.size() : e) - p - 3));
        }
      }
      return true;
    }
    std::string_view cols[8];
    std::size_t n_cols = 0, p = 0;
    while (n_cols < 8) {
      const char* tab = static_cast<const char*>(std::memchr(line.data() + p, '\t', line.size() - p));
      const std::size_t e = tab ? (std::size_t)(tab - line.data()) : line.size();
      cols[n_cols++] = line.substr(p, e - p);
      if (!tab) break;
      p = e + 1;
    }
    if (n_cols < 8) return fail("fewer than 8 columns");
    InfoFields info;
    scan_info(cols[7], info);
    if (!info.svtype.data()) return fail("INFO/SVTYPE missing");
    const std::string_view type = info.svtype;
    const bool bnd = type == "BND";
    const std::string_view end = bnd ? info.pos2 : (nls ? info.svend : info.end);
    if (!end.data()) return fail(std::string("INFO/") + (bnd ? "POS2" : (nls ? "SVEND" : "END")) + " missing");
    std::uint32_t c2 = kNoChrom;
    if (bnd || type == "TRA") {
      if (!info.chr2.data()) return fail("INFO/CHR2 missing");
      c2 = chroms.id(info.chr2);
    }
    // 0 = '-', 1 = '+', kInherit = "whatever the previous INV record of the file left" (resolved by
    // carry_strands once all pieces are in file order): the reference updates one record object in place and
    // swallows the exception of a missing strand tag (vcf.hpp:305-310, vcf_info.cpp:17-31)
    std::uint8_t s1 = kInherit, s2 = kInherit;
    if (type == "INV" && info.strand1.data()) {
      s1 = info.strand1 == "+";
      if (info.strand2.data()) s2 = info.strand2 == "+";
    }
    long long pos = 0, svend = 0;
    if (!parse_i64(cols[1], pos)) return fail("POS is not a number");
    if (!parse_i64(end, svend)) return fail("end coordinate is not a number");
    const std::uint32_t type_id = types.id(type);
    if (type_id > 255) return fail("more than 256 distinct SVTYPE values");
    t.chrom.push_back(chroms.id(cols[0]));
    t.pos.push_back(static_cast<std::uint32_t>(pos - 1));
    t.svend.push_back(static_cast<std::uint32_t>(svend));
    t.svtype.push_back(static_cast<std::uint8_t>(type_id));
    t.chr2.push_back(c2);
    t.strand1.push_back(s1);
    t.strand2.push_back(s2);
    return true;
  };
  while (start < text.size()) {
    const char* nl = static_cast<const char*>(std::memchr(text.data() + start, '\n', text.size() - start));
    const std::size_t e = nl ? (std::size_t)(nl - text.data()) : text.size();
    if (!handle_line(text.substr(start, e - start))) break;
    start = e + 1;
  }
  n_lines = line_no;
  return error;
}

// appends `part` to `dst`, translating the part's dictionary ids into dst's (first-seen order is kept)
inline bool append_table(VcfTable& dst, Interner& chroms, Interner& types, const VcfTable& part) {
  std::vector<std::uint32_t> cmap(part.chrom_names.size()), tmap(part.type_names.size());
  for (std::size_t i = 0; i < cmap.size(); ++i) cmap[i] = chroms.id(part.chrom_names[i]);
  for (std::size_t i = 0; i < tmap.size(); ++i) tmap[i] = types.id(part.type_names[i]);
  if (dst.type_names.size() > 256) return false;
  dst.contigs.insert(dst.contigs.end(), part.contigs.begin(), part.contigs.end());
  const std::size_t at = dst.size(), n = part.size();
  dst.chrom.resize(at + n); dst.chr2.resize(at + n); dst.svtype.resize(at + n);
  for (std::size_t i = 0; i < n; ++i) {
    dst.chrom[at + i] = cmap[part.chrom[i]];
    dst.chr2[at + i] = part.chr2[i] == kNoChrom ? kNoChrom : cmap[part.chr2[i]];
    dst.svtype[at + i] = (std::uint8_t)tmap[part.svtype[i]];
  }
  dst.pos.insert(dst.pos.end(), part.pos.begin(), part.pos.end());
  dst.svend.insert(dst.svend.end(), part.svend.begin(), part.svend.end());
  dst.strand1.insert(dst.strand1.end(), part.strand1.begin(), part.strand1.end());
  dst.strand2.insert(dst.strand2.end(), part.strand2.begin(), part.strand2.end());
  return true;
}

inline unsigned parse_threads() {  // SV2NL_PARSE_THREADS overrides; default: up to 8 hardware threads
  if (const char* e = std::getenv("SV2NL_PARSE_THREADS")) {
    const long v = std::atol(e);
    if (v >= 1) return (unsigned)std::min<long>(v, 64);
  }
  const unsigned hw = std::thread::hardware_concurrency();
  return std::max(1u, std::min(hw ? hw : 1u, 8u));
}
}  // namespace detail

// The file is read in blocks of 4 MB per thread (a block's unfinished last line is carried into the next one); every
// block is cut at line ends into one piece per thread, the pieces are parsed concurrently into private
// tables and appended in order -- the result does not depend on the thread count.
inline VcfTable read_vcf(const std::string& path, std::string_view source) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) throw binary::VcfReaderError("cannot open " + path);
  gzbuffer(f, 1 << 20);
  VcfTable t;
  detail::Interner chroms(t.chrom_names), types(t.type_names);
  const bool nls = source == "nls";
  const unsigned n_threads = detail::parse_threads();
  std::size_t lines_before = 0;  // lines of the blocks already consumed (for error messages)
  auto fail = [&](std::size_t line_no, const std::string& what) {
    gzclose(f);
    throw binary::VcfReaderError(path + ":" + std::to_string(line_no) + ": " + what);
  };

  std::vector<char> buf((std::size_t)(4u << 20) * n_threads);  // ~4 MB of text per thread and block
  std::size_t have = 0;
  for (;;) {
    if (have == buf.size()) buf.resize(buf.size() * 2);  // one line longer than the block
    const int got = gzread(f, buf.data() + have, (unsigned)std::min<std::size_t>(buf.size() - have, 1u << 30));
    if (got < 0) fail(lines_before, "read error");
    have += (std::size_t)got;
    // whole lines available: everything up to the last newline (at end of file: everything)
    std::size_t whole = have;
    if (got != 0) {
      const void* last = memrchr(buf.data(), '\n', have);
      if (!last) continue;  // no complete line yet: read more
      whole = (std::size_t)(static_cast<const char*>(last) - buf.data()) + 1;
    }
    // cut [0, whole) into pieces at line ends
    std::vector<std::size_t> cut{0};
    for (unsigned k = 1; k < n_threads; ++k) {
      std::size_t at = whole / n_threads * k;
      if (at <= cut.back()) continue;
      const char* nl = static_cast<const char*>(std::memchr(buf.data() + at, '\n', whole - at));
      if (!nl) break;
      at = (std::size_t)(nl - buf.data()) + 1;
      if (at > cut.back() && at < whole) cut.push_back(at);
    }
    cut.push_back(whole);
    const std::size_t n_parts = cut.size() - 1;
    if (n_parts == 1) {  // one piece: straight into the result
      std::size_t n_lines = 0, bad = 0;
      const std::string err = detail::parse_lines(std::string_view(buf.data(), whole), nls, t, chroms, types, n_lines, bad);
      if (!err.empty()) fail(lines_before + bad, err);
      lines_before += n_lines;
    } else {
      std::vector<VcfTable> parts(n_parts);
      std::vector<std::string> errors(n_parts);
      std::vector<std::size_t> bad(n_parts, 0), n_lines(n_parts, 0);
      auto work = [&](std::size_t k) {
        detail::Interner c(parts[k].chrom_names), ty(parts[k].type_names);
        errors[k] = detail::parse_lines(std::string_view(buf.data() + cut[k], cut[k + 1] - cut[k]), nls, parts[k], c, ty,
                                        n_lines[k], bad[k]);
      };
      std::vector<std::thread> pool;
      for (std::size_t k = 1; k < n_parts; ++k) pool.emplace_back(work, k);
      work(0);
      for (auto& th : pool) th.join();
      for (std::size_t k = 0; k < n_parts; ++k) {
        if (!errors[k].empty()) fail(lines_before + bad[k], errors[k]);
        if (!detail::append_table(t, chroms, types, parts[k])) fail(lines_before, "more than 256 distinct SVTYPE values");
        lines_before += n_lines[k];
      }
    }
    if (got == 0) break;
    std::memmove(buf.data(), buf.data() + whole, have - whole);
    have -= whole;
  }
  gzclose(f);
  detail::carry_strands(t);
  return t;
}

}  // namespace sv2nl
