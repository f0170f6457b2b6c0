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
// most): the table holds ids, the strings exist once in the dictionaries. The file is read in 4 MB blocks
// and split with memchr; numbers go through std::from_chars. Nothing is allocated per record.
#pragma once

#include <zlib.h>

#include <charconv>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <string_view>
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

inline VcfTable read_vcf(const std::string& path, std::string_view source) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) throw binary::VcfReaderError("cannot open " + path);
  gzbuffer(f, 1 << 20);
  VcfTable t;
  detail::Interner chroms(t.chrom_names), types(t.type_names);
  const bool nls = source == "nls";
  std::size_t line_no = 0;
  auto fail = [&](const std::string& what) {
    gzclose(f);
    throw binary::VcfReaderError(path + ":" + std::to_string(line_no) + ": " + what);
  };

  auto handle_line = [&](std::string_view line) {
    ++line_no;
    if (!line.empty() && line.back() == '\r') line.remove_suffix(1);
    if (line.empty()) return;
    if (line[0] == '#') {
      if (line.rfind("##contig=<", 0) == 0) {
        const std::size_t p = line.find("ID=");
        if (p != std::string_view::npos) {
          const std::size_t e = line.find_first_of(",>", p);
          t.contigs.emplace_back(line.substr(p + 3, (e == std::string_view::npos ? line.size() : e) - p - 3));
        }
      }
      return;
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
    if (n_cols < 8) fail("fewer than 8 columns");
    detail::InfoFields info;
    detail::scan_info(cols[7], info);
    if (!info.svtype.data()) fail("INFO/SVTYPE missing");
    const std::string_view type = info.svtype;
    const bool bnd = type == "BND";
    const std::string_view end = bnd ? info.pos2 : (nls ? info.svend : info.end);
    if (!end.data()) fail(std::string("INFO/") + (bnd ? "POS2" : (nls ? "SVEND" : "END")) + " missing");
    std::uint32_t c2 = kNoChrom;
    if (bnd || type == "TRA") {
      if (!info.chr2.data()) fail("INFO/CHR2 missing");
      c2 = chroms.id(info.chr2);
    }
    std::uint8_t s1 = 1, s2 = 1;
    if (type == "INV" && info.strand1.data()) {  // a missing STRAND1 keeps both defaults (vcf_info.cpp:17-31)
      s1 = info.strand1 == "+";
      if (info.strand2.data()) s2 = info.strand2 == "+";
    }
    long long pos = 0, svend = 0;
    if (!detail::parse_i64(cols[1], pos)) fail("POS is not a number");
    if (!detail::parse_i64(end, svend)) fail("end coordinate is not a number");
    const std::uint32_t type_id = types.id(type);
    if (type_id > 255) fail("more than 256 distinct SVTYPE values");
    t.chrom.push_back(chroms.id(cols[0]));
    t.pos.push_back(static_cast<std::uint32_t>(pos - 1));
    t.svend.push_back(static_cast<std::uint32_t>(svend));
    t.svtype.push_back(static_cast<std::uint8_t>(type_id));
    t.chr2.push_back(c2);
    t.strand1.push_back(s1);
    t.strand2.push_back(s2);
  };

  // 4 MB blocks; the unfinished tail of a block is carried to the front of the next one
  std::vector<char> buf(4u << 20);
  std::size_t have = 0;
  for (;;) {
    if (have == buf.size()) buf.resize(buf.size() * 2);  // one line longer than the block
    const int got = gzread(f, buf.data() + have, (unsigned)(buf.size() - have));
    if (got < 0) fail("read error");
    have += (std::size_t)got;
    std::size_t start = 0;
    for (;;) {
      const char* nl = static_cast<const char*>(std::memchr(buf.data() + start, '\n', have - start));
      if (!nl) break;
      const std::size_t e = (std::size_t)(nl - buf.data());
      handle_line(std::string_view(buf.data() + start, e - start));
      start = e + 1;
    }
    if (got == 0) {  // end of file: a last line without newline
      if (start < have) handle_line(std::string_view(buf.data() + start, have - start));
      break;
    }
    std::memmove(buf.data(), buf.data() + start, have - start);
    have -= start;
  }
  gzclose(f);
  return t;
}

}  // namespace sv2nl
