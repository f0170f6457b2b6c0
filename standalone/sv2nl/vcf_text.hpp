// Minimal text-VCF reader (plain or gzip/bgzip via zlib) for sv2nl, columnar output, parse-once.
//
// Stands in for the reference's htslib-backed VcfRanges (library/include/binary/parser/vcf.hpp), which
// cannot be built in this image; it reads exactly the fields sv2nl uses:
//   chrom = column 1, pos = POS - 1 (0-based, vcf.hpp:305-310, test_vcf.cpp:100)
//   INFO SVTYPE (required), CHR2 (TRA/BND), STRAND1/STRAND2 == "+" (INV, missing keeps the default true),
//   end = POS2 if SVTYPE == BND, else SVEND for source "nls", else END   (sv2nl vcf_info.cpp:9-43)
//   contigs = ##contig=<ID=...> lines in header order                    (vcf.hpp:577-589)
#pragma once

#include <zlib.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

namespace binary {
class VcfReaderError : public std::runtime_error {  // reference: library/include/binary/exception.hpp:13
public:
  using std::runtime_error::runtime_error;
};
}  // namespace binary

namespace sv2nl {

struct VcfTable {
  std::vector<std::string> contigs;
  std::vector<std::string> chrom, chr2, svtype;
  std::vector<std::uint32_t> pos, svend;
  std::vector<std::uint8_t> strand1, strand2;  // 1 = "+"
  [[nodiscard]] std::size_t size() const { return pos.size(); }
};

namespace detail {
inline bool info_value(std::string_view info, std::string_view key, std::string_view& out) {
  std::size_t p = 0;
  while (p < info.size()) {
    std::size_t e = info.find(';', p);
    if (e == std::string_view::npos) e = info.size();
    std::string_view kv = info.substr(p, e - p);
    std::size_t eq = kv.find('=');
    if (eq != std::string_view::npos && kv.substr(0, eq) == key) {
      out = kv.substr(eq + 1);
      return true;
    }
    p = e + 1;
  }
  return false;
}
inline bool read_line(gzFile f, std::string& line) {
  line.clear();
  char buf[1 << 16];
  while (gzgets(f, buf, sizeof(buf)) != nullptr) {
    line += buf;
    if (!line.empty() && line.back() == '\n') {
      line.pop_back();
      if (!line.empty() && line.back() == '\r') line.pop_back();
      return true;
    }
  }
  return !line.empty();
}
}  // namespace detail

inline VcfTable read_vcf(const std::string& path, std::string_view source) {
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) throw binary::VcfReaderError("cannot open " + path);
  VcfTable t;
  std::string line;
  std::size_t line_no = 0;
  auto fail = [&](const std::string& what) {
    gzclose(f);
    throw binary::VcfReaderError(path + ":" + std::to_string(line_no) + ": " + what);
  };
  while (detail::read_line(f, line)) {
    ++line_no;
    if (line.rfind("##contig=<", 0) == 0) {
      std::size_t p = line.find("ID=");
      if (p != std::string::npos) {
        std::size_t e = line.find_first_of(",>", p);
        t.contigs.push_back(line.substr(p + 3, e - p - 3));
      }
      continue;
    }
    if (line.empty() || line[0] == '#') continue;
    std::vector<std::string_view> cols;
    std::string_view sv{line};
    for (std::size_t p = 0; cols.size() < 8;) {
      std::size_t e = sv.find('\t', p);
      cols.push_back(sv.substr(p, e == std::string_view::npos ? sv.size() - p : e - p));
      if (e == std::string_view::npos) break;
      p = e + 1;
    }
    if (cols.size() < 8) fail("fewer than 8 columns");
    std::string_view info = cols[7], type, end, v;
    if (!detail::info_value(info, "SVTYPE", type)) fail("INFO/SVTYPE missing");
    const char* end_key = type == "BND" ? "POS2" : (source == "nls" ? "SVEND" : "END");
    if (!detail::info_value(info, end_key, end)) fail(std::string("INFO/") + end_key + " missing");
    std::string c2;
    if (type == "TRA" || type == "BND") {
      if (!detail::info_value(info, "CHR2", v)) fail("INFO/CHR2 missing");
      c2 = std::string(v);
    }
    std::uint8_t s1 = 1, s2 = 1;
    if (type == "INV" && detail::info_value(info, "STRAND1", v)) {  // a missing key keeps the defaults
      s1 = v == "+";
      if (detail::info_value(info, "STRAND2", v)) s2 = v == "+";
    }
    t.chrom.emplace_back(cols[0]);
    t.pos.push_back(static_cast<std::uint32_t>(std::stoll(std::string(cols[1])) - 1));
    t.svend.push_back(static_cast<std::uint32_t>(std::stoll(std::string(end))));
    t.svtype.emplace_back(type);
    t.chr2.push_back(std::move(c2));
    t.strand1.push_back(s1);
    t.strand2.push_back(s2);
  }
  gzclose(f);
  return t;
}

}  // namespace sv2nl
