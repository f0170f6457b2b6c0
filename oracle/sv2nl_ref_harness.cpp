// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// Drives the UNMODIFIED reference sv2nl sources
//   /root/reference/standalone/sv2nl/include/{mapper,helper,vcf_info,writer}.hpp
//   /root/reference/standalone/sv2nl/source/{mapper,vcf_info,writer,helper}.cpp
//   /root/reference/library/include/binary/parser/vcf.hpp, library/source/utils.cpp
// (compiled from where they lie by oracle/Makefile; nothing is copied into this repo) behind a C ABI:
//
//   sv2nl_ref_run[_some]    = what the reference's main.cpp `run()` does (main.cpp:47-84): DupMapper, InvMapper
//                             and TraMapper over one thread pool, outputs <prefix>.dup/.inv/.tra. main.cpp itself
//                             needs cxxopts (absent from the image), so its ten lines of set-up are repeated here.
//   sv2nl_ref_check         = {Dup,Inv,Tra}Mapper::check_condition (mapper.cpp:50-79,144-156) on two records
//                             given as plain fields -- the predicates the CUDA `accept<>`, the C++ tool and the
//                             Python restatement are pinned against.
//   sv2nl_ref_validate      = validate_record (helper.hpp:52-63)
//   sv2nl_ref_map_key       = format_map_key (helper.hpp:84-91), sv2nl_ref_format_keys = Writer::format_keys
//                             (writer.cpp:21-27)
//
// htslib is replaced by the text-VCF stand-in under oracle/stubs/ (see oracle/stubs/htslib/hts.h): the
// reference's OWN parser layer (VcfRanges, Sv2nlInfoField::update), mappers, cache rule and writer run as
// written; only the field extraction below them is this repo's. Compiled with -fno-access-control so that
// the mappers' private check_condition members can be called directly.
#include <spdlog/fmt/ostr.h>
#include <spdlog/spdlog.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "mapper.hpp"

namespace {

sv2nl::Sv2nlVcfRecord make_record(const char* chrom, std::uint32_t pos, std::uint32_t svend, const char* svtype,
                                  const char* chr2, int strand1, int strand2) {
  sv2nl::Sv2nlVcfRecord r;
  r.chrom = chrom ? chrom : "";
  r.pos = pos;
  r.info->svend = svend;
  r.info->svtype = svtype ? svtype : "";
  r.info->chr2 = chr2 ? chr2 : "";
  r.info->strand1 = strand1 != 0;
  r.info->strand2 = strand2 != 0;
  return r;
}

int copy_out(const std::string& s, char* out, int cap) {
  if (out && cap > 0) {
    const int n = (int)std::min<size_t>(s.size(), (size_t)cap - 1);
    std::memcpy(out, s.data(), (size_t)n);
    out[n] = 0;
  }
  return (int)s.size();
}

sv2nl::mapper_options opts(const char* nl, const char* sv, const std::string& out, const char* nl_type,
                           const char* sv_type, std::uint32_t diff, bool use_strand) {
  return sv2nl::mapper_options()
      .nl_file(nl)
      .sv_file(sv)
      .output_file(out)
      .nl_type(nl_type)
      .sv_type(sv_type)
      .diff(diff)
      .use_strand(use_strand);
}

}  // namespace

extern "C" {

// kind: 0 = DupMapper, 1 = InvMapper, 2 = TraMapper. Returns check_condition(nl, sv) as 0/1.
int sv2nl_ref_check(int kind, std::uint32_t diff, int use_strand,                                    //
                    const char* nl_chrom, std::uint32_t nl_pos, std::uint32_t nl_end, const char* nl_type,
                    const char* nl_chr2, int nl_s1, int nl_s2,                                       //
                    const char* sv_chrom, std::uint32_t sv_pos, std::uint32_t sv_end, const char* sv_type,
                    const char* sv_chr2) {
  const auto nl = make_record(nl_chrom, nl_pos, nl_end, nl_type, nl_chr2, nl_s1, nl_s2);
  const auto sv = make_record(sv_chrom, sv_pos, sv_end, sv_type, sv_chr2, 1, 1);
  const std::string sink = "/dev/null";
  const auto o = opts("", "", sink, "", "", diff, use_strand != 0);
  switch (kind) {
    case 0: return sv2nl::DupMapper(o).check_condition(nl, sv) ? 1 : 0;
    case 1: return sv2nl::InvMapper(o).check_condition(nl, sv) ? 1 : 0;
    case 2: return sv2nl::TraMapper(o).check_condition(nl, sv) ? 1 : 0;
    default: return -1;
  }
}

// validate_record: the (possibly swapped) fields come back through the out parameters; chrom/chr2 as a flag
// (1 = they were exchanged).
int sv2nl_ref_validate(const char* chrom, std::uint32_t pos, std::uint32_t svend, const char* svtype,
                       const char* chr2, std::uint32_t* out_pos, std::uint32_t* out_end, int* swapped_chroms) {
  const auto in = make_record(chrom, pos, svend, svtype, chr2, 1, 1);
  const auto r = sv2nl::validate_record(in);
  *out_pos = r.pos;
  *out_end = r.info->svend;
  *swapped_chroms = (r.chrom != in.chrom || r.info->chr2 != in.info->chr2) ? 1 : 0;
  return 0;
}

int sv2nl_ref_map_key(const char* chrom, std::uint32_t pos, std::uint32_t svend, const char* svtype,
                      const char* chr2, char* out, int cap) {
  return copy_out(sv2nl::format_map_key(make_record(chrom, pos, svend, svtype, chr2, 1, 1)), out, cap);
}

int sv2nl_ref_format_keys(const char* chrom, std::uint32_t pos, std::uint32_t svend, const char* svtype,
                          const char* chr2, char* out, int cap) {
  return copy_out(sv2nl::Writer::format_keys(make_record(chrom, pos, svend, svtype, chr2, 1, 1)), out, cap);
}

// The reference tool's run(): three mappers, one pool (destroyed -- i.e. drained -- before the writers close).
// mappers: bit 0 DupMapper, bit 1 InvMapper, bit 2 TraMapper (a mapper left out still writes its header line). The
// reference's TraMapper visits a large fraction of ALL BND records per NL record (one tree over raw [POS, POS2]
// intervals, mapper.cpp:103): at config E's full size that is ~3e11 pair visits, so the full-size comparison leaves it out.
int sv2nl_ref_run_some(const char* nl, const char* sv, const char* out_prefix, std::uint32_t diff, int threads,
                       int use_strand, int mappers) {
  try {
    const std::string dup_out = std::string(out_prefix) + ".dup", inv_out = std::string(out_prefix) + ".inv",
                      tra_out = std::string(out_prefix) + ".tra";
    auto dup = sv2nl::DupMapper(opts(nl, sv, dup_out, "TDUP", "DUP", diff, true));
    auto inv = sv2nl::InvMapper(opts(nl, sv, inv_out, "INV", "INV", diff, use_strand != 0));
    auto tra = sv2nl::TraMapper(opts(nl, sv, tra_out, "TRA", "BND", diff, true));
    {
      auto pool = dp::thread_pool(threads > 0 ? (unsigned)threads : 4u);
      if (mappers & 1) dup.map(pool);
      if (mappers & 2) inv.map(pool);
      if (mappers & 4) tra.map(pool);
    }
    tra.close_writer();
    inv.close_writer();
    dup.close_writer();
    return 0;
  } catch (std::exception const& e) {
    std::fprintf(stderr, "sv2nl_ref_run: %s\n", e.what());
    return -1;
  }
}
int sv2nl_ref_run(const char* nl, const char* sv, const char* out_prefix, std::uint32_t diff, int threads,
                  int use_strand) {
  return sv2nl_ref_run_some(nl, sv, out_prefix, diff, threads, use_strand, 7);
}

}  // extern "C"
