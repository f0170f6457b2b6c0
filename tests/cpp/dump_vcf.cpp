// Test helper: parse a VCF with the sv2nl tool's reader (standalone/sv2nl/vcf_text.hpp) and print the table
// as TSV, one record per line, so that tests/test_sv2nl.py can compare it with the Python reader.
// usage: dump_vcf <file> <nls|delly>
#include <cstdio>

#include "../../standalone/sv2nl/vcf_text.hpp"

int main(int argc, char** argv) {
  if (argc != 3) { std::fprintf(stderr, "usage: dump_vcf <file> <nls|delly>\n"); return 2; }
  try {
    const sv2nl::VcfTable t = sv2nl::read_vcf(argv[1], argv[2]);
    for (auto const& c : t.contigs) std::printf("contig\t%s\n", c.c_str());
    for (std::size_t i = 0; i < t.size(); ++i)
      std::printf("rec\t%s\t%u\t%u\t%s\t%s\t%d\t%d\n", t.chrom_name(i).c_str(), t.pos[i], t.svend[i],
                  t.type_name(i).c_str(), t.chr2[i] == sv2nl::kNoChrom ? "" : t.chrom_names[t.chr2[i]].c_str(),
                  (int)t.strand1[i], (int)t.strand2[i]);
  } catch (const binary::VcfReaderError& e) {
    std::fprintf(stderr, "VcfReaderError: %s\n", e.what());
    return 1;
  }
  return 0;
}
