/* TEST INFRASTRUCTURE ONLY.
 *
 * Minimal stand-in for the htslib headers, written for this repo (htslib itself -- samtools/htslib, pinned
 * 1.15.1 by the reference's cmake/htslib.cmake:64 -- is not in the image and cannot be fetched). It declares
 * exactly the types and functions that the reference's binary/parser/vcf.hpp names (vcf.hpp:7-8, 35-50,
 * 120-149, 265, 306-308, 493-550, 585), so that the UNMODIFIED reference sv2nl sources compile here; the
 * functions are implemented over plain/gzip TEXT VCF in oracle/stubs/htslib_text.cpp following htslib's
 * published behaviour for those calls (vcf_parse: pos = POS-1, INFO types from the ##INFO header lines,
 * undeclared tags become String; bcf_get_info_values return codes -1/-2/-3). Tabix queries are not
 * implemented (sv2nl never calls them): they report failure.
 */
#ifndef BCU_STUB_HTSLIB_HTS_H
#define BCU_STUB_HTSLIB_HTS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hts_pos_t;

typedef struct kstring_t {
  size_t l, m;
  char* s;
} kstring_t;

typedef struct htsFile htsFile;       /* opaque: a text VCF being read */
typedef struct hts_itr_t hts_itr_t;   /* opaque: never created by the stub */

htsFile* hts_open(const char* fn, const char* mode);
int hts_close(htsFile* fp);
void hts_itr_destroy(hts_itr_t* iter);

#ifdef __cplusplus
}
#endif
#endif
