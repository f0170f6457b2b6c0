/* TEST INFRASTRUCTURE ONLY -- see oracle/stubs/htslib/hts.h. */
#ifndef BCU_STUB_HTSLIB_VCF_H
#define BCU_STUB_HTSLIB_VCF_H
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif

#define BCF_HT_FLAG 0
#define BCF_HT_INT 1
#define BCF_HT_REAL 2
#define BCF_HT_STR 3
#define BCF_HT_LONG (BCF_HT_INT | 0x100)

#define BCF_DT_ID 0
#define BCF_DT_CTG 1
#define BCF_DT_SAMPLE 2

typedef struct bcf_idinfo_t bcf_idinfo_t;
typedef struct bcf_idpair_t {
  const char* key;
  const bcf_idinfo_t* val;
} bcf_idpair_t;

typedef struct bcf_hdr_t {
  int32_t n[3];        /* n[BCF_DT_CTG] = number of ##contig lines (plus contigs met in records) */
  bcf_idpair_t* id[3]; /* id[BCF_DT_CTG][i].key = contig name, header order */
  void* impl;          /* stub-private: INFO tag types, name -> id maps */
} bcf_hdr_t;

typedef struct bcf1_t {
  hts_pos_t pos;  /* 0-based: POS - 1 */
  hts_pos_t rlen; /* END - pos when INFO/END is given, else strlen(REF) */
  int32_t rid;    /* index into id[BCF_DT_CTG] */
  void* impl;     /* stub-private: the parsed INFO column */
} bcf1_t;

bcf_hdr_t* bcf_hdr_read(htsFile* fp);
void bcf_hdr_destroy(bcf_hdr_t* h);
bcf1_t* bcf_init(void);
void bcf_destroy(bcf1_t* v);
#define bcf_init1() bcf_init()
#define bcf_destroy1(v) bcf_destroy(v)
/* 0 = a record was read, -1 = end of file, < -1 = error */
int bcf_read(htsFile* fp, const bcf_hdr_t* h, bcf1_t* v);
#define bcf_read1(fp, h, v) bcf_read((fp), (h), (v))
int vcf_parse(kstring_t* s, const bcf_hdr_t* h, bcf1_t* v);
#define vcf_parse1(s, h, v) vcf_parse((s), (h), (v))
const char* bcf_seqname_safe(const bcf_hdr_t* hdr, const bcf1_t* rec);
int bcf_hdr_name2id(const bcf_hdr_t* hdr, const char* id);
/* returns the number of values (string: its length); -1 tag not defined in the header, -2 type clash,
 * -3 tag absent from this record. *dst is (re)allocated with malloc, *ndst = elements allocated. */
int bcf_get_info_values(const bcf_hdr_t* hdr, bcf1_t* line, const char* tag, void** dst, int* ndst, int type);

#ifdef __cplusplus
}
#endif
#endif
