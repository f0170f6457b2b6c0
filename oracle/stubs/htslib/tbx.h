/* TEST INFRASTRUCTURE ONLY -- see oracle/stubs/htslib/hts.h. Tabix is declared, not implemented. */
#ifndef BCU_STUB_HTSLIB_TBX_H
#define BCU_STUB_HTSLIB_TBX_H
#include "hts.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct tbx_t tbx_t;
tbx_t* tbx_index_load(const char* fn);
void tbx_destroy(tbx_t* tbx);
int tbx_name2id(tbx_t* tbx, const char* ss);
hts_itr_t* tbx_itr_queryi(tbx_t* tbx, int tid, hts_pos_t beg, hts_pos_t end);
hts_itr_t* tbx_itr_querys(tbx_t* tbx, const char* reg);
int tbx_itr_next(htsFile* fp, tbx_t* tbx, hts_itr_t* iter, void* data);
#ifdef __cplusplus
}
#endif
#endif
