/* oracle/ref_stubs: lets the reference's src/ext/bcftools/bcftools-ext.c compile without htslib.
 * That file only needs fixed-width integer types from <htslib/sam.h>. TEST INFRASTRUCTURE ONLY. */
#ifndef STUB_HTSLIB_SAM_H
#define STUB_HTSLIB_SAM_H
#include <stdint.h>
#endif
