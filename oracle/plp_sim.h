/*
 * oracle/plp_sim.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restatement of the htslib multi-file pileup iterator the reference drives at
 * src/plp.c:892-915 and src/sc-plp.c:673-703 (bam_mplp_init / bam_mplp_init_overlaps /
 * bam_mplp_set_maxcnt / bam_mplp64_auto). htslib is NOT vendored in /root/reference
 * (Rhtslib, version unpinned: DESCRIPTION:60-61), so this follows the published behaviour of
 * htslib sam.c (1.13 .. 1.18: bam_plp_push, bam_plp64_next, resolve_cigar2, overlap_push,
 * tweak_overlap_quality); see SURVEY.md Appendix C.
 */
#ifndef ORACLE_PLP_SIM_H
#define ORACLE_PLP_SIM_H

#include "hts_min.h"

typedef struct {
    const rec_t* b;
    int qpos;
    int is_del, is_refskip;
    int ordinal; /* per-file arrival index of the read (for tests of order-dependent logic) */
} plp1_t;

/* read source: returns 0 and fills r, -1 at EOF, < -1 error (same contract as readaln) */
typedef int (*plp_read_f)(void* data, rec_t* r);

typedef struct plp_iter plp_iter_t;
typedef struct mplp mplp_t;

/* overlap_mode: 0 = no overlap handling, 1 = htslib >= 1.13 rule (qname hash picks the kept mate,
 * deletion catch-up), 2 = htslib <= 1.12 rule (first-seen mate always kept) */
mplp_t* mplp_init(int n, plp_read_f func, void** data, int overlap_mode, int maxcnt);
void mplp_destroy(mplp_t* m);
/* returns number of files with a column at (tid,pos) (>0), 0 at end, <0 on error */
int mplp_auto(mplp_t* m, int* tid, int* pos, int* n_plp, const plp1_t** plp);

#endif
