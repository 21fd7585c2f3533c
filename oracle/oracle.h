/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of raer's pileup hot path (reference: /root/reference/src/plp.c,
 * src/sc-plp.c, src/plp_utils.c, src/regfile.c, src/ext/{htsbox,bcftools,samtools}).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker. The product (libraer_b200.so) never links it.
 *
 * Parity status: PINNED for counts, row sets, strands, REF/ALT on the reference's own golden
 * values (tests/test_oracle_golden.py). rpbz/vdb/sor arithmetic is pinned
 * against the reference's own bcftools-ext.c compiled from /root/reference (oracle/_ref), but the
 * reference has no golden values for the statistics themselves nor for the htslib mate-overlap
 * tie-break; those two items are "parity unpinned" (SURVEY.md section 8c).
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Arguments of .Call(".pileup", ...): src/plp.c:1143-1145, unpacked as in set_mplp_conf :1094-1138 */
typedef struct {
    int nbam;
    const char** bampaths;
    const char** indexes;    /* may be NULL: <bam>.bai is tried */
    const char* fapath;
    const char* region;      /* NULL when character(0) */
    int n_regions;           /* 0 when lst is NULL */
    const char** reg_seqnames;
    const int* reg_start;    /* 1-based inclusive */
    const int* reg_end;
    int int_args[14];
    double dbl_args[5];
    int lgl_args[2];
    const int* libtype;
    const int* only_keep_variants;
    const int* min_mapq;
    const char* umi_tag;     /* NULL when character(0) */
    int overlap_mode;        /* 1: htslib>=1.13 (default), 2: htslib<=1.12 */
} oracle_plp_args;

typedef struct {
    int64_t nrows;
    int nbam;
    int n_targets;
    char** target_names;
    /* per row (identical for every file, src/plp.c:445-459) */
    int32_t* tid;
    int32_t* pos;            /* 1-based, src/plp.c:258 */
    char* strand;            /* '+' / '-' */
    char* ref;
    double* rpbz;
    double* vdb;
    double* sor;
    /* per file, [nbam][nrows] flattened file-major */
    char* alt;               /* 12 bytes per cell, NUL terminated (src/plp.c:241) */
    int32_t* counts;         /* 8 ints per cell: nRef nAlt nA nT nC nG nN nX */
} oracle_plp_result;

int oracle_pileup(const oracle_plp_args* a, oracle_plp_result* out);
void oracle_plp_result_free(oracle_plp_result* r);

/* Arguments of .Call(".scpileup", ...): src/sc-plp.c:945-948 */
typedef struct {
    int nbam;
    const char** bampaths;
    const char** indexes;
    const char* region;
    int n_sites;
    const char** site_seqnames;
    const int* site_pos;     /* 1-based */
    const int* site_strand;  /* 1 = +, 2 = - */
    const char** site_ref;
    const char** site_alt;
    const int* site_idx;     /* 1-based row index */
    int n_barcodes;
    const char** barcodes;
    const char* cb_tag;      /* NULL when character(0) */
    int int_args[14];
    double dbl_args[5];
    int libtype;
    const char* outfns[3];   /* mtx, sites, barcodes */
    const char* umi_tag;
    int remove_overlaps;
    int min_mapq;
    int min_counts;
    int overlap_mode;
} oracle_sc_args;

int oracle_scpileup(const oracle_sc_args* a);

/* statistics restated from src/ext/bcftools/bcftools-ext.c and src/plp_utils.c:375-387 */
double oracle_calc_vdb(const int* pos, int npos);
double oracle_calc_mwu_biasZ(const int* a, const int* b, int n, int left_only, int do_Z);
double oracle_calc_sor(int fwd_ref, int rev_ref, int fwd_alt, int rev_alt);
double oracle_kf_erfc(double x);

/* aligned-base count of a BAM (M/=/X bases of mapped records): unit of the bench metric */
int64_t oracle_count_aligned_bases(const char* bampath, int64_t* n_records);

#ifdef __cplusplus
}
#endif
#endif
