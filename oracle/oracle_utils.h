/*
 * oracle/oracle_utils.h -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 * Read/site helper restatements (reference: src/plp_utils.c, src/ext/htsbox/htsbox-ext.c,
 * src/regfile.c) and a khash emulation for the order-dependent ALT string (src/plp.c:193-215).
 */
#ifndef ORACLE_UTILS_H
#define ORACLE_UTILS_H

#include "hts_min.h"

unsigned char o_comp(unsigned char c); /* src/plp_utils.c:438-455 */

typedef struct {
    int nmer, splice_dist, indel_dist, trim_i5p, trim_i3p;
    int n_mm_type, n_mm, min_overhang, min_var_reads;
    double trim_f5p, trim_f3p;
} ofilter_t;

int o_query_start(const rec_t* b);
int o_query_end(const rec_t* b);
int o_check_simple_repeat(const char* ref, int64_t ref_len, int pos, int nmer);
int o_check_variant_pos(const rec_t* b, int qpos, int d5, int d3);
int o_check_variant_fpos(const rec_t* b, int qpos, double f5, double f3);
int o_dist_to_splice(const rec_t* b, int qpos, int dist);
int o_check_splice_overhang(const rec_t* b, int qpos, int dist);
int o_dist_to_indel(const rec_t* b, int qpos, int dist);
int o_read_base_quality(const rec_t* b, float pc, int mq);
int o_invert_read_orientation(const rec_t* b, int libtype);
int o_relative_position(const rec_t* b, int qpos);
int o_parse_mismatches(const rec_t* b, int n_types, int n_mis);
/* shared per-(read,position) filter cascade: src/plp.c:568-615 == src/sc-plp.c:171-216 */
int o_check_read_filters(const rec_t* b, int qpos, int is_del, int is_refskip, const ofilter_t* ef,
                         int min_bq, int min_mq);

/* ---- khash emulation: string keys (only the first char distinguishes keys here), int values ---- */
typedef struct {
    unsigned n_buckets, size, n_occupied, upper_bound;
    unsigned char* flags; /* 0 = live, 1 = deleted, 2 = empty (one byte per bucket) */
    unsigned char* keys;
    int* vals;
} okh_t;
void okh_init(okh_t* h);
void okh_free(okh_t* h);
void okh_clear(okh_t* h);
/* returns bucket; *ret = 1 new, 2 reused deleted slot, 0 present */
unsigned okh_put(okh_t* h, unsigned char key, int* ret);
void okh_del(okh_t* h, unsigned x);
#define okh_exist(h, x) ((h)->flags[x] == 0)

/* ---- region index (htslib regidx as used by src/regfile.c): 0-based inclusive intervals ---- */
typedef struct {
    int n;
    int *beg, *end, *payload; /* sorted by (beg,end); payload = caller index */
    int* maxend;              /* prefix maximum of end */
} oreg_chr_t;
typedef struct {
    int n_chr;
    char** names;
    oreg_chr_t* chr;
} oregidx_t;
oregidx_t* oreg_build(int n, const char** seqnames, const int* beg0, const int* end0);
void oreg_free(oregidx_t* r);
const oreg_chr_t* oreg_chr(const oregidx_t* r, const char* name);
int oreg_overlap(const oreg_chr_t* c, int from, int to);
/* first index i (in sorted order) that may overlap; iterate while c->beg[i] <= to */
int oreg_first(const oreg_chr_t* c, int from, int to);

#endif
