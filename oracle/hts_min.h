/*
 * oracle/hts_min.h -- TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
 *
 * Minimal, single-threaded BGZF/BAM/BAI/FASTA access for the CPU oracle. raer links htslib
 * (via Rhtslib, not vendored in /root/reference), so the pieces of
 * htslib the pileup path relies on are restated here from the published SAM/BAM/BAI/BGZF
 * specification (SAMv1.pdf sections 4.1, 4.2, 5.2) and the documented faidx layout.
 */
#ifndef ORACLE_HTS_MIN_H
#define ORACLE_HTS_MIN_H

#include <stdint.h>
#include <stdio.h>

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128

enum { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X, OP_B };
#define cig_op(c) ((int)((c) & 0xf))
#define cig_len(c) ((int)((c) >> 4))
/* bit0: consumes query, bit1: consumes reference (htslib bam_cigar_type table 0x3C1A7) */
#define cig_type(o) ((0x3C1A7 >> ((o) << 1)) & 3)

extern const char nt16_str[17]; /* "=ACMGRSVTWYHKDBN" */

typedef struct {
    int32_t tid, pos, mtid, mpos, isize;
    uint16_t flag;
    uint8_t mapq;
    int l_qname, n_cigar, l_qseq, l_aux;
    char* qname;     /* NUL terminated */
    uint32_t* cigar; /* n_cigar */
    uint8_t* seq;    /* (l_qseq+1)/2, 4-bit packed */
    uint8_t* qual;   /* l_qseq */
    uint8_t* aux;    /* l_aux */
    uint8_t* data;   /* owning buffer */
    int m_data;
} rec_t;

typedef struct {
    FILE* fp;
    uint8_t* cbuf;    /* compressed block */
    uint8_t* ubuf;    /* inflated block */
    int ulen, uoff;   /* bytes in ubuf, read offset */
    int64_t blk_addr; /* file offset of current block */
    int eof;
    /* header */
    int n_targets;
    char** target_name;
    int32_t* target_len;
} bam_file_t;

bam_file_t* bamf_open(const char* path);
void bamf_close(bam_file_t* f);
int bamf_seek(bam_file_t* f, uint64_t voffset);
/* returns 0 on success, -1 on EOF, < -1 on error */
int bamf_read1(bam_file_t* f, rec_t* r);
int bamf_name2tid(const bam_file_t* f, const char* name);
void rec_free(rec_t* r);
void rec_copy(rec_t* dst, const rec_t* src);

#define rec_seqi(r, i) (((r)->seq[(i) >> 1] >> ((~(i) & 1) << 2)) & 0xf)
int rec_rlen(const rec_t* r);          /* reference length of CIGAR (bam_cigar2rlen) */
int rec_endpos(const rec_t* r);        /* bam_endpos: pos+rlen, or pos+1 when rlen==0 */
const uint8_t* rec_aux_get(const rec_t* r, const char tag[2]); /* points at type byte */
const char* rec_aux_z(const rec_t* r, const char tag[2]);      /* Z/H tag text or NULL */

/* BAI: smallest virtual offset where records overlapping [beg, ...) on tid can start.
 * returns 0 ok (and *voff), 1 if the index holds no alignments for tid, -1 if unreadable */
int bai_query_start(const char* bai_path, int tid, int beg, uint64_t* voff);

/* FASTA via .fai: whole contig, case preserved. returns malloc'd buffer, length in *len, or NULL */
char* fasta_fetch(const char* fa_path, const char* name, int64_t* len);

/* parse "chr", "chr:beg-end", "chr:beg" (1-based inclusive, commas allowed) to 0-based half open.
 * returns length of the name part or -1 */
int parse_region(const char* s, int* beg, int* end);

#endif
