/*
 * raer_b200.h -- C ABI of the B200-native pileup engines (libraer_b200.so).
 *
 * Two layers, both plain C (pointers and sizes only, no torch / R types):
 *
 *  1. The ".Call" layer: raer_pileup() / raer_scpileup() take exactly the arguments the reference
 *     passes through .Call(".pileup", 13 args) and .Call(".scpileup", 14 args)
 *     (reference: src/init.c:12-18, src/plp.c:1143-1145, src/sc-plp.c:945-948) and return what
 *     those calls return (column vectors per BAM file + rpbz/vdb/sor; three text files).
 *     raer_get_region() / raer_fisher_exact() cover the two remaining registered routines
 *     (src/plp_utils.c:14-46, src/utils.c:14-44) so that one shared object can replace raer.so.
 *     The R shim that binds these to SEXPs is raer_b200/csrc/r_shim.c.
 *
 *  2. The batch layer: host code (BGZF inflate + record parsing, src/plp.c:32-78 semantics)
 *     packs alignment records into a read batch (raer_read_batch) which raer_batch_*() hands to
 *     the CUDA kernels. This is the boundary between the C host side and the device code; bench.py
 *     drives it directly for the device-resident and host-buffer (e2e) measurements.
 *
 * There is no CPU fallback: every compute entry point returns RAER_ERR_NO_DEVICE when no CUDA
 * device is usable.
 */
#ifndef RAER_B200_H
#define RAER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAER_OK 0
#define RAER_ERR -1            /* generic failure ("[raer internal] error detected during pileup") */
#define RAER_ERR_ARG -2        /* argument validation failed (Rf_error before any allocation) */
#define RAER_ERR_IO -3         /* cannot open / parse BAM, BAI or FASTA */
#define RAER_ERR_NO_DEVICE -4  /* no usable CUDA device: the product has no CPU path */
#define RAER_ERR_CUDA -5       /* CUDA runtime error, see raer_last_error() */
#define RAER_ERR_MD -6         /* MD tag inconsistent with CIGAR while max_mismatch_type is active */
#define RAER_ERR_LIMIT -7      /* input exceeds a documented limit (files per call, read length) */
#define RAER_ERR_INTERRUPT -8  /* the caller's poll callback asked to stop (user interrupt); everything was released */

/* Interrupt hook. The reference polls R_CheckUserInterrupt every 2^18 pileup columns (src/plp.c:925-931,
 * src/plp_utils.c:7-12). Here the CALLING thread invokes poll(poll_ctx) between read batches (never a worker thread:
 * the R API is single threaded); a nonzero return stops the call, which releases its resources and returns
 * RAER_ERR_INTERRUPT. NULL = never polled. */
typedef int (*raer_poll_fn)(void* poll_ctx);

#define RAER_MAX_FILES 32      /* BAM files per raer_pileup call (shared-memory counter budget) */

const char* raer_last_error(void);
const char* raer_version(void);
int raer_device_count(void);

/* ------------------------------------------------------------------ layer 1: .Call equivalents */

/* .Call(".pileup", bampaths, indexes, n, fapath, region, lst, int_args, dbl_args, lgl_args,
 *       libtype, only_keep_variants, min_mapQ, umi)       -- reference src/plp.c:1143-1198 */
typedef struct raer_pileup_args {
    int32_t nbam;
    const char* const* bampaths;
    const char* const* indexes;      /* nbam entries; an entry may be NULL => "<bam>.bai" */
    const char* fapath;
    const char* region;              /* NULL == character(0) */
    int32_t n_regions;               /* 0 == lst NULL (or empty) */
    const char* const* reg_seqnames; /* lst[[1]] */
    const int32_t* reg_start;        /* lst[[2]], 1-based */
    const int32_t* reg_end;          /* lst[[3]], 1-based inclusive */
    int32_t int_args[14];            /* max_depth, min_depth, min_base_quality, trim_5p, trim_3p, indel_dist,
                                        splice_dist, min_splice_overhang, homopolymer_len, min_variant_reads,
                                        max_mismatch_type[2], bam_flags[2]  (R/pileup.R:434-477) */
    double dbl_args[5];              /* ftrim_5p, ftrim_3p, min_allelic_freq, read_bqual[2] */
    int32_t lgl_args[2];             /* report_multiallelic, remove_overlaps */
    const int32_t* libtype;          /* nbam: 0 unstranded, 1 fr-first-strand, 2 fr-second-strand */
    const int32_t* only_keep_variants; /* nbam */
    const int32_t* min_mapq;         /* nbam */
    const char* umi_tag;             /* NULL == character(0) */
    /* extensions (zero-initialise for reference behaviour) */
    int32_t overlap_mode;            /* 0/1: htslib >= 1.13 mate-overlap rule (default); 2: htslib <= 1.12 */
    int32_t device;                  /* CUDA device ordinal */
    int32_t host_threads;            /* BGZF inflate threads; 0 = hardware concurrency */
    int32_t shard_index;             /* multi-GPU: this process handles contig-tiles i where       */
    int32_t shard_count;             /*   i % shard_count == shard_index; 0/1 = everything         */
    int32_t aei_mode;                /* 1: no rows; raer_pileup_result.aei = per file, the counts of the rows that would have
                                        been written, summed by (REF of the row, counted base): all calc_AEI() needs
                                        (reference R/calc_AEI.R:274-323). Needs the fast tile kernel. */
    int32_t pad1;
    raer_poll_fn poll;               /* interrupt hook, see raer_poll_fn; NULL = none */
    void* poll_ctx;
} raer_pileup_args;

/* Column layout of the .Call(".pileup") result (src/plp_data.c:285-289, 321-361). Every file has
 * the same rows (src/plp.c:445-459); row order = contig order of the first BAM header, ascending
 * position, '+' before '-'. */
typedef struct raer_pileup_result {
    int64_t nrows;
    int32_t nbam;
    int32_t n_targets;
    char** target_names;
    int32_t* tid;
    int32_t* pos;     /* 1-based */
    char* strand;     /* '+' | '-' */
    char* ref;        /* REF on that strand */
    double* rpbz;
    double* vdb;
    double* sor;
    char* alt;        /* [nbam][nrows][12], NUL terminated: "-" or comma-joined bases */
    int32_t* counts;  /* [nbam][nrows][8]: nRef nAlt nA nT nC nG nN nX */
    /* timing breakdown of this call, seconds (BASELINE.json: "host decode and H2D broken out") */
    double t_decode, t_pack, t_h2d, t_kernel, t_d2h, t_total;
    int64_t n_records, n_aligned_bases, gpu_launches;
    int64_t* aei;     /* aei_mode: [nbam][4][4], index [REF of the row: A C G T][nA nC nG nT], else NULL */
    int64_t n_inflated_bytes; /* BGZF payload inflated by the call, all files (drops when a site list lets the ingest skip
                                 through the BAI index) */
    int64_t n_seeks;          /* index jumps taken between the intervals of a site / region list */
    int64_t n_device_contigs; /* contigs whose BGZF blocks were inflated, parsed and packed by kernels (raer_gingest.cu) */
    int64_t n_host_contigs;   /* contigs the device ingest handed back to the host stream (device ingest) */
} raer_pileup_result;

int raer_pileup(const raer_pileup_args* args, raer_pileup_result* out);
void raer_pileup_result_free(raer_pileup_result* r);

/* .Call(".scpileup", bampaths, indexes, query_region, lst, barcodes, cbtag, int_args, dbl_args,
 *       libtype, outfns, umi, remove_overlaps, min_mapq, min_counts) -- src/sc-plp.c:945-1031 */
typedef struct raer_scpileup_args {
    int32_t nbam;                    /* > 1 => Smart-seq2 mode: barcode i names BAM i (src/sc-plp.c:760) */
    const char* const* bampaths;
    const char* const* indexes;
    const char* region;
    int32_t n_sites;
    const char* const* site_seqnames;
    const int32_t* site_pos;         /* 1-based */
    const int32_t* site_strand;      /* 1 '+', 2 '-' */
    const char* const* site_ref;
    const char* const* site_alt;
    const int32_t* site_idx;         /* 1-based row index written to the matrix file */
    int32_t n_barcodes;
    const char* const* barcodes;
    const char* cb_tag;              /* NULL == character(0) */
    int32_t int_args[14];
    double dbl_args[5];
    int32_t libtype;
    const char* outfns[3];           /* matrix, sites, barcodes */
    const char* umi_tag;             /* NULL == character(0) */
    int32_t remove_overlaps;
    int32_t min_mapq;
    int32_t min_counts;
    int32_t overlap_mode;
    int32_t device;
    int32_t host_threads;
    int32_t pad1;
    raer_poll_fn poll;               /* interrupt hook, see raer_poll_fn; NULL = none */
    void* poll_ctx;
} raer_scpileup_args;

/* returns the status the reference returns through ScalarInteger (>= 0 ok, < 0 failure) */
int raer_scpileup(const raer_scpileup_args* args);

/* The same call with the nonzero cells handed back in memory, in the order of the lines of the counts file
 * (src/sc-plp.c:410-437 writes "<site_idx> <cb_idx> <nRef> <nAlt>"; R re-reads that file with read.table,
 * R/sc-pileup.R:390-467: this entry point removes the round trip). The three text files are
 * written as well when all of args->outfns are set, and skipped when they are NULL. */
typedef struct raer_sc_triplets {
    int64_t n;
    int32_t* site_idx; /* 1-based index column of the site (raer_scpileup_args.site_idx) */
    int32_t* cb_idx;   /* 1-based barcode column */
    int32_t* n_ref;
    int32_t* n_alt;
} raer_sc_triplets;
int raer_scpileup_triplets(const raer_scpileup_args* args, raer_sc_triplets* out);
void raer_sc_triplets_free(raer_sc_triplets* t);

/* .Call(".get_region", region): hts_parse_reg semantics; *beg 0-based, *end exclusive */
int raer_get_region(const char* region, char* chrom, int32_t chrom_cap, int32_t* beg, int32_t* end);
/* .Call(".fisher_exact", mat): mat is 4 x ncol column-major (a_ref, a_alt, b_ref, b_alt) */
int raer_fisher_exact(const int32_t* mat, int32_t ncol, double* pvalue);

/* ------------------------------------------------------------------ layer 2: read batches */

/* One alignment record, 32 bytes, one 32-byte sector per warp-wide broadcast load. */
typedef struct raer_read_hdr {
    int32_t pos;        /* 0-based leftmost reference position */
    int32_t end;        /* pos + reference length of the CIGAR (exclusive) */
    uint16_t flag;      /* BAM flag */
    uint8_t mapq;
    uint8_t bits;       /* RAER_RB_* */
    uint16_t l_qseq;
    uint16_t n_cigar;
    uint32_t cigar_off; /* first CIGAR word in raer_read_batch.cigar */
    uint32_t sq_off;    /* byte offset of the packed sequence; qualities start at 2*sq_off */
    int32_t mate;       /* batch index of the earlier mate this read resolves its overlap with, or -1 */
    uint32_t aux;       /* single cell: index into cb/umi arrays == read index. Bulk UMI mode (src/plp.c:622-646): positions
                           below aux are taken by an earlier read with the same UMI on the same strand of the same file
                           (0 = none, 0xffffffff = the read has no UMI tag: it is skipped everywhere); else 0 */
} raer_read_hdr;

#define RAER_RB_INVERT 0x01  /* library type + flag put this read on the minus strand (plp_utils.c:331-371) */
#define RAER_RB_MMFAIL 0x02  /* parse_mismatches() > 0 for the configured max_mismatch_type */
#define RAER_RB_MMERR 0x04   /* MD tag inconsistent: hard error if the mismatch filter consults this read */
#define RAER_RB_KEEP_A 0x08  /* mate overlap: the earlier mate keeps its qualities (qname hash, htslib >= 1.13) */
#define RAER_RB_OHANG_N 0x10 /* the word after the CIGAR reads as a ref-skip (plp_utils.c:245-251 over-read) */
#define RAER_RB_UMI_HOLES 0x20 /* bulk UMI mode: raer_read_batch.umi_holes lists positions below hdr.aux that stay counted */

/* Struct-of-arrays read batch: all reads of all files that overlap the genomic window
 * [beg, end) of contig tid, file-major, coordinate sorted inside a file. Host or device pointers. */
typedef struct raer_read_batch {
    int32_t n_files;
    int32_t tid;
    int32_t beg, end;          /* positions this batch finalises */
    int64_t n_reads;
    const int64_t* file_off;   /* n_files + 1 (host memory, always) */
    const raer_read_hdr* hdr;  /* n_reads */
    const int32_t* maxend;     /* n_reads: running maximum of hdr.end inside each file range */
    const uint32_t* cigar;     /* n_cigar_words */
    const uint8_t* seq;        /* n_sq_bytes: 4-bit packed, every read starts on a byte */
    const uint8_t* qual;       /* 2 * n_sq_bytes */
    int64_t n_cigar_words;
    int64_t n_sq_bytes;
    const int32_t* pair_b;     /* n_pairs: indices of reads whose hdr.mate >= 0 */
    int64_t n_pairs;
    const int32_t* cb;         /* single cell: whitelist column (1-based) per read, 0 = none; else NULL */
    const uint32_t* umi;       /* single cell: UMI dictionary id per read, 0 = no tag; else NULL */
    int64_t n_aligned_bases;   /* M/=/X bases of the batch (metric unit), informational */
    const int32_t* umi_holes;  /* bulk UMI mode: (read index, position) pairs sorted by read index, see RAER_RB_UMI_HOLES */
    int64_t n_umi_holes;       /* number of pairs */
} raer_read_batch;

/* Per-call configuration of the bulk kernels (the unpacked FilterParam, src/plp.c:1094-1138). */
typedef struct raer_bulk_conf {
    int32_t n_files;
    int32_t min_depth, min_bq;
    int32_t trim_5p, trim_3p, indel_dist, splice_dist, min_overhang, nmer, min_var_reads;
    int32_t n_mm_type, n_mm;
    int32_t report_multiallelic;
    int32_t remove_overlaps;   /* 0 none, 1 htslib >= 1.13, 2 htslib <= 1.12 */
    int32_t want_stats;        /* 1: rpbz/vdb/sor + ALT order (reference behaviour); 0 skips the second pass */
    int32_t umi_mode;          /* 1: bulk UMI mode, hdr.aux / umi_holes carry the per-read verdicts */
    int32_t aei_mode;          /* 1: no rows, only the 4 x 4 (REF, base) sums per file: raer_ctx_last_aei() */
    int32_t pad0;
    double ftrim_5p, ftrim_3p, min_af;
    int32_t min_mapq[RAER_MAX_FILES];
    int32_t only_keep_variants[RAER_MAX_FILES];
    /* positions eligible for output: [reg_beg, reg_end) and, if site_mask != NULL, bit (p - mask_beg) set */
    int32_t reg_beg, reg_end;
    const uint32_t* site_mask; /* same memory space as the batch */
    int32_t mask_beg, mask_bits;
} raer_bulk_conf;

/* Rows produced by one batch, in host memory, ordered by (pos, strand). altcode encodes the ALT
 * string (see raer_altcode_to_string). */
typedef struct raer_rows {
    int64_t n_rows;
    int32_t n_files;
    int32_t* pos;        /* 0-based */
    uint32_t* meta;      /* bit0: 1 = '-' strand; bits 8-15: REF character */
    double* stats;       /* [n_rows][3]: rpbz, vdb, sor */
    int32_t* counts;     /* [n_rows][n_files][8] */
    uint32_t* altcode;   /* [n_rows][n_files] */
    /* [file * 2 + strand]: first 0-based position of the batch at which the variant table of that file and strand
     * grew from 4 to 8 khash buckets at a site that wrote no row for the strand (src/plp.c:487: every variant read is a
     * kh_put, written or not); INT32_MAX if none. raer_altcode_to_string callers apply it to the rows behind it. */
    int32_t grow_pos[64];
} raer_rows;
void raer_rows_free(raer_rows* r);

typedef struct raer_ctx raer_ctx; /* one per device: stream, scratch buffers, reference cache */
raer_ctx* raer_ctx_create(int32_t device);
void raer_ctx_destroy(raer_ctx* ctx);
/* stream used by the context (cudaStream_t) so that callers can time with events on it */
void* raer_ctx_stream(raer_ctx* ctx);
/* per-launch CUDA-event timing of the dominant kernel (the tile kernel: k_pileup_fast or k_pileup_tiles) for the roofline figure */
void raer_ctx_profile(raer_ctx* ctx, int on);
int raer_ctx_profile_read(raer_ctx* ctx, double* ms_sum, int64_t* n_launches);
/* install the reference bases of one contig (raw FASTA bytes, case preserved) for later batches */
int raer_ctx_set_reference(raer_ctx* ctx, int32_t tid, const char* seq, int64_t len);
/* same, from a device pointer that stays valid (bench: synthetic reference generated in HBM) */
int raer_ctx_set_reference_device(raer_ctx* ctx, int32_t tid, const void* dseq, int64_t len);

/* e2e path: host batch -> H2D -> kernels -> D2H rows. */
int raer_batch_pileup_host(raer_ctx* ctx, const raer_bulk_conf* conf, const raer_read_batch* host_batch,
                           raer_rows* out);
/* Pipelined e2e path: two contexts used alternately so that the H2D copy of one batch overlaps the kernels of
 * the previous one and the D2H of the one before. At most two batches in flight (submit returns RAER_ERR_LIMIT
 * otherwise); batches complete in submission order; the host arrays of a batch and its reference bases (raw
 * FASTA bytes of contig tid, host memory) must stay valid until its rows were collected. */
typedef struct raer_pipe raer_pipe;
raer_pipe* raer_pipe_create(int32_t device);
void raer_pipe_destroy(raer_pipe* pipe);
int raer_pipe_submit(raer_pipe* pipe, const raer_bulk_conf* conf, const raer_read_batch* host_batch, int32_t tid,
                     const char* ref, int64_t ref_len);
int raer_pipe_collect(raer_pipe* pipe, raer_rows* out);
/* aei_mode: the sums of the batch collected last on this context, n_files x 16 values */
int raer_ctx_last_aei(raer_ctx* ctx, int64_t* sums, int32_t n_files);
/* device-resident path (bench "value"): every pointer of dev_batch (except file_off) is device memory.
 * Runs the kernels on the context stream; rows stay on the device. *n_rows is read back at the end
 * unless NULL. Counts launched kernels into *launches when not NULL. */
int raer_batch_pileup_device(raer_ctx* ctx, const raer_bulk_conf* conf, const raer_read_batch* dev_batch,
                             int64_t* n_rows, int64_t* launches);
/* algorithmic bytes of a batch + its output rows (SURVEY.md section 8d constants) */
int64_t raer_batch_algorithmic_bytes(const raer_read_batch* b, int64_t n_rows, int64_t ref_bases);

/* decode an ALT code to "-" / "A,G" ... given and updating the khash size state of that
 * (file, strand) variant table (4 or 8 buckets, src/plp.c:193-215). out needs 12 bytes. */
void raer_altcode_to_string(uint32_t altcode, int ref_char, int32_t* khash_buckets, char* out);

/* ------------------------------------------------------------------ layer 2, single cell (pileup_cells)
 * The batch layer under raer_scpileup(): a read batch with the cb / umi arrays filled plus the sites of its contig go
 * to the device kernels that replace process_one_site / count_record / count_consensus_base / write_counts
 * (src/sc-plp.c:528-597, 235-309, 321-348, 410-437); what comes back are the nonzero (site, barcode) cells. */
typedef struct raer_sc_conf {
    int32_t min_bq, min_mapq;
    int32_t trim_5p, trim_3p, indel_dist, splice_dist, min_overhang;
    int32_t remove_overlaps;    /* 0 none, 1 htslib >= 1.13, 2 htslib <= 1.12 */
    int32_t has_umi;            /* 1: consensus per (site, barcode, UMI) on summed base qualities (src/sc-plp.c:321-348) */
    int32_t n_barcodes;         /* whitelist size: raer_read_batch.cb holds 1 .. n_barcodes (0 = not on the list) */
    uint32_t n_umi;             /* raer_read_batch.umi holds 1 .. n_umi (0 = read without the tag) */
    int32_t n_sites;            /* sites of the batch's contig, sorted by position; payloads may share a position */
    double ftrim_5p, ftrim_3p;
    const int32_t* site_pos;    /* 0-based; same memory space as the batch */
    const uint8_t* site_strand; /* 1 '+', 2 '-' */
    const uint8_t* site_ref;    /* first character of REF */
    const uint8_t* site_alt;    /* first character of ALT */
} raer_sc_conf;

/* nonzero cells of one batch in host memory, sorted by (site, barcode) */
typedef struct raer_sc_cells {
    int64_t n;
    int32_t* cells;             /* [n][4]: index into the site list, 1-based barcode column, nRef, nAlt */
    int64_t n_sites;
    int32_t* site_totals;       /* [n_sites][2]: nRef, nAlt summed over the cells of a site (the min_counts rule, src/sc-plp.c:586-588) */
} raer_sc_cells;
void raer_sc_cells_free(raer_sc_cells* c);

typedef struct raer_sc_ctx raer_sc_ctx; /* one per device: stream and scratch buffers of the single-cell kernels */
raer_sc_ctx* raer_sc_ctx_create(int32_t device);
void raer_sc_ctx_destroy(raer_sc_ctx* ctx);
void* raer_sc_ctx_stream(raer_sc_ctx* ctx);
/* device-resident path: every pointer of dev_batch (except file_off) and of conf is device memory; the cells stay on
 * the device. *n_cells and *n_tuples (emitted (site, barcode, UMI) observations) are read back unless NULL. */
int raer_sc_batch_device(raer_sc_ctx* ctx, const raer_sc_conf* conf, const raer_read_batch* dev_batch, int64_t* n_cells,
                         int64_t* n_tuples, int64_t* launches);
/* host path: H2D of the batch and the sites, kernels, D2H of the cells */
int raer_sc_batch_host(raer_sc_ctx* ctx, const raer_sc_conf* conf, const raer_read_batch* host_batch, raer_sc_cells* out);
/* host site arrays in conf, batch already in device memory (what the device ingest produces): kernels, D2H of the cells */
int raer_sc_batch_device_cells(raer_sc_ctx* ctx, const raer_sc_conf* conf, const raer_read_batch* dev_batch, raer_sc_cells* out);
/* contigs the last raer_scpileup() / raer_scpileup_triplets() call of this process took through the device ingest */
int64_t raer_sc_last_device_contigs(void);
/* per-kernel-class CUDA-event timing since raer_sc_ctx_profile(ctx, 1): ms[0..3] / n[0..3] = tuple emission, radix
 * sort (all passes), consensus + cell reduction, mate overlap */
void raer_sc_ctx_profile(raer_sc_ctx* ctx, int on);
int raer_sc_ctx_profile_read(raer_sc_ctx* ctx, double* ms, int64_t* n);
/* algorithmic bytes of a single-cell batch (2.04 B per aligned base model + 16 B per output cell) */
int64_t raer_sc_algorithmic_bytes(const raer_read_batch* b, int64_t n_cells);

/* Host-only probe of the ingest used by raer_pileup() (no device needed): streams the BAMs, applies
 * the read-level prefilter, packs batches of about target_reads reads and reports
 * out[0..7] = reads handed to the device (halo reads once per batch), mate pairs, CIGAR words,
 * packed sequence bytes, batches, aligned bases of all mapped records, records read, distinct reads;
 * out[8] = BGZF bytes inflated, out[9] = index jumps taken (site lists). out must hold 16 values. */
int raer_ingest_summary(const raer_pileup_args* args, int64_t target_reads, int64_t* out);
/* seconds spent in H2D copies, kernels and D2H copies by the last raer_batch_pileup_host() call */
void raer_ctx_last_timing(raer_ctx* ctx, double* h2d_s, double* kernel_s, double* d2h_s);


/* Device DEFLATE decoder used by the device ingest (one warp per BGZF block), exposed for tests and measurements:
 * n_blocks raw DEFLATE payloads at comp + c_off[i] (c_len[i] bytes, host memory) are inflated to out + u_off[i]
 * (u_len[i] bytes expected, host memory); crc[i] (NULL: not checked) is the CRC-32 the inflated bytes must have.
 * status[i] = 0 or a negative decoder code (bad block type, bad code lengths, distance too far, output length
 * mismatch, -21 CRC mismatch ...). Replaces what htslib's bgzf_read_block / zlib inflate do per block under
 * sam_itr_next (reference: src/plp.c:32-78 readaln -> sam_itr_next / sam_read1). */
int raer_gpu_inflate(const uint8_t* comp, int64_t comp_len, const int64_t* c_off, const int32_t* c_len, const int64_t* u_off,
                     const int32_t* u_len, const uint32_t* crc, int32_t n_blocks, uint8_t* out, int64_t out_len, int32_t* status,
                     double* kernel_ms);
/* The device ingest allocates from the device's stream-ordered memory pool and keeps up to 8 GB of it, and a few
 * page-locked staging buffers, for the next call of the process. This hands both back to the driver (e.g. from the
 * unload hook of a binding). */
void raer_trim_device_cache(int32_t device);
/* bam_plp_push's max_depth rule over the passing reads of one file and contig in stream order (pos / end, 0-based, end
 * exclusive): dropped[i] = 1 for the reads htslib drops (a read that starts where the previous one did while the buffer
 * holds max_depth reads). Host only; what the device ingest runs when the depth can reach max_depth. Returns the number
 * of dropped reads. */
int64_t raer_replay_max_depth(const int32_t* pos, const int32_t* end, int64_t n, int32_t max_depth, uint8_t* dropped);
/* same batch entry as raer_batch_pileup_device(), rows downloaded (what the device ingest calls per contig) */
int raer_batch_pileup_device_rows(raer_ctx* ctx, const raer_bulk_conf* conf, const raer_read_batch* dev_batch, raer_rows* out,
                                  int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* RAER_B200_H */
