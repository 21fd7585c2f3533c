// raer_dev.h -- internal structs shared by the host runtime and the CUDA kernels.
#ifndef RAER_DEV_H
#define RAER_DEV_H

#include <stdint.h>

#include "raer_b200.h"

#define RAER_NCLS 6           // shared-memory counter classes per (file, strand): A C G T other X
#define RAER_CLS_OTHER 4
#define RAER_CLS_X 5
#define RAER_BLOCK 1024       // threads per CTA of the tile kernel (one CTA per SM, 32 warps)
#define RAER_NHIST 100        // read-position bins (src/plp.c:17 NBASE_POS)

// ALT code layout (uint32 per row per file), decoded by raer_altcode_to_string()
#define RAER_ALT_MASK 0x0fu        // bit c: variant base c in {A,C,G,T} survives min_allelic_freq pruning
#define RAER_ALT_OTHER 0x10u       // a non-ACGT (IUPAC) read base was counted as variant and survives pruning
#define RAER_ALT_NINS_SHIFT 5      // 3 bits: number of distinct variant keys inserted (ACGT + one IUPAC key)
#define RAER_ALT_ORDER_SHIFT 8     // 5 x 3 bits: keys (0..3 = ACGT, 4 = the IUPAC key) in first-seen (BAM) order
#define RAER_ALT_GROW 0x800000u    // the puts at this site grew the khash from 4 to 8 buckets
#define RAER_ALT_OCODE_SHIFT 24    // 4 bits: nt16 code of the IUPAC key (complemented on the minus strand)
#define RAER_ALT_OAMBIG 0x10000000u // more than one distinct IUPAC key at the site: ALT shows the smallest code only
#define RAER_SLOT_FS 8             // words per (file, strand) in a second-pass slot

// internal device -> host conditions (never returned through the C ABI): the host re-runs the batch
#define RAER_RETRY_DEEP (-100)     // k_pileup_fast met coverage > 65535: re-run with k_pileup_tiles
#define RAER_LANE_MW 5             // k_pileup_lane: mask words per read
#define RAER_RETRY_ROWCAP (-102)   // more rows than the row buffers hold (they are sized from the reads, not the span): re-run with the count
#define RAER_RETRY_LISTCAP (-101)  // the tile read lists did not fit their allocation: re-run with the exact size

struct PlpParams {
    // read batch (device pointers)
    const raer_read_hdr* hdr;
    const int32_t* maxend;
    const uint32_t* cigar;
    const uint8_t* seq;
    uint8_t* qual;
    long long file_off[RAER_MAX_FILES + 1];
    int n_files;
    // tiling
    int t_beg, t_end, W, n_tiles;
    int2* tile_ranges;               // [n_tiles][n_files] {lo, hi}
    const int* tile_off;             // k_pileup_fast: [n_tiles * n_files + 1] offsets into tile_list
    const int* tile_list;            //   indices of the reads overlapping each (tile, file)
    int tile_list_cap;               //   entries allocated (the lists are sized before their total is known)
    unsigned int* tile_counter;      // dynamic tile scheduler
    // reference of the contig
    const char* ref;
    int ref_len;
    // configuration
    int min_depth, min_bq, trim_5p, trim_3p, indel_dist, splice_dist, min_overhang, nmer, min_var_reads;
    int n_mm_type, n_mm, report_multiallelic, want_stats, any_okv;
    int umi_mode;                    // bulk UMI mode: hdr.aux = end of the positions an earlier read of the same UMI takes
    const int32_t* umi_holes;        //   (read index, position) pairs sorted by read index: positions that stay counted
    int n_umi_holes;
    int aei_mode;                    // no rows: per file the 4 x 4 (REF of the row, counted base) sums of the rows a tile would write
    unsigned long long* aei_out;     //   [n_files][16]
    int fast;                        // only mapq / base-quality filters configured: branch-free item path
    int use_fast;                    // launch k_pileup_fast (raer_fast.cuh) instead of k_pileup_tiles
    int use_lane;                    // launch k_pileup_lane (raer_lane.cuh): lane per read, TMA-staged chunks
    int carry_only;                  //   the tile lists hold only the reads that started in an earlier tile
    const int* tile_first;           //   [(n_tiles + 1) * n_files] first read that starts at or behind the tile's first position
    uint32_t* pmask_own;             //   [n_reads][RAER_LANE_MW] counted-reference-match masks of the tiles' own reads
    uint32_t* pmask_cin;             //   [tile_list_cap][RAER_LANE_MW] ... of the carry-in list entries
    int stage_bytes;                 //   per-warp staging slice (sequence span | quality span | masks)
    double ftrim_5p, ftrim_3p, min_af;
    int min_mapq[RAER_MAX_FILES];
    int okv[RAER_MAX_FILES];
    int reg_beg, reg_end;
    const uint32_t* site_mask;
    int mask_beg, mask_bits;
    // outputs
    int* row_pos;
    unsigned int* row_meta;
    double* row_stats;
    int* row_counts;
    unsigned int* row_alt;
    long long row_cap;
    unsigned long long* row_counter;
    int2* tile_rows;                 // per tile {first row, n rows}
    int* err_flag;
    unsigned long long* phase_clk;   // development (RAER_PHASE_PROF): per-phase clock sums of the fast tile kernel, or nullptr
    int* grow;                       // [n_files * 2]: 0x7fffffff - first position at which the puts of an UNWRITTEN
                                     // (file, strand) grew its khash from 4 to 8 buckets (0 = none)
    // shared memory plan
    int S;                           // second-pass site slots per round
    int cnt_words;                   // words of the counter region (>= S * slot_words)
    int slot_words;
};

#endif
