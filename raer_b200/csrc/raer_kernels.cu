// raer_kernels.cu -- hand-written sm_100a kernels of the bulk pileup path.
//
//   k_mate_overlap   htslib bam_mplp overlap logic (overlap_push + tweak_overlap_quality), one
//                    thread per overlapping pair, qualities rewritten in place before counting.
//   k_tile_ranges    per (tile, file): first/last read that can overlap the tile (binary search).
//   k_pileup_tiles   persistent CTAs pull genomic tiles from an atomic queue; per tile:
//        phase 1  warp-per-read CIGAR expansion, read/base filters (src/plp.c:568-615,
//                 src/plp_utils.c), shared-memory counters [file][strand][class][pos] updated with
//                 conflict-free atomics (a warp's lanes sit on consecutive positions of one read)
//        phase 2  per-position site decision (src/plp.c:671-704 + store_counts :358-461),
//                 block scan, coalesced row stores
//        phase 3  second, sparse pass over the tile's reads for the rows that carry a variant:
//                 100-bin read-position histograms, strand tallies, first-seen order of variant bases
//        phase 4  rpbz / vdb / sor (src/ext/bcftools/bcftools-ext.c, src/plp_utils.c:375-387) and ALT codes
//
// Nothing here is a dense contraction: no tensor cores. The bound is HBM traffic of the read batch
// (about 1.9 B per aligned base) and shared-memory atomic throughput; see DESIGN.md.
// Compiled with -fmad=false so that the fp32/fp64 statistics round exactly like the reference's C.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "raer_dev.h"

#define FULL 0xffffffffu
// counter slot of tile position pidx: one pad word every 32 positions skews the banks, so the 8-base
// items of one read (lanes 8 positions apart: 8l + (8l >> 5) covers all 32 banks) do not conflict, and
// the 8 positions of an item stay consecutive words.
#define CIDX(pidx, W) ((pidx) + ((pidx) >> 5))
#define WPAD(W) ((W) + ((W) >> 5))  // words per counter class

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int cg_op(uint32_t c) { return (int)(c & 0xf); }
__device__ __forceinline__ int cg_len(uint32_t c) { return (int)(c >> 4); }
// bit0 consumes query, bit1 consumes reference (M I D N S H P = X B)
__device__ __forceinline__ int cg_type(int op) { return (0x3C1A7 >> (op << 1)) & 3; }
enum { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X };

// 4-bit base code -> counter class: A0 C1 G2 T3, 4 = other IUPAC, -1 = N (code 15)
__device__ __forceinline__ int cls_of_code(int code) {
    const unsigned long long lut = 0x0555555455535215ull;  // (class + 1) per nibble, code 0 first
    return (int)((lut >> (code << 2)) & 0xf) - 1;
}
__device__ __forceinline__ int d_upper(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }
__device__ __forceinline__ int cls_of_ref(int up) {
    return up == 'A' ? 0 : up == 'C' ? 1 : up == 'G' ? 2 : up == 'T' ? 3 : 4;
}
__device__ __forceinline__ int d_comp_char(int c) {  // src/plp_utils.c:438-455 (upper case part)
    switch (c) {
    case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
    case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
    case 'K': return 'M'; case 'M': return 'K'; case 'R': return 'Y'; case 'Y': return 'R';
    case 'U': return 'A';
    default: return c;
    }
}
// BAM packs base 2b in the high nibble of byte b; after the swap base j sits in nibble j of the word
__device__ __forceinline__ uint32_t nib_swap(uint32_t w) { return ((w & 0x0f0f0f0fu) << 4) | ((w >> 4) & 0x0f0f0f0fu); }
__device__ __forceinline__ uint32_t nib_nonzero(uint32_t x) {  // bit 4j set iff nibble j != 0
    uint32_t t = x | (x >> 2);
    t |= t >> 1;
    return t & 0x11111111u;
}
__device__ __forceinline__ int seq_code(const uint8_t* seq, int q) { return (seq[q >> 1] >> ((~q & 1) << 2)) & 0xf; }

struct ReadCtx {
    const uint32_t* cig;
    const uint8_t* seq;
    const uint8_t* qual;
    int pos, end, flag, mapq, bits, l_qseq, n_cigar;
    int qs, qe, rev, invert, d5f, d3f, ftrim_all;
};

__device__ __forceinline__ void finish_ctx(const PlpParams& P, ReadCtx& c) {
    c.rev = (c.flag >> 4) & 1;
    c.invert = c.bits & RAER_RB_INVERT;
    // query_start / query_end (src/plp_utils.c:52-104): soft clips at either end, hard clips skipped
    int qs = 0, k;
    for (k = 0; k < c.n_cigar; ++k) {
        uint32_t w = c.cig[k];
        int op = cg_op(w);
        if (op == OP_H) continue;
        if (op == OP_S) qs += cg_len(w);
        else break;
    }
    int qe = c.l_qseq;
    for (k = c.n_cigar - 1; k >= 0; --k) {
        uint32_t w = c.cig[k];
        int op = cg_op(w);
        if (op == OP_H) continue;
        if (op == OP_S) qe -= cg_len(w);
        else break;
    }
    c.qs = qs;
    c.qe = qe;
    c.d5f = c.d3f = 0;
    c.ftrim_all = 0;
    if (P.ftrim_5p > 0 || P.ftrim_3p > 0) {  // check_variant_fpos, src/plp_utils.c:163-193
        int ql = qe - qs;
        if (ql <= 0) c.ftrim_all = 1;
        c.d5f = (int)floor(P.ftrim_5p * ql);
        c.d3f = (int)ceil(P.ftrim_3p * ql);
    }
}

// src/plp_utils.c:196-221
static __device__ int d_dist_to_splice(const ReadCtx& c, int q, int dist) {
    int ip = 0;
    for (int i = 0; i < c.n_cigar; ++i) {
        uint32_t w = c.cig[i];
        int op = cg_op(w);
        if (cg_type(op) & 1) { ip += cg_len(w); continue; }
        if (op == OP_N && ip >= q - dist && ip <= q + dist) return abs(ip - q);
    }
    return -1;
}
// src/plp_utils.c:228-266 (inclusive upper bound, BAM_CMATCH only, one-past-the-end read modelled by RAER_RB_OHANG_N)
static __device__ int d_check_splice_overhang(const ReadCtx& c, int q, int dist) {
    int ip = 0, p_op = -1;
    for (int i = 0; i < c.n_cigar; ++i) {
        uint32_t w = c.cig[i];
        int op = cg_op(w), cl = cg_len(w);
        if (op == OP_M && q >= ip && q <= cl + ip) {
            if (i > 0 && p_op == OP_N) return cl >= dist ? -1 : cl;
            int nop = (i + 1 < c.n_cigar) ? cg_op(c.cig[i + 1]) : ((c.bits & RAER_RB_OHANG_N) ? OP_N : -1);
            if (nop == OP_N) return cl >= dist ? -1 : cl;
            return 0;
        }
        p_op = op;
        if (cg_type(op) & 1) ip += cl;
    }
    return -2;
}
// src/plp_utils.c:269-306
static __device__ int d_dist_to_indel(const ReadCtx& c, int q, int dist) {
    int rp = 0, sp = q - dist, ep = q + dist;
    for (int i = 0; i < c.n_cigar; ++i) {
        uint32_t w = c.cig[i];
        int op = cg_op(w), cl = cg_len(w);
        if (cg_type(op) & 1) {
            if (op == OP_I) {
                int is = rp, ie = rp + cl;
                if (sp <= ie && is <= ep) {
                    int l = q - ie, r = is - q;
                    return l > r ? l : r;
                }
            }
            rp += cl;
        }
        if (op == OP_D && rp >= sp && rp <= ep) return abs(rp - q);
    }
    return -1;
}

// check_read_filters for an aligned base (src/plp.c:568-615): 0 count, 1 count as nX, 2 drop silently
__device__ __forceinline__ int base_verdict(const PlpParams& P, const ReadCtx& c, int mapq_fail, int q, int bq) {
    if (mapq_fail) return 1;
    if (bq < P.min_bq) return bq == 0 ? 2 : 1;
    if (P.ftrim_5p > 0 || P.ftrim_3p > 0) {
        if (c.ftrim_all) return 1;
        if (!c.rev) { if (q < c.d5f + c.qs || c.qe - q <= c.d3f) return 1; }
        else { if (c.qe - q <= c.d5f || q < c.d3f + c.qs) return 1; }
    }
    if (P.trim_5p > 0 || P.trim_3p > 0) {  // check_variant_pos, src/plp_utils.c:137-159
        if (!c.rev) { if (q < P.trim_5p + c.qs || c.qe - q <= P.trim_3p) return 1; }
        else { if (c.qe - q <= P.trim_5p || q < P.trim_3p + c.qs) return 1; }
    }
    if (P.splice_dist && d_dist_to_splice(c, q, P.splice_dist) >= 0) return 1;
    if (P.min_overhang) {
        int r = d_check_splice_overhang(c, q, P.min_overhang);
        if (r > 0) return 1;  // r == -2 propagates as -1 upstream, which the caller treats as pass (src/plp.c:735)
    }
    if (P.indel_dist && d_dist_to_indel(c, q, P.indel_dist) >= 0) return 1;
    return 0;
}

// raer_sc.cu includes this file with RAER_HELPERS_ONLY to share the read/filter helpers above.
#ifndef RAER_HELPERS_ONLY

// ---------------------------------------------------------------------------------------------
// mate overlap: htslib tweak_overlap_quality (>= 1.13 when mode == 1, <= 1.12 when mode == 2)
// ---------------------------------------------------------------------------------------------
struct CigCur {
    const uint32_t* cig;
    int k, n;      // current op, number of ops
    int icig, iseq, iref;
};
__device__ int cur_set(CigCur& c, int iref_in) {
    int pos = iref_in;
    if (pos < 0) return -1;
    c.icig = c.iseq = c.iref = 0;
    while (c.k < c.n) {
        uint32_t w = c.cig[c.k];
        int op = cg_op(w), n = cg_len(w);
        if (op == OP_S) { c.k++; c.iseq += n; c.icig = 0; continue; }
        if (op == OP_H || op == OP_P) { c.k++; c.icig = 0; continue; }
        if (op == OP_M || op == OP_EQ || op == OP_X) {
            pos -= n;
            if (pos < 0) { c.icig = n + pos; c.iseq += c.icig; c.iref += c.icig; return OP_M; }
            c.k++; c.iseq += n; c.icig = 0; c.iref += n;
            continue;
        }
        if (op == OP_I) { c.k++; c.iseq += n; c.icig = 0; continue; }
        if (op == OP_D || op == OP_N) {
            pos -= n;
            if (pos < 0) pos = 0;
            c.k++; c.icig = 0; c.iref += n;
            continue;
        }
        return -2;
    }
    c.iseq = -1;
    return -1;
}
__device__ int cur_next(CigCur& c) {
    while (c.k < c.n) {
        uint32_t w = c.cig[c.k];
        int op = cg_op(w), n = cg_len(w);
        if (op == OP_M || op == OP_EQ || op == OP_X) {
            if (c.icig >= n - 1) { c.icig = -1; c.k++; continue; }
            c.iseq++; c.icig++; c.iref++;
            return OP_M;
        }
        if (op == OP_D || op == OP_N) { c.k++; c.iref += n; c.icig = -1; continue; }
        if (op == OP_I || op == OP_S) { c.k++; c.iseq += n; c.icig = -1; continue; }
        if (op == OP_H || op == OP_P) { c.k++; c.icig = -1; continue; }
        return -2;
    }
    c.iseq = -1;
    c.iref = -1;
    return -1;
}

// Same state as calling cur_next() until c.iref >= target (or the read ends), without visiting the bases of an
// aligned block one by one: spliced mates make the literal loop step through whole reads.
__device__ __forceinline__ int cur_skip_to(CigCur& c, int target, int ret) {
    while (ret >= 0 && c.iref >= 0 && c.iref < target) {
        if (c.k < c.n && c.icig >= 0) {
            const uint32_t w = c.cig[c.k];
            const int op = cg_op(w);
            if (op == OP_M || op == OP_EQ || op == OP_X) {
                const int st = min(cg_len(w) - 1 - c.icig, target - c.iref);
                if (st > 0) { c.iseq += st; c.icig += st; c.iref += st; continue; }
            }
        }
        ret = cur_next(c);
    }
    return ret;
}

// per byte: min(a + b, 200), the quality two agreeing mates share (htslib tweak_overlap_quality)
__device__ __forceinline__ uint32_t qsum200(uint32_t a, uint32_t b) {
    const uint32_t e = __vminu2((a & 0x00ff00ffu) + (b & 0x00ff00ffu), 0x00c800c8u);
    const uint32_t o = __vminu2(((a >> 8) & 0x00ff00ffu) + ((b >> 8) & 0x00ff00ffu), 0x00c800c8u);
    return e | (o << 8);
}

// Apply the overlap rule to run + 1 consecutive positions: base ia0 + i of read A faces base ib0 + i of read B.
// 8 qualities at a time: one aligned 64-bit word of A, the matching bytes of B funnelled out of two aligned
// words, 8 packed bases of each read compared by XOR, agreeing bases summed byte-parallel; only disagreeing bases
// run scalar code. Whole words are written back (a read owns its 8-byte quality words, raer_read_hdr.sq_off).
// nbp <= 8 positions that share one aligned 8-byte quality word of A: base ia + j of A faces base ib + j of B.
// B_SHARED: other lanes work on the neighbouring chunks of the same pair at the same time. A's word belongs to this
// chunk alone, but B's bytes straddle words that the neighbours write too: B is then stored with byte / 32-bit
// stores that cover exactly this chunk's bytes instead of a read-modify-write of whole words.
template <bool B_SHARED>
__device__ __forceinline__ void overlap_chunk(uint8_t* a_q, uint8_t* b_q, const uint8_t* a_seq, const uint8_t* b_seq, int ia,
                                              int ib, int nbp, int amul, int bmul, int mode) {
    const int oa = ia & 7, ob = ib & 7;
    uint8_t* pa = a_q + (ia - oa);
    uint8_t* pb = b_q + (ib - ob);
    const bool two = ob + nbp > 8;
    const unsigned long long wa = *reinterpret_cast<const unsigned long long*>(pa);
    const unsigned long long wb0 = *reinterpret_cast<const unsigned long long*>(pb);
    const unsigned long long wb1 = two ? *reinterpret_cast<const unsigned long long*>(pb + 8) : 0ull;
    const unsigned long long QA = wa >> (8 * oa);
    const unsigned long long QB = ob ? ((wb0 >> (8 * ob)) | (wb1 << (64 - 8 * ob))) : wb0;
    // base j of the chunk in nibble j
    const uint32_t sa = nib_swap(*reinterpret_cast<const uint32_t*>(a_seq + ((ia - oa) >> 1))) >> (4 * oa);
    const uint32_t sb0 = nib_swap(*reinterpret_cast<const uint32_t*>(b_seq + ((ib - ob) >> 1)));
    const uint32_t sb1 = two ? nib_swap(*reinterpret_cast<const uint32_t*>(b_seq + ((ib - ob) >> 1) + 4)) : 0u;
    const uint32_t xs = sa ^ (ob ? ((sb0 >> (4 * ob)) | (sb1 << (32 - 4 * ob))) : sb0);
    const unsigned long long mk = nbp == 8 ? ~0ull : ((1ull << (8 * nbp)) - 1ull);
    // agreeing bases: one mate gets min(qa + qb, 200), the other 0 (mode 2: always the earlier mate)
    const unsigned long long QS = (unsigned long long)qsum200((uint32_t)QA, (uint32_t)QB) |
                                  ((unsigned long long)qsum200((uint32_t)(QA >> 32), (uint32_t)(QB >> 32)) << 32);
    unsigned long long NA = amul ? QS : 0ull, NB = bmul ? QS : 0ull;
    uint32_t mm = nib_nonzero(xs) & (nbp == 8 ? 0x11111111u : ((1u << (4 * nbp)) - 1u) & 0x11111111u);
    while (mm) {  // disagreeing bases
        const int j = (__ffs(mm) - 1) >> 2;
        mm &= mm - 1;
        const int qa = (int)((QA >> (8 * j)) & 0xff), qb = (int)((QB >> (8 * j)) & 0xff);
        int na, nb;
        if (mode == 2) {
            if (qa >= qb) { na = (uint8_t)(0.8 * qa); nb = 0; }
            else { nb = (uint8_t)(0.8 * qb); na = 0; }
        } else {
            if (qa > qb) { na = (uint8_t)(0.8 * qa); nb = 0; }
            else if (qa < qb) { nb = (uint8_t)(0.8 * qb); na = 0; }
            else { na = (uint8_t)(amul * 0.8 * qa); nb = (uint8_t)(bmul * 0.8 * qb); }
        }
        const unsigned long long bm = 0xffull << (8 * j);
        NA = (NA & ~bm) | ((unsigned long long)(na & 0xff) << (8 * j));
        NB = (NB & ~bm) | ((unsigned long long)(nb & 0xff) << (8 * j));
    }
    NA &= mk;
    NB &= mk;
    *reinterpret_cast<unsigned long long*>(pa) = (wa & ~(mk << (8 * oa))) | (NA << (8 * oa));
    if (B_SHARED) {
        uint8_t* d = b_q + ib;
        int k = 0;
        for (; k < nbp && ((ib + k) & 3); ++k) d[k] = (uint8_t)(NB >> (8 * k));
        for (; k + 4 <= nbp; k += 4) *reinterpret_cast<uint32_t*>(d + k) = (uint32_t)(NB >> (8 * k));
        for (; k < nbp; ++k) d[k] = (uint8_t)(NB >> (8 * k));
    } else {
        *reinterpret_cast<unsigned long long*>(pb) = (wb0 & ~(mk << (8 * ob))) | (NB << (8 * ob));
        if (two) *reinterpret_cast<unsigned long long*>(pb + 8) = (wb1 & ~(mk >> (64 - 8 * ob))) | (NB >> (64 - 8 * ob));
    }
}

// Apply the overlap rule to run + 1 consecutive positions: base ia0 + i of read A faces base ib0 + i of read B.
// 8 qualities at a time: one aligned 64-bit word of A, the matching bytes of B funnelled out of two aligned
// words, 8 packed bases of each read compared by XOR, agreeing bases summed byte-parallel; only disagreeing bases
// run scalar code. Whole words are written back (a read owns its 8-byte quality words, raer_read_hdr.sq_off).
__device__ __forceinline__ void overlap_run(uint8_t* a_q, uint8_t* b_q, const uint8_t* a_seq, const uint8_t* b_seq, int ia0,
                                            int ib0, int run, int amul, int bmul, int mode) {
    for (int i = 0; i <= run;) {
        const int nbp = min(8 - ((ia0 + i) & 7), run + 1 - i);
        overlap_chunk<false>(a_q, b_q, a_seq, b_seq, ia0 + i, ib0 + i, nbp, amul, bmul, mode);
        i += nbp;
    }
}

// true iff the CIGAR is one aligned block with nothing but clips around it; qs = leading soft clip, len = block length
__device__ __forceinline__ bool single_block(const uint32_t* cig, int n, int& qs, int& len) {
    int nblk = 0, q = 0;
    bool other = false;
    len = 0;
    for (int k = 0; k < n; ++k) {
        const uint32_t w = cig[k];
        const int op = cg_op(w);
        if (op == OP_M || op == OP_EQ || op == OP_X) { ++nblk; len = cg_len(w); }
        else if (op == OP_S) { if (!nblk) q += cg_len(w); }
        else if (op != OP_H) other = true;
    }
    qs = q;
    return nblk == 1 && !other;
}

// Pairs whose mates are both one aligned block: the overlap is one run, known up front. The lanes of a warp pool
// the 8-position chunks of their pairs and take one chunk each (overlaps of 1 .. 100 bases would otherwise leave
// most lanes waiting for the longest one). Every other pair goes to `general` (compacted, warp-aggregated
// append) for k_mate_overlap.
__global__ void k_mate_overlap_simple(const raer_read_hdr* __restrict__ hdr, const uint32_t* __restrict__ cigar,
                                      const uint8_t* __restrict__ seq, uint8_t* qual, const int32_t* __restrict__ pair_b,
                                      long long n_pairs, int mode, int* general, int* n_general) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int ia0 = 0, ib0 = 0, run = -1, amul = 0, bmul = 0;
    unsigned asq = 0, bsq = 0;
    bool gen = false;
    if (t < n_pairs) {
        const raer_read_hdr hb = hdr[pair_b[t]];
        const raer_read_hdr ha = hdr[hb.mate];
        if (mode == 2 || (hb.bits & RAER_RB_KEEP_A)) { amul = 1; bmul = 0; }
        else { amul = 0; bmul = 1; }
        int qsa, lena, qsb, lenb;
        if (single_block(cigar + ha.cigar_off, ha.n_cigar, qsa, lena) && single_block(cigar + hb.cigar_off, hb.n_cigar, qsb, lenb)) {
            if (hb.pos >= ha.pos) {  // cur_set(A, negative) otherwise: nothing to do
                const int lo = hb.pos, hi = min(ha.pos + lena, hb.pos + lenb);
                ia0 = qsa + (lo - ha.pos);
                ib0 = qsb;
                run = min(hi - lo - 1, min((int)ha.l_qseq - 1 - ia0, (int)hb.l_qseq - 1 - ib0));
                asq = ha.sq_off;
                bsq = hb.sq_off;
            }
        } else {
            gen = true;
        }
    }
    {   // compacted list of the remaining pairs
        const unsigned m = __ballot_sync(FULL, gen);
        int base = 0;
        if (m && lane == (int)(__ffs(m) - 1)) base = atomicAdd(n_general, __popc(m));
        base = __shfl_sync(FULL, base, m ? __ffs(m) - 1 : 0);
        if (gen) general[base + __popc(m & ((1u << lane) - 1u))] = (int)t;
    }
    // chunks are cut at the 8-byte boundaries of A's qualities
    const int nch = run >= 0 ? ((ia0 + run) >> 3) - (ia0 >> 3) + 1 : 0;
    int inc = nch;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += v;
    }
    const int tot = __shfl_sync(FULL, inc, 31);
    for (int itb = 0; itb < tot; itb += 32) {
        const int it = itb + lane;
        // owner of chunk `it`: the first lane whose inclusive prefix exceeds it
        int owner = 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int v = __shfl_sync(FULL, inc, owner + o - 1);
            if (v <= it) owner += o;
        }
        owner = min(owner, 31);
        const int o_inc = __shfl_sync(FULL, inc, owner), o_nch = __shfl_sync(FULL, nch, owner);
        const int o_ia0 = __shfl_sync(FULL, ia0, owner), o_ib0 = __shfl_sync(FULL, ib0, owner), o_run = __shfl_sync(FULL, run, owner);
        const int o_am = __shfl_sync(FULL, amul, owner), o_bm = __shfl_sync(FULL, bmul, owner);
        const unsigned o_asq = __shfl_sync(FULL, asq, owner), o_bsq = __shfl_sync(FULL, bsq, owner);
        if (it < tot) {
            const int c = it - (o_inc - o_nch);                    // chunk index inside the pair
            const int w0 = (o_ia0 >> 3) + c;                       // A's quality word
            const int ia = max(w0 << 3, o_ia0), ie = min((w0 << 3) + 8, o_ia0 + o_run + 1);
            overlap_chunk<true>(qual + 2ull * o_asq, qual + 2ull * o_bsq, seq + o_asq, seq + o_bsq, ia, o_ib0 + (ia - o_ia0), ie - ia,
                          o_am, o_bm, mode);
        }
    }
}

// Pairs with a CIGAR to walk: htslib's two cursors, literally (thread per pair; `list` = indices into pair_b, or
// nullptr for all pairs).
__global__ void k_mate_overlap(const raer_read_hdr* __restrict__ hdr, const uint32_t* __restrict__ cigar,
                               const uint8_t* __restrict__ seq, uint8_t* qual, const int32_t* __restrict__ pair_b,
                               long long n_pairs, int mode, const int* __restrict__ list, const int* __restrict__ n_list) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (list) {
        if (t >= *n_list) return;
        t = list[t];
    }
    if (t >= n_pairs) return;
    const raer_read_hdr hb = hdr[pair_b[t]];
    const raer_read_hdr ha = hdr[hb.mate];
    CigCur A, B;
    A.cig = cigar + ha.cigar_off; A.k = 0; A.n = ha.n_cigar;
    B.cig = cigar + hb.cigar_off; B.k = 0; B.n = hb.n_cigar;
    const uint8_t* a_seq = seq + ha.sq_off;
    const uint8_t* b_seq = seq + hb.sq_off;
    uint8_t* a_q = qual + 2ull * ha.sq_off;
    uint8_t* b_q = qual + 2ull * hb.sq_off;
    const int apos = ha.pos, bpos = hb.pos;
    int amul, bmul;
    if (mode == 2 || (hb.bits & RAER_RB_KEEP_A)) { amul = 1; bmul = 0; }
    else { amul = 0; bmul = 1; }
    {
        // Both mates are one aligned block: the cursors below would visit every position of the intersection in
        // order with equal positions, i.e. one run.
        int qsa, lena, qsb, lenb;
        if (single_block(A.cig, A.n, qsa, lena) && single_block(B.cig, B.n, qsb, lenb)) {
            if (bpos < apos) return;  // cur_set(A, negative): nothing to do
            const int lo = bpos, hi = min(apos + lena, bpos + lenb);
            const int ia0 = qsa + (lo - apos), ib0 = qsb;
            const int run = min(hi - lo - 1, min((int)ha.l_qseq - 1 - ia0, (int)hb.l_qseq - 1 - ib0));
            if (run >= 0) overlap_run(a_q, b_q, a_seq, b_seq, ia0, ib0, run, amul, bmul, mode);
            return;
        }
    }
    int iref = bpos;
    int a_ret = cur_set(A, iref - apos);
    if (a_ret < 0) return;
    int b_ret = cur_set(B, iref - bpos);
    if (b_ret < 0) return;
    while (true) {
        a_ret = cur_skip_to(A, iref - apos, a_ret);
        if (a_ret < 0) break;
        if (iref < A.iref + apos) iref = A.iref + apos;
        b_ret = cur_skip_to(B, iref - bpos, b_ret);
        if (b_ret < 0) break;
        if (iref < B.iref + bpos) iref = B.iref + bpos;
        iref++;
        if (A.iref + apos != B.iref + bpos) {
            if (mode == 2) continue;
            if (A.iref + apos < B.iref + bpos && B.k > 0 && cg_op(B.cig[B.k - 1]) == OP_D) {
                bool done = false;
                do {
                    a_q[A.iseq] = amul ? (uint8_t)(a_q[A.iseq] * 0.8) : 0;
                    a_ret = cur_next(A);
                    if (a_ret < 0) { done = true; break; }
                } while (A.iref + apos < B.iref + bpos);
                if (done) return;
            } else if (A.k > 0 && cg_op(A.cig[A.k - 1]) == OP_D) {
                bool done = false;
                do {
                    b_q[B.iseq] = bmul ? (uint8_t)(b_q[B.iseq] * 0.8) : 0;
                    b_ret = cur_next(B);
                    if (b_ret < 0) { done = true; break; }
                } while (B.iref + bpos < A.iref + apos);
                if (done) return;
            } else {
                continue;
            }
        }
        if (A.iseq > ha.l_qseq || B.iseq > hb.l_qseq) return;
        // Both cursors sit on the same reference position inside an aligned block. The literal loop would now
        // step both by one base per iteration until one of the blocks ends; do those `run` extra positions here
        // without the cursor machinery (same positions, same order, same arithmetic).
        int run = min(cg_len(A.cig[A.k]) - 1 - A.icig, cg_len(B.cig[B.k]) - 1 - B.icig);
        run = min(run, min((int)ha.l_qseq - 1 - A.iseq, (int)hb.l_qseq - 1 - B.iseq));
        // (after a deletion catch-up the two cursors may sit on different positions: no run then)
        if (run < 0 || A.iref + apos != B.iref + bpos) run = 0;
        overlap_run(a_q, b_q, a_seq, b_seq, A.iseq, B.iseq, run, amul, bmul, mode);
        A.iseq += run; A.icig += run; A.iref += run;
        B.iseq += run; B.icig += run; B.iref += run;
        iref += run;
    }
}

// ---------------------------------------------------------------------------------------------
// tile read ranges
// ---------------------------------------------------------------------------------------------
__global__ void k_tile_ranges(PlpParams P) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n_tiles * P.n_files) return;
    int tile = t / P.n_files, f = t % P.n_files;
    int t0 = P.t_beg + tile * P.W;
    int t1 = min(t0 + P.W, P.t_end);
    long long fo = P.file_off[f], fe = P.file_off[f + 1];
    long long lo = fo, hi = fe;
    while (lo < hi) {  // first read whose running max end exceeds t0
        long long mid = (lo + hi) >> 1;
        if (P.maxend[mid] > t0) hi = mid; else lo = mid + 1;
    }
    long long first = lo;
    lo = first; hi = fe;
    while (lo < hi) {  // first read starting at or after t1
        long long mid = (lo + hi) >> 1;
        if (P.hdr[mid].pos >= t1) hi = mid; else lo = mid + 1;
    }
    P.tile_ranges[t] = make_int2((int)first, (int)lo);
}

// ---------------------------------------------------------------------------------------------
// statistics (device restatements; exactness notes in DESIGN.md)
// ---------------------------------------------------------------------------------------------
__device__ double d_kf_erfc(double x) {  // htslib kfunc.c kf_erfc (AS66)
    const double p0 = 220.2068679123761, p1 = 221.2135961699311, p2 = 112.0792914978709, p3 = 33.912866078383,
                 p4 = 6.37396220353165, p5 = .7003830644436881, p6 = .03526249659989109;
    const double q0 = 440.4137358247522, q1 = 793.8265125199484, q2 = 637.3336333788311, q3 = 296.5642487796737,
                 q4 = 86.78073220294608, q5 = 16.06417757920695, q6 = 1.755667163182642, q7 = .08838834764831844;
    double z = fabs(x) * 1.41421356237309504880;
    if (z > 37.) return x > 0. ? 0. : 2.;
    double e = exp(z * z * -.5), r;
    if (z < 10. / 1.41421356237309504880)
        r = e * ((((((p6 * z + p5) * z + p4) * z + p3) * z + p2) * z + p1) * z + p0) /
            (((((((q7 * z + q6) * z + q5) * z + q4) * z + q3) * z + q2) * z + q1) * z + q0);
    else
        r = e / 2.506628274631001 / (z + 1. / (z + 2. / (z + 3. / (z + 4. / (z + .65)))));
    return x > 0. ? 2. * r : 2. * (1. - r);
}

// 100-bin read-position histogram inside a second-pass slot: 32-bit bins, or two 16-bit bins per word
struct HistView {
    const uint32_t* w;
    int base, narrow;
    __device__ __forceinline__ int operator[](int i) const {
        int b = base + i;
        return narrow ? (int)((w[b >> 1] >> ((b & 1) << 4)) & 0xffffu) : (int)w[b];
    }
};

__device__ double d_calc_vdb(const HistView pos) {  // src/ext/bcftools/bcftools-ext.c:51-111
    const float prm[15][3] = {{3, 0.079f, 18},    {4, 0.09f, 19.8f},  {5, 0.1f, 20.5f},   {6, 0.11f, 21.5f},
                              {7, 0.125f, 21.6f}, {8, 0.135f, 22},    {9, 0.14f, 22.2f},  {10, 0.153f, 22.3f},
                              {15, 0.19f, 22.8f}, {20, 0.22f, 23.2f}, {30, 0.26f, 23.4f}, {40, 0.29f, 23.5f},
                              {50, 0.35f, 23.65f}, {100, 0.5f, 23.7f}, {200, 0.7f, 23.7f}};
    int i, dp = 0;
    float mean_pos = 0, mean_diff = 0;
    for (i = 0; i < RAER_NHIST; ++i) {
        int v = (int)pos[i];
        if (!v) continue;
        dp += v;
        mean_pos += (float)(v * i);
    }
    if (dp < 2) return CUDART_INF;
    mean_pos /= (float)dp;
    for (i = 0; i < RAER_NHIST; ++i) {
        int v = (int)pos[i];
        if (!v) continue;
        // float += int * fabs(float): evaluated in double, rounded back to float
        mean_diff = (float)((double)mean_diff + (double)v * fabs((double)((float)i - mean_pos)));
    }
    mean_diff /= (float)dp;
    int ipos = (int)mean_diff;
    if (dp == 2) return (double)((2 * 100 - 2 * (ipos + 1) - 1) * (ipos + 1) / 99) / (100 * 0.5);
    if (dp >= 200) i = 15;
    else
        for (i = 0; i < 15; ++i)
            if (prm[i][0] >= (float)dp) break;
    float pshift, pscale;
    if (i == 15) { pscale = prm[14][1]; pshift = prm[14][2]; }
    else if (i > 0 && prm[i][0] != (float)dp) {
        pscale = (float)((double)(prm[i - 1][1] + prm[i][1]) * 0.5);
        pshift = (float)((double)(prm[i - 1][2] + prm[i][2]) * 0.5);
    } else { pscale = prm[i][1]; pshift = prm[i][2]; }
    return 0.5 * d_kf_erfc((double)(-(mean_diff - pshift) * pscale));
}

__device__ double d_calc_mwu_z(const HistView a, const HistView b) {  // bcftools-ext.c:144-212, do_Z = 1
    int i;
    long long t = 0;
    for (i = 0; i < RAER_NHIST; ++i)
        if (b[i]) break;
    int b_empty = (i == RAER_NHIST);
    int e = 0, l = 0, na = 0, nb = 0;
    if (b_empty) {
        for (i = RAER_NHIST - 1; i >= 0; --i) {
            int ai = (int)a[i];
            na += ai;
            t += (ai * ai - 1) * ai;
        }
    } else {
        for (i = RAER_NHIST - 1; i >= 0; --i) {
            int ai = (int)a[i], bi = (int)b[i];
            e += ai * bi;
            l += ai * nb;
            na += ai;
            nb += bi;
            int p = ai + bi;
            t += (p * p - 1) * p;
        }
    }
    if (!na || !nb) return CUDART_INF;
    double U = l + e * 0.5;
    double m = na * nb / 2.0;
    double var2 = (na * nb) / 12.0 * ((na + nb + 1) - t / (double)((na + nb) * (na + nb - 1)));
    if (var2 <= 0) return 0;
    return (U - m) / sqrt(var2);
}

__device__ double d_calc_sor(int fwd_ref, int rev_ref, int fwd_alt, int rev_alt) {  // src/plp_utils.c:375-387
    double t00 = (double)fwd_ref + 1, t01 = (double)rev_ref + 1, t11 = (double)fwd_alt + 1, t10 = (double)rev_alt + 1;
    double ratio = (t00 / t01) * (t11 / t10) + (t01 / t00) * (t10 / t11);
    double rr = fmin(t00, t01) / fmax(t00, t01);
    double ar = fmin(t10, t11) / fmax(t10, t11);
    return log(ratio) + log(rr) - log(ar);
}

// ---------------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------------
struct TileSmem {
    uint32_t* cnt;        // counters, later second-pass slots
    uint32_t* covered;    // W/32 words
    uint32_t* fbits;      // W/32 + 4 words: flagged positions of the current second-pass round, one pad word in front
    int* flag_rowp;       // W
    int* flag_rowm;       // W
    uint16_t* flag_pidx;  // W
    uint16_t* slot_of;    // W
    uint8_t* dec;         // W
    uint32_t* cigs;       // per warp: staged read records + work prefix sums (raer_walk.cuh, ST_STRIDE words)
    int hist_words;       // second pass: 400 (32-bit bins) or 200 (packed 16-bit bins), chosen per tile
    int slot_words;       // hist_words + 4 tallies + RAER_SLOT_FS words per (file, strand)
};

__device__ __forceinline__ bool pos_eligible(const PlpParams& P, int p) {
    if (p < P.reg_beg || p >= P.reg_end) return false;
    if (P.site_mask) {
        int b = p - P.mask_beg;
        if (b < 0 || b >= P.mask_bits) return false;
        return (__ldg(P.site_mask + (b >> 5)) >> (b & 31)) & 1u;
    }
    return true;
}

// src/plp_utils.c:106-133 on the raw FASTA bytes
__device__ bool d_simple_repeat(const char* ref, int ref_len, int pos, int nmer) {
    int n_pos = nmer * 2 - 1;
    if (n_pos < 1 || nmer < 1 || ref_len < nmer || pos < 0) return false;
    int start = pos - nmer + 1;
    if (start < 0) { n_pos += start; start = 0; }
    if (n_pos + start > ref_len) n_pos = ref_len - start;
    int run = 0;
    char prev = 0;
    for (int i = 0; i < n_pos; ++i) {
        char c = ref[start + i];
        run = (i > 0 && c == prev) ? run + 1 : 1;
        prev = c;
        if (run >= nmer) return true;
    }
    return false;
}

struct SiteCounts {  // per (file, strand)
    int c[RAER_NCLS];
};

// Site decision for one position; mirrors process_one_site + store_counts. Returns bit0 write '+',
// bit1 write '-', bit2 '+' has a variant in some file, bit3 '-' has a variant.
// Counter layouts behind one accessor: get() returns the six class counts A C G T other X of (file, strand)
// at tile position pidx; pc is the class of the reference base on the plus strand.
struct CntPlain {  // k_pileup_tiles: 32-bit words [file][strand][class][W], bank-skewed
    const uint32_t* cnt;
    int W;
    __device__ __forceinline__ void get(int f, int s, int pidx, int pc, int c[RAER_NCLS]) const {
        const int WP = WPAD(W);
        const uint32_t* base = cnt + (size_t)((f * 2 + s) * RAER_NCLS) * WP + CIDX(pidx, W);
#pragma unroll
        for (int k = 0; k < RAER_NCLS; ++k) c[k] = (int)base[k * WP];
    }
};
struct CntPacked {  // k_pileup_fast: [file][strand]{coverage, A|C<<16, G|T<<16, other|X<<16}[WS]; matches are implicit
    const uint32_t* cnt;
    int WS;
    __device__ __forceinline__ void get(int f, int s, int pidx, int pc, int c[RAER_NCLS]) const {
        const uint32_t* base = cnt + (size_t)((f * 2 + s) * 4) * WS + pidx;
        const uint32_t cov = base[0], ac = base[WS], gt = base[2 * WS], ox = base[3 * WS];
        c[0] = (int)(ac & 0xffffu); c[1] = (int)(ac >> 16);
        c[2] = (int)(gt & 0xffffu); c[3] = (int)(gt >> 16);
        c[4] = (int)(ox & 0xffffu); c[5] = (int)(ox >> 16);
        // bases that equal the reference never touched a class counter: coverage minus everything explicit
        const int match = (int)cov - (c[0] + c[1] + c[2] + c[3] + c[4] + c[5]);
        const int rc = pc < 4 ? (s == 0 ? pc : 3 - pc) : 4;
#pragma unroll
        for (int k = 0; k < 5; ++k) c[k] += (k == rc) ? match : 0;
    }
};

template <class CNT>
__device__ int decide_site(const PlpParams& P, const CNT cnt, const uint32_t* covered, int pidx, int p) {
    if (!pos_eligible(P, p)) return 0;
    int r = (P.ref && p < P.ref_len) ? (unsigned char)P.ref[p] : 'N';
    int pref = d_upper(r);
    if (pref == 'N') return 0;
    if (P.nmer > 0 && d_simple_repeat(P.ref, P.ref_len, p, P.nmer)) return 0;
    if (P.min_depth <= 0 && !((covered[pidx >> 5] >> (pidx & 31)) & 1u)) return 0;  // no pileup column here
    int pc = cls_of_ref(pref);
    int write_s[2] = {0, 0}, has_v[2] = {0, 0}, is_ma[2] = {0, 0}, any_var[2] = {0, 0};
    int three[2] = {0, 0}, raw_ok = 0;
    for (int f = 0; f < P.n_files; ++f) {
        int raw = 0;
        for (int s = 0; s < 2; ++s) {
            int cc[RAER_NCLS];
            cnt.get(f, s, pidx, pc, cc);
            const int co = cc[4];
            int total = cc[0] + cc[1] + cc[2] + cc[3] + co;
            int rc = s == 0 ? pc : (pc < 4 ? 3 - pc : 4);
            int nr = rc < 4 ? cc[rc] : 0;
            int nv = total - nr;
            raw += total + cc[5];
            // three distinct variant keys and a further put resize the file's khash (src/plp.c:487, khash.h)
            // whether or not the site is written: such sites take the second pass for the put order
            int keys = co > 0;
            for (int b = 0; b < 4; ++b) keys += (b != rc && cc[b] > 0);
            if (keys >= 3 && nv > 3) three[s] = 1;
            if (total >= P.min_depth && nv >= P.min_var_reads) write_s[s] = 1;
            if (nv > 0) any_var[s] = 1;
            int vs = 0;
            for (int b = 0; b < 4; ++b) {
                if (b == rc || cc[b] == 0) continue;
                if (P.min_af > 0 && (double)cc[b] / (double)total < P.min_af) continue;
                ++vs;
            }
            if (co > 0 && !(P.min_af > 0 && (double)co / (double)total < P.min_af)) ++vs;
            if (vs > 0) {
                if (P.okv[f]) has_v[s] = 1;
                if (vs > 1) is_ma[s] = 1;
            }
        }
        // the reference counts a site when one file's column holds min_depth entries (src/plp.c:697-704); the
        // counted bases of the file are a lower bound of its column
        if (raw >= P.min_depth) raw_ok = 1;
    }
    int out = 0;
    for (int s = 0; s < 2; ++s) {
        int w = write_s[s];
        if (P.any_okv) w = w && has_v[s];
        if (!P.report_multiallelic) w = w && !is_ma[s];
        if (w) out |= 1 << s;
        if (any_var[s]) out |= 4 << s;
        if (!w && three[s] && raw_ok && P.want_stats && !P.aei_mode) out |= 16 << s;
    }
    return out;
}

template <class CNT>
__device__ void write_row(const PlpParams& P, const CNT cnt, int pidx, int p, int s, long long row) {
    int pref = d_upper((unsigned char)P.ref[p]);
    int pc = cls_of_ref(pref);
    int rc = s == 0 ? pc : (pc < 4 ? 3 - pc : 4);
    int refc = s == 0 ? pref : d_comp_char(pref);
    P.row_pos[row] = p;
    P.row_meta[row] = (unsigned)s | ((unsigned)refc << 8);
    P.row_stats[row * 3 + 0] = 0.0;
    P.row_stats[row * 3 + 1] = 0.0;
    P.row_stats[row * 3 + 2] = 0.0;
    for (int f = 0; f < P.n_files; ++f) {
        int cc[RAER_NCLS];
        cnt.get(f, s, pidx, pc, cc);
        const int co = cc[4], cx = cc[5];
        int total = cc[0] + cc[1] + cc[2] + cc[3] + co;
        int nr = rc < 4 ? cc[rc] : 0;
        unsigned alt = 0;
        for (int b = 0; b < 4; ++b) {
            if (b == rc || cc[b] == 0) continue;
            if (P.min_af > 0 && (double)cc[b] / (double)total < P.min_af) continue;
            alt |= 1u << b;
        }
        if (co > 0 && !(P.min_af > 0 && (double)co / (double)total < P.min_af)) alt |= RAER_ALT_OTHER;
        int4* o = reinterpret_cast<int4*>(P.row_counts + (row * P.n_files + f) * 8);
        o[0] = make_int4(nr, total - nr, cc[0], cc[3]);  // nRef nAlt nA nT
        o[1] = make_int4(cc[1], cc[2], co, cx);          // nC nG nN nX
        P.row_alt[row * P.n_files + f] = alt;
    }
}

// Second-pass accumulation of one counted base at a flagged site: read-position histogram, strand
// tallies, first-seen order of the variant keys.
__device__ __forceinline__ void pass2_accumulate_core(const PlpParams& P, const TileSmem& S, const ReadCtx& c, int f, int s,
                                                      int ordinal, int slot_index, int q, int code, int isref,
                                                      uint32_t rcp = 0) {
    const int HW = S.hist_words;  // 400: 32-bit bins; 200: two 16-bit bins per word (tile depth <= 65535)
    uint32_t* slot = S.cnt + (size_t)slot_index * S.slot_words;
    int cls = cls_of_code(code);
    // strand-bias tally on the uncomplemented base (src/plp.c:762-774)
    atomicAdd(&slot[HW + (isref ? 0 : 2) + c.rev], 1u);  // ref_fwd ref_rev alt_fwd alt_rev
    // get_relative_position (src/plp_utils.c:401-412)
    int tail = c.l_qseq - c.qe;
    int alen = c.l_qseq - c.qs - tail;
    // (int)((double)a / b * 99): the two roundings are ~1e-14 away from 99a/b, which is at least 1/b away from
    // an integer unless it is one, so the integer quotient is exact except when b divides 99a (fp64 then).
    const int ra = q + 1 - c.qs, rb = alen + 1;
    int rpos;
    if (ra > 0 && rb > 0 && ra < (1 << 24)) {
        // rcp = ceil(2^32 / rb), staged once per read: exact quotient while 99 * ra * rb < 2^32
        if (rcp && ra < 4096 && rb < 4096) rpos = (int)__umulhi((uint32_t)(99 * ra), rcp);
        else rpos = (99 * ra) / rb;
        if (rpos * rb == 99 * ra) rpos = (int)((double)ra / (double)rb * 99.0);
    } else {
        rpos = (int)((double)ra / (double)rb * 99.0);
    }
    if (rpos < 0 || rpos >= RAER_NHIST) { atomicExch(P.err_flag, RAER_ERR); return; }
    const int bin = (s * 2 + (isref ? 0 : 1)) * RAER_NHIST + rpos;
    if (HW == 400) atomicAdd(&slot[bin], 1u);
    else atomicAdd(&slot[bin >> 1], 1u << ((bin & 1) << 4));
    if (!isref) {
        int cb = (c.invert && cls < 4) ? 3 - cls : cls;
        uint32_t* fs = slot + HW + 4 + (f * 2 + s) * RAER_SLOT_FS;
        if (cb < 4) {
            atomicMin(&fs[cb], (unsigned)ordinal);
        } else {
            // IUPAC key: complement of an nt16 code is its 4-bit reversal
            unsigned oc = c.invert ? (__brev((unsigned)code) >> 28) : (unsigned)code;
            atomicMin(&fs[5], (unsigned)ordinal);
            atomicMin(&fs[6], oc);
            atomicMax(&fs[7], oc);
        }
        atomicMax(&fs[4], (unsigned)ordinal + 1u);  // last variant read (ordinal + 1; 0 = none)
    }
}

__device__ __forceinline__ void pass2_accumulate_slot(const PlpParams& P, const TileSmem& S, const ReadCtx& c, int f, int s,
                                                      int ordinal, int p, int slot_index, int q, int code) {
    int pref = (p < P.ref_len) ? d_upper((unsigned char)P.ref[p]) : 'N';
    int letter = "=ACMGRSVTWYHKDBN"[code];
    pass2_accumulate_core(P, S, c, f, s, ordinal, slot_index, q, code, letter == pref);
}

__device__ __forceinline__ void pass2_accumulate(const PlpParams& P, const TileSmem& S, const ReadCtx& c, int f, int s,
                                                 int ordinal, int p, int pidx, int q, int code) {
    pass2_accumulate_slot(P, S, c, f, s, ordinal, p, (int)S.slot_of[pidx], q, code);
}

// Per-base work. PASS2 == false: phase 1 (counters). PASS2 == true: phase 3 (histograms of flagged sites).
template <bool PASS2>
__device__ __forceinline__ void process_base(const PlpParams& P, const TileSmem& S, const ReadCtx& c, int f, int s,
                                             int mapq_fail, bool mm_on, int ordinal, int p, int pidx, int q, int code,
                                             int bq) {
    int v = base_verdict(P, c, mapq_fail, q, bq);
    if (v == 2) return;
    int cls = cls_of_code(code);
    int pref = 0;
    if (PASS2 || (mm_on && v == 0)) pref = (p < P.ref_len) ? d_upper((unsigned char)P.ref[p]) : 'N';
    if (v == 0 && mm_on) {
        // src/plp.c:750: inverted reads always consult the MD verdict, others only on a mismatch
        int letter = "=ACMGRSVTWYHKDBN"[code];
        if (c.invert || letter != pref) {
            if (c.bits & RAER_RB_MMERR) { atomicExch(P.err_flag, RAER_ERR_MD); return; }
            if (c.bits & RAER_RB_MMFAIL) return;
        }
    }
    if (!PASS2) {
        int k2 = v == 1 ? RAER_CLS_X : ((c.invert && cls < 4) ? 3 - cls : cls);
        atomicAdd(&S.cnt[(size_t)((f * 2 + s) * RAER_NCLS + k2) * WPAD(P.W) + CIDX(pidx, P.W)], 1u);
    } else {
        if (v != 0) return;
        pass2_accumulate(P, S, c, f, s, ordinal, p, pidx, q, code);
    }
}

// ---- item-parallel walker ----------------------------------------------------------------------
// A warp fetches 32 consecutive read headers (one per lane, coalesced), the owning lane derives every
// per-read quantity (clips, fractional trim, filter bits) and stages a 48-byte record in the warp's
// shared-memory scratch. The reads are then cut into items of 8 consecutive QUERY bases; lanes take
// items from the warp-wide item list, so there is no warp-uniform per-read work and short reads do not
// leave lanes idle. An item costs one 32-bit sequence load and one 64-bit quality load.
#define ST_WORDS 13  // 12 used + 1: an odd stride keeps lanes reading the same field of different records off one bank
enum { ST_POS = 0, ST_LEN, ST_SQ, ST_FLAGS, ST_QSQE, ST_FTRIM, ST_CIGOFF, ST_ORD, ST_CIG0 };
#define STF_MAPQ_FAIL 0x100u
#define STF_FTRIM_ALL 0x200u

__device__ __forceinline__ void stage_ctx(const PlpParams& P, const uint32_t* rec, ReadCtx& c) {
    c.pos = (int)rec[ST_POS];
    c.l_qseq = (int)(rec[ST_LEN] & 0xffff);
    c.n_cigar = (int)(rec[ST_LEN] >> 16);
    unsigned fl = rec[ST_FLAGS];
    c.bits = (int)(fl & 0xff);
    c.invert = c.bits & RAER_RB_INVERT;
    c.rev = (fl >> 16) & 1;
    c.ftrim_all = (fl & STF_FTRIM_ALL) != 0;
    c.qs = (int)(rec[ST_QSQE] & 0xffff);
    c.qe = (int)(rec[ST_QSQE] >> 16);
    c.d5f = (int)(rec[ST_FTRIM] & 0xffff);
    c.d3f = (int)(rec[ST_FTRIM] >> 16);
    c.cig = c.n_cigar <= 4 ? rec + ST_CIG0 : P.cigar + rec[ST_CIGOFF];
    c.seq = P.seq + rec[ST_SQ];
    c.qual = P.qual + 2ull * rec[ST_SQ];
    c.end = 0; c.flag = 0; c.mapq = 0;
}

#include "raer_walk.cuh"  // scan_reads<PASS2, FAST>: the item-parallel walker


__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* scratch) {
    // RAER_BLOCK threads, warp shuffle scan + smem across warps
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(FULL, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) scratch[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int nw = blockDim.x >> 5;
        int wv = lane < nw ? scratch[lane] : 0;
        int z = wv;
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(FULL, z, o);
            if (lane >= o) z += y;
        }
        if (lane < nw) scratch[lane] = z - wv;
        if (lane == nw - 1) scratch[32] = z;
    }
    __syncthreads();
    int res = x - v + scratch[wid];
    *total = scratch[32];
    __syncthreads();
    return res;
}

#include "raer_fast.cuh"  // k_pileup_fast: item-parallel tile kernel, every filter
#include "raer_lane.cuh"  // k_pileup_lane: lane per read, TMA-staged chunks (mapq / base-quality filters only)

__global__ void __launch_bounds__(RAER_BLOCK, 1) k_pileup_tiles(PlpParams P) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int s_scan[40];
    __shared__ int s_tile, s_nflag, s_rowbase, s_deep;
    const int W = P.W;
    TileSmem S;
    S.cnt = smem;
    S.covered = S.cnt + P.cnt_words;
    S.fbits = S.covered + (W >> 5);
    S.flag_rowp = (int*)(S.fbits + (W >> 5) + 4);
    S.flag_rowm = S.flag_rowp + W;
    S.flag_pidx = (uint16_t*)(S.flag_rowm + W);
    S.slot_of = S.flag_pidx + W;
    S.dec = (uint8_t*)(S.slot_of + W);
    S.cigs = (uint32_t*)(S.dec + W);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int cnt_used = P.n_files * 2 * RAER_NCLS * WPAD(W);

    while (true) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(P.tile_counter, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= P.n_tiles) break;
        const int t0 = P.t_beg + tile * W;
        const int t1 = min(t0 + W, P.t_end);
        for (int i = threadIdx.x; i < cnt_used; i += blockDim.x) S.cnt[i] = 0;
        for (int i = threadIdx.x; i < (W >> 5); i += blockDim.x) S.covered[i] = 0;
        if (threadIdx.x == 0) { s_nflag = 0; s_deep = 0; }
        __syncthreads();

        // ---- phase 1: count
        for (int f = 0; f < P.n_files; ++f) {
            int2 rg = P.tile_ranges[tile * P.n_files + f];
            if (P.fast) scan_reads<false, true>(P, S, f, rg, t0, t1, t0, t1, lane, wid, nwarp);
            else scan_reads<false, false>(P, S, f, rg, t0, t1, t0, t1, lane, wid, nwarp);
        }
        __syncthreads();

        // ---- phase 2: decide, reserve rows, write them in (pos, strand) order
        int my_rows = 0;
        for (int pidx = threadIdx.x; pidx < W; pidx += blockDim.x) {
            int d = 0;
            if (t0 + pidx < t1) d = decide_site(P, CntPlain{S.cnt, W}, S.covered, pidx, t0 + pidx);
            S.dec[pidx] = (uint8_t)d;
            my_rows += (d & 1) + ((d >> 1) & 1) + ((d & 48) ? 0x10000 : 0);
        }
        int tile_tot;
        block_exclusive_scan(my_rows, &tile_tot, s_scan);
        const int tile_rows_total = tile_tot & 0xffff;
        const bool tile_grow = (tile_tot >> 16) != 0;
        if (threadIdx.x == 0) {
            long long base = (long long)atomicAdd(P.row_counter, (unsigned long long)tile_rows_total);
            if (base + tile_rows_total > P.row_cap) { atomicCAS(P.err_flag, 0, RAER_RETRY_ROWCAP); base = -1; }  // row_counter keeps the total: the host re-runs with that many
            s_rowbase = (int)base;
            P.tile_rows[tile] = make_int2((int)base, base < 0 ? 0 : tile_rows_total);
        }
        __syncthreads();
        const long long rowbase = s_rowbase;
        if (rowbase >= 0 && (tile_rows_total > 0 || tile_grow)) {
            int carried = 0, fcarried = 0;
            for (int chunk = 0; chunk < W; chunk += blockDim.x) {
                int pidx = chunk + threadIdx.x;
                int d = pidx < W ? S.dec[pidx] : 0;
                int n = (d & 1) + ((d >> 1) & 1);
                bool fl = P.want_stats && (((d & 1) && (d & 4)) || ((d & 2) && (d & 8)) || (d & 48));
                int tot, ftot;
                int off = block_exclusive_scan(n, &tot, s_scan) + carried;
                int slot = block_exclusive_scan(fl ? 1 : 0, &ftot, s_scan) + fcarried;  // position-sorted flag list
                carried += tot;
                fcarried += ftot;
                if (n || fl) {
                    long long row = rowbase + off;
                    int rp = -1, rm = -1;
                    if (d & 1) { write_row(P, CntPlain{S.cnt, W}, pidx, t0 + pidx, 0, row); rp = (int)row; ++row; }
                    if (d & 2) { write_row(P, CntPlain{S.cnt, W}, pidx, t0 + pidx, 1, row); rm = (int)row; }
                    if (fl) {
                        S.flag_pidx[slot] = (uint16_t)pidx;
                        S.flag_rowp[slot] = ((d & 1) && (d & 4)) ? rp : -1;
                        S.flag_rowm[slot] = ((d & 2) && (d & 8)) ? rm : -1;
                        // 16-bit histogram bins are safe while no flagged site is deeper than 65535 counted bases
                        unsigned depth = 0;
                        for (int k = 0; k < P.n_files * 2 * RAER_NCLS; ++k) depth += S.cnt[(size_t)k * WPAD(W) + CIDX(pidx, W)];
                        if (depth > 65535u) s_deep = 1;
                    }
                }
            }
            if (threadIdx.x == 0) s_nflag = fcarried;
        }
        __syncthreads();

        // ---- phases 3+4: second pass for the variant rows, `spr` sites per round
        const int nflag = s_nflag;
        S.hist_words = s_deep ? 4 * RAER_NHIST : 2 * RAER_NHIST;
        S.slot_words = S.hist_words + 4 + 2 * RAER_SLOT_FS * P.n_files;
        const int HW = S.hist_words, SW = S.slot_words;
        const int spr = min(P.cnt_words / SW, 4096);
        for (int round0 = 0; round0 < nflag; round0 += spr) {
            const int ns = min(spr, nflag - round0);
            for (int i = threadIdx.x; i < W; i += blockDim.x) S.slot_of[i] = 0xffff;
            for (int i = threadIdx.x; i < ns * SW; i += blockDim.x) {
                int wsl = i % SW;
                // first-seen ordinals (and the minimum IUPAC code) start at +inf, everything else at 0
                int fw = (wsl - (HW + 4)) % RAER_SLOT_FS;
                bool is_fs = wsl >= HW + 4 && (fw < 4 || fw == 5 || fw == 6);
                S.cnt[i] = is_fs ? 0xffffffffu : 0u;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < (W >> 5) + 4; i += blockDim.x) S.fbits[i] = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < ns; i += blockDim.x) {
                int fp = S.flag_pidx[round0 + i];
                S.slot_of[fp] = (uint16_t)i;
                atomicOr(&S.fbits[1 + (fp >> 5)], 1u << (fp & 31));
            }
            __syncthreads();
            // the flag list is position sorted: only reads overlapping this round's span need a second look
            const int span_lo = t0 + S.flag_pidx[round0];
            const int span_hi = t0 + S.flag_pidx[round0 + ns - 1] + 1;
            for (int f = 0; f < P.n_files; ++f) {
                int2 rg = P.tile_ranges[tile * P.n_files + f];
                if (P.fast) scan_reads<true, true>(P, S, f, rg, t0, t1, span_lo, span_hi, lane, wid, nwarp, S.flag_pidx + round0, ns);
                else scan_reads<true, false>(P, S, f, rg, t0, t1, span_lo, span_hi, lane, wid, nwarp, S.flag_pidx + round0, ns);
            }
            __syncthreads();
            for (int i = threadIdx.x; i < ns; i += blockDim.x)
                slot_stats(P, S.cnt + (size_t)i * SW, HW, S.flag_rowp[round0 + i], S.flag_rowm[round0 + i],
                           S.dec[S.flag_pidx[round0 + i]] >> 4, t0 + (int)S.flag_pidx[round0 + i]);
            __syncthreads();
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers (called from raer_runtime.cpp)
// ---------------------------------------------------------------------------------------------
// scratch: n_pairs + 1 ints of device memory for the compacted list of CIGAR-walking pairs, or nullptr (one kernel
// handles every pair then)
extern "C" int raer_launch_overlap(const raer_read_hdr* hdr, const uint32_t* cigar, const uint8_t* seq, uint8_t* qual,
                                   const int32_t* pair_b, long long n_pairs, int mode, cudaStream_t st, int* scratch) {
    if (n_pairs <= 0) return 0;
    const int block = 128;
    const long long grid = (n_pairs + block - 1) / block;
    if (!scratch) {
        k_mate_overlap<<<(unsigned)grid, block, 0, st>>>(hdr, cigar, seq, qual, pair_b, n_pairs, mode, nullptr, nullptr);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    cudaMemsetAsync(scratch, 0, sizeof(int), st);
    k_mate_overlap_simple<<<(unsigned)grid, block, 0, st>>>(hdr, cigar, seq, qual, pair_b, n_pairs, mode, scratch + 1, scratch);
    // the size of the list stays on the device: launch for the worst case, surplus threads leave at once
    k_mate_overlap<<<(unsigned)grid, block, 0, st>>>(hdr, cigar, seq, qual, pair_b, n_pairs, mode, scratch + 1, scratch);
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
}

extern "C" int raer_lane_first(PlpParams* P, int* first, cudaStream_t st) {
    const int n = (P->n_tiles + 1) * P->n_files;
    k_tile_first<<<(n + 127) / 128, 128, 0, st>>>(*P, first);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

extern "C" size_t raer_tile_smem_bytes(int n_files, int W, int S, int* cnt_words, int* slot_words) {
    int sw = 404 + 2 * RAER_SLOT_FS * n_files;
    int cw = n_files * 2 * RAER_NCLS * WPAD(W);
    if (S * sw > cw) cw = S * sw;
    *cnt_words = cw;
    *slot_words = sw;
    return (size_t)cw * 4 + ((W >> 5) * 2 + 4) * 4 + (size_t)W * (4 + 4 + 2 + 2 + 1) + (RAER_BLOCK / 32) * ST_STRIDE * 4 + 64;
}

extern "C" int raer_launch_tiles(PlpParams* P, int n_sm, cudaStream_t st, int* launches, cudaEvent_t e0, cudaEvent_t e1) {
    int nl = 0;
    if (P->n_tiles <= 0) { *launches = 0; return 0; }
    if (!P->use_fast) {
        int n = P->n_tiles * P->n_files;
        k_tile_ranges<<<(n + 255) / 256, 256, 0, st>>>(*P);
        ++nl;
    }
    if (P->use_lane) {
        const size_t lsmem = raer_lane_fixed_bytes(P->n_files, P->W) + (size_t)LN_WARPS * P->stage_bytes;
        static size_t lconfigured_dev[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        size_t& lconfigured = lconfigured_dev[dev & 63];
        if (lsmem > lconfigured) {
            if (cudaFuncSetAttribute(k_pileup_lane, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem) != cudaSuccess) return -1;
            lconfigured = lsmem;
        }
        int grid = n_sm;
        if (grid > P->n_tiles) grid = P->n_tiles;
        if (e0) cudaEventRecord(e0, st);
        k_pileup_lane<<<grid, LN_THREADS, lsmem, st>>>(*P);
        if (e1) cudaEventRecord(e1, st);
        ++nl;
        *launches = nl;
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    if (P->use_fast) {
        const size_t fsmem = raer_fast_smem_bytes(P->n_files, P->W);
        // the opt-in shared-memory size is a per-device attribute of the function
        static size_t fconfigured_dev[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        size_t& fconfigured = fconfigured_dev[dev & 63];
        if (fsmem > fconfigured) {
            if (cudaFuncSetAttribute(k_pileup_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem) != cudaSuccess ||
                cudaFuncSetAttribute(k_pileup_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem) != cudaSuccess)
                return -1;
            fconfigured = fsmem;
        }
        int grid = n_sm * 2;
        if (grid > P->n_tiles) grid = P->n_tiles;
        if (e0) cudaEventRecord(e0, st);
        if (P->fast) k_pileup_fast<false><<<grid, FB_THREADS, fsmem, st>>>(*P);
        else k_pileup_fast<true><<<grid, FB_THREADS, fsmem, st>>>(*P);
        if (e1) cudaEventRecord(e1, st);
        ++nl;
        *launches = nl;
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    size_t smem = raer_tile_smem_bytes(P->n_files, P->W, P->S, &P->cnt_words, &P->slot_words);
    static size_t configured_dev[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& configured = configured_dev[dev & 63];
    if (smem > configured) {
        if (cudaFuncSetAttribute(k_pileup_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -1;
        configured = smem;
    }
    int per_sm = smem > 110 * 1024 ? 1 : 2;
    int grid = n_sm * per_sm;
    if (grid > P->n_tiles) grid = P->n_tiles;
    if (e0) cudaEventRecord(e0, st);
    k_pileup_tiles<<<grid, RAER_BLOCK, smem, st>>>(*P);
    if (e1) cudaEventRecord(e1, st);
    ++nl;
    *launches = nl;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

#endif  // RAER_HELPERS_ONLY
