// raer_gingest.cu -- BGZF/BAM ingest on the device: what the reference gets from htslib's
// sam_read1 + readaln() (src/plp.c:32-78) + the stream-order parts of bam_plp_push / overlap_push, for one contig of
// one BAM file at a time, without the inflated bytes ever visiting the host:
//
//   host   reads the compressed byte range of the contig (BAI: first chunk begin .. last chunk end), walks the BGZF
//          block headers (18 bytes each) and turns the linear index of the contig into record-aligned start points
//   k_gi_inflate   one warp per BGZF block: raw DEFLATE into one contiguous buffer of inflated records (raer_ginflate.cuh)
//   k_gw_*         one thread per start point chains through block_size fields: record offsets, in file order
//   k_gp_pass      one thread per record: the read-level prefilter of readaln (flags, region / site bitmap, mapq,
//                  paired-not-proper, read_bqual), reference span and aligned bases from the CIGAR
//   k_gp_kept      bam_plp_push's "can still produce a column" rule against the previous pushed read; sizes of what the
//                  read will occupy in the batch; sort-order and max_depth guards
//   k_gp_write     struct-of-arrays read batch (raer_read_hdr, CIGAR words, 4-bit sequence, qualities) written in place
//   k_gp_pair      overlap_push: reads sharing a name meet through a sort of the name hashes; every name's records are
//                  replayed in file order against the one-entry-per-name table htslib keeps
// Exclusive scans between the stages keep everything in file (= coordinate) order. Anything this path does not cover
// (bulk UMI mode, the MD mismatch filter, reads longer than 65535, a contig whose depth could reach max_depth, hash
// collisions between different names) is reported as RAER_GI_FALLBACK and the caller streams that contig on the host.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "raer_dev.h"
#include "raer_host.h"
#include <zlib.h>

#include "raer_ginflate.cuh"
#include "raer_inflate.h"
#include "raer_radix.cuh"

using raer::set_error;

#define GCK(call)                                                                             \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            set_error("CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__);  \
            return RAER_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define RAER_GI_FALLBACK 1  // not an error: this contig goes through the host ingest
// what the kernels report through the `flags` word: 1 corrupt record layout, 2 read beyond the documented limits,
// 4 reads out of order, 8 depth may reach max_depth, 16 two read names with one hash, 32 a record chain did not land on
// the next start point, 64 a record of another contig inside the contig's byte range (chromosomes out of order),
// 256 a UMI the 2-bit code does not cover

// ------------------------------------------------------------------------------------- scans
#define SC_BLOCK 256
#define SC_ITEMS 16
template <bool MAX>
__global__ void k_sc_sum(const uint32_t* __restrict__ in, long long n, uint32_t* __restrict__ bsum) {
    __shared__ uint32_t s[SC_BLOCK / 32];
    const long long base = (long long)blockIdx.x * SC_BLOCK * SC_ITEMS + (long long)threadIdx.x * SC_ITEMS;
    uint32_t v = 0;
    for (int k = 0; k < SC_ITEMS; ++k)
        if (base + k < n) v = MAX ? max(v, in[base + k]) : v + in[base + k];
    v = MAX ? __reduce_max_sync(0xffffffffu, v) : __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SC_BLOCK / 32; ++w) t = MAX ? max(t, s[w]) : t + s[w];
        bsum[blockIdx.x] = t;
    }
}
// exclusive scan (sum) or running maximum of the block values in place, one CTA; *total = sum / maximum of all
template <bool MAX>
__global__ void k_sc_top(uint32_t* bsum, int nb, unsigned long long* total) {
    __shared__ uint32_t part[256];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 256) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nb ? bsum[i] : 0;
        part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const uint32_t t = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
            __syncthreads();
            part[threadIdx.x] = MAX ? max(part[threadIdx.x], t) : part[threadIdx.x] + t;
            __syncthreads();
        }
        const uint32_t inc = part[threadIdx.x];
        // exclusive value: everything before this block
        const uint32_t before = threadIdx.x ? part[threadIdx.x - 1] : 0;
        if (i < nb) bsum[i] = MAX ? max(carry, before) : carry + before;
        __syncthreads();
        if (threadIdx.x == 255) carry = MAX ? max(carry, inc) : carry + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}
// out[i] = exclusive sum / inclusive running maximum
template <bool MAX>
__global__ void k_sc_apply(const uint32_t* __restrict__ in, long long n, const uint32_t* __restrict__ bsum, uint32_t* __restrict__ out) {
    __shared__ uint32_t part[SC_BLOCK];
    const long long base = (long long)blockIdx.x * SC_BLOCK * SC_ITEMS + (long long)threadIdx.x * SC_ITEMS;
    uint32_t v = 0;
    for (int k = 0; k < SC_ITEMS; ++k)
        if (base + k < n) v = MAX ? max(v, in[base + k]) : v + in[base + k];
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SC_BLOCK; o <<= 1) {
        const uint32_t t = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] = MAX ? max(part[threadIdx.x], t) : part[threadIdx.x] + t;
        __syncthreads();
    }
    const uint32_t before_t = threadIdx.x ? part[threadIdx.x - 1] : 0;
    uint32_t run = MAX ? max(bsum[blockIdx.x], before_t) : bsum[blockIdx.x] + before_t;
    for (int k = 0; k < SC_ITEMS; ++k) {
        const long long i = base + k;
        if (i >= n) break;
        const uint32_t x = in[i];
        if (MAX) { run = max(run, x); out[i] = run; }
        else { out[i] = run; run += x; }
    }
}

// ------------------------------------------------------------------------------------- records
typedef unsigned long long recoff_t;  // offset of a record in the inflated buffer of its file range (may exceed 4 GB)
__device__ __forceinline__ uint32_t ld32u(const uint8_t* p) {  // unaligned little-endian load
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ uint32_t ld16u(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// start point k chains through the records that begin in [sp[k], sp[k + 1]); FILL writes their offsets
template <bool FILL>
__global__ void k_gw_walk(const uint8_t* __restrict__ ub, unsigned long long total, const unsigned long long* __restrict__ sp, int nsp,
                          uint32_t* cnt, const uint32_t* __restrict__ base, recoff_t* __restrict__ rec_off, int* flags) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nsp) return;
    unsigned long long off = sp[k];
    const unsigned long long end = sp[k + 1];
    uint32_t n = 0;
    const uint32_t b0 = FILL ? base[k] : 0;
    while (off < end) {
        if (off + 36 > total) { atomicOr(flags, 1); break; }  // a record header does not fit: corrupt or truncated
        const uint32_t bs = ld32u(ub + off);
        if (bs < 32 || off + 4 + bs > total) { atomicOr(flags, 1); break; }
        if (FILL) rec_off[b0 + n] = off;
        ++n;
        off += 4ull + bs;
    }
    if (off != end) atomicOr(flags, 32);  // the chain must land exactly on the next start point
    if (!FILL) cnt[k] = n;
}

struct GpConf {
    int tid;                 // contig of this pass (the file's own numbering)
    uint32_t keep0, keep1;   // scanBamFlag masks
    int min_global_mq;
    int rq_minq;
    float rq_pct;
    int has_rq;
    int libtype;
    int remove_overlaps;
    int min_overhang;
    int max_depth;
    // positions some interval of the call's region list covers, as a bitmap over the contig + the number of set bits in
    // front of every word (a read is kept if [pos, endpos] holds a set bit: inclusive end, src/plp.c:56-60)
    const uint32_t* site_bits;
    const uint32_t* site_pre;
    int site_len;
    // single cell (screadaln, src/sc-plp.c:449-511): the cell-barcode tag has to be on the whitelist (open-addressing table
    // of FNV-1a name hashes, strings compared on a hit), or every read of the file belongs to one column (Smart-seq2)
    int sc, has_cb_tag, has_umi_tag, fixed_cb;
    int n_mm_type, n_mm;               // mismatch filter (max_mismatch_type): the MD verdict is one bit per read
    int reg_on, reg_beg, reg_end;      // a coordinate range of the contig: only the reads that overlap [reg_beg, reg_end)
    uint8_t cb_tag[2], umi_tag[2];
    const unsigned long long* wl_key;  // 0 = empty slot
    const int32_t* wl_col;             // 1-based whitelist column
    const uint32_t* wl_off;            // offset of the barcode string in wl_pool (NUL terminated)
    const uint8_t* wl_pool;
    uint32_t wl_mask;
    int* umi_len;                      // length of the UMIs seen so far (0: none yet); all have to agree
};

// BAM auxiliary fields: the value of tag t0 t1 (pointer to its type byte) or nullptr; same walk as the host's aux_get
__device__ const uint8_t* gp_aux_find(const uint8_t* p, const uint8_t* end, uint8_t t0, uint8_t t1) {
    while (p + 3 <= end) {
        const bool hit = p[0] == t0 && p[1] == t1;
        const int t = p[2];
        const uint8_t* v = p + 2;
        p += 3;
        if (t == 'Z' || t == 'H') {
            while (p < end && *p) ++p;
            ++p;
        } else if (t == 'B') {
            if (p + 5 > end) return nullptr;
            const int e = p[0];
            const int sz = (e == 'c' || e == 'C' || e == 'A') ? 1 : (e == 's' || e == 'S') ? 2 : (e == 'i' || e == 'I' || e == 'f') ? 4 : e == 'd' ? 8 : 0;
            const uint32_t cnt = ld32u(p + 1);
            p += 5 + (size_t)sz * cnt;
        } else {
            const int sz = (t == 'c' || t == 'C' || t == 'A') ? 1 : (t == 's' || t == 'S') ? 2 : (t == 'i' || t == 'I' || t == 'f') ? 4 : t == 'd' ? 8 : 0;
            if (!sz) return nullptr;
            p += sz;
        }
        if (p > end) return nullptr;
        if (hit) return v;
    }
    return nullptr;
}
__device__ __forceinline__ unsigned long long gp_fnv(const uint8_t* s) {
    unsigned long long h = 1469598103934665603ull;
    for (; *s; ++s) { h ^= *s; h *= 1099511628211ull; }
    return h ? h : 1ull;
}
// whitelist column of the NUL-terminated barcode s, 0 if it is not listed
__device__ int gp_wl_lookup(const GpConf& C, const uint8_t* s) {
    const unsigned long long h = gp_fnv(s);
    for (uint32_t slot = (uint32_t)h & C.wl_mask;; slot = (slot + 1) & C.wl_mask) {
        const unsigned long long k = C.wl_key[slot];
        if (k == 0) return 0;
        if (k != h) continue;
        const uint8_t* w = C.wl_pool + C.wl_off[slot];
        int i = 0;
        while (w[i] && w[i] == s[i]) ++i;
        if (w[i] == s[i]) return C.wl_col[slot];
    }
}

// per record: does it pass readaln? pos / end / aligned bases. One thread per record.
__global__ void k_gp_pass(const uint8_t* __restrict__ ub, const recoff_t* __restrict__ rec_off, long long n_rec, GpConf C,
                          uint32_t* __restrict__ pass, int32_t* __restrict__ r_pos, int32_t* __restrict__ r_end,
                          unsigned long long* stats, int* flags, int32_t* __restrict__ rec_cb) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    const uint8_t* p = ub + rec_off[i] + 4;
    const int tid = (int)ld32u(p), pos = (int)ld32u(p + 4);
    const int l_qname = p[8], mapq = p[9];
    const int n_cigar = (int)ld16u(p + 12), flag = (int)ld16u(p + 14);
    const int l_qseq = (int)ld32u(p + 16);
    const uint32_t bs = ld32u(p - 4);
    uint32_t ok = 1;
    if (l_qseq < 0 || 32ull + l_qname + 4ull * n_cigar + ((unsigned long long)l_qseq + 1) / 2 + (unsigned long long)l_qseq > bs) {
        atomicOr(flags, 1);  // corrupt record layout
        pass[i] = 0; r_pos[i] = 0; r_end[i] = 0;
        if (C.sc) rec_cb[i] = 0;
        return;
    }
    // In a coordinate-sorted file the byte range of a contig (first record .. end of the last one) holds nothing else:
    // a record of another contig means the input is not sorted (htslib: "chromosomes out of order")
    if (tid != C.tid) { atomicOr(flags, 64); ok = 0; }
    if (flag & 4) ok = 0;
    int rlen = 0;
    unsigned nal = 0;
    const uint8_t* cg = p + 32 + l_qname;
    for (int k = 0; k < n_cigar; ++k) {
        const uint32_t w = ld32u(cg + 4 * k);
        const int op = (int)(w & 0xf);
        if ((0x3C1A7 >> (op << 1)) & 2) rlen += (int)(w >> 4);
        if (op == 0 || op == 7 || op == 8) nal += (w >> 4);
    }
    if (C.reg_on) {  // what the region iterator hands out (sam_itr_next): reads that start before the end and reach the begin
        const int span = ((flag & 4) || n_cigar == 0 || rlen == 0) ? 1 : rlen;
        if (!(pos < C.reg_end && pos + span > C.reg_beg)) ok = 0;
    }
    {   // unit of the bench metric: aligned bases of every mapped record, before filters; longest reference span
        const unsigned am = __activemask();
        const unsigned tot = __reduce_add_sync(am, ok ? nal : 0u);
        const unsigned mx = __reduce_max_sync(am, (unsigned)max(rlen, 1));
        if ((int)(threadIdx.x & 31) == __ffs(am) - 1) {
            if (tot) atomicAdd(&stats[0], (unsigned long long)tot);
            atomicMax(&stats[1], (unsigned long long)mx);
        }
    }
    if (C.sc) {  // src/sc-plp.c:464-477: reads without a listed cell barcode are skipped before anything else
        int col = 0;
        if (C.has_cb_tag) {
            if (ok) {
                const uint8_t* aux = cg + 4 * n_cigar + (l_qseq + 1) / 2 + l_qseq;
                const uint8_t* v = gp_aux_find(aux, p + bs, C.cb_tag[0], C.cb_tag[1]);
                if (v && (*v == 'Z' || *v == 'H')) col = gp_wl_lookup(C, v + 1);
                if (!col) ok = 0;
            }
        }
        if (C.fixed_cb) col = C.fixed_cb;
        rec_cb[i] = col;
    }
    const uint32_t test = (C.keep0 & ~(uint32_t)flag) | (C.keep1 & (uint32_t)flag);
    if (~test & 2047u) ok = 0;
    const int endpos = pos + ((n_cigar == 0 || rlen == 0) ? 1 : rlen);
    if (ok && C.site_bits) {
        // any listed position in [pos, endpos] (inclusive)?
        const int lo = max(pos, 0), hi = min(endpos, C.site_len - 1);
        bool hit = false;
        if (lo <= hi) {
            const int w0 = lo >> 5, w1 = hi >> 5;
            const uint32_t m0 = 0xffffffffu << (lo & 31), m1 = 0xffffffffu >> (31 - (hi & 31));
            if (w0 == w1) hit = (C.site_bits[w0] & m0 & m1) != 0;
            else hit = (C.site_bits[w0] & m0) || (C.site_bits[w1] & m1) || (C.site_pre[w1] - C.site_pre[w0 + 1]) != 0;
        }
        if (!hit) ok = 0;
    }
    if (mapq < C.min_global_mq) ok = 0;
    if ((flag & 1) && !(flag & 2)) ok = 0;
    if (ok && C.has_rq) {  // read_base_quality (src/plp_utils.c:310-328): double ratio against a float threshold
        if (!l_qseq) ok = 0;
        else {
            const uint8_t* q = cg + 4 * n_cigar + (l_qseq + 1) / 2;
            int lq = 0;
            for (int k = 0; k < l_qseq; ++k) lq += q[k] < C.rq_minq;
            const double frac = (double)lq / l_qseq;
            if (!(frac <= C.rq_pct)) ok = 0;
        }
    }
    if (ok && (l_qseq > 65535 || n_cigar > 65535)) { atomicOr(flags, 2); ok = 0; }  // documented limit: host path reports it
    pass[i] = ok;
    r_pos[i] = pos;
    r_end[i] = pos + rlen;
}

// positions of the passing reads in their own (compact) numbering: the "previous pushed read" of read P is P - 1
__global__ void k_gp_cpos(const uint32_t* __restrict__ pass, const uint32_t* __restrict__ pidx, const int32_t* __restrict__ r_pos,
                          const int32_t* __restrict__ r_end, long long n_rec, int32_t* __restrict__ cpos, int32_t* __restrict__ cend) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n_rec && pass[i]) { cpos[pidx[i]] = r_pos[i]; cend[pidx[i]] = r_end[i]; }
}

// bam_plp_push keeps a read only if it can still produce a column (end > position of the previous pushed read);
// sizes of the kept reads; guards
__global__ void k_gp_kept(const uint8_t* __restrict__ ub, const recoff_t* __restrict__ rec_off, const uint32_t* __restrict__ pass,
                          const uint32_t* __restrict__ pidx, const int32_t* __restrict__ r_pos, const int32_t* __restrict__ r_end,
                          const int32_t* __restrict__ cpos, long long n_rec, long long n_pass, GpConf C, int max_span,
                          uint32_t* __restrict__ kept, uint32_t* __restrict__ ncig, uint32_t* __restrict__ sqb,
                          uint32_t* __restrict__ elig, int32_t* __restrict__ prevpos, int* flags,
                          const uint8_t* __restrict__ dropped) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    uint32_t k = 0, nc = 0, sb = 0, el = 0;
    int pp = INT_MIN;
    if (pass[i]) {
        const long long P = pidx[i];
        const int pos = r_pos[i], end = r_end[i];
        pp = P > 0 ? cpos[P - 1] : INT_MIN;
        if (P > 0 && pos < pp) atomicOr(flags, 4);  // htslib: "the input is not sorted"
        k = (P == 0) || (end > pp);
        if (dropped && dropped[P]) k = 0;  // max_depth: decided by the replay of the stream (gi_replay_max_depth)
        // max_depth (bam_plp_push drops a read that starts where the previous one did once the buffer holds more than
        // max_depth reads): reads still in the buffer all started within max_span positions in front of this one
        if (!dropped && P > 0 && pos == pp) {
            long long lo = 0, hi = P;  // first passing read with position >= pos - max_span
            const int from = pos - max_span;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if (cpos[mid] < from) lo = mid + 1; else hi = mid;
            }
            if (P - lo + 1 > C.max_depth) atomicOr(flags, 8);  // cannot be ruled out here: the host replays the stream
        }
        if (k) {
            const uint8_t* p = ub + rec_off[i] + 4;
            const int flag = (int)ld16u(p + 14), l_qseq = (int)ld32u(p + 16);
            nc = ld16u(p + 12);
            sb = (uint32_t)((((l_qseq + 1) >> 1) + 3) & ~3);
            if (C.remove_overlaps) {  // overlap_push: who takes part
                const int mtid = (int)ld32u(p + 20), mpos = (int)ld32u(p + 24);
                const int isz0 = (int)ld32u(p + 28);
                const long long isz = isz0 < 0 ? -(long long)isz0 : isz0;
                el = !(flag & 8) && (flag & 2);
                if ((mtid >= 0 && C.tid != mtid) || (isz >= 2LL * l_qseq && mpos >= end)) el = 0;
            }
        }
    }
    kept[i] = k;
    ncig[i] = nc;
    sqb[i] = sb;
    elig[i] = el;
    prevpos[i] = pp;
}

__device__ __forceinline__ uint32_t x31_hash(const uint8_t* s) {  // khash __ac_X31_hash_string (htslib overlap hash)
    uint32_t h = (uint32_t)(int)(signed char)*s;
    if (h)
        for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)(int)(signed char)*s;
    return h;
}
__device__ __forceinline__ uint32_t wang_hash(uint32_t key) {
    key += ~(key << 15);
    key ^= (key >> 10);
    key += (key << 3);
    key ^= (key >> 6);
    key += ~(key << 11);
    key ^= (key >> 16);
    return key;
}

// Per kept read of a file (index K = arrival order among the kept reads): what the later stages need without going back
// to the record table. A contig is handed to the pileup kernels in position windows (each window is one read batch with
// 32-bit offsets); everything that depends on the order of the stream -- which reads are kept, who pairs with whom --
// is decided here, once, over the whole contig.
struct GpMeta {
    recoff_t* krec;     // record offset in the inflated buffer
    int32_t* kprev;     // position of the read pushed before it (overlap_push: when does an entry die)
    int32_t* kpos;
    int32_t* kend;
    uint32_t* kcoff;    // exclusive sums of CIGAR words / packed sequence bytes (one extra element: the totals)
    uint32_t* ksoff;
    int32_t* kcb;       // single cell: whitelist column, UMI code
    uint32_t* kumi;
    int32_t* mateK;     // K of the earlier mate this read resolves an overlap with, or -1
    int32_t* keepend;   // a read stays in the windows until this position: its own end or that of the mate paired with it
    uint8_t* keepA;     // bit 0: the earlier mate keeps its qualities (htslib >= 1.13 coin); bits 1, 2: MD verdict fail / inconsistent
};

// parse_mismatches (src/ext/htsbox/htsbox-ext.c:17-97) as one verdict per read, like the host's md_verdict: 0 pass,
// > 0 fail, -1 MD tag inconsistent with the CIGAR
__device__ int gp_md_verdict(const uint8_t* p, int n_types, int n_mis) {
    const int l_qname = p[8], n_cigar = (int)ld16u(p + 12), l_qseq = (int)ld32u(p + 16);
    const uint8_t* cg = p + 32 + l_qname;
    const uint8_t* sq = cg + 4 * n_cigar;
    const uint8_t* aux = sq + ((l_qseq + 1) >> 1) + l_qseq;
    const uint8_t* md = gp_aux_find(aux, p + ld32u(p - 4), 'M', 'D');
    if (!md || (*md != 'Z' && *md != 'H')) return 0;
    int x = 0;
    for (int k = 0; k < n_cigar; ++k) {
        const uint32_t w = ld32u(cg + 4 * k);
        const int op = (int)(w & 0xf);
        if (op == 0 || op == 7 || op == 8) x += (int)(w >> 4);
    }
    const uint8_t* q = md + 1;
    int y = 0, nm = 0, ntypes = 0;
    unsigned seen = 0;
    auto is_alpha = [](uint8_t c) { return (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z'); };
    while (*q >= '0' && *q <= '9') {
        long long v = 0;
        while (*q >= '0' && *q <= '9') { v = v * 10 + (*q - '0'); if (v > INT_MAX) v = INT_MAX; ++q; }
        y += (int)v;
        if (*q == 0) break;
        if (*q == '^') {
            ++q;
            while (is_alpha(*q)) ++q;
        } else {
            while (is_alpha(*q)) {
                if (y >= x) { y = -1; break; }
                const int code = y < l_qseq ? (sq[y >> 1] >> ((~y & 1) << 2)) & 0xf : 15;
                ++nm;
                if (!(seen & (1u << code))) { seen |= 1u << code; ++ntypes; }
                ++y;
                ++q;
            }
            if (y == -1) break;
        }
    }
    if (x != y) return -1;
    return (ntypes >= n_types && nm >= n_mis) ? (ntypes > 0 ? ntypes : 0) : 0;
}

__global__ void k_gp_meta(const uint8_t* __restrict__ ub, const recoff_t* __restrict__ rec_off, const uint32_t* __restrict__ kept,
                          const uint32_t* __restrict__ kidx, const uint32_t* __restrict__ coff, const uint32_t* __restrict__ soff,
                          const uint32_t* __restrict__ elig, const uint32_t* __restrict__ eidx, const int32_t* __restrict__ r_pos,
                          const int32_t* __restrict__ r_end, const int32_t* __restrict__ prevpos, long long n_rec, GpConf C, GpMeta M,
                          unsigned long long* __restrict__ pkey, uint32_t* __restrict__ pval, const int32_t* __restrict__ rec_cb,
                          int* flags) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_rec || !kept[i]) return;
    const uint8_t* p = ub + rec_off[i] + 4;
    const uint32_t K = kidx[i];
    M.krec[K] = rec_off[i];
    M.kprev[K] = prevpos[i];
    M.kpos[K] = r_pos[i];
    M.kend[K] = r_end[i];
    M.kcoff[K] = coff[i];
    M.ksoff[K] = soff[i];
    M.mateK[K] = -1;
    M.keepend[K] = r_end[i];
    uint8_t rb = 0;  // bit 0: the earlier mate keeps its qualities (set by k_gp_pair); bits 1, 2: MD verdict
    if (C.n_mm_type > 0 || C.n_mm > 0) {
        const int m = gp_md_verdict(p, C.n_mm_type, C.n_mm);
        if (m > 0) rb |= 2;
        else if (m < 0) rb |= 4;
    }
    M.keepA[K] = rb;
    if (C.sc) {
        M.kcb[K] = rec_cb[i];
        uint32_t code = 0;
        if (C.has_umi_tag) {
            // The host keeps a dictionary string -> id; all the kernels need is that equal UMIs get equal ids. UMIs are
            // short runs of ACGT of one length: 2 bits per base + 1 is such an id without a dictionary. Anything else
            // (other letters, more than 15 bases, mixed lengths) sends the contig to the host.
            const int l_qname = p[8], n_cigar = (int)ld16u(p + 12), l_qseq = (int)ld32u(p + 16);
            const uint8_t* aux = p + 32 + l_qname + 4 * n_cigar + ((l_qseq + 1) >> 1) + l_qseq;
            const uint8_t* v = gp_aux_find(aux, p + ld32u(p - 4), C.umi_tag[0], C.umi_tag[1]);
            if (v && (*v == 'Z' || *v == 'H')) {
                uint32_t x = 0;
                int n = 0;
                bool good = true;
                for (const uint8_t* q = v + 1; *q; ++q, ++n) {
                    const int c = *q == 'A' ? 0 : *q == 'C' ? 1 : *q == 'G' ? 2 : *q == 'T' ? 3 : -1;
                    if (c < 0 || n >= 15) { good = false; break; }
                    x = (x << 2) | (uint32_t)c;
                }
                if (good && n > 0) {
                    const int seen = atomicCAS(C.umi_len, 0, n);
                    if (seen != 0 && seen != n) good = false;
                }
                if (!good || n == 0) atomicOr(flags, 256);
                code = x + 1u;
            }
        }
        M.kumi[K] = code;
    }
    if (elig[i]) {
        // 64-bit FNV-1a of the read name: the sort key under which the records of one name meet
        unsigned long long hh = 1469598103934665603ull;
        for (const uint8_t* q = p + 32; *q; ++q) { hh ^= *q; hh *= 1099511628211ull; }
        pkey[eidx[i]] = hh ^ (hh >> 29);
        pval[eidx[i]] = K;
    }
}

// overlap_push (htslib): per read name one table entry at a time. The records of a name arrive here side by side
// (sorted by name hash, stable: file order inside a name) and are replayed: a record that finds a live entry pairs
// with it and deletes it; otherwise it becomes the entry if its mate lies ahead (mpos >= pos, or mpos == -1 of a
// paired read). An entry is dead once its read has left the pileup buffer (end < position of the previous pushed read).
__global__ void k_gp_pair(const unsigned long long* __restrict__ pkey, const uint32_t* __restrict__ pval, long long n,
                          const uint8_t* __restrict__ ub, GpMeta M, int* flags) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned long long key = pkey[e];
    if (e > 0 && pkey[e - 1] == key) return;  // not the head of its run
    long long entry = -1;                     // kept index of the record waiting for its mate
    for (long long j = e; j < n && pkey[j] == key; ++j) {
        const uint32_t K = pval[j];
        const uint8_t* p = ub + M.krec[K] + 4;
        if (j > e) {  // same hash: must be the same name, else this is a collision the host has to sort out
            const uint8_t* a = ub + M.krec[pval[e]] + 4 + 32;
            const uint8_t* b = p + 32;
            bool same = true;
            for (int k = 0;; ++k) {
                if (a[k] != b[k]) { same = false; break; }
                if (!a[k]) break;
            }
            if (!same) { atomicOr(flags, 16); return; }
        }
        if (entry >= 0 && M.kend[entry] < M.kprev[K]) entry = -1;  // its read left the buffer before this one came
        if (entry < 0) {
            const int pos = (int)ld32u(p + 4), flag = (int)ld16u(p + 14), mpos = (int)ld32u(p + 24);
            if (mpos >= pos || ((flag & 1) && mpos == -1)) entry = K;
        } else {
            M.mateK[K] = (int32_t)entry;
            // htslib >= 1.13: the earlier mate keeps its qualities iff Wang(X31(name)) & 1
            M.keepA[K] |= (uint8_t)(wang_hash(x31_hash(ub + M.krec[entry] + 4 + 32)) & 1u);
            // htslib's overlap pass can rewrite qualities of the later mate far beyond the earlier mate's end (deletion
            // catch-up across a ref-skip), so the earlier mate stays packed for as long as the later one is
            atomicMax(&M.keepend[entry], M.kend[K]);
            entry = -1;
        }
    }
}

// window planning helpers (one thread each): first K with kpos[K] >= w; kpos of the first K whose ksoff reaches `bytes`
__global__ void k_gp_lower_pos(const int32_t* __restrict__ kpos, long long n, int w, unsigned long long* out) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (kpos[mid] < w) lo = mid + 1; else hi = mid;
    }
    *out = (unsigned long long)lo;
}
__global__ void k_gp_pos_at_bytes(const int32_t* __restrict__ kpos, const uint32_t* __restrict__ ksoff, long long n,
                                  unsigned long long bytes, unsigned long long* out) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((unsigned long long)ksoff[mid] < bytes) lo = mid + 1; else hi = mid;
    }
    *out = lo < n ? (unsigned long long)(unsigned)max(kpos[lo], 0) : ~0ull;
}
// first K in [from, to) that has to be in a window starting at w (keepend > w); *out preset to `to`
__global__ void k_gp_first_live(const int32_t* __restrict__ keepend, long long from, long long to, int w, unsigned long long* out) {
    const long long K = from + blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (K < to && keepend[K] > w) atomicMin(out, (unsigned long long)K);
}

struct GpOut {  // the batch arrays (device) of a window and where this file's share starts in them
    raer_read_hdr* hdr;
    uint32_t* cigar;
    uint8_t* seq;
    uint8_t* qual;
    int32_t* cb;     // single cell: whitelist column / UMI code of every read
    uint32_t* umi;
    int32_t* pair_b;
    unsigned long long* n_pairs;
    long long r0, c0, s0;
};

// eight lanes per kept read K in [k_lo, k_hi): header, CIGAR words, packed sequence, qualities, mate link. The copies are
// byte-wise (records lie at arbitrary offsets); spreading them over eight lanes cuts the chain of dependent accesses a
// single thread per read would walk (measured with one thread per read: 1.17 ms per 10^6 reads at 3 % issue utilisation)
#define GB_LANES 8
__global__ void k_gp_batch(const uint8_t* __restrict__ ub, GpMeta M, long long k_lo, long long k_hi, GpConf C, GpOut O) {
    const long long K = k_lo + (blockIdx.x * (long long)blockDim.x + threadIdx.x) / GB_LANES;
    const int sub = (int)(threadIdx.x & (GB_LANES - 1));
    if (K >= k_hi) return;
    const uint8_t* p = ub + M.krec[K] + 4;
    const int l_qname = p[8], mapq = p[9];
    const int n_cigar = (int)ld16u(p + 12), flag = (int)ld16u(p + 14);
    const int l_qseq = (int)ld32u(p + 16);
    const uint8_t* cg = p + 32 + l_qname;
    const uint8_t* sq = cg + 4 * n_cigar;
    const uint8_t* ql = sq + ((l_qseq + 1) >> 1);
    const long long bi = O.r0 + (K - k_lo);
    const long long co_ = O.c0 + (long long)(M.kcoff[K] - M.kcoff[k_lo]);
    const long long so_ = O.s0 + (long long)(M.ksoff[K] - M.ksoff[k_lo]);
    int bits = 0;
    {  // invert_read_orientation (src/plp_utils.c:331-371)
        const int lt = C.libtype, rev = (flag & 16) != 0;
        int inv = 0;
        if (lt == 1 || lt == 2) {
            int r1 = -1;
            if (!(flag & 1)) r1 = 1;
            else if (flag & 64) r1 = 1;
            else if (flag & 128) r1 = 0;
            if (r1 >= 0) {
                const int when_rev = (lt == 2);
                inv = r1 ? (rev == when_rev) : (rev != when_rev);
            }
        }
        if (inv) bits |= RAER_RB_INVERT;
    }
    if (C.min_overhang && l_qseq >= 2 && (sq[0] & 0xf) == 3) bits |= RAER_RB_OHANG_N;
    const long long mk = M.mateK[K];
    const bool paired = mk >= k_lo;  // an earlier mate in front of the window has nothing left in it
    const int rb = M.keepA[K];
    if (paired && (rb & 1)) bits |= RAER_RB_KEEP_A;
    if (rb & 2) bits |= RAER_RB_MMFAIL;
    if (rb & 4) bits |= RAER_RB_MMERR;
    raer_read_hdr h;
    h.pos = M.kpos[K];
    h.end = M.kend[K];
    h.flag = (uint16_t)flag;
    h.mapq = (uint8_t)mapq;
    h.bits = (uint8_t)bits;
    h.l_qseq = (uint16_t)l_qseq;
    h.n_cigar = (uint16_t)n_cigar;
    h.cigar_off = (uint32_t)co_;
    h.sq_off = (uint32_t)so_;
    h.mate = paired ? (int32_t)(O.r0 + (mk - k_lo)) : -1;
    h.aux = (uint32_t)bi;  // its own batch index, like the host packer (bulk UMI mode never comes here)
    if (sub == 0) {
        O.hdr[bi] = h;
        if (paired) O.pair_b[atomicAdd(O.n_pairs, 1ull)] = (int32_t)bi;
        if (C.sc) {
            O.cb[bi] = M.kcb[K];
            O.umi[bi] = M.kumi[K];
        }
    }
    uint32_t* co = O.cigar + co_;
    for (int k = sub; k < n_cigar; k += GB_LANES) co[k] = ld32u(cg + 4 * k);
    const int nsb = (l_qseq + 1) >> 1, padded = (nsb + 3) & ~3;
    uint8_t* so = O.seq + so_;
    for (int k = sub; k < padded; k += GB_LANES) so[k] = k < nsb ? sq[k] : (uint8_t)0xff;
    uint8_t* qo = O.qual + 2 * so_;
    for (int k = sub; k < 2 * padded; k += GB_LANES) qo[k] = k < l_qseq ? ql[k] : (uint8_t)0;
}

// ------------------------------------------------------------------------------------- host side
namespace {

static inline uint32_t h_le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint64_t h_le64(const uint8_t* p) { return (uint64_t)h_le32(p) | ((uint64_t)h_le32(p + 4) << 32); }

struct BaiRef {
    uint64_t vbeg = 0, vend = 0;  // first chunk begin / last chunk end of the contig (0 / 0: no records)
    std::vector<uint64_t> lin;
};

// the whole .bai: per reference the extent of its records in the file and its linear index
static bool load_bai(const char* path, std::vector<BaiRef>& out) {
    FILE* fp = fopen(path, "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    const long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<uint8_t> d((size_t)std::max<long>(sz, 0));
    const bool rd = sz >= 8 && fread(d.data(), 1, (size_t)sz, fp) == (size_t)sz;
    fclose(fp);
    if (!rd || memcmp(d.data(), "BAI\1", 4)) return false;
    const int n_ref = (int)h_le32(d.data() + 4);
    if (n_ref < 0) return false;
    long o = 8;
    auto room = [&](long bytes) { return bytes >= 0 && o >= 0 && o <= sz && bytes <= sz - o; };
    out.assign((size_t)n_ref, BaiRef());
    for (int t = 0; t < n_ref; ++t) {
        if (!room(4)) return false;
        const int n_bin = (int)h_le32(d.data() + o);
        o += 4;
        if (n_bin < 0) return false;
        uint64_t vb = UINT64_MAX, ve = 0;
        for (int b = 0; b < n_bin; ++b) {
            if (!room(8)) return false;
            const uint32_t bin = h_le32(d.data() + o);
            const int n_chunk = (int)h_le32(d.data() + o + 4);
            o += 8;
            if (n_chunk < 0 || !room(16L * n_chunk)) return false;
            for (int c = 0; c < n_chunk; ++c) {
                if (bin != 37450) {
                    vb = std::min(vb, h_le64(d.data() + o));
                    ve = std::max(ve, h_le64(d.data() + o + 8));
                }
                o += 16;
            }
        }
        if (!room(4)) return false;
        const int n_intv = (int)h_le32(d.data() + o);
        o += 4;
        if (n_intv < 0 || !room(8L * n_intv)) return false;
        BaiRef& R = out[(size_t)t];
        if (vb != UINT64_MAX && ve > vb) { R.vbeg = vb; R.vend = ve; }
        R.lin.resize((size_t)n_intv);
        for (int k = 0; k < n_intv; ++k) R.lin[(size_t)k] = h_le64(d.data() + o + 8L * k);
        o += 8L * n_intv;
    }
    return true;
}

// Device buffers come from the stream-ordered allocator: a contig needs some sixty arrays whose sizes are only known
// stage by stage, and cudaMalloc / cudaFree synchronise the device on every call (measured: 48 ms to set up and 23 ms to
// tear down per call of the public entry point). The pool keeps freed memory for the next contig and is trimmed when the
// call ends. The stream is the owning raer_gi's.
static thread_local cudaStream_t t_alloc_stream = nullptr;
struct DBufG {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap && p) return 0;
        if (p) cudaFreeAsync(p, t_alloc_stream);
        const size_t want = n + n / 8 + 4096;
        if (cudaMallocAsync(&p, want, t_alloc_stream) != cudaSuccess) { p = nullptr; cap = 0; cudaGetLastError(); set_error("cudaMallocAsync(%zu) failed", want); return RAER_ERR_CUDA; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeAsync(p, t_alloc_stream); p = nullptr; cap = 0; }
    ~DBufG() { release(); }
};

struct PinnedBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
    bool ensure(size_t n) {
        if (n <= cap) return true;
        if (p) cudaFreeHost(p);
        cap = n + n / 8 + 4096;
        if (cudaHostAlloc((void**)&p, cap, cudaHostAllocDefault) != cudaSuccess) { p = nullptr; cap = 0; cudaGetLastError(); return false; }
        return true;
    }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

// The readers' staging buffers (a few MB each) outlive the call: page-locking costs about a millisecond per MB, every call
static std::mutex g_pin_mu;
static std::vector<uint8_t*> g_pin_pool;
#define GI_CHUNK ((size_t)4 << 20)
#define GI_OVERLAP ((size_t)65536 + 4096)
#define GI_HOST_INFLATE_BYTES ((size_t)384 << 10)  // compressed ranges up to this size are inflated on the host
static uint8_t* pin_acquire() {
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (!g_pin_pool.empty()) { uint8_t* p = g_pin_pool.back(); g_pin_pool.pop_back(); return p; }
    }
    uint8_t* p = nullptr;
    if (cudaHostAlloc((void**)&p, GI_CHUNK + GI_OVERLAP, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
static void pin_release(uint8_t* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (g_pin_pool.size() < 16) g_pin_pool.push_back(p);
    else cudaFreeHost(p);
}

// everything one file contributes to the batch of a contig, still in its own buffers
struct FileStage {
    DBufG comp, blk, ubuf, status, sp, cnt, base, rec_off, pass, pidx, r_pos, r_end, cpos, kept, kidx, ncig, coff, sqb, soff, elig, eidx,
        prevpos, bsum, cend, dropped, krec, kprev, pkey, pval, pkey2, pval2, hist, queue, rec_cb, kpos, kend, kcoff, ksoff, kcb, kumi, mateK, keepend,
        keepA;
    long long k_lo_prev = 0;  // windows: where the previous window's live reads began
    PinnedBuf hblk;
    long long n_rec = 0, n_pass = 0, n_kept = 0, n_cig = 0, n_sq = 0, n_elig = 0;
    unsigned long long u_total = 0;
    int64_t aligned_bases = 0;
    int max_span = 0;
};

}  // namespace

struct GiReader {
    uint8_t* buf = nullptr;
    cudaStream_t st = nullptr;
};

struct raer_gi {
    int device = 0;
    cudaStream_t st = nullptr;
    int nf = 0;
    int n_sm = 148;
    bool device_inflate_only = false;  // RAER_GI_DEVICE_INFLATE=1: tests send even tiny ranges through the inflate kernel
    uint32_t check_crc = 1;  // RAER_SKIP_CRC=1 turns the CRC-32 check of the inflated blocks off, as in the host ingest
    std::vector<GiReader> readers;
    std::vector<FILE*> fp;
    std::vector<std::vector<BaiRef>> bai;
    std::vector<long long> fsize;
    std::vector<FileStage> stage;
    DBufG small, site_bits, site_pre, hdr, maxend, cigar, seq, qual, pair_b, endtmp, cb, umi, wl_key, wl_col, wl_off, wl_pool, umi_len;
    uint32_t wl_mask = 0;
    bool wl_built = false;
    // the contig being served: position windows [win[j], win[j + 1]) and what a window's batch needs
    std::vector<int> win;
    std::vector<GpConf> fconf;  // per file (library type, contig number)
    int cur_tid = -1, cur_len = 0, cur_beg = 0;  // rows of [cur_beg, cur_len) are wanted
    bool cur_sc = false;
    int64_t cur_bases = 0;
    size_t pool_bytes() const {  // what the device pool holds for reuse (freed buffers of earlier contigs count as free memory)
        cudaMemPool_t pool;
        unsigned long long reserved = 0, used = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess) return 0;
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
        return reserved > used ? (size_t)(reserved - used) : 0;
    }
    unsigned long long* h_small = nullptr;  // pinned, 8 counters
    int64_t launches = 0, inflated = 0, compressed = 0, records = 0, walk_retries = 0, n_dropped = 0;
    double t_read = 0, t_table = 0, t_alloc = 0, t_inflate = 0, t_pass = 0, t_kept = 0, t_write = 0;  // host wall seconds per stage
    ~raer_gi() {
        for (FILE* f : fp) if (f) fclose(f);
        for (GiReader& r : readers) {
            if (r.st) cudaStreamDestroy(r.st);
            pin_release(r.buf);
        }
        if (h_small) cudaFreeHost(h_small);
        if (st) cudaStreamDestroy(st);
    }
};

extern "C" void raer_gi_destroy(raer_gi* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    t_alloc_stream = g->st;
    if (g->st) cudaStreamSynchronize(g->st);
    if (getenv("RAER_TIMING"))
        fprintf(stderr, "[device ingest] %.1f MB compressed -> %.1f MB, %lld records: read %.3f s, block table %.3f, alloc %.3f, "
                        "h2d+inflate+walk %.3f, prefilter %.3f, kept+scans %.3f, write+sort+pair %.3f; %lld walk(s) redone from the index\n",
                g->compressed / 1e6, g->inflated / 1e6, (long long)g->records, g->t_read, g->t_table, g->t_alloc, g->t_inflate, g->t_pass,
                g->t_kept, g->t_write, (long long)g->walk_retries);
    const int device = g->device;
    cudaStream_t st = g->st;
    g->st = nullptr;  // the buffers are freed on it by the destructors
    delete g;
    if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    // The pool keeps up to 8 GB for the next call of the process (handing memory back to the driver costs far more than
    // the ingest of a small file: 0.35 s measured for 3 GB); raer_trim_device_cache() returns it on request.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = 8ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    t_alloc_stream = nullptr;
}

// opens the files and their indexes; nullptr (with the error set) if any index is missing or unreadable
extern "C" raer_gi* raer_gi_create(int device, int nf, const char* const* paths, const char* const* indexes) {
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    raer_gi* g = new raer_gi();
    g->device = device;
    g->nf = nf;
    g->stage.resize((size_t)nf);
    g->bai.resize((size_t)nf);
    for (int i = 0; i < nf; ++i) {
        FILE* f = fopen(paths[i], "rb");
        g->fp.push_back(f);
        if (!f) { set_error("failed to open %s", paths[i]); delete g; return nullptr; }
        fseek(f, 0, SEEK_END);
        g->fsize.push_back(ftell(f));
        const std::string bp = (indexes && indexes[i]) ? indexes[i] : std::string(paths[i]) + ".bai";
        if (!load_bai(bp.c_str(), g->bai[(size_t)i])) { set_error("fail to load bamfile index %s", bp.c_str()); delete g; return nullptr; }
    }
    {   // freed buffers stay in the pool until the call ends (raer_gi_destroy trims it)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if (cudaStreamCreateWithFlags(&g->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc((void**)&g->h_small, 64, cudaHostAllocDefault) != cudaSuccess) {
        set_error("stream / pinned memory");
        delete g;
        return nullptr;
    }
    if (getenv("RAER_SKIP_CRC")) g->check_crc = 0;
    if (const char* e = getenv("RAER_GI_DEVICE_INFLATE")) g->device_inflate_only = atoi(e) != 0;
    {   // per device: SM count, and the opt-in for the inflate kernel's shared memory
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) g->n_sm = v;
        if (cudaFuncSetAttribute(k_gi_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, GI_SMEM_BYTES) != cudaSuccess) {
            cudaGetLastError();
            set_error("the device does not offer %d bytes of shared memory per block", (int)GI_SMEM_BYTES);
            delete g;
            return nullptr;
        }
    }
    int n_readers = 4;
    if (const char* e = getenv("RAER_GI_READERS")) n_readers = std::max(1, std::min(16, atoi(e)));
    g->readers.resize((size_t)n_readers);
    for (GiReader& r : g->readers)
        if (!(r.buf = pin_acquire()) || cudaStreamCreateWithFlags(&r.st, cudaStreamNonBlocking) != cudaSuccess) {
            set_error("stream / pinned memory");
            delete g;
            return nullptr;
        }
    return g;
}

template <bool MAX>
static int g_scan(raer_gi* g, FileStage& S, const uint32_t* in, uint32_t* out, long long n, unsigned long long* d_total) {
    if (n <= 0) {
        if (d_total) cudaMemsetAsync(d_total, 0, 8, g->st);
        return 0;
    }
    const unsigned nb = (unsigned)((n + (long long)SC_BLOCK * SC_ITEMS - 1) / ((long long)SC_BLOCK * SC_ITEMS));
    int rc = S.bsum.ensure((size_t)nb * 4 + 16);
    if (rc) return rc;
    k_sc_sum<MAX><<<nb, SC_BLOCK, 0, g->st>>>(in, n, (uint32_t*)S.bsum.p);
    k_sc_top<MAX><<<1, 256, 0, g->st>>>((uint32_t*)S.bsum.p, (int)nb, d_total);
    k_sc_apply<MAX><<<nb, SC_BLOCK, 0, g->st>>>(in, n, (const uint32_t*)S.bsum.p, out);
    g->launches += 3;
    return 0;
}

// bam_plp_push's max_depth rule over the passing reads of one file and contig (pos / end in stream order): a read that
// starts where the previous one did is dropped when the buffer already holds max_depth reads (reads that were pushed and
// whose end has not fallen behind the iterator). Same decisions as Ingest::parse_one; a min-heap of ends instead of its
// lazily compacted vector. Returns the number of dropped reads.
static long long gi_replay_max_depth(const int32_t* pos, const int32_t* end, long long n, int max_depth, uint8_t* dropped) {
    std::priority_queue<int32_t, std::vector<int32_t>, std::greater<int32_t>> live;
    long long n_drop = 0;
    int32_t prev = INT_MIN;
    for (long long i = 0; i < n; ++i) {
        if (i > 0 && pos[i] == prev && (long long)live.size() + 1 > max_depth) {
            while (!live.empty() && live.top() < prev) live.pop();
            if ((long long)live.size() + 1 > max_depth) { dropped[i] = 1; ++n_drop; continue; }
        }
        const bool kept = i == 0 || end[i] > prev;
        prev = pos[i];
        if (kept) live.push(end[i]);
    }
    return n_drop;
}

// the replay as a C-ABI entry (host only, no device needed): tests hold it to a literal restatement of the rule
extern "C" int64_t raer_replay_max_depth(const int32_t* pos, const int32_t* end, int64_t n, int32_t max_depth, uint8_t* dropped) {
    memset(dropped, 0, (size_t)n);
    return gi_replay_max_depth(pos, end, n, max_depth, dropped);
}

// Stage 1 for one file: compressed range -> inflated records -> record table -> prefilter -> kept reads and their sizes.
// Returns RAER_OK, RAER_GI_FALLBACK or an error.
static int gi_file_stage(raer_gi* g, int f, int ftid, const GpConf& C0, bool to_contig_end = false) {
    FileStage& S = g->stage[(size_t)f];
    S.n_rec = S.n_pass = S.n_kept = S.n_cig = S.n_sq = S.n_elig = 0;
    S.aligned_bases = 0;
    S.u_total = 0;
    if (ftid < 0 || ftid >= (int)g->bai[(size_t)f].size()) return RAER_OK;
    const BaiRef& R = g->bai[(size_t)f][(size_t)ftid];
    if (R.vend <= R.vbeg) return RAER_OK;  // no records of this contig in this file
    cudaStream_t st = g->st;
    int rc;
    // ---- the virtual-offset range to read: the whole contig, or -- for a coordinate range -- from the linear-index entry of
    // the 16 kb window the range begins in (the first record that reaches that window, like sam_itr_querys) to the entry of
    // a window behind its end. The records are sorted by position, so the range holds every read that starts before
    // reg_end if the first record behind it starts at or after reg_end: that record is looked at once the blocks are
    // inflated (peek), and if it does not, the stage is redone to the end of the contig.
    uint64_t vb = R.vbeg, ve = R.vend;
    bool peek = false;
    if (C0.reg_on) {
        const size_t n_intv = R.lin.size();
        if (n_intv) {
            size_t w = std::min<size_t>((size_t)(std::max(C0.reg_beg, 0) >> 14), n_intv - 1);
            uint64_t v = R.lin[w];
            while (v == 0 && w > 0) v = R.lin[--w];
            if (v > vb && v < ve) vb = v;
        }
        if (!to_contig_end && C0.reg_end < INT_MAX - (1 << 16)) {
            for (size_t k = (size_t)(C0.reg_end >> 14) + 2; k < n_intv; ++k)
                if (R.lin[k] > vb && R.lin[k] < ve) { ve = R.lin[k]; peek = true; break; }
        }
    }
    // ---- compressed bytes of the range: from the block of the first record to the block its last record ends in
    const long long c_beg = (long long)(vb >> 16);
    long long c_last = (long long)(ve >> 16);
    if ((ve & 0xffff) == 0 && !peek) c_last -= 1;  // the range ends exactly where a block starts: that block is not needed
    long long c_end = std::min<long long>(c_last + 65536 + 32, g->fsize[(size_t)f]);
    if (c_end <= c_beg || c_last >= g->fsize[(size_t)f]) {  // the index describes records the file does not hold
        set_error("truncated BGZF file (the index points beyond the end of the file)");
        return RAER_ERR_IO;
    }
    const size_t nbytes = (size_t)(c_end - c_beg);
    {   // compressed bytes + inflated records (about 4x) + per-record tables + this file's share of the batch: about 8x
        size_t free_b = 0, total_b = 0;
        if (nbytes > ((size_t)64 << 20) && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && nbytes * 8 > free_b + g->pool_bytes())
            return RAER_GI_FALLBACK;
        cudaGetLastError();
    }
    std::vector<GiBlock> blocks;
    std::vector<long long> blk_c;  // file offset of every block
    blocks.reserve(nbytes / 12000 + 64);
    blk_c.reserve(nbytes / 12000 + 64);
    unsigned long long u = 0;
    size_t o = 0;  // next block header, relative to c_beg
    // A range of a few blocks (a contig with few reads; assemblies have thousands of those) is inflated right here: one
    // block takes the device decoder 4 ms however few there are, the host decoder a quarter of a millisecond.
    const bool small = nbytes <= GI_HOST_INFLATE_BYTES && !g->device_inflate_only;
    if (small) {
        const double t0s = raer::now_seconds();
        std::vector<uint8_t> cb(nbytes + 64, 0);
        size_t got = 0;
        while (got < nbytes) {
            const ssize_t r = pread(fileno(g->fp[(size_t)f]), cb.data() + got, nbytes - got, (off_t)(c_beg + (long long)got));
            if (r <= 0) break;
            got += (size_t)r;
        }
        if (got != nbytes) { set_error("read error in BAM file"); return RAER_ERR_IO; }
        std::vector<uint32_t> crcs;
        while ((long long)o + c_beg <= c_last) {
            if (o + 18 > nbytes) { set_error("truncated BGZF file (the index points beyond the end of the file)"); return RAER_ERR_IO; }
            const uint8_t* h = cb.data() + o;
            if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { set_error("corrupt BGZF block header"); return RAER_ERR_IO; }
            const int xlen = h[10] | (h[11] << 8);
            if (o + 12 + (size_t)xlen > nbytes) { set_error("truncated BGZF file (incomplete block at the end)"); return RAER_ERR_IO; }
            int bsize = -1;
            for (int e = 0; e + 4 <= xlen;) {
                const int slen = h[12 + e + 2] | (h[12 + e + 3] << 8);
                if (h[12 + e] == 'B' && h[13 + e] == 'C' && slen == 2) bsize = h[12 + e + 4] | (h[12 + e + 5] << 8);
                e += 4 + slen;
            }
            if (bsize < 0 || bsize + 1 < 12 + xlen + 2 + 8) { set_error("corrupt BGZF block (BSIZE)"); return RAER_ERR_IO; }
            if (o + (size_t)bsize + 1 > nbytes) { set_error("truncated BGZF file (incomplete block at the end)"); return RAER_ERR_IO; }
            const uint32_t isize = h_le32(h + bsize + 1 - 4);
            if (isize > 65536u) { set_error("corrupt BGZF block (ISIZE above 64 KiB)"); return RAER_ERR_IO; }
            GiBlock b;
            b.c_off = o + 12 + (size_t)xlen;
            b.c_len = (uint32_t)(bsize + 1 - 12 - xlen - 8);
            b.u_off = u;
            b.u_len = isize;
            b.crc = h_le32(h + bsize + 1 - 8);
            b.check_crc = g->check_crc;
            blocks.push_back(b);
            blk_c.push_back(c_beg + (long long)o);
            u += isize;
            o += (size_t)bsize + 1;
        }
        if (blocks.empty()) return RAER_OK;
        if (!S.hblk.ensure((size_t)u + 64)) { set_error("pinned allocation failed"); return RAER_ERR_CUDA; }
        for (const GiBlock& b : blocks) {
            if (!b.u_len) continue;
            uint8_t* dst = S.hblk.p + b.u_off;
            bool ok = raer::inflate_raw(cb.data() + b.c_off, b.c_len, dst, b.u_len) == 0;
            if (!ok) {  // what the own decoder refuses goes to zlib, as in the host ingest
                z_stream zs;
                memset(&zs, 0, sizeof(zs));
                if (inflateInit2(&zs, -15) == Z_OK) {
                    zs.next_in = cb.data() + b.c_off;
                    zs.avail_in = b.c_len;
                    zs.next_out = dst;
                    zs.avail_out = b.u_len;
                    ok = inflate(&zs, Z_FINISH) == Z_STREAM_END && zs.avail_out == 0;
                    inflateEnd(&zs);
                }
            }
            if (!ok) { set_error("inflate failed (corrupt BGZF block)"); return RAER_ERR_IO; }
            if (b.check_crc && (uint32_t)crc32(0L, dst, b.u_len) != b.crc) { set_error("inflate failed (BGZF block CRC32 mismatch)"); return RAER_ERR_IO; }
        }
        if ((rc = S.ubuf.ensure((size_t)u + 64)) || (rc = g->small.ensure(256))) return rc;
        memset(S.hblk.p + u, 0, 64);
        GCK(cudaMemcpyAsync(S.ubuf.p, S.hblk.p, (size_t)u + 64, cudaMemcpyHostToDevice, st));
        g->t_read += raer::now_seconds() - t0s;
    } else {
    // The range is read in chunks of 8 MB by a few threads. Each has one small pinned buffer and a stream of its own:
    // pread -> H2D copy -> (in chunk order) the BGZF block headers of its chunk, 18 bytes each, chained through BSIZE. A
    // chunk is read with 64 KiB of overlap so that every block that starts in it is complete. This thread collects the
    // block tables and launches the inflate kernel whenever a wave of blocks is ready, so the device inflates while the
    // rest of the range is still being read.
    const double t0 = raer::now_seconds();
    const size_t CH = GI_CHUNK, OV = GI_OVERLAP;
    const int n_chunks = (int)((nbytes + CH - 1) / CH);
    const int n_readers = std::max(1, std::min((int)g->readers.size(), n_chunks));
    const size_t blk_guess = nbytes / 4096 + 1024;
    if ((rc = S.comp.ensure(nbytes + OV + 64)) || (rc = S.ubuf.ensure(std::max(S.ubuf.cap, nbytes * 5 + (1u << 20)))) ||
        (rc = S.blk.ensure(blk_guess * sizeof(GiBlock))) || (rc = S.status.ensure(blk_guess * 4)) ||
        (rc = S.queue.ensure(((size_t)n_chunks + 2) * 4)) || (rc = g->small.ensure(256)))
        return rc;
    if (!S.hblk.ensure(blk_guess * sizeof(GiBlock))) { set_error("pinned allocation failed"); return RAER_ERR_CUDA; }
    GCK(cudaMemsetAsync(S.queue.p, 0, ((size_t)n_chunks + 2) * 4, st));
    GCK(cudaMemsetAsync((uint8_t*)S.comp.p + nbytes, 0, 64, st));
    GCK(cudaStreamSynchronize(st));  // the readers copy on their own streams
    g->t_alloc += raer::now_seconds() - t0;
    const double t1 = raer::now_seconds();
    int turn = 0;            // chunk whose headers are scanned next
    bool scan_done = false;  // the block that holds the last record of the contig has been seen
    int fail = 0;
    std::string fail_msg;
    std::mutex mu;
    std::condition_variable cv;
    const int fd = fileno(g->fp[(size_t)f]);
    uint8_t* const d_comp = (uint8_t*)S.comp.p;
    auto reader = [&](int t) {
        cudaSetDevice(g->device);
        GiReader& RD = g->readers[(size_t)t];
        for (int i = t; i < n_chunks; i += n_readers) {
            const size_t lo = (size_t)i * CH, len = std::min(CH + OV, nbytes - lo);
            const char* why = nullptr;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (fail || scan_done) break;
            }
            size_t got = 0;
            while (got < len) {
                const ssize_t r = pread(fd, RD.buf + got, len - got, (off_t)(c_beg + (long long)(lo + got)));
                if (r <= 0) break;
                got += (size_t)r;
            }
            if (got != len) why = "read error in BAM file";
            if (!why && (cudaMemcpyAsync(d_comp + lo, RD.buf, len, cudaMemcpyHostToDevice, RD.st) != cudaSuccess ||
                         cudaStreamSynchronize(RD.st) != cudaSuccess))
                why = "H2D copy of compressed blocks failed";
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&]() { return turn == i || fail || scan_done; });
            if (fail || scan_done) break;
            const size_t hi = std::min(lo + CH, nbytes);  // blocks that start before hi are this chunk's
            while (!why && o < hi) {
                if ((long long)o + c_beg > c_last) { scan_done = true; break; }  // the block that ends the contig has been seen
                if (o + 18 > nbytes) { why = "truncated BGZF file (the index points beyond the end of the file)"; break; }
                const uint8_t* h = RD.buf + (o - lo);
                const size_t avail = lo + len;
                if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) { why = "corrupt BGZF block header"; break; }
                const int xlen = h[10] | (h[11] << 8);
                if (o + 12 + (size_t)xlen > avail) { why = "truncated BGZF file (incomplete block at the end)"; break; }
                int bsize = -1;
                for (int e = 0; e + 4 <= xlen;) {
                    const int slen = h[12 + e + 2] | (h[12 + e + 3] << 8);
                    if (h[12 + e] == 'B' && h[13 + e] == 'C' && slen == 2) bsize = h[12 + e + 4] | (h[12 + e + 5] << 8);
                    e += 4 + slen;
                }
                if (bsize < 0 || bsize + 1 < 12 + xlen + 2 + 8) { why = "corrupt BGZF block (BSIZE)"; break; }
                if (o + (size_t)bsize + 1 > avail) { why = "truncated BGZF file (incomplete block at the end)"; break; }
                const uint32_t isize = h_le32(h + bsize + 1 - 4);
                if (isize > 65536u) { why = "corrupt BGZF block (ISIZE above 64 KiB)"; break; }
                GiBlock b;
                b.c_off = o + 12 + (size_t)xlen;
                b.c_len = (uint32_t)(bsize + 1 - 12 - xlen - 8);
                b.u_off = u;
                b.u_len = isize;
                b.crc = h_le32(h + bsize + 1 - 8);
                b.check_crc = g->check_crc;
                blocks.push_back(b);
                blk_c.push_back(c_beg + (long long)o);
                u += isize;
                o += (size_t)bsize + 1;
            }
            if (why) { fail = 1; fail_msg = why; }
            turn = i + 1;
            lk.unlock();
            cv.notify_all();
        }
    };
    std::vector<std::thread> readers;
    for (int t = 0; t < n_readers; ++t) readers.emplace_back(reader, t);
    struct Joiner {
        std::vector<std::thread>& th;
        std::mutex& mu;
        std::condition_variable& cv;
        int& fail;
        ~Joiner() {
            { std::lock_guard<std::mutex> lk(mu); if (!fail) fail = 2; }
            cv.notify_all();
            for (auto& t : th) if (t.joinable()) t.join();
        }
    };
    const int n_sm = g->n_sm;
    // b_lo .. b_hi: entries already staged in S.hblk (or, beyond its capacity, passed in `extra`)
    int n_launch = 0;
    auto launch_inflate = [&](size_t b_lo, size_t b_hi, const GiBlock* extra) -> int {
        if (b_hi <= b_lo) return 0;
        int r2;
        if (b_hi * sizeof(GiBlock) > S.blk.cap || b_hi * 4 > S.status.cap) {  // more (smaller) blocks than guessed: regrow, keep contents
            DBufG nb, ns;
            if ((r2 = nb.ensure(b_hi * 2 * sizeof(GiBlock))) || (r2 = ns.ensure(b_hi * 2 * 4))) return r2;
            GCK(cudaStreamSynchronize(st));
            GCK(cudaMemcpy(nb.p, S.blk.p, b_lo * sizeof(GiBlock), cudaMemcpyDeviceToDevice));
            GCK(cudaMemcpy(ns.p, S.status.p, b_lo * 4, cudaMemcpyDeviceToDevice));
            std::swap(nb.p, S.blk.p); std::swap(nb.cap, S.blk.cap);
            std::swap(ns.p, S.status.p); std::swap(ns.cap, S.status.cap);
        }
        const GiBlock* src = extra ? extra : (const GiBlock*)S.hblk.p + b_lo;  // pinned: a pageable source would wait for the stream
        GCK(cudaMemcpyAsync((GiBlock*)S.blk.p + b_lo, src, (b_hi - b_lo) * sizeof(GiBlock), cudaMemcpyHostToDevice, st));
        const int nbk = (int)(b_hi - b_lo);
        const int grid = std::min<int>(n_sm, nbk);
        k_gi_inflate<<<grid, GI_WARPS * 32, GI_SMEM_BYTES, st>>>((const uint8_t*)S.comp.p, (const GiBlock*)S.blk.p + b_lo, nbk,
                                                                              (uint8_t*)S.ubuf.p, (int*)S.status.p + b_lo,
                                                                              (unsigned*)S.queue.p + n_launch);
        ++n_launch;
        ++g->launches;
        return 0;
    };
    size_t launched = 0;
    bool too_small = false;
    double t_wait = 0;
    {
        Joiner joiner{readers, mu, cv, fail};
        const size_t wave = (size_t)n_sm * GI_WARPS;
        while (true) {
            size_t have;
            bool finished;
            std::vector<GiBlock> extra;
            {
                const double tw = raer::now_seconds();
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&]() { return fail || scan_done || turn == n_chunks || blocks.size() - launched >= wave; });
                t_wait += raer::now_seconds() - tw;
                if (fail) { set_error("%s", fail_msg.c_str()); return fail_msg.find("H2D") != std::string::npos ? RAER_ERR_CUDA : RAER_ERR_IO; }
                have = blocks.size();
                finished = scan_done || turn == n_chunks;
                if (u + 64 > S.ubuf.cap) too_small = true;
                if (!too_small && have > launched) {
                    if (have * sizeof(GiBlock) <= S.hblk.cap)
                        memcpy(S.hblk.p + launched * sizeof(GiBlock), blocks.data() + launched, (have - launched) * sizeof(GiBlock));
                    else
                        extra.assign(blocks.begin() + (long)launched, blocks.begin() + (long)have);
                }
            }
            if (!too_small && have > launched && (finished || have - launched >= wave)) {
                if (n_launch >= n_chunks + 1) { GCK(cudaStreamSynchronize(st)); GCK(cudaMemsetAsync(S.queue.p, 0, ((size_t)n_chunks + 2) * 4, st)); n_launch = 0; }
                if ((rc = launch_inflate(launched, have, extra.empty() ? nullptr : extra.data()))) return rc;
                if (!extra.empty()) GCK(cudaStreamSynchronize(st));
                launched = have;
            }
            if (finished) break;
        }
    }
    if (!scan_done && (long long)o + c_beg <= c_last) { set_error("truncated BGZF file (incomplete block at the end)"); return RAER_ERR_IO; }
    g->t_read += t_wait;
    g->t_table += raer::now_seconds() - t1 - t_wait;
    if (blocks.empty()) return RAER_OK;
    if (too_small) {  // inflated size beyond the guess: the compressed bytes are on the device, inflate everything again
        GCK(cudaStreamSynchronize(st));
        if ((rc = S.ubuf.ensure((size_t)u + 64))) return rc;
        GCK(cudaMemsetAsync(S.queue.p, 0, 8, st));
        n_launch = 0;
        if ((rc = launch_inflate(0, blocks.size(), blocks.data()))) return rc;
        GCK(cudaStreamSynchronize(st));
    }
    }  // device inflate
    S.u_total = u;
    g->compressed += (int64_t)o;
    g->inflated += (int64_t)u;
    // ---- start points: the first record, and every linear-index entry inside the range, as inflated offsets
    auto v2u = [&](uint64_t v, unsigned long long* out) -> bool {
        const long long c = (long long)(v >> 16);
        const size_t k = std::lower_bound(blk_c.begin(), blk_c.end(), c) - blk_c.begin();
        if (k == blk_c.size() || blk_c[k] != c) {
            if (k == blk_c.size() && (v & 0xffff) == 0 && c >= blk_c.back()) { *out = u; return true; }  // behind the last block
            return false;
        }
        if ((v & 0xffff) > blocks[k].u_len) return false;
        *out = blocks[k].u_off + (v & 0xffff);
        return true;
    };
    unsigned long long u_first = 0, u_last = u;
    if (!v2u(vb, &u_first)) return RAER_GI_FALLBACK;
    if (!v2u(ve, &u_last)) { if (peek) return gi_file_stage(g, f, ftid, C0, true); u_last = u; }
    // Start points of the record walk. htslib-style writers never let a record straddle two BGZF blocks (bgzf_flush_try
    // before every record), so every block start is a record start: try those first -- the walk verifies the guess,
    // every chain has to land exactly on the next start point -- and fall back to the linear index of the .bai, whose
    // entries are record starts by construction (one per 16 kb window that has reads).
    std::vector<unsigned long long> sp_blk, sp_lin;
    sp_blk.push_back(u_first);
    for (const GiBlock& b : blocks)
        if (b.u_off > u_first && b.u_off < u_last && b.u_len) sp_blk.push_back(b.u_off);
    sp_lin.push_back(u_first);
    for (uint64_t v : R.lin) {
        unsigned long long x;
        if (v > vb && v < ve && v2u(v, &x) && x > u_first && x < u_last) sp_lin.push_back(x);
    }
    std::sort(sp_lin.begin(), sp_lin.end());
    sp_lin.erase(std::unique(sp_lin.begin(), sp_lin.end()), sp_lin.end());
    sp_blk.push_back(u_last);
    sp_lin.push_back(u_last);
    const bool try_blocks = !getenv("RAER_GI_LINEAR_WALK");
    std::vector<unsigned long long>& sp = try_blocks ? sp_blk : sp_lin;
    int nsp = (int)sp.size() - 1;
    const size_t sp_cap = std::max(sp_blk.size(), sp_lin.size());
    const double t2 = raer::now_seconds();
    if ((rc = S.sp.ensure(sp_cap * 8)) || (rc = S.cnt.ensure(sp_cap * 4 + 16)) || (rc = S.base.ensure(sp_cap * 4 + 16))) return rc;
    g->t_alloc += raer::now_seconds() - t2;
    const double t3 = raer::now_seconds();
    GCK(cudaMemcpyAsync(S.sp.p, sp.data(), sp.size() * 8, cudaMemcpyHostToDevice, st));
    GCK(cudaMemsetAsync(g->small.p, 0, 256, st));
    unsigned long long* d_small = (unsigned long long*)g->small.p;  // [1] flags, [2] aligned bases, [3] max span, [4..] totals
    // ---- record table
    int* d_flags = (int*)(d_small + 1);
    std::vector<int> status(small ? 0 : blocks.size());  // verdicts of the inflate kernel (a host-inflated range has been checked)
    if (!small) GCK(cudaMemcpyAsync(status.data(), S.status.p, blocks.size() * 4, cudaMemcpyDeviceToHost, st));
    for (int attempt = try_blocks ? 0 : 1; attempt < 2; ++attempt) {
        std::vector<unsigned long long>& pts = attempt == 0 ? sp_blk : sp_lin;
        nsp = (int)pts.size() - 1;
        if (attempt == 1) {
            GCK(cudaMemcpyAsync(S.sp.p, pts.data(), pts.size() * 8, cudaMemcpyHostToDevice, st));
            GCK(cudaMemsetAsync(d_flags, 0, 8, st));
        }
        k_gw_walk<false><<<(nsp + 127) / 128, 128, 0, st>>>((const uint8_t*)S.ubuf.p, u_last, (const unsigned long long*)S.sp.p, nsp,
                                                             (uint32_t*)S.cnt.p, nullptr, nullptr, d_flags);
        ++g->launches;
        if ((rc = g_scan<false>(g, S, (const uint32_t*)S.cnt.p, (uint32_t*)S.base.p, nsp, d_small + 4))) return rc;
        GCK(cudaMemcpyAsync(g->h_small, d_small, 64, cudaMemcpyDeviceToHost, st));
        GCK(cudaStreamSynchronize(st));
        if (attempt == 0)  // the inflate results are in by now
            for (size_t k = 0; k < status.size(); ++k)
                if (status[k] != 0) { set_error("inflate failed (corrupt BGZF block, device code %d)", status[k]); return RAER_ERR_IO; }
        if (!(g->h_small[1] & (1 | 32))) break;
        if (attempt == 0) { ++g->walk_retries; continue; }  // some record straddles two blocks: walk from the index entries
        if (g->h_small[1] & 1) { set_error("corrupt BAM record"); return RAER_ERR_IO; }
        return RAER_GI_FALLBACK;  // the index and the records disagree: the host stream reads them one by one
    }
    if (!try_blocks)
        for (size_t k = 0; k < status.size(); ++k)
            if (status[k] != 0) { set_error("inflate failed (corrupt BGZF block, device code %d)", status[k]); return RAER_ERR_IO; }
    if (peek) {  // the first record behind the range has to start at or after the end of the coordinate range
        uint8_t hd[12];
        bool ok = u_last + 12 <= u;
        if (ok) {
            GCK(cudaMemcpy(hd, (const uint8_t*)S.ubuf.p + u_last, 12, cudaMemcpyDeviceToHost));
            ok = (int)h_le32(hd + 4) == ftid && (int)h_le32(hd + 8) >= C0.reg_end;
        }
        if (!ok) return gi_file_stage(g, f, ftid, C0, true);
    }
    S.n_rec = (long long)g->h_small[4];
    g->records += S.n_rec;
    g->t_inflate += raer::now_seconds() - t3;
    if (S.n_rec == 0) return RAER_OK;
    const double t4 = raer::now_seconds();
    const size_t nr = (size_t)S.n_rec;
    if ((rc = S.rec_off.ensure(nr * 8)) || (rc = S.pass.ensure(nr * 4)) || (rc = S.pidx.ensure(nr * 4)) || (rc = S.r_pos.ensure(nr * 4)) ||
        (rc = S.r_end.ensure(nr * 4)) || (rc = S.cpos.ensure(nr * 4)) || (rc = S.kept.ensure(nr * 4)) || (rc = S.kidx.ensure(nr * 4)) ||
        (rc = S.ncig.ensure(nr * 4)) || (rc = S.coff.ensure(nr * 4)) || (rc = S.sqb.ensure(nr * 4)) || (rc = S.soff.ensure(nr * 4)) ||
        (rc = S.elig.ensure(nr * 4)) || (rc = S.eidx.ensure(nr * 4)) || (rc = S.prevpos.ensure(nr * 4)) || (rc = S.cend.ensure(nr * 4)) ||
        (C0.sc && (rc = S.rec_cb.ensure(nr * 4))))
        return rc;
    g->t_alloc += raer::now_seconds() - t4;
    const double t5 = raer::now_seconds();
    k_gw_walk<true><<<(nsp + 127) / 128, 128, 0, st>>>((const uint8_t*)S.ubuf.p, u_last, (const unsigned long long*)S.sp.p, nsp, nullptr,
                                                        (const uint32_t*)S.base.p, (recoff_t*)S.rec_off.p, d_flags);
    GpConf C = C0;
    C.tid = ftid;
    const unsigned gr = (unsigned)((nr + 255) / 256);
    k_gp_pass<<<gr, 256, 0, st>>>((const uint8_t*)S.ubuf.p, (const recoff_t*)S.rec_off.p, S.n_rec, C, (uint32_t*)S.pass.p,
                                  (int32_t*)S.r_pos.p, (int32_t*)S.r_end.p, d_small + 2, d_flags, (int32_t*)S.rec_cb.p);
    g->launches += 2;
    if ((rc = g_scan<false>(g, S, (const uint32_t*)S.pass.p, (uint32_t*)S.pidx.p, S.n_rec, d_small + 5))) return rc;
    k_gp_cpos<<<gr, 256, 0, st>>>((const uint32_t*)S.pass.p, (const uint32_t*)S.pidx.p, (const int32_t*)S.r_pos.p,
                                  (const int32_t*)S.r_end.p, S.n_rec, (int32_t*)S.cpos.p, (int32_t*)S.cend.p);
    ++g->launches;
    GCK(cudaMemcpyAsync(g->h_small, d_small, 64, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    if (g->h_small[1] & 1) { set_error("corrupt BAM record"); return RAER_ERR_IO; }
    if (g->h_small[1] & 64) { set_error("the input is not sorted (chromosomes out of order)"); return RAER_ERR_IO; }
    if (g->h_small[1] & 2) return RAER_GI_FALLBACK;  // a read beyond the documented limits: the host path reports it
    S.aligned_bases = (int64_t)g->h_small[2];
    S.max_span = (int)g->h_small[3];
    S.n_pass = (long long)g->h_small[5];
    g->t_pass += raer::now_seconds() - t5;
    const double t6 = raer::now_seconds();
    auto launch_kept = [&](const uint8_t* dropped) {
        k_gp_kept<<<gr, 256, 0, st>>>((const uint8_t*)S.ubuf.p, (const recoff_t*)S.rec_off.p, (const uint32_t*)S.pass.p,
                                      (const uint32_t*)S.pidx.p, (const int32_t*)S.r_pos.p, (const int32_t*)S.r_end.p,
                                      (const int32_t*)S.cpos.p, S.n_rec, S.n_pass, C, S.max_span, (uint32_t*)S.kept.p,
                                      (uint32_t*)S.ncig.p, (uint32_t*)S.sqb.p, (uint32_t*)S.elig.p, (int32_t*)S.prevpos.p, d_flags,
                                      dropped);
        ++g->launches;
    };
    launch_kept(nullptr);
    GCK(cudaMemcpyAsync(g->h_small, d_small, 64, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    if (g->h_small[1] & 4) { set_error("the input is not sorted (reads out of order)"); return RAER_ERR_IO; }
    if (g->h_small[1] & 8) {
        // The depth may reach max_depth somewhere on this contig. bam_plp_push then drops reads that start where the previous
        // one did, which depends on what was dropped before: the stream is replayed -- over two integers per read, on the
        // host -- and the verdicts go back to the device. (A dropped read also deletes the overlap entry of its name; with
        // flags that let secondary or supplementary records through a name can have a third record that would notice: host.)
        if (C.remove_overlaps && ((C.keep1 & 0x900u) != 0)) return RAER_GI_FALLBACK;
        std::vector<int32_t> hp((size_t)S.n_pass), he((size_t)S.n_pass);
        GCK(cudaMemcpyAsync(hp.data(), S.cpos.p, (size_t)S.n_pass * 4, cudaMemcpyDeviceToHost, st));
        GCK(cudaMemcpyAsync(he.data(), S.cend.p, (size_t)S.n_pass * 4, cudaMemcpyDeviceToHost, st));
        GCK(cudaStreamSynchronize(st));
        std::vector<uint8_t> drop((size_t)S.n_pass, 0);
        const long long n_drop = gi_replay_max_depth(hp.data(), he.data(), S.n_pass, C.max_depth, drop.data());
        g->n_dropped += n_drop;
        if ((rc = S.dropped.ensure((size_t)S.n_pass + 16))) return rc;
        GCK(cudaMemcpyAsync(S.dropped.p, drop.data(), (size_t)S.n_pass, cudaMemcpyHostToDevice, st));
        GCK(cudaMemsetAsync(d_flags, 0, 4, st));
        launch_kept((const uint8_t*)S.dropped.p);
        GCK(cudaStreamSynchronize(st));  // drop lives on this frame
    }
    if ((rc = g_scan<false>(g, S, (const uint32_t*)S.kept.p, (uint32_t*)S.kidx.p, S.n_rec, d_small + 4)) ||
        (rc = g_scan<false>(g, S, (const uint32_t*)S.ncig.p, (uint32_t*)S.coff.p, S.n_rec, d_small + 5)) ||
        (rc = g_scan<false>(g, S, (const uint32_t*)S.sqb.p, (uint32_t*)S.soff.p, S.n_rec, d_small + 6)) ||
        (rc = g_scan<false>(g, S, (const uint32_t*)S.elig.p, (uint32_t*)S.eidx.p, S.n_rec, d_small + 7)))
        return rc;
    GCK(cudaMemcpyAsync(g->h_small, d_small, 64, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    S.n_kept = (long long)g->h_small[4];
    S.n_cig = (long long)g->h_small[5];
    S.n_sq = (long long)g->h_small[6];
    S.n_elig = (long long)g->h_small[7];
    g->t_kept += raer::now_seconds() - t6;
    return RAER_OK;
}

static int gi_maxend(raer_gi* g, raer_read_batch* b);

static GpMeta gi_meta(FileStage& S) {
    GpMeta M;
    M.krec = (recoff_t*)S.krec.p; M.kprev = (int32_t*)S.kprev.p; M.kpos = (int32_t*)S.kpos.p; M.kend = (int32_t*)S.kend.p;
    M.kcoff = (uint32_t*)S.kcoff.p; M.ksoff = (uint32_t*)S.ksoff.p; M.kcb = (int32_t*)S.kcb.p; M.kumi = (uint32_t*)S.kumi.p;
    M.mateK = (int32_t*)S.mateK.p; M.keepend = (int32_t*)S.keepend.p; M.keepA = (uint8_t*)S.keepA.p;
    return M;
}

// Prepares contig `tid` (numbering of file 0; `ftid[f]` = the same contig in file f's header): every file's records are
// inflated, parsed, filtered and paired on the device, and the contig is cut into position windows whose read batches fit
// the 32-bit offsets of raer_read_hdr. site: the contig's intervals of the call's region list or nullptr;
// *site_mask_dev is the device bitmap of the listed positions (nullptr without a list). *n_windows batches are then
// fetched with raer_gi_window(), in order.
extern "C" int raer_gi_contig(raer_gi* g, int tid, const int* ftid, int contig_len, const raer::IngestConf* ic,
                              const raer::RegionIndex::Chr* site, int reg_beg, int reg_end, int* n_windows,
                              const uint32_t** site_mask_dev, int64_t* aligned_bases) {
    cudaSetDevice(g->device);
    cudaStream_t st = g->st;
    t_alloc_stream = st;
    const int nf = g->nf;
    int rc;
    *n_windows = 0;
    g->win.clear();
    GpConf C;
    memset(&C, 0, sizeof(C));
    C.keep0 = ic->keep_flag[0];
    C.keep1 = ic->keep_flag[1];
    C.min_global_mq = ic->min_global_mq;
    C.n_mm_type = ic->n_mm_type;
    C.n_mm = ic->n_mm;
    C.reg_on = reg_beg > 0 || reg_end < contig_len;  // [reg_beg, reg_end) = [0, INT_MAX) for a whole contig
    C.reg_beg = reg_beg;
    C.reg_end = reg_end;
    C.has_rq = ic->rq_pct && ic->rq_minq;
    C.rq_minq = ic->rq_minq;
    C.rq_pct = (float)ic->rq_pct;
    C.remove_overlaps = ic->remove_overlaps;
    C.min_overhang = ic->min_overhang;
    C.max_depth = ic->max_depth;
    if (ic->sc) {
        C.sc = 1;
        C.has_cb_tag = ic->cb_index && ic->cb_tag.size() == 2;
        C.has_umi_tag = ic->umi_tag.size() == 2;
        C.fixed_cb = ic->fixed_cb;
        if (C.has_cb_tag) { C.cb_tag[0] = (uint8_t)ic->cb_tag[0]; C.cb_tag[1] = (uint8_t)ic->cb_tag[1]; }
        if (C.has_umi_tag) { C.umi_tag[0] = (uint8_t)ic->umi_tag[0]; C.umi_tag[1] = (uint8_t)ic->umi_tag[1]; }
        if (C.has_cb_tag && !g->wl_built) {  // the whitelist as an open-addressing table, once per call
            const size_t n = ic->cb_index->size();
            size_t cap = 64;
            while (cap < 2 * n + 2) cap <<= 1;
            std::vector<unsigned long long> key(cap, 0);
            std::vector<int32_t> col(cap, 0);
            std::vector<uint32_t> off(cap, 0);
            std::vector<uint8_t> pool;
            for (const auto& kv : *ic->cb_index) {
                unsigned long long h = 1469598103934665603ull;
                for (unsigned char ch : kv.first) { if (!ch) break; h ^= ch; h *= 1099511628211ull; }
                if (!h) h = 1;
                size_t slot = (size_t)((uint32_t)h & (uint32_t)(cap - 1));
                while (key[slot]) slot = (slot + 1) & (cap - 1);
                key[slot] = h;
                col[slot] = kv.second;
                off[slot] = (uint32_t)pool.size();
                pool.insert(pool.end(), kv.first.begin(), kv.first.end());
                pool.push_back(0);
            }
            pool.push_back(0);
            if ((rc = g->wl_key.ensure(cap * 8)) || (rc = g->wl_col.ensure(cap * 4)) || (rc = g->wl_off.ensure(cap * 4)) ||
                (rc = g->wl_pool.ensure(pool.size())))
                return rc;
            GCK(cudaMemcpyAsync(g->wl_key.p, key.data(), cap * 8, cudaMemcpyHostToDevice, st));
            GCK(cudaMemcpyAsync(g->wl_col.p, col.data(), cap * 4, cudaMemcpyHostToDevice, st));
            GCK(cudaMemcpyAsync(g->wl_off.p, off.data(), cap * 4, cudaMemcpyHostToDevice, st));
            GCK(cudaMemcpyAsync(g->wl_pool.p, pool.data(), pool.size(), cudaMemcpyHostToDevice, st));
            GCK(cudaStreamSynchronize(st));
            g->wl_mask = (uint32_t)(cap - 1);
            g->wl_built = true;
        }
        C.wl_key = (const unsigned long long*)g->wl_key.p;
        C.wl_col = (const int32_t*)g->wl_col.p;
        C.wl_off = (const uint32_t*)g->wl_off.p;
        C.wl_pool = (const uint8_t*)g->wl_pool.p;
        C.wl_mask = g->wl_mask;
        if (!g->umi_len.p) {
            if ((rc = g->umi_len.ensure(16))) return rc;
            GCK(cudaMemsetAsync(g->umi_len.p, 0, 16, st));
        }
        C.umi_len = (int*)g->umi_len.p;
    }
    *site_mask_dev = nullptr;
    if (site) {
        const size_t words = ((size_t)contig_len + 31) / 32 + 1;
        std::vector<uint32_t> bits(words, 0), pre(words + 1, 0);
        for (size_t k = 0; k < site->beg.size(); ++k) {
            const int lo = std::max(site->beg[k], 0), hi = std::min(site->end[k], contig_len - 1);
            for (int p = lo; p <= hi; ++p) bits[(size_t)p >> 5] |= 1u << (p & 31);
        }
        for (size_t w = 0; w < words; ++w) pre[w + 1] = pre[w] + (uint32_t)__builtin_popcount(bits[w]);
        if ((rc = g->site_bits.ensure(words * 4)) || (rc = g->site_pre.ensure((words + 1) * 4))) return rc;
        GCK(cudaMemcpyAsync(g->site_bits.p, bits.data(), words * 4, cudaMemcpyHostToDevice, st));
        GCK(cudaMemcpyAsync(g->site_pre.p, pre.data(), (words + 1) * 4, cudaMemcpyHostToDevice, st));
        GCK(cudaStreamSynchronize(st));  // the vectors go out of scope
        C.site_bits = (const uint32_t*)g->site_bits.p;
        C.site_pre = (const uint32_t*)g->site_pre.p;
        C.site_len = contig_len;
        *site_mask_dev = C.site_bits;
    }
    long long n_reads = 0;
    unsigned long long sq_total = 0, sq_max = 0;
    int f_max = 0;
    *aligned_bases = 0;
    g->fconf.assign((size_t)nf, C);
    for (int f = 0; f < nf; ++f) {
        C.libtype = ic->libtype[(size_t)f];
        rc = gi_file_stage(g, f, ftid[f], C);
        if (rc) return rc;
        FileStage& S = g->stage[(size_t)f];
        S.k_lo_prev = 0;
        g->fconf[(size_t)f] = C;
        g->fconf[(size_t)f].tid = ftid[f];
        n_reads += S.n_kept;
        sq_total += (unsigned long long)S.n_sq;
        if ((unsigned long long)S.n_sq > sq_max) { sq_max = (unsigned long long)S.n_sq; f_max = f; }
        *aligned_bases += S.aligned_bases;
    }
    g->cur_tid = tid;
    g->cur_len = std::min(contig_len, reg_end);
    g->cur_beg = std::max(reg_beg, 0);
    g->cur_sc = ic->sc;
    g->cur_bases = *aligned_bases;
    if (n_reads == 0) return RAER_OK;
    // a file whose own prefix sums overflow 32 bits cannot be described by the per-read tables either
    for (int f = 0; f < nf; ++f)
        if (g->stage[(size_t)f].n_kept >= INT_MAX || g->stage[(size_t)f].n_cig >= (1ll << 32) || g->stage[(size_t)f].n_sq >= (1ll << 32))
            return RAER_GI_FALLBACK;
    unsigned long long* d_small = (unsigned long long*)g->small.p;
    GCK(cudaMemsetAsync(d_small, 0, 256, st));
    const double t7 = raer::now_seconds();
    for (int f = 0; f < nf; ++f) {
        FileStage& S = g->stage[(size_t)f];
        if (S.n_kept == 0) continue;
        const size_t nk = (size_t)S.n_kept, ne = (size_t)S.n_elig;
        if ((rc = S.krec.ensure(nk * 8)) || (rc = S.kprev.ensure(nk * 4)) || (rc = S.kpos.ensure(nk * 4)) || (rc = S.kend.ensure(nk * 4)) ||
            (rc = S.kcoff.ensure((nk + 1) * 4)) || (rc = S.ksoff.ensure((nk + 1) * 4)) || (rc = S.mateK.ensure(nk * 4)) ||
            (rc = S.keepend.ensure(nk * 4)) || (rc = S.keepA.ensure(nk)) ||
            (ic->sc && ((rc = S.kcb.ensure(nk * 4)) || (rc = S.kumi.ensure(nk * 4)))) || (rc = S.pkey.ensure(ne * 8 + 16)) ||
            (rc = S.pval.ensure(ne * 4 + 16)) || (rc = S.pkey2.ensure(ne * 8 + 16)) || (rc = S.pval2.ensure(ne * 4 + 16)))
            return rc;
        const GpMeta M = gi_meta(S);
        const uint32_t tot[2] = {(uint32_t)S.n_cig, (uint32_t)S.n_sq};
        GCK(cudaMemcpyAsync(M.kcoff + nk, &tot[0], 4, cudaMemcpyHostToDevice, st));
        GCK(cudaMemcpyAsync(M.ksoff + nk, &tot[1], 4, cudaMemcpyHostToDevice, st));
        GCK(cudaStreamSynchronize(st));  // tot lives on this frame
        const unsigned gr = (unsigned)((S.n_rec + 255) / 256);
        k_gp_meta<<<gr, 256, 0, st>>>((const uint8_t*)S.ubuf.p, (const recoff_t*)S.rec_off.p, (const uint32_t*)S.kept.p,
                                      (const uint32_t*)S.kidx.p, (const uint32_t*)S.coff.p, (const uint32_t*)S.soff.p,
                                      (const uint32_t*)S.elig.p, (const uint32_t*)S.eidx.p, (const int32_t*)S.r_pos.p,
                                      (const int32_t*)S.r_end.p, (const int32_t*)S.prevpos.p, S.n_rec, g->fconf[(size_t)f], M,
                                      (unsigned long long*)S.pkey.p, (uint32_t*)S.pval.p, (const int32_t*)S.rec_cb.p,
                                      (int*)(d_small + 1));
        ++g->launches;
        if (S.n_elig > 0) {
            const unsigned nblk = (unsigned)((S.n_elig + (long long)RS_BLOCK * RS_ITEMS - 1) / ((long long)RS_BLOCK * RS_ITEMS));
            if ((rc = S.hist.ensure((size_t)nblk * 256 * 4 + 1024))) return rc;
            unsigned long long* k_in = (unsigned long long*)S.pkey.p;
            unsigned* v_in = (unsigned*)S.pval.p;
            unsigned long long* k_out = (unsigned long long*)S.pkey2.p;
            unsigned* v_out = (unsigned*)S.pval2.p;
            g->launches += rs_sort(k_in, v_in, k_out, v_out, S.n_elig, 64, (unsigned*)S.hist.p, st);
            k_gp_pair<<<(unsigned)((S.n_elig + 255) / 256), 256, 0, st>>>(k_in, v_in, S.n_elig, (const uint8_t*)S.ubuf.p, M,
                                                                          (int*)(d_small + 1));
            ++g->launches;
        }
    }
    GCK(cudaMemcpyAsync(g->h_small, d_small, 64, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    if (cudaGetLastError() != cudaSuccess) { set_error("device ingest kernels failed"); return RAER_ERR_CUDA; }
    if (g->h_small[1] & 16) return RAER_GI_FALLBACK;  // two read names with one 64-bit hash: the host's string table decides
    if (g->h_small[1] & 256) return RAER_GI_FALLBACK;  // a UMI the 2-bit code does not cover: the host's dictionary takes it
    // ---- position windows: boundaries at the positions where the largest file's packed sequence crosses equal shares
    unsigned long long budget = 3ull << 29;  // packed sequence bytes per window, all files (the offsets are 31 bits)
    if (const char* e = getenv("RAER_GI_WINDOW_BYTES")) budget = std::max<unsigned long long>(4096, strtoull(e, nullptr, 10));
    const int n_cut = (int)std::min<unsigned long long>((sq_total + budget - 1) / budget, 1u << 20);
    g->win.push_back(g->cur_beg);
    if (n_cut > 1) {
        FileStage& S = g->stage[(size_t)f_max];
        const GpMeta M = gi_meta(S);
        for (int j = 1; j < n_cut; ++j) {
            const unsigned long long target = sq_max / (unsigned long long)n_cut * (unsigned long long)j;
            k_gp_pos_at_bytes<<<1, 1, 0, st>>>(M.kpos, M.ksoff, S.n_kept, target, d_small + 8);
            GCK(cudaMemcpyAsync(g->h_small, d_small + 8, 8, cudaMemcpyDeviceToHost, st));
            GCK(cudaStreamSynchronize(st));
            const unsigned long long w = g->h_small[0];
            if (w != ~0ull && (int)w > g->win.back() && (int)w < g->cur_len) g->win.push_back((int)w);
        }
        g->launches += n_cut - 1;
    }
    g->win.push_back(std::max(g->cur_len, g->win.back() + 1));
    if (g->win.size() > 2) {
        // the boundaries follow one file: make sure no window is too large over ALL files (a window that is gets cut in half,
        // before any row of the contig is out)
        auto bytes_before = [&](int w, unsigned long long* out) -> int {  // packed sequence of the reads that start before w
            unsigned long long tot = 0;
            for (int f = 0; f < nf; ++f) {
                FileStage& S = g->stage[(size_t)f];
                if (S.n_kept == 0) continue;
                const GpMeta M = gi_meta(S);
                k_gp_lower_pos<<<1, 1, 0, st>>>(M.kpos, S.n_kept, w, d_small + 8);
                GCK(cudaMemcpyAsync(g->h_small, d_small + 8, 8, cudaMemcpyDeviceToHost, st));
                GCK(cudaStreamSynchronize(st));
                uint32_t v = 0;
                GCK(cudaMemcpy(&v, M.ksoff + g->h_small[0], 4, cudaMemcpyDeviceToHost));
                tot += v;
                ++g->launches;
            }
            *out = tot;
            return 0;
        };
        const unsigned long long hard = std::min<unsigned long long>(7ull << 28, std::max<unsigned long long>(budget + budget / 4, 8192));
        std::vector<int> W = g->win;
        std::vector<unsigned long long> B(W.size(), 0);
        for (size_t k = 0; k < W.size(); ++k)
            if ((rc = bytes_before(k + 1 == W.size() ? INT_MAX : W[k], &B[k]))) return rc;
        for (size_t k = 0; k + 1 < W.size() && W.size() < (1u << 20);) {
            if (B[k + 1] - B[k] > hard && W[k + 1] - W[k] > 1) {
                const int mid = W[k] + (W[k + 1] - W[k]) / 2;
                unsigned long long bm = 0;
                if ((rc = bytes_before(mid, &bm))) return rc;
                W.insert(W.begin() + (long)k + 1, mid);
                B.insert(B.begin() + (long)k + 1, bm);
                continue;
            }
            ++k;
        }
        g->win = W;
    }
    *n_windows = (int)g->win.size() - 1;
    g->t_write += raer::now_seconds() - t7;
    return RAER_OK;
}

// The read batch of window j of the prepared contig: *b describes device arrays owned by g (valid until the next call).
// Windows have to be fetched in ascending order. A window no read reaches comes back with n_reads == 0.
extern "C" int raer_gi_window(raer_gi* g, int j, raer_read_batch* b, int64_t* file_off /* nf + 1 */) {
    cudaSetDevice(g->device);
    cudaStream_t st = g->st;
    t_alloc_stream = st;
    const int nf = g->nf;
    int rc;
    if (j < 0 || j + 1 >= (int)g->win.size()) { set_error("no such window"); return RAER_ERR_ARG; }
    const int w0 = g->win[(size_t)j], w1 = g->win[(size_t)j + 1];
    const bool last = j + 2 == (int)g->win.size();
    const double t7 = raer::now_seconds();
    unsigned long long* d_small = (unsigned long long*)g->small.p;
    // per file: [k_lo, k_hi) = the reads that start before the window ends, from the first one that is still needed
    std::vector<long long> k_lo((size_t)nf, 0), k_hi((size_t)nf, 0);
    for (int f = 0; f < nf; ++f) {
        FileStage& S = g->stage[(size_t)f];
        if (S.n_kept == 0) continue;
        const GpMeta M = gi_meta(S);
        if (last) {
            k_hi[(size_t)f] = S.n_kept;
        } else {
            k_gp_lower_pos<<<1, 1, 0, st>>>(M.kpos, S.n_kept, w1, d_small + 8);
            GCK(cudaMemcpyAsync(g->h_small, d_small + 8, 8, cudaMemcpyDeviceToHost, st));
            GCK(cudaStreamSynchronize(st));
            k_hi[(size_t)f] = (long long)g->h_small[0];
        }
        const long long from = S.k_lo_prev, to = k_hi[(size_t)f];
        k_lo[(size_t)f] = to;
        if (to > from) {
            const unsigned long long init = (unsigned long long)to;
            GCK(cudaMemcpyAsync(d_small + 9, &init, 8, cudaMemcpyHostToDevice, st));
            GCK(cudaStreamSynchronize(st));
            k_gp_first_live<<<(unsigned)((to - from + 255) / 256), 256, 0, st>>>(M.keepend, from, to, w0, d_small + 9);
            GCK(cudaMemcpyAsync(g->h_small, d_small + 9, 8, cudaMemcpyDeviceToHost, st));
            GCK(cudaStreamSynchronize(st));
            k_lo[(size_t)f] = (long long)g->h_small[0];
        }
        S.k_lo_prev = k_lo[(size_t)f];
        g->launches += 2;
    }
    // sizes of the window: prefix sums at k_lo / k_hi
    long long n_reads = 0, n_cig = 0, n_sq = 0;
    std::vector<long long> c_of((size_t)nf, 0), s_of((size_t)nf, 0);
    for (int f = 0; f < nf; ++f) {
        FileStage& S = g->stage[(size_t)f];
        file_off[f] = n_reads;
        if (S.n_kept == 0 || k_hi[(size_t)f] <= k_lo[(size_t)f]) continue;
        const GpMeta M = gi_meta(S);
        uint32_t v[4];
        GCK(cudaMemcpyAsync(&v[0], M.kcoff + k_lo[(size_t)f], 4, cudaMemcpyDeviceToHost, st));
        GCK(cudaMemcpyAsync(&v[1], M.kcoff + k_hi[(size_t)f], 4, cudaMemcpyDeviceToHost, st));
        GCK(cudaMemcpyAsync(&v[2], M.ksoff + k_lo[(size_t)f], 4, cudaMemcpyDeviceToHost, st));
        GCK(cudaMemcpyAsync(&v[3], M.ksoff + k_hi[(size_t)f], 4, cudaMemcpyDeviceToHost, st));
        GCK(cudaStreamSynchronize(st));
        c_of[(size_t)f] = n_cig;
        s_of[(size_t)f] = n_sq;
        n_reads += k_hi[(size_t)f] - k_lo[(size_t)f];
        n_cig += (long long)(v[1] - v[0]);
        n_sq += (long long)(v[3] - v[2]);
    }
    file_off[nf] = n_reads;
    memset(b, 0, sizeof(*b));
    b->n_files = nf;
    b->tid = g->cur_tid;
    b->beg = w0;
    b->end = last ? std::max(g->cur_len, w0 + 1) : w1;
    b->n_reads = n_reads;
    b->file_off = file_off;
    if (n_reads == 0) return RAER_OK;
    if (n_reads >= INT_MAX || n_cig >= (1ll << 32) || n_sq >= (1ll << 31)) {
        // a window denser than the plan assumed (the boundaries follow the largest file only): not recoverable here, the
        // rows of earlier windows are out already
        set_error("device ingest: a window of %lld sequence bytes exceeds the 32-bit offsets of the read batch; set "
                  "RAER_GI_WINDOW_BYTES lower", n_sq);
        return RAER_ERR_LIMIT;
    }
    if ((rc = g->hdr.ensure((size_t)n_reads * sizeof(raer_read_hdr))) || (rc = g->maxend.ensure((size_t)n_reads * 4)) ||
        (rc = g->endtmp.ensure((size_t)n_reads * 4)) || (rc = g->cigar.ensure((size_t)n_cig * 4 + 16)) ||
        (rc = g->seq.ensure((size_t)n_sq + 16)) || (rc = g->qual.ensure((size_t)n_sq * 2 + 16)) ||
        (rc = g->pair_b.ensure((size_t)n_reads * 4 + 16)) ||
        (g->cur_sc && ((rc = g->cb.ensure((size_t)n_reads * 4)) || (rc = g->umi.ensure((size_t)n_reads * 4)))))
        return rc;
    GCK(cudaMemsetAsync(d_small + 4, 0, 8, st));
    for (int f = 0; f < nf; ++f) {
        FileStage& S = g->stage[(size_t)f];
        const long long n = k_hi[(size_t)f] - k_lo[(size_t)f];
        if (S.n_kept == 0 || n <= 0) continue;
        GpOut O;
        O.hdr = (raer_read_hdr*)g->hdr.p;
        O.cigar = (uint32_t*)g->cigar.p;
        O.seq = (uint8_t*)g->seq.p;
        O.qual = (uint8_t*)g->qual.p;
        O.cb = (int32_t*)g->cb.p;
        O.umi = (uint32_t*)g->umi.p;
        O.pair_b = (int32_t*)g->pair_b.p;
        O.n_pairs = d_small + 4;
        O.r0 = file_off[f];
        O.c0 = c_of[(size_t)f];
        O.s0 = s_of[(size_t)f];
        k_gp_batch<<<(unsigned)((n * GB_LANES + 255) / 256), 256, 0, st>>>((const uint8_t*)S.ubuf.p, gi_meta(S), k_lo[(size_t)f], k_hi[(size_t)f],
                                                               g->fconf[(size_t)f], O);
        ++g->launches;
    }
    GCK(cudaMemcpyAsync(g->h_small, d_small, 64, cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    if (cudaGetLastError() != cudaSuccess) { set_error("device ingest kernels failed"); return RAER_ERR_CUDA; }
    b->hdr = (const raer_read_hdr*)g->hdr.p;
    b->cigar = (const uint32_t*)g->cigar.p;
    b->seq = (const uint8_t*)g->seq.p;
    b->qual = (const uint8_t*)g->qual.p;
    b->n_cigar_words = n_cig;
    b->n_sq_bytes = n_sq;
    b->pair_b = (const int32_t*)g->pair_b.p;
    b->n_pairs = (int64_t)g->h_small[4];
    b->n_aligned_bases = j == 0 ? g->cur_bases : 0;
    if (g->cur_sc) {
        b->cb = (const int32_t*)g->cb.p;
        b->umi = (const uint32_t*)g->umi.p;
    }
    rc = gi_maxend(g, b);
    g->t_write += raer::now_seconds() - t7;
    return rc;
}

__global__ void k_gp_ends(const raer_read_hdr* __restrict__ hdr, long long n, uint32_t* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)max(hdr[i].end, 0);
}
// running maximum of hdr.end inside every file range (what k_tile_ranges of the generic tile kernel searches)
static int gi_maxend(raer_gi* g, raer_read_batch* b) {
    cudaSetDevice(g->device);
    for (int f = 0; f < b->n_files; ++f) {
        const long long lo = b->file_off[f], n = b->file_off[f + 1] - lo;
        if (n <= 0) continue;
        k_gp_ends<<<(unsigned)((n + 255) / 256), 256, 0, g->st>>>(b->hdr + lo, n, (uint32_t*)g->endtmp.p + lo);
        int rc = g_scan<true>(g, g->stage[0], (const uint32_t*)g->endtmp.p + lo, (uint32_t*)g->maxend.p + lo, n, nullptr);
        if (rc) return rc;
    }
    GCK(cudaStreamSynchronize(g->st));
    b->maxend = (const int32_t*)g->maxend.p;
    return RAER_OK;
}

// single cell: number of ids the UMI code can take (4^length of the UMIs seen so far), 0 without UMIs
extern "C" uint32_t raer_gi_umi_space(raer_gi* g) {
    if (!g->umi_len.p) return 0;
    int len = 0;
    cudaSetDevice(g->device);
    if (cudaMemcpy(&len, g->umi_len.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (len <= 0) return 0;
    return len >= 15 ? (1u << 30) : (1u << (2 * len));
}

extern "C" void raer_trim_device_cache(int32_t device) {
    cudaMemPool_t pool;
    if (cudaSetDevice(device) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        cudaDeviceSynchronize();
        cudaMemPoolTrimTo(pool, 0);
    }
    cudaGetLastError();
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (uint8_t* p : g_pin_pool) cudaFreeHost(p);
    g_pin_pool.clear();
}

extern "C" void raer_gi_stats(raer_gi* g, int64_t* launches, int64_t* inflated, int64_t* compressed, int64_t* records) {
    *launches = g->launches;
    *inflated = g->inflated;
    *compressed = g->compressed;
    *records = g->records;
}

// test / bench entry: inflate n raw DEFLATE payloads that lie in one host buffer; status[i] = 0 or the decoder's code
extern "C" int raer_gpu_inflate(const uint8_t* comp, int64_t comp_len, const int64_t* c_off, const int32_t* c_len, const int64_t* u_off,
                                const int32_t* u_len, const uint32_t* crc, int32_t n_blocks, uint8_t* out, int64_t out_len,
                                int32_t* status, double* kernel_ms) {
    if (n_blocks <= 0) return RAER_OK;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); set_error("no usable CUDA device (the product has no CPU path)"); return RAER_ERR_NO_DEVICE; }
    std::vector<GiBlock> blk((size_t)n_blocks);
    for (int i = 0; i < n_blocks; ++i) {
        blk[(size_t)i].c_off = (unsigned long long)c_off[i];
        blk[(size_t)i].c_len = (uint32_t)c_len[i];
        blk[(size_t)i].u_off = (unsigned long long)u_off[i];
        blk[(size_t)i].u_len = (uint32_t)u_len[i];
        blk[(size_t)i].crc = crc ? crc[i] : 0u;
        blk[(size_t)i].check_crc = crc ? 1u : 0u;
    }
    DBufG d_comp, d_blk, d_out, d_status, d_q;
    int rc;
    if ((rc = d_comp.ensure((size_t)comp_len + 64)) || (rc = d_blk.ensure(blk.size() * sizeof(GiBlock))) ||
        (rc = d_out.ensure((size_t)out_len + 64)) || (rc = d_status.ensure((size_t)n_blocks * 4)) || (rc = d_q.ensure(64)))
        return rc;
    GCK(cudaMemset(d_comp.p, 0, (size_t)comp_len + 64));
    GCK(cudaMemcpy(d_comp.p, comp, (size_t)comp_len, cudaMemcpyHostToDevice));
    GCK(cudaMemcpy(d_blk.p, blk.data(), blk.size() * sizeof(GiBlock), cudaMemcpyHostToDevice));
    GCK(cudaMemset(d_q.p, 0, 64));
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    const int n_sm = cudaGetDeviceProperties(&prop, dev) == cudaSuccess ? prop.multiProcessorCount : 148;
    cudaFuncSetAttribute(k_gi_inflate, cudaFuncAttributeMaxDynamicSharedMemorySize, GI_SMEM_BYTES);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = std::min<int>(n_sm, (n_blocks + GI_WARPS - 1) / GI_WARPS);
    cudaEventRecord(e0);
    k_gi_inflate<<<grid, GI_WARPS * 32, GI_SMEM_BYTES>>>((const uint8_t*)d_comp.p, (const GiBlock*)d_blk.p, n_blocks,
                                                                       (uint8_t*)d_out.p, (int*)d_status.p, (unsigned*)d_q.p);
    cudaEventRecord(e1);
    GCK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms;
    GCK(cudaMemcpy(out, d_out.p, (size_t)out_len, cudaMemcpyDeviceToHost));
    GCK(cudaMemcpy(status, d_status.p, (size_t)n_blocks * 4, cudaMemcpyDeviceToHost));
    return RAER_OK;
}
