// raer_sc.cu -- single-cell pileup path (pileup_cells): .Call(".scpileup") equivalent.
//
// Reference: src/sc-plp.c (scpileup :945, run_scpileup :612, screadaln :449, process_one_site :528,
// count_record :235, count_consensus_base :321, write_counts :410, write_all_sites :852).
//
// Device pipeline per read batch (sites are sparse, so the unit of work is a read, not a base):
//   k_sc_emit      one thread per read: binary-search the sorted site list for positions inside the
//                  read's span, locate the query base through the CIGAR, apply the shared read/base
//                  filters, and emit one tuple (site, cell barcode, UMI | class, base quality) per
//                  matching payload. One pass: a warp reserves room for its tuples with one atomic; a
//                  buffer that turns out too small re-runs the batch with the size the device counted.
//   rs_sort        LSD radix sort of the 64-bit tuple keys over the used key bits (raer_radix.cuh:
//                  8-bit digits, stable block scatter).
//   k_sc_umi/cell  segmented reduce over the sorted tuples: (site, CB, UMI) quality sums -> consensus
//                  vote (ties go to alt, src/sc-plp.c:338) -> (site, CB) nRef/nAlt + per-site totals;
//                  k_cp_* compact the nonzero cells in order.
// Batch layer (raer_sc_ctx, raer_sc_batch_device / _host): the C-ABI entry bench.py drives.
// The host applies the min_counts / min_variant_reads site rule (src/sc-plp.c:586-596) and writes the
// three text files the R side reads back.
#define RAER_HELPERS_ONLY
#include "raer_kernels.cu"

#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "raer_host.h"

using raer::set_error;

#define SCK(call)                                                                             \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            set_error("CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__);  \
            return RAER_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

struct ScSites {  // sites of one contig, sorted by position (device pointers)
    const int32_t* pos;      // 0-based
    const uint8_t* strand;   // 1 '+', 2 '-'
    const uint8_t* ref;      // first character of REF
    const uint8_t* alt;      // first character of ALT
    int n;
};

struct ScParams {
    PlpParams P;   // batch pointers + filter configuration (shared helpers read it)
    ScSites S;
    const int32_t* cb;
    const uint32_t* umi;
    long long n_reads;
    int win_beg, win_end;
    int has_umi;
    int bits_cb, bits_umi;
    unsigned long long* counter;
    unsigned long long* keys;
    unsigned int* vals;
    unsigned long long cap;
};

template <bool WRITE>
__global__ void k_sc_emit(ScParams A) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= A.n_reads) return;
    const PlpParams& P = A.P;
    const int4* h = reinterpret_cast<const int4*>(P.hdr + r);
    int4 a = __ldg(h);
    const int pos = a.x, end = a.y;
    if (end <= A.win_beg || pos >= A.win_end) return;
    const int cb = A.cb[r];
    if (cb <= 0) return;                     // count_record: CB missing / not in the whitelist (src/sc-plp.c:248-257)
    const unsigned umi = A.has_umi ? A.umi[r] : 0u;
    // first site at or after the read start
    int lo = 0, hi = A.S.n;
    const int from = max(pos, A.win_beg);
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (A.S.pos[mid] < from) lo = mid + 1; else hi = mid;
    }
    int k = lo;
    const int upto = min(end, A.win_end);
    if (k >= A.S.n || A.S.pos[k] >= upto) return;
    int4 b = __ldg(h + 1);
    ReadCtx c;
    c.pos = pos; c.end = end;
    c.flag = a.z & 0xffff; c.mapq = (a.z >> 16) & 0xff; c.bits = (a.z >> 24) & 0xff;
    c.l_qseq = a.w & 0xffff; c.n_cigar = (a.w >> 16) & 0xffff;
    c.cig = P.cigar + (unsigned)b.x;
    c.seq = P.seq + (unsigned)b.y;
    c.qual = P.qual + 2ull * (unsigned)b.y;
    finish_ctx(P, c);
    const int mapq_fail = c.mapq < P.min_mapq[0];
    const int strand = c.invert ? 2 : 1;
    while (k < A.S.n) {
        const int p = A.S.pos[k];
        if (p >= upto) break;
        int k_end = k + 1;
        while (k_end < A.S.n && A.S.pos[k_end] == p) ++k_end;
        // locate p in the read (resolve_cigar2: deletions and ref-skips give is_del, filtered with code 2)
        int x = pos, y = 0, q = -1;
        for (int i = 0; i < c.n_cigar; ++i) {
            uint32_t w = c.cig[i];
            int op = cg_op(w), len = cg_len(w);
            if (op == OP_M || op == OP_EQ || op == OP_X) {
                if (p < x + len) { q = y + (p - x); break; }
                x += len; y += len;
            } else if (op == OP_D || op == OP_N) {
                if (p < x + len) break;
                x += len;
            } else if (op == OP_I || op == OP_S) {
                y += len;
            }
        }
        if (q >= 0) {
            // sc process_one_site does not skip read base N (src/sc-plp.c:545-560); qpos beyond l_qseq reads as N / q0
            int code = q < c.l_qseq ? seq_code(c.seq, q) : 15;
            int bq = q < c.l_qseq ? c.qual[q] : 0;
            int v = base_verdict(P, c, mapq_fail, q, bq);
            if (v == 0 && !(A.has_umi && umi == 0)) {
                int letter = "=ACMGRSVTWYHKDBN"[code];
                if (c.invert) letter = d_comp_char(letter);
                for (int kk = k; kk < k_end; ++kk) {
                    if (A.S.strand[kk] != strand) continue;
                    int cls = letter == A.S.ref[kk] ? 0 : (letter == A.S.alt[kk] ? 1 : 2);
                    if (!A.has_umi && cls == 2) continue;  // no UMI: "other" bases are not tallied (src/sc-plp.c:300-308)
                    // one atomic per group of lanes that arrive here together
                    const unsigned am = __activemask();
                    const int leader = __ffs(am) - 1, ln = threadIdx.x & 31;
                    unsigned long long idx = 0;
                    if (ln == leader) idx = atomicAdd(A.counter, (unsigned long long)__popc(am));
                    idx = __shfl_sync(am, idx, leader) + __popc(am & ((1u << ln) - 1u));
                    if (WRITE && idx < A.cap) {
                        A.keys[idx] = ((unsigned long long)kk << (A.bits_cb + A.bits_umi)) |
                                      ((unsigned long long)cb << A.bits_umi) | (unsigned long long)umi;
                        A.vals[idx] = ((unsigned)cls << 8) | (unsigned)bq;
                    }
                }
            }
        }
        k = k_end;
    }
}

#include "raer_radix.cuh"  // k_rs_*: LSD radix sort of 64-bit keys with 32-bit values

// ------------------------------------------------------------------------------------- segmented reduce
// pass 1: one vote per (site, CB, UMI) segment (UMI mode) or per read (no UMI): 1 = ref, 2 = alt, 0 = none
__global__ void k_sc_umi(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals, long long n,
                         int has_umi, unsigned char* __restrict__ vote) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!has_umi) {
        unsigned cls = vals[i] >> 8;
        vote[i] = cls == 0 ? 1 : (cls == 1 ? 2 : 0);
        return;
    }
    unsigned long long k = keys[i];
    if (i > 0 && keys[i - 1] == k) { vote[i] = 0; return; }
    int ref = 0, alt = 0, oth = 0;
    for (long long j = i; j < n && keys[j] == k; ++j) {
        unsigned v = vals[j];
        int bq = (int)(v & 0xff);
        unsigned cls = v >> 8;
        if (cls == 0) ref += bq; else if (cls == 1) alt += bq; else oth += bq;
    }
    // count_consensus_base, src/sc-plp.c:321-348
    unsigned char r = 0;
    if (!((oth > alt) && (oth > ref)) && !(alt == 0 && ref == 0)) r = alt >= ref ? 2 : 1;
    vote[i] = r;
}

// pass 2: one thread per (site, CB) segment head sums the votes; a cell with a nonzero count is flagged and its
// counts parked at the head (the ordered compaction below turns the flags into output slots)
__global__ void k_sc_cell(const unsigned long long* __restrict__ keys, const unsigned char* __restrict__ vote, long long n,
                          int bits_umi, unsigned* __restrict__ flag, int2* __restrict__ cnt, int bits_cb, int2* __restrict__ site_tot) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long cell = keys[i] >> bits_umi;
    if (i > 0 && (keys[i - 1] >> bits_umi) == cell) { flag[i] = 0; return; }
    int nref = 0, nalt = 0;
    for (long long j = i; j < n && (keys[j] >> bits_umi) == cell; ++j) {
        unsigned char v = vote[j];
        nref += v == 1;
        nalt += v == 2;
    }
    int site = (int)(cell >> bits_cb);
    if (nref) atomicAdd(&site_tot[site].x, nref);
    if (nalt) atomicAdd(&site_tot[site].y, nalt);
    flag[i] = (nref || nalt) ? 1u : 0u;
    cnt[i] = make_int2(nref, nalt);
}

// ordered compaction of the flagged heads: block sums, one-block scan of the sums, block-local scan + write. The
// cells come out sorted by (site, barcode) because the tuples are.
#define CP_BLOCK 256
#define CP_ITEMS 16
__global__ void k_cp_sum(const unsigned* __restrict__ flag, long long n, unsigned* __restrict__ bsum) {
    __shared__ unsigned s[CP_BLOCK / 32];
    const long long base = (long long)blockIdx.x * CP_BLOCK * CP_ITEMS + (long long)threadIdx.x * CP_ITEMS;
    unsigned v = 0;
    for (int k = 0; k < CP_ITEMS; ++k)
        if (base + k < n) v += flag[base + k];
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < CP_BLOCK / 32; ++w) t += s[w];
        bsum[blockIdx.x] = t;
    }
}
// exclusive scan of nb block sums in place (one CTA); the total goes to *total
__global__ void k_cp_scan(unsigned* bsum, int nb, unsigned long long* total) {
    __shared__ unsigned part[256];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 256) {
        int i = base + threadIdx.x;
        unsigned v = i < nb ? bsum[i] : 0;
        unsigned inc = block_incl_scan_256(v, part);
        if (i < nb) bsum[i] = carry + inc - v;
        __syncthreads();
        if (threadIdx.x == 255) carry += inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void k_cp_write(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ flag,
                           const int2* __restrict__ cnt, long long n, const unsigned* __restrict__ bsum, int bits_umi, int bits_cb,
                           int4* __restrict__ out) {
    __shared__ unsigned part[CP_BLOCK];
    const long long base = (long long)blockIdx.x * CP_BLOCK * CP_ITEMS + (long long)threadIdx.x * CP_ITEMS;
    unsigned v = 0;
    for (int k = 0; k < CP_ITEMS; ++k)
        if (base + k < n) v += flag[base + k];
    unsigned inc = block_incl_scan_256(v, part);
    unsigned o = bsum[blockIdx.x] + inc - v;
    for (int k = 0; k < CP_ITEMS; ++k) {
        const long long i = base + k;
        if (i < n && flag[i]) {
            const unsigned long long cell = keys[i] >> bits_umi;
            const int2 c = cnt[i];
            out[o++] = make_int4((int)(cell >> bits_cb), (int)(cell & ((1ull << bits_cb) - 1ull)), c.x, c.y);
        }
    }
}

// ------------------------------------------------------------------------------------- host side
extern "C" int raer_launch_overlap(const raer_read_hdr* hdr, const uint32_t* cigar, const uint8_t* seq, uint8_t* qual,
                                   const int32_t* pair_b, long long n_pairs, int mode, cudaStream_t st, int* scratch);
extern "C" int raer_device_count(void);

struct DBuf {  // stream-ordered pool allocation on the owning context's stream (see DevBuf in raer_runtime.cu)
    void* p = nullptr;
    size_t cap = 0;
    cudaStream_t st = nullptr;
    int ensure(size_t n) {
        if (n <= cap && p) return 0;
        if (p) cudaFreeAsync(p, st);
        size_t want = n + n / 4 + 256;
        if (cudaMallocAsync(&p, want, st) != cudaSuccess) { p = nullptr; cap = 0; cudaGetLastError(); set_error("cudaMallocAsync(%zu) failed", want); return RAER_ERR_CUDA; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeAsync(p, st); p = nullptr; cap = 0; }
};

static int bits_for(unsigned long long maxval) {
    int b = 1;
    while ((maxval >> b) != 0) ++b;
    return b;
}

// ------------------------------------------------------------------------------------- batch engine
struct raer_sc_ctx {
    int device = 0;
    cudaStream_t st = nullptr;
    DBuf hdr, cigar, seq, qual, pair, cb, umi, spos, sstr, sref, salt, small, keys, vals, keys2, vals2, hist, vote, flag, cnt, bsum,
        cells, tot, ov;
    unsigned long long tup_cap = 0;  // tuples the key / value buffers hold
    // results of the last batch (device): cells sorted by (site, barcode), totals per site
    long long n_cells = 0, n_tuples = 0;
    int n_sites = 0;
    int64_t launches = 0;
    // optional per-kernel-class timing
    bool prof = false;
    std::vector<cudaEvent_t> ev;       // pairs
    std::vector<int> ev_class;
    size_t ev_used = 0;
    unsigned long long* h_small = nullptr;  // pinned: tuple counter | cell counter
    std::vector<DBuf*> bufs() {
        return {&hdr, &cigar, &seq, &qual, &pair, &cb, &umi, &spos, &sstr, &sref, &salt, &small, &keys, &vals, &keys2, &vals2, &hist,
                &vote, &flag, &cnt, &bsum, &cells, &tot, &ov};
    }
    ~raer_sc_ctx() {
        for (DBuf* b : bufs()) b->release();
        for (auto e : ev) cudaEventDestroy(e);
        if (h_small) cudaFreeHost(h_small);
        if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    }
};

extern "C" raer_sc_ctx* raer_sc_ctx_create(int32_t device) {
    const int n = raer_device_count();
    if (n <= 0 || device >= n) { set_error("no usable CUDA device (the product has no CPU path)"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    raer_sc_ctx* c = new raer_sc_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaHostAlloc((void**)&c->h_small, 64, cudaHostAllocDefault) != cudaSuccess) {
        delete c;
        set_error("stream / pinned memory");
        return nullptr;
    }
    for (DBuf* b : c->bufs()) b->st = c->st;
    return c;
}
extern "C" void raer_sc_ctx_destroy(raer_sc_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    delete c;
}
extern "C" void* raer_sc_ctx_stream(raer_sc_ctx* c) { return (void*)c->st; }
extern "C" void raer_sc_ctx_profile(raer_sc_ctx* c, int on) { c->prof = on != 0; c->ev_used = 0; }
extern "C" int raer_sc_ctx_profile_read(raer_sc_ctx* c, double* ms, int64_t* n) {
    for (int k = 0; k < 4; ++k) { ms[k] = 0; n[k] = 0; }
    for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
        float t = 0;
        if (cudaEventElapsedTime(&t, c->ev[i], c->ev[i + 1]) != cudaSuccess) { cudaGetLastError(); return RAER_ERR_CUDA; }
        ms[c->ev_class[i / 2]] += t;
        n[c->ev_class[i / 2]] += 1;
    }
    return RAER_OK;
}
extern "C" int64_t raer_sc_algorithmic_bytes(const raer_read_batch* b, int64_t n_cells) {
    // the bulk record model plus 16 B per read (CB index, UMI id, padding) and 16 B per output cell
    const int64_t bases = b->n_sq_bytes * 2;
    return (32 + 16) * b->n_reads + 4 * b->n_cigar_words + bases / 2 + bases + 16 * n_cells;
}
extern "C" void raer_sc_cells_free(raer_sc_cells* c) {
    if (!c) return;
    free(c->cells);
    free(c->site_totals);
    memset(c, 0, sizeof(*c));
}

struct ScTimer {  // brackets a group of launches with two events when profiling is on
    raer_sc_ctx* c;
    ScTimer(raer_sc_ctx* c_, int cls) : c(c_) {
        if (!c->prof) return;
        while (c->ev.size() < c->ev_used + 2) { cudaEvent_t e; cudaEventCreate(&e); c->ev.push_back(e); }
        if (c->ev_class.size() < c->ev.size() / 2) c->ev_class.resize(c->ev.size() / 2);
        c->ev_class[c->ev_used / 2] = cls;
        cudaEventRecord(c->ev[c->ev_used], c->st);
    }
    ~ScTimer() {
        if (!c->prof) return;
        cudaEventRecord(c->ev[c->ev_used + 1], c->st);
        c->ev_used += 2;
    }
};

// The kernels of one batch, everything in device memory: mate overlap, tuple emission, radix sort, consensus votes,
// cell reduction with ordered compaction. One host wait at the end (tuple and cell counters); a tuple buffer that
// turns out too small re-runs the batch with the size the device counted. Results stay in c->cells / c->tot.
static int sc_device(raer_sc_ctx* c, const raer_sc_conf* conf, const raer_read_batch* b, const int32_t* d_spos,
                     const uint8_t* d_sstr, const uint8_t* d_sref, const uint8_t* d_salt) {
    cudaStream_t st = c->st;
    int rc;
    const int ns = conf->n_sites;
    c->n_cells = c->n_tuples = 0;
    c->n_sites = ns;
    if (b->n_reads <= 0 || ns <= 0 || b->end <= b->beg) return RAER_OK;
    if (!b->cb || (conf->has_umi && !b->umi)) { set_error("single-cell batch without cb / umi arrays"); return RAER_ERR_ARG; }
    const size_t nr = (size_t)b->n_reads;
    if ((rc = c->small.ensure(64)) || (rc = c->tot.ensure((size_t)ns * 8))) return rc;
    if (conf->remove_overlaps && b->n_pairs > 0) {
        ScTimer t(c, 3);
        if ((rc = c->ov.ensure(((size_t)b->n_pairs + 1) * sizeof(int)))) return rc;
        int r = raer_launch_overlap(b->hdr, b->cigar, b->seq, (uint8_t*)b->qual, b->pair_b, b->n_pairs, conf->remove_overlaps, st,
                                    (int*)c->ov.p);
        if (r < 0) { set_error("overlap kernel launch failed"); return RAER_ERR_CUDA; }
        c->launches += r;
    }
    ScParams A;
    memset(&A, 0, sizeof(A));
    A.P.hdr = b->hdr; A.P.cigar = b->cigar; A.P.seq = b->seq; A.P.qual = (uint8_t*)b->qual;
    A.P.min_bq = conf->min_bq; A.P.trim_5p = conf->trim_5p; A.P.trim_3p = conf->trim_3p;
    A.P.indel_dist = conf->indel_dist; A.P.splice_dist = conf->splice_dist; A.P.min_overhang = conf->min_overhang;
    A.P.ftrim_5p = conf->ftrim_5p; A.P.ftrim_3p = conf->ftrim_3p;
    A.P.min_mapq[0] = conf->min_mapq;
    A.S.pos = d_spos; A.S.strand = d_sstr; A.S.ref = d_sref; A.S.alt = d_salt; A.S.n = ns;
    A.cb = b->cb; A.umi = b->umi;
    A.n_reads = b->n_reads; A.win_beg = b->beg; A.win_end = b->end;
    A.has_umi = conf->has_umi;
    A.bits_cb = bits_for((unsigned long long)conf->n_barcodes + 1);
    A.bits_umi = conf->has_umi ? bits_for((unsigned long long)conf->n_umi + 1) : 1;
    const int bits_site = bits_for((unsigned long long)ns);
    if (bits_site + A.bits_cb + A.bits_umi > 64) { set_error("site x barcode x UMI key exceeds 64 bits"); return RAER_ERR_LIMIT; }
    const int key_bits = bits_site + A.bits_cb + A.bits_umi;
    unsigned long long want = std::max<unsigned long long>(c->tup_cap, std::max<unsigned long long>(2 * nr, 1ull << 16));
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (want > c->tup_cap) {
            if ((rc = c->keys.ensure(want * 8)) || (rc = c->vals.ensure(want * 4)) || (rc = c->keys2.ensure(want * 8)) ||
                (rc = c->vals2.ensure(want * 4)) || (rc = c->vote.ensure(want)) || (rc = c->flag.ensure(want * 4)) ||
                (rc = c->cnt.ensure(want * 8)) || (rc = c->cells.ensure(want * 16)))
                return rc;
            c->tup_cap = want;
        }
        SCK(cudaMemsetAsync(c->small.p, 0, 64, st));
        SCK(cudaMemsetAsync(c->tot.p, 0, (size_t)ns * 8, st));
        A.counter = (unsigned long long*)c->small.p;
        A.keys = (unsigned long long*)c->keys.p; A.vals = (unsigned*)c->vals.p; A.cap = c->tup_cap;
        {
            ScTimer t(c, 0);
            k_sc_emit<true><<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(A);
            ++c->launches;
        }
        // the sort needs the number of tuples on the host: the one wait of the batch that cannot move to its end
        SCK(cudaMemcpyAsync(c->h_small, c->small.p, 8, cudaMemcpyDeviceToHost, st));
        SCK(cudaStreamSynchronize(st));
        const unsigned long long ntup = c->h_small[0];
        if (ntup > c->tup_cap) { want = ntup + ntup / 8; continue; }  // the buffers were too small: exact size now
        c->n_tuples = (long long)ntup;
        if (ntup == 0) return RAER_OK;
        const unsigned nblk = (unsigned)((ntup + (unsigned long long)RS_BLOCK * RS_ITEMS - 1) / ((unsigned long long)RS_BLOCK * RS_ITEMS));
        if ((rc = c->hist.ensure((size_t)nblk * 256 * 4 + 1024))) return rc;
        unsigned long long* k_in = (unsigned long long*)c->keys.p; unsigned* v_in = (unsigned*)c->vals.p;
        unsigned long long* k_out = (unsigned long long*)c->keys2.p; unsigned* v_out = (unsigned*)c->vals2.p;
        unsigned* d_dtot = (unsigned*)c->hist.p + (size_t)nblk * 256;
        {
            ScTimer t(c, 1);
            // LSD radix sort over the used key bits
            for (int shift = 0; shift < key_bits; shift += 8) {
                k_rs_hist<<<nblk, RS_BLOCK, 0, st>>>(k_in, (long long)ntup, shift, (unsigned*)c->hist.p);
                k_rs_scan_digit<<<256, 256, 0, st>>>((unsigned*)c->hist.p, (int)nblk, d_dtot);
                k_rs_scan_tot<<<1, 256, 0, st>>>(d_dtot);
                k_rs_scatter<<<nblk, RS_BLOCK, 0, st>>>(k_in, v_in, (long long)ntup, shift, (const unsigned*)c->hist.p, d_dtot, k_out,
                                                        v_out);
                std::swap(k_in, k_out);
                std::swap(v_in, v_out);
                c->launches += 4;
            }
        }
        {
            ScTimer t(c, 2);
            const unsigned g2 = (unsigned)((ntup + 255) / 256);
            k_sc_umi<<<g2, 256, 0, st>>>(k_in, v_in, (long long)ntup, conf->has_umi ? 1 : 0, (unsigned char*)c->vote.p);
            k_sc_cell<<<g2, 256, 0, st>>>(k_in, (const unsigned char*)c->vote.p, (long long)ntup, A.bits_umi, (unsigned*)c->flag.p,
                                          (int2*)c->cnt.p, A.bits_cb, (int2*)c->tot.p);
            const unsigned nb = (unsigned)((ntup + (unsigned long long)CP_BLOCK * CP_ITEMS - 1) / ((unsigned long long)CP_BLOCK * CP_ITEMS));
            if ((rc = c->bsum.ensure((size_t)nb * 4 + 16))) return rc;
            k_cp_sum<<<nb, CP_BLOCK, 0, st>>>((const unsigned*)c->flag.p, (long long)ntup, (unsigned*)c->bsum.p);
            k_cp_scan<<<1, 256, 0, st>>>((unsigned*)c->bsum.p, (int)nb, (unsigned long long*)c->small.p + 1);
            k_cp_write<<<nb, CP_BLOCK, 0, st>>>(k_in, (const unsigned*)c->flag.p, (const int2*)c->cnt.p, (long long)ntup,
                                                (const unsigned*)c->bsum.p, A.bits_umi, A.bits_cb, (int4*)c->cells.p);
            c->launches += 5;
        }
        SCK(cudaMemcpyAsync(c->h_small, c->small.p, 16, cudaMemcpyDeviceToHost, st));
        SCK(cudaStreamSynchronize(st));
        if (cudaGetLastError() != cudaSuccess) { set_error("single-cell kernels failed"); return RAER_ERR_CUDA; }
        c->n_cells = (long long)c->h_small[1];
        return RAER_OK;
    }
    set_error("tuple buffers kept overflowing");
    return RAER_ERR;
}

extern "C" int raer_sc_batch_device(raer_sc_ctx* c, const raer_sc_conf* conf, const raer_read_batch* b, int64_t* n_cells,
                                    int64_t* n_tuples, int64_t* launches) {
    cudaSetDevice(c->device);
    const int64_t l0 = c->launches;
    int rc = sc_device(c, conf, b, conf->site_pos, conf->site_strand, conf->site_ref, conf->site_alt);
    if (rc) return rc;
    if (n_cells) *n_cells = c->n_cells;
    if (n_tuples) *n_tuples = c->n_tuples;
    if (launches) *launches += c->launches - l0;
    return RAER_OK;
}

// host batch + host site arrays -> device; the site arrays of a contig are uploaded once (same pointer, same count)
static int sc_upload(raer_sc_ctx* c, const raer_sc_conf* conf, const raer_read_batch* hb, raer_read_batch* db) {
    cudaStream_t st = c->st;
    int rc;
    const size_t nr = (size_t)hb->n_reads;
    const int ns = conf->n_sites;
    if ((rc = c->spos.ensure((size_t)ns * 4 + 16)) || (rc = c->sstr.ensure(ns + 16)) || (rc = c->sref.ensure(ns + 16)) ||
        (rc = c->salt.ensure(ns + 16)))
        return rc;
    if (ns > 0) {
        SCK(cudaMemcpyAsync(c->spos.p, conf->site_pos, (size_t)ns * 4, cudaMemcpyHostToDevice, st));
        SCK(cudaMemcpyAsync(c->sstr.p, conf->site_strand, ns, cudaMemcpyHostToDevice, st));
        SCK(cudaMemcpyAsync(c->sref.p, conf->site_ref, ns, cudaMemcpyHostToDevice, st));
        SCK(cudaMemcpyAsync(c->salt.p, conf->site_alt, ns, cudaMemcpyHostToDevice, st));
    }
    *db = *hb;
    if (nr == 0) return RAER_OK;
    if ((rc = c->hdr.ensure(nr * 32)) || (rc = c->cigar.ensure((size_t)hb->n_cigar_words * 4 + 16)) ||
        (rc = c->seq.ensure((size_t)hb->n_sq_bytes + 16)) || (rc = c->qual.ensure((size_t)hb->n_sq_bytes * 2 + 16)) ||
        (rc = c->pair.ensure((size_t)hb->n_pairs * 4 + 16)) || (rc = c->cb.ensure(nr * 4)) || (rc = c->umi.ensure(nr * 4)))
        return rc;
    SCK(cudaMemcpyAsync(c->hdr.p, hb->hdr, nr * 32, cudaMemcpyHostToDevice, st));
    if (hb->n_cigar_words) SCK(cudaMemcpyAsync(c->cigar.p, hb->cigar, (size_t)hb->n_cigar_words * 4, cudaMemcpyHostToDevice, st));
    if (hb->n_sq_bytes) {
        SCK(cudaMemcpyAsync(c->seq.p, hb->seq, (size_t)hb->n_sq_bytes, cudaMemcpyHostToDevice, st));
        SCK(cudaMemcpyAsync(c->qual.p, hb->qual, (size_t)hb->n_sq_bytes * 2, cudaMemcpyHostToDevice, st));
    }
    if (hb->n_pairs) SCK(cudaMemcpyAsync(c->pair.p, hb->pair_b, (size_t)hb->n_pairs * 4, cudaMemcpyHostToDevice, st));
    if (hb->cb) SCK(cudaMemcpyAsync(c->cb.p, hb->cb, nr * 4, cudaMemcpyHostToDevice, st));
    if (hb->umi) SCK(cudaMemcpyAsync(c->umi.p, hb->umi, nr * 4, cudaMemcpyHostToDevice, st));
    db->hdr = (const raer_read_hdr*)c->hdr.p; db->cigar = (const uint32_t*)c->cigar.p; db->seq = (const uint8_t*)c->seq.p;
    db->qual = (const uint8_t*)c->qual.p; db->pair_b = (const int32_t*)c->pair.p;
    db->cb = hb->cb ? (const int32_t*)c->cb.p : nullptr;
    db->umi = hb->umi ? (const uint32_t*)c->umi.p : nullptr;
    db->maxend = nullptr;
    return RAER_OK;
}

extern "C" int raer_sc_batch_host(raer_sc_ctx* c, const raer_sc_conf* conf, const raer_read_batch* hb, raer_sc_cells* out) {
    cudaSetDevice(c->device);
    memset(out, 0, sizeof(*out));
    raer_read_batch db;
    int rc = sc_upload(c, conf, hb, &db);
    if (rc) return rc;
    rc = sc_device(c, conf, &db, (const int32_t*)c->spos.p, (const uint8_t*)c->sstr.p, (const uint8_t*)c->sref.p,
                   (const uint8_t*)c->salt.p);
    if (rc) return rc;
    const size_t n = (size_t)c->n_cells, ns = (size_t)conf->n_sites;
    out->n = (int64_t)n;
    out->n_sites = (int64_t)ns;
    out->cells = (int32_t*)malloc((n ? n : 1) * 16);
    out->site_totals = (int32_t*)calloc(ns ? ns : 1, 8);
    if (!out->cells || !out->site_totals) { raer_sc_cells_free(out); set_error("out of memory"); return RAER_ERR; }
    if (n) SCK(cudaMemcpyAsync(out->cells, c->cells.p, n * 16, cudaMemcpyDeviceToHost, c->st));
    if (ns && c->n_tuples > 0) SCK(cudaMemcpyAsync(out->site_totals, c->tot.p, ns * 8, cudaMemcpyDeviceToHost, c->st));
    SCK(cudaStreamSynchronize(c->st));
    return RAER_OK;
}

// the same for a batch that already lives in device memory (device ingest): only the sites are uploaded
extern "C" int raer_sc_batch_device_cells(raer_sc_ctx* c, const raer_sc_conf* conf, const raer_read_batch* db, raer_sc_cells* out) {
    cudaSetDevice(c->device);
    memset(out, 0, sizeof(*out));
    raer_read_batch none;
    memset(&none, 0, sizeof(none));
    raer_read_batch dummy;
    int rc = sc_upload(c, conf, &none, &dummy);  // sites only: the empty batch uploads nothing else
    if (rc) return rc;
    rc = sc_device(c, conf, db, (const int32_t*)c->spos.p, (const uint8_t*)c->sstr.p, (const uint8_t*)c->sref.p,
                   (const uint8_t*)c->salt.p);
    if (rc) return rc;
    const size_t n = (size_t)c->n_cells, ns = (size_t)conf->n_sites;
    out->n = (int64_t)n;
    out->n_sites = (int64_t)ns;
    out->cells = (int32_t*)malloc((n ? n : 1) * 16);
    out->site_totals = (int32_t*)calloc(ns ? ns : 1, 8);
    if (!out->cells || !out->site_totals) { raer_sc_cells_free(out); set_error("out of memory"); return RAER_ERR; }
    if (n) SCK(cudaMemcpyAsync(out->cells, c->cells.p, n * 16, cudaMemcpyDeviceToHost, c->st));
    if (ns && c->n_tuples > 0) SCK(cudaMemcpyAsync(out->site_totals, c->tot.p, ns * 8, cudaMemcpyDeviceToHost, c->st));
    SCK(cudaStreamSynchronize(c->st));
    return RAER_OK;
}

struct raer_gi;
extern "C" raer_gi* raer_gi_create(int device, int nf, const char* const* paths, const char* const* indexes);
extern "C" void raer_gi_destroy(raer_gi* g);
extern "C" int raer_gi_contig(raer_gi* g, int tid, const int* ftid, int contig_len, const raer::IngestConf* ic,
                              const raer::RegionIndex::Chr* site, int reg_beg, int reg_end, int* n_windows,
                              const uint32_t** site_mask_dev, int64_t* aligned_bases);
extern "C" int raer_gi_window(raer_gi* g, int j, raer_read_batch* b, int64_t* file_off);
extern "C" uint32_t raer_gi_umi_space(raer_gi* g);

static int64_t g_sc_last_device_contigs = 0;
// contigs the last raer_scpileup() / raer_scpileup_triplets() call of this process took through the device ingest
extern "C" int64_t raer_sc_last_device_contigs(void) { return g_sc_last_device_contigs; }

// One engine behind both entry points: text files (the reference's contract), triplets in memory, or both.
static int sc_run(const raer_scpileup_args* a, bool want_files, raer_sc_triplets* trip_out) {
    // check_sc_plp_args, src/sc-plp.c:874-940
    if (!a || a->nbam < 1 || !a->bampaths || a->n_sites < 1 || a->n_barcodes < 1 || !a->barcodes ||
        (want_files && (!a->outfns[0] || !a->outfns[1] || !a->outfns[2]))) {
        set_error("invalid arguments");
        return RAER_ERR_ARG;
    }
    std::vector<int32_t> mem_site, mem_cb, mem_ref, mem_alt;
    if (raer_device_count() <= 0) { set_error("no usable CUDA device (the product has no CPU path)"); return RAER_ERR_NO_DEVICE; }
    const bool is_ss2 = a->nbam > 1;
    if (is_ss2 && a->n_barcodes < a->nbam) { set_error("one barcode per BAM is required"); return RAER_ERR_ARG; }
    // every resource is released by its owner's destructor, whatever path leaves the function
    struct Files {
        FILE* f[3] = {nullptr, nullptr, nullptr};
        ~Files() { for (FILE* x : f) if (x) fclose(x); }
    } fps;
    if (want_files) {
        for (int i = 0; i < 3; ++i) fps.f[i] = fopen(a->outfns[i], "w");
        if (!fps.f[0] || !fps.f[1] || !fps.f[2]) { set_error("cannot open output files"); return RAER_ERR_IO; }
    }
    struct CtxHolder {
        raer_sc_ctx* c = nullptr;
        ~CtxHolder() { raer_sc_ctx_destroy(c); }
    } ch;

    // barcode whitelist: first occurrence wins (src/sc-plp.c:814-826)
    std::unordered_map<std::string, int> cbidx;
    for (int i = 0; i < a->n_barcodes; ++i) cbidx.emplace(a->barcodes[i], i + 1);
    if (fps.f[2])
        for (int i = 0; i < a->n_barcodes; ++i) fprintf(fps.f[2], "%s\n", a->barcodes[i]);

    // sites per contig, sorted by position (stable), as regidx keeps them
    std::vector<std::string> chr_order;
    std::unordered_map<std::string, std::vector<int>> by_chr;
    for (int i = 0; i < a->n_sites; ++i) {
        auto it = by_chr.find(a->site_seqnames[i]);
        if (it == by_chr.end()) { chr_order.push_back(a->site_seqnames[i]); it = by_chr.emplace(a->site_seqnames[i], std::vector<int>()).first; }
        it->second.push_back(i);
    }
    for (auto& kv : by_chr)
        std::stable_sort(kv.second.begin(), kv.second.end(), [&](int x, int y) { return a->site_pos[x] < a->site_pos[y]; });
    if (fps.f[1])
        for (auto& nm : chr_order)
            for (int s : by_chr[nm])
                fprintf(fps.f[1], "%d\t%s\t%d\t%d\t%s\t%s\n", a->site_idx[s], nm.c_str(), a->site_pos[s], a->site_strand[s],
                        a->site_ref[s], a->site_alt[s]);
    raer::RegionIndex ridx;
    {
        std::vector<int32_t> b0(a->n_sites);
        for (int i = 0; i < a->n_sites; ++i) b0[i] = a->site_pos[i] - 1;
        ridx.build(a->n_sites, a->site_seqnames, b0.data(), b0.data());
    }

    const int max_depth = a->int_args[0] ? a->int_args[0] : 10000;
    const int min_counts = a->min_counts < 0 ? 0 : a->min_counts;
    const int min_var_reads = a->int_args[9];
    const bool has_umi = a->umi_tag != nullptr;
    if (a->libtype < 0 || a->libtype > 2) { set_error("invert read orientation failure"); return RAER_ERR; }

    ch.c = raer_sc_ctx_create(a->device);
    if (!ch.c) return RAER_ERR_CUDA;
    raer_sc_ctx* c = ch.c;
    raer_sc_conf conf;
    memset(&conf, 0, sizeof(conf));
    conf.min_bq = a->int_args[2] < 0 ? 0 : a->int_args[2];
    conf.min_mapq = a->min_mapq < 0 ? 0 : a->min_mapq;
    conf.trim_5p = a->int_args[3]; conf.trim_3p = a->int_args[4];
    conf.indel_dist = a->int_args[5]; conf.splice_dist = a->int_args[6]; conf.min_overhang = a->int_args[7];
    conf.ftrim_5p = a->dbl_args[0]; conf.ftrim_3p = a->dbl_args[1];
    conf.remove_overlaps = a->remove_overlaps ? (a->overlap_mode == 2 ? 2 : 1) : 0;
    conf.has_umi = has_umi ? 1 : 0;
    conf.n_barcodes = a->n_barcodes;

    int rc = RAER_OK;
    int64_t n_device_contigs = 0;
    for (int bi = 0; bi < (is_ss2 ? a->nbam : 1) && rc == RAER_OK; ++bi) {
        raer::IngestConf ic;
        ic.n_files = 1;
        ic.keep_flag[0] = (uint32_t)a->int_args[12];
        ic.keep_flag[1] = (uint32_t)a->int_args[13];
        ic.max_depth = max_depth;
        ic.min_global_mq = conf.min_mapq;
        ic.rq_pct = a->dbl_args[3];
        ic.rq_minq = (int)a->dbl_args[4];
        ic.remove_overlaps = conf.remove_overlaps;
        ic.min_overhang = a->int_args[7];
        ic.libtype.assign(1, a->libtype);
        ic.regidx = &ridx;
        ic.sc = true;
        ic.cb_index = &cbidx;
        ic.cb_tag = a->cb_tag ? a->cb_tag : "";
        ic.umi_tag = a->umi_tag ? a->umi_tag : "";
        ic.fixed_cb = 0;
        if (is_ss2) {
            auto it = cbidx.find(a->barcodes[bi]);
            ic.fixed_cb = it == cbidx.end() ? 0 : it->second;
        }
        const char* one_bam[1] = {a->bampaths[bi]};
        const char* one_idx[1] = {a->indexes ? a->indexes[bi] : nullptr};
        // the sites of a contig as the arrays the kernels take
        std::vector<int32_t> sp;
        std::vector<uint8_t> ss, sr, sa;
        auto contig_sites = [&](const std::string& name) -> const std::vector<int>* {
            auto cit = by_chr.find(name);
            const std::vector<int>* sites = cit == by_chr.end() ? nullptr : &cit->second;
            const int ns = sites ? (int)sites->size() : 0;
            sp.resize(ns); ss.resize(ns); sr.resize(ns); sa.resize(ns);
            for (int k = 0; k < ns; ++k) {
                const int s = (*sites)[k];
                sp[k] = a->site_pos[s] - 1;
                ss[k] = (uint8_t)a->site_strand[s];
                sr[k] = (uint8_t)a->site_ref[s][0];
                sa[k] = (uint8_t)a->site_alt[s][0];
            }
            conf.n_sites = ns;
            conf.site_pos = sp.data(); conf.site_strand = ss.data(); conf.site_ref = sr.data(); conf.site_alt = sa.data();
            return sites;
        };
        // site rule, src/sc-plp.c:586-596; the cells arrive sorted by (site, barcode)
        auto emit_cells = [&](const std::vector<int>* sites, const raer_sc_cells& cells) {
            for (int64_t i = 0; i < cells.n; ++i) {
                const int32_t* t = cells.cells + 4 * i;
                const int32_t* tt = cells.site_totals + 2 * t[0];
                if (!(min_counts == 0 || (tt[0] + tt[1] >= min_counts && tt[1] >= min_var_reads))) continue;
                if (fps.f[0]) fprintf(fps.f[0], "%d %d %d %d\n", a->site_idx[(*sites)[t[0]]], t[1], t[2], t[3]);
                if (trip_out) {
                    mem_site.push_back(a->site_idx[(*sites)[t[0]]]); mem_cb.push_back(t[1]);
                    mem_ref.push_back(t[2]); mem_alt.push_back(t[3]);
                }
            }
        };
        // One pass over an opened ingest: its contigs, their sites, their batches.
        auto drain = [&](raer::Ingest& ing) -> int {
            int rc2 = RAER_OK;
            const std::vector<std::string>& names = ing.names();
            raer::PackedBatch pb;
            int tid;
            while (rc2 == RAER_OK && (tid = ing.next_contig()) >= 0) {
                const std::vector<int>* sites = contig_sites(names[tid]);
                const int ns = conf.n_sites;
                while (rc2 == RAER_OK && ing.next_batch(pb)) {
                    const raer_read_batch& b = pb.b;
                    if (a->poll && a->poll(a->poll_ctx)) {  // the calling thread polls between batches (src/sc-plp.c:707-712)
                        set_error("interrupted");
                        return RAER_ERR_INTERRUPT;
                    }
                    if (!ns || b.n_reads == 0 || b.end <= b.beg) continue;
                    conf.n_umi = (uint32_t)ing.umi_dict().size();
                    raer_sc_cells cells;
                    rc2 = raer_sc_batch_host(c, &conf, &b, &cells);
                    if (rc2) break;
                    emit_cells(sites, cells);
                    raer_sc_cells_free(&cells);
                }
            }
            if (rc2 == RAER_OK && ing.failed()) {  // the stream stopped on an error, not at the end of the file
                set_error("%s", ing.error().c_str());
                rc2 = RAER_ERR_IO;
            }
            return rc2;
        };
        // a region that is a plain contig name (the per-chromosome calls of R/sc-pileup.R:185-230) restricts the loop below;
        // a coordinate range goes through the single stream
        const bool whole_contig = a->region && !strchr(a->region, ':');
        if ((!a->region || whole_contig) && raer::indexes_usable(one_bam, one_idx, 1) && !getenv("RAER_SERIAL_INGEST")) {
            // the sites are a list of positions: contig by contig through the BAI index, skipping what lies between them
            // (src/sc-plp.c:648-669 queries the index too)
            raer::BamStream hs;
            if (!hs.open(one_bam[0], 1, true)) { rc = RAER_ERR_IO; break; }
            const std::vector<std::string> names = hs.names();
            const std::vector<int32_t> lens = hs.lengths();
            // Device ingest (raer_gingest.cu): a contig whose sites are spread over most of it is inflated, parsed, filtered
            // (cell-barcode whitelist included) and packed by kernels; the others, and whatever that path hands back,
            // go through the host stream and the index.
            bool use_gi = !(getenv("RAER_GPU_INGEST") && atoi(getenv("RAER_GPU_INGEST")) == 0) && !getenv("RAER_BATCH_READS");
            if (use_gi) {
                static const unsigned char eof_blk[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                                          0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
                unsigned char tail[28];
                FILE* fp = fopen(one_bam[0], "rb");
                use_gi = fp && fseek(fp, -28, SEEK_END) == 0 && fread(tail, 1, 28, fp) == 28 && memcmp(tail, eof_blk, 28) == 0;
                if (fp) fclose(fp);
            }
            struct GiHolder {
                raer_gi* g = nullptr;
                ~GiHolder() { if (g) raer_gi_destroy(g); }
            } gh;
            if (use_gi) {
                gh.g = raer_gi_create(a->device, 1, one_bam, one_idx);
                if (!gh.g) { rc = RAER_ERR_IO; break; }
            }
            if (whole_contig && std::find(names.begin(), names.end(), std::string(a->region)) == names.end()) {
                set_error("fail to parse region '%s'", a->region);
                rc = RAER_ERR_IO;
                break;
            }
            for (size_t ci = 0; ci < names.size() && rc == RAER_OK; ++ci) {
                if (whole_contig && names[ci] != a->region) continue;
                const raer::RegionIndex::Chr* rchr = ridx.find(names[ci]);
                if (!rchr) continue;
                const std::vector<std::pair<int32_t, int32_t>> jl = raer::jump_intervals(rchr);
                bool on_device = use_gi;
                if (on_device) {
                    int64_t covered = 0;
                    for (const auto& iv : jl) covered += (int64_t)iv.second - iv.first;
                    on_device = covered * 2 >= (int64_t)lens[ci];
                }
                if (on_device) {
                    if (a->poll && a->poll(a->poll_ctx)) { set_error("interrupted"); rc = RAER_ERR_INTERRUPT; break; }
                    int64_t file_off[2] = {0, 0}, bases = 0;
                    const uint32_t* d_sites = nullptr;
                    const int ftid = (int)ci;
                    int n_windows = 0;
                    const int r = raer_gi_contig(gh.g, (int)ci, &ftid, lens[ci], &ic, rchr, 0, INT_MAX, &n_windows, &d_sites, &bases);
                    if (r < 0) { rc = r; break; }
                    if (r == 0) {
                        const std::vector<int>* sites = contig_sites(names[ci]);
                        for (int j = 0; j < n_windows && rc == RAER_OK; ++j) {
                            raer_read_batch b;
                            if ((rc = raer_gi_window(gh.g, j, &b, file_off))) break;
                            if (!conf.n_sites || b.n_reads == 0) continue;
                            conf.n_umi = raer_gi_umi_space(gh.g);
                            raer_sc_cells cells;
                            rc = raer_sc_batch_device_cells(c, &conf, &b, &cells);
                            if (rc) break;
                            emit_cells(sites, cells);
                            raer_sc_cells_free(&cells);
                        }
                        if (rc) break;
                        ++n_device_contigs;
                        continue;
                    }
                    // r == 1: handed back (UMIs the 2-bit code does not cover, a name-hash collision, not enough device memory, ...)
                }
                raer::Ingest ing(ic);
                if (!ing.open(one_bam, one_idx, a->host_threads, names[ci].c_str(), &jl)) { rc = RAER_ERR_IO; break; }
                rc = drain(ing);
            }
        } else {
            raer::Ingest ing(ic);
            if (!ing.open(one_bam, one_idx, a->host_threads, a->region)) { rc = RAER_ERR_IO; break; }
            rc = drain(ing);
        }
    }
    if (rc != RAER_OK) return rc < 0 ? rc : RAER_ERR;
    if (getenv("RAER_TIMING")) fprintf(stderr, "[raer_scpileup] %lld contig(s) through the device ingest\n", (long long)n_device_contigs);
    g_sc_last_device_contigs = n_device_contigs;
    if (trip_out) {
        const size_t n = mem_site.size();
        trip_out->n = (int64_t)n;
        trip_out->site_idx = (int32_t*)malloc((n ? n : 1) * 4);
        trip_out->cb_idx = (int32_t*)malloc((n ? n : 1) * 4);
        trip_out->n_ref = (int32_t*)malloc((n ? n : 1) * 4);
        trip_out->n_alt = (int32_t*)malloc((n ? n : 1) * 4);
        if (n) {
            memcpy(trip_out->site_idx, mem_site.data(), n * 4); memcpy(trip_out->cb_idx, mem_cb.data(), n * 4);
            memcpy(trip_out->n_ref, mem_ref.data(), n * 4); memcpy(trip_out->n_alt, mem_alt.data(), n * 4);
        }
    }
    return 0;
}

extern "C" int raer_scpileup(const raer_scpileup_args* a) { return sc_run(a, true, nullptr); }

extern "C" int raer_scpileup_triplets(const raer_scpileup_args* a, raer_sc_triplets* out) {
    if (!out) { set_error("invalid arguments"); return RAER_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    const bool files = a && a->outfns[0] && a->outfns[1] && a->outfns[2];
    return sc_run(a, files, out);
}

extern "C" void raer_sc_triplets_free(raer_sc_triplets* t) {
    if (!t) return;
    free(t->site_idx); free(t->cb_idx); free(t->n_ref); free(t->n_alt);
    memset(t, 0, sizeof(*t));
}
