// raer_sc.cu -- single-cell pileup path (pileup_cells): .Call(".scpileup") equivalent.
//
// Reference: src/sc-plp.c (scpileup :945, run_scpileup :612, screadaln :449, process_one_site :528,
// count_record :235, count_consensus_base :321, write_counts :410, write_all_sites :852).
//
// Device pipeline per read batch (sites are sparse, so the unit of work is a read, not a base):
//   k_sc_emit      one thread per read: binary-search the sorted site list for positions inside the
//                  read's span, locate the query base through the CIGAR, apply the shared read/base
//                  filters, and emit one tuple (site, cell barcode, UMI | class, base quality) per
//                  matching payload. Runs twice: count, then write (exact allocation).
//   k_rs_*         LSD radix sort of the 64-bit tuple keys (8-bit digits, stable block scatter).
//   k_sc_umi/cell  segmented reduce over the sorted tuples: (site, CB, UMI) quality sums -> consensus
//                  vote (ties go to alt, src/sc-plp.c:338) -> (site, CB) nRef/nAlt + per-site totals.
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
                    unsigned long long idx = atomicAdd(A.counter, 1ull);
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

// ------------------------------------------------------------------------------------- radix sort
#define RS_BLOCK 256
#define RS_ITEMS 16  // keys per thread per block chunk

__global__ void k_rs_hist(const unsigned long long* __restrict__ keys, long long n, int shift, unsigned* __restrict__ hist) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    long long base = (long long)blockIdx.x * RS_BLOCK * RS_ITEMS;
    for (int i = 0; i < RS_ITEMS; ++i) {
        long long j = base + (long long)i * RS_BLOCK + threadIdx.x;
        if (j < n) atomicAdd(&h[(keys[j] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];  // digit-major
}

__device__ __forceinline__ unsigned block_incl_scan_256(unsigned v, unsigned* part) {
    part[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        unsigned t = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    return part[threadIdx.x];
}

// hist is digit-major [256][nblk]. Block d turns row d into an exclusive scan and writes the row total;
// k_rs_scan_tot then turns the 256 totals into digit bases.
__global__ void k_rs_scan_digit(unsigned* hist, int nblk, unsigned* tot) {
    __shared__ unsigned part[256];
    __shared__ unsigned carry;
    unsigned* row = hist + (size_t)blockIdx.x * nblk;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 256) {
        int i = base + threadIdx.x;
        unsigned v = i < nblk ? row[i] : 0;
        unsigned inc = block_incl_scan_256(v, part);
        if (i < nblk) row[i] = carry + inc - v;
        __syncthreads();
        if (threadIdx.x == 255) carry += inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) tot[blockIdx.x] = carry;
}
__global__ void k_rs_scan_tot(unsigned* tot) {
    __shared__ unsigned part[256];
    unsigned v = tot[threadIdx.x];
    unsigned inc = block_incl_scan_256(v, part);
    tot[threadIdx.x] = inc - v;
}

__global__ void k_rs_scatter(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ vals, long long n,
                             int shift, const unsigned* __restrict__ offs, const unsigned* __restrict__ dbase,
                             unsigned long long* __restrict__ okeys, unsigned* __restrict__ ovals) {
    __shared__ unsigned wcnt[RS_BLOCK / 32][256];
    __shared__ unsigned run[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    run[threadIdx.x] = dbase[threadIdx.x] + offs[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
    long long base = (long long)blockIdx.x * RS_BLOCK * RS_ITEMS;
    for (int i = 0; i < RS_ITEMS; ++i) {
        for (int d = threadIdx.x; d < (RS_BLOCK / 32) * 256; d += RS_BLOCK) (&wcnt[0][0])[d] = 0;
        __syncthreads();
        long long j = base + (long long)i * RS_BLOCK + threadIdx.x;
        bool valid = j < n;
        unsigned long long key = valid ? keys[j] : 0ull;
        unsigned dig = valid ? (unsigned)((key >> shift) & 0xff) : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, dig);
        unsigned rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) wcnt[w][dig] = __popc(peers);
        __syncthreads();
        if (valid) {
            unsigned before = 0;
            for (int ww = 0; ww < w; ++ww) before += wcnt[ww][dig];
            unsigned dst = run[dig] + before + rank;
            okeys[dst] = key;
            ovals[dst] = vals[j];
        }
        __syncthreads();
        {
            unsigned tot = 0;
            for (int ww = 0; ww < RS_BLOCK / 32; ++ww) tot += wcnt[ww][threadIdx.x];
            run[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}

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

// pass 2: one thread per (site, CB) segment head sums the votes and appends a triplet
__global__ void k_sc_cell(const unsigned long long* __restrict__ keys, const unsigned char* __restrict__ vote, long long n,
                          int bits_umi, int bits_cb, int4* __restrict__ trip, unsigned long long* __restrict__ ntrip,
                          int2* __restrict__ site_tot) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long cell = keys[i] >> bits_umi;
    if (i > 0 && (keys[i - 1] >> bits_umi) == cell) return;
    int nref = 0, nalt = 0;
    for (long long j = i; j < n && (keys[j] >> bits_umi) == cell; ++j) {
        unsigned char v = vote[j];
        nref += v == 1;
        nalt += v == 2;
    }
    int site = (int)(cell >> bits_cb);
    int cb = (int)(cell & ((1ull << bits_cb) - 1ull));
    if (nref) atomicAdd(&site_tot[site].x, nref);
    if (nalt) atomicAdd(&site_tot[site].y, nalt);
    if (nref || nalt) {
        unsigned long long o = atomicAdd(ntrip, 1ull);
        trip[o] = make_int4(site, cb, nref, nalt);
    }
}

// ------------------------------------------------------------------------------------- host side
extern "C" int raer_launch_overlap(const raer_read_hdr* hdr, const uint32_t* cigar, const uint8_t* seq, uint8_t* qual,
                                   const int32_t* pair_b, long long n_pairs, int mode, cudaStream_t st, int* scratch);

struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap && p) return 0;
        if (p) cudaFree(p);
        size_t want = n + n / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) { p = nullptr; cap = 0; set_error("cudaMalloc(%zu) failed", want); return RAER_ERR_CUDA; }
        cap = want;
        return 0;
    }
    ~DBuf() { if (p) cudaFree(p); }
};

static int bits_for(unsigned long long maxval) {
    int b = 1;
    while ((maxval >> b) != 0) ++b;
    return b;
}

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
    FILE* fps[3] = {nullptr, nullptr, nullptr};
    if (want_files)
        for (int i = 0; i < 3; ++i) fps[i] = fopen(a->outfns[i], "w");
    auto close_all = [&]() { for (int i = 0; i < 3; ++i) if (fps[i]) fclose(fps[i]); };
    if (want_files && (!fps[0] || !fps[1] || !fps[2])) { close_all(); set_error("cannot open output files"); return RAER_ERR_IO; }

    // barcode whitelist: first occurrence wins (src/sc-plp.c:814-826)
    std::unordered_map<std::string, int> cbidx;
    for (int i = 0; i < a->n_barcodes; ++i) cbidx.emplace(a->barcodes[i], i + 1);
    if (fps[2])
        for (int i = 0; i < a->n_barcodes; ++i) fprintf(fps[2], "%s\n", a->barcodes[i]);

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
    if (fps[1])
        for (auto& nm : chr_order)
            for (int s : by_chr[nm])
                fprintf(fps[1], "%d\t%s\t%d\t%d\t%s\t%s\n", a->site_idx[s], nm.c_str(), a->site_pos[s], a->site_strand[s],
                        a->site_ref[s], a->site_alt[s]);
    raer::RegionIndex ridx;
    {
        std::vector<int32_t> b0(a->n_sites);
        for (int i = 0; i < a->n_sites; ++i) b0[i] = a->site_pos[i] - 1;
        ridx.build(a->n_sites, a->site_seqnames, b0.data(), b0.data());
    }

    const int min_bq = a->int_args[2] < 0 ? 0 : a->int_args[2];
    const int max_depth = a->int_args[0] ? a->int_args[0] : 10000;
    const int min_mq = a->min_mapq < 0 ? 0 : a->min_mapq;
    const int min_counts = a->min_counts < 0 ? 0 : a->min_counts;
    const int min_var_reads = a->int_args[9];
    const bool has_umi = a->umi_tag != nullptr;

    cudaSetDevice(a->device);
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { close_all(); set_error("stream"); return RAER_ERR_CUDA; }
    int rc = RAER_OK;
    DBuf d_hdr, d_cigar, d_seq, d_qual, d_pair, d_cb, d_umi, d_spos, d_sstr, d_sref, d_salt, d_small, d_keys, d_vals, d_keys2,
        d_vals2, d_hist, d_vote, d_trip, d_tot;

    for (int bi = 0; bi < (is_ss2 ? a->nbam : 1) && rc == RAER_OK; ++bi) {
        raer::IngestConf ic;
        ic.n_files = 1;
        ic.keep_flag[0] = (uint32_t)a->int_args[12];
        ic.keep_flag[1] = (uint32_t)a->int_args[13];
        ic.max_depth = max_depth;
        ic.min_global_mq = min_mq;
        ic.rq_pct = a->dbl_args[3];
        ic.rq_minq = (int)a->dbl_args[4];
        ic.remove_overlaps = a->remove_overlaps ? (a->overlap_mode == 2 ? 2 : 1) : 0;
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
        if (a->libtype < 0 || a->libtype > 2) { rc = RAER_ERR; set_error("invert read orientation failure"); break; }
        raer::Ingest ing(ic);
        const char* one_bam[1] = {a->bampaths[bi]};
        const char* one_idx[1] = {a->indexes ? a->indexes[bi] : nullptr};
        if (!ing.open(one_bam, one_idx, a->host_threads, a->region)) { rc = RAER_ERR_IO; break; }
        const std::vector<std::string>& names = ing.names();
        raer::PackedBatch pb;
        int tid;
        while (rc == RAER_OK && (tid = ing.next_contig()) >= 0) {
            auto cit = by_chr.find(names[tid]);
            const std::vector<int>* sites = cit == by_chr.end() ? nullptr : &cit->second;
            int ns = sites ? (int)sites->size() : 0;
            if (ns) {
                std::vector<int32_t> sp(ns);
                std::vector<uint8_t> ss(ns), sr(ns), sa(ns);
                for (int k = 0; k < ns; ++k) {
                    int s = (*sites)[k];
                    sp[k] = a->site_pos[s] - 1;
                    ss[k] = (uint8_t)a->site_strand[s];
                    sr[k] = (uint8_t)a->site_ref[s][0];
                    sa[k] = (uint8_t)a->site_alt[s][0];
                }
                if ((rc = d_spos.ensure(ns * 4)) || (rc = d_sstr.ensure(ns)) || (rc = d_sref.ensure(ns)) || (rc = d_salt.ensure(ns)) ||
                    (rc = d_tot.ensure((size_t)ns * 8))) break;
                SCK(cudaMemcpyAsync(d_spos.p, sp.data(), ns * 4, cudaMemcpyHostToDevice, st));
                SCK(cudaMemcpyAsync(d_sstr.p, ss.data(), ns, cudaMemcpyHostToDevice, st));
                SCK(cudaMemcpyAsync(d_sref.p, sr.data(), ns, cudaMemcpyHostToDevice, st));
                SCK(cudaMemcpyAsync(d_salt.p, sa.data(), ns, cudaMemcpyHostToDevice, st));
                SCK(cudaStreamSynchronize(st));
            }
            while (rc == RAER_OK && ing.next_batch(pb)) {
                const raer_read_batch& b = pb.b;
                if (!ns || b.n_reads == 0 || b.end <= b.beg) continue;
                size_t nr = (size_t)b.n_reads;
                if ((rc = d_hdr.ensure(nr * 32)) || (rc = d_cigar.ensure((size_t)b.n_cigar_words * 4 + 16)) ||
                    (rc = d_seq.ensure((size_t)b.n_sq_bytes + 16)) || (rc = d_qual.ensure((size_t)b.n_sq_bytes * 2 + 16)) ||
                    (rc = d_pair.ensure((size_t)b.n_pairs * 4 + 16)) || (rc = d_cb.ensure(nr * 4)) || (rc = d_umi.ensure(nr * 4)) ||
                    (rc = d_small.ensure(64))) break;
                SCK(cudaMemcpyAsync(d_hdr.p, b.hdr, nr * 32, cudaMemcpyHostToDevice, st));
                if (b.n_cigar_words) SCK(cudaMemcpyAsync(d_cigar.p, b.cigar, (size_t)b.n_cigar_words * 4, cudaMemcpyHostToDevice, st));
                if (b.n_sq_bytes) {
                    SCK(cudaMemcpyAsync(d_seq.p, b.seq, (size_t)b.n_sq_bytes, cudaMemcpyHostToDevice, st));
                    SCK(cudaMemcpyAsync(d_qual.p, b.qual, (size_t)b.n_sq_bytes * 2, cudaMemcpyHostToDevice, st));
                }
                if (b.n_pairs) SCK(cudaMemcpyAsync(d_pair.p, b.pair_b, (size_t)b.n_pairs * 4, cudaMemcpyHostToDevice, st));
                SCK(cudaMemcpyAsync(d_cb.p, b.cb, nr * 4, cudaMemcpyHostToDevice, st));
                SCK(cudaMemcpyAsync(d_umi.p, b.umi, nr * 4, cudaMemcpyHostToDevice, st));
                SCK(cudaMemsetAsync(d_small.p, 0, 64, st));
                SCK(cudaMemsetAsync(d_tot.p, 0, (size_t)ns * 8, st));
                if (ic.remove_overlaps && b.n_pairs > 0)
                    raer_launch_overlap((const raer_read_hdr*)d_hdr.p, (const uint32_t*)d_cigar.p, (const uint8_t*)d_seq.p,
                                        (uint8_t*)d_qual.p, (const int32_t*)d_pair.p, b.n_pairs, ic.remove_overlaps, st, nullptr);
                ScParams A;
                memset(&A, 0, sizeof(A));
                A.P.hdr = (const raer_read_hdr*)d_hdr.p; A.P.cigar = (const uint32_t*)d_cigar.p;
                A.P.seq = (const uint8_t*)d_seq.p; A.P.qual = (uint8_t*)d_qual.p;
                A.P.min_bq = min_bq; A.P.trim_5p = a->int_args[3]; A.P.trim_3p = a->int_args[4];
                A.P.indel_dist = a->int_args[5]; A.P.splice_dist = a->int_args[6]; A.P.min_overhang = a->int_args[7];
                A.P.ftrim_5p = a->dbl_args[0]; A.P.ftrim_3p = a->dbl_args[1];
                A.P.min_mapq[0] = min_mq;
                A.S.pos = (const int32_t*)d_spos.p; A.S.strand = (const uint8_t*)d_sstr.p;
                A.S.ref = (const uint8_t*)d_sref.p; A.S.alt = (const uint8_t*)d_salt.p; A.S.n = ns;
                A.cb = (const int32_t*)d_cb.p; A.umi = (const uint32_t*)d_umi.p;
                A.n_reads = b.n_reads; A.win_beg = b.beg; A.win_end = b.end;
                A.has_umi = has_umi;
                A.bits_cb = bits_for((unsigned long long)a->n_barcodes + 1);
                A.bits_umi = has_umi ? bits_for((unsigned long long)ing.umi_dict().size() + 1) : 1;
                int bits_site = bits_for((unsigned long long)ns);
                if (bits_site + A.bits_cb + A.bits_umi > 64) { rc = RAER_ERR_LIMIT; set_error("site x barcode x UMI key exceeds 64 bits"); break; }
                A.counter = (unsigned long long*)d_small.p;
                unsigned grid = (unsigned)((nr + 255) / 256);
                k_sc_emit<false><<<grid, 256, 0, st>>>(A);
                unsigned long long ntup = 0;
                SCK(cudaMemcpyAsync(&ntup, d_small.p, 8, cudaMemcpyDeviceToHost, st));
                SCK(cudaStreamSynchronize(st));
                if (ntup == 0) continue;
                if ((rc = d_keys.ensure(ntup * 8)) || (rc = d_vals.ensure(ntup * 4)) || (rc = d_keys2.ensure(ntup * 8)) ||
                    (rc = d_vals2.ensure(ntup * 4)) || (rc = d_vote.ensure(ntup)) || (rc = d_trip.ensure(ntup * 16))) break;
                SCK(cudaMemsetAsync(d_small.p, 0, 64, st));
                A.keys = (unsigned long long*)d_keys.p; A.vals = (unsigned*)d_vals.p; A.cap = ntup;
                k_sc_emit<true><<<grid, 256, 0, st>>>(A);
                // LSD radix sort over the used key bits
                int key_bits = bits_site + A.bits_cb + A.bits_umi;
                unsigned nblk = (unsigned)((ntup + (unsigned long long)RS_BLOCK * RS_ITEMS - 1) / ((unsigned long long)RS_BLOCK * RS_ITEMS));
                if ((rc = d_hist.ensure((size_t)nblk * 256 * 4 + 1024))) break;
                unsigned long long* k_in = (unsigned long long*)d_keys.p; unsigned* v_in = (unsigned*)d_vals.p;
                unsigned long long* k_out = (unsigned long long*)d_keys2.p; unsigned* v_out = (unsigned*)d_vals2.p;
                unsigned* d_dtot = (unsigned*)d_hist.p + (size_t)nblk * 256;
                for (int shift = 0; shift < key_bits; shift += 8) {
                    k_rs_hist<<<nblk, RS_BLOCK, 0, st>>>(k_in, (long long)ntup, shift, (unsigned*)d_hist.p);
                    k_rs_scan_digit<<<256, 256, 0, st>>>((unsigned*)d_hist.p, (int)nblk, d_dtot);
                    k_rs_scan_tot<<<1, 256, 0, st>>>(d_dtot);
                    k_rs_scatter<<<nblk, RS_BLOCK, 0, st>>>(k_in, v_in, (long long)ntup, shift, (const unsigned*)d_hist.p, d_dtot,
                                                            k_out, v_out);
                    std::swap(k_in, k_out);
                    std::swap(v_in, v_out);
                }
                unsigned g2 = (unsigned)((ntup + 255) / 256);
                k_sc_umi<<<g2, 256, 0, st>>>(k_in, v_in, (long long)ntup, has_umi ? 1 : 0, (unsigned char*)d_vote.p);
                SCK(cudaMemsetAsync(d_small.p, 0, 64, st));
                k_sc_cell<<<g2, 256, 0, st>>>(k_in, (const unsigned char*)d_vote.p, (long long)ntup, A.bits_umi, A.bits_cb,
                                              (int4*)d_trip.p, (unsigned long long*)d_small.p, (int2*)d_tot.p);
                unsigned long long ntrip = 0;
                SCK(cudaMemcpyAsync(&ntrip, d_small.p, 8, cudaMemcpyDeviceToHost, st));
                SCK(cudaStreamSynchronize(st));
                if (cudaGetLastError() != cudaSuccess) { rc = RAER_ERR_CUDA; set_error("single-cell kernels failed"); break; }
                std::vector<int4> trip(ntrip);
                std::vector<int2> tot(ns);
                if (ntrip) SCK(cudaMemcpyAsync(trip.data(), d_trip.p, ntrip * 16, cudaMemcpyDeviceToHost, st));
                SCK(cudaMemcpyAsync(tot.data(), d_tot.p, (size_t)ns * 8, cudaMemcpyDeviceToHost, st));
                SCK(cudaStreamSynchronize(st));
                // site rule, src/sc-plp.c:586-596; order of lines carries no meaning (R builds a sparse matrix)
                std::sort(trip.begin(), trip.end(), [](const int4& x, const int4& y) { return x.x != y.x ? x.x < y.x : x.y < y.y; });
                for (const int4& t : trip) {
                    int2 tt = tot[t.x];
                    if (!(min_counts == 0 || (tt.x + tt.y >= min_counts && tt.y >= min_var_reads))) continue;
                    if (fps[0]) fprintf(fps[0], "%d %d %d %d\n", a->site_idx[(*sites)[t.x]], t.y, t.z, t.w);
                    if (trip_out) {
                        mem_site.push_back(a->site_idx[(*sites)[t.x]]); mem_cb.push_back(t.y);
                        mem_ref.push_back(t.z); mem_alt.push_back(t.w);
                    }
                }
            }
        }
        if (rc == RAER_OK && ing.failed()) {  // the stream stopped on an error, not at the end of the file
            set_error("%s", ing.error().c_str());
            rc = RAER_ERR_IO;
        }
    }
    cudaStreamDestroy(st);
    close_all();
    if (rc != RAER_OK) return rc < 0 ? rc : RAER_ERR;
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
