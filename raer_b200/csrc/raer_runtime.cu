// raer_runtime.cu -- device context, batch transfer, kernel orchestration and the C ABI of the
// bulk path (include/raer_b200.h). Host logic above it: raer_ingest.cpp. Kernels: raer_kernels.cu.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

#include "raer_dev.h"
#include "raer_host.h"

namespace raer {
const char* last_error();
}
using raer::set_error;

extern "C" int raer_launch_overlap(const raer_read_hdr* hdr, const uint32_t* cigar, const uint8_t* seq, uint8_t* qual,
                                   const int32_t* pair_b, long long n_pairs, int mode, cudaStream_t st, int* scratch);
extern "C" size_t raer_tile_smem_bytes(int n_files, int W, int S, int* cnt_words, int* slot_words);
extern "C" int raer_fast_width(int n_files);
extern "C" int raer_fast_index_count(PlpParams* P, long long n_reads, int* cnt, int* off, int* cur, int* total, cudaStream_t st);
extern "C" int raer_fast_index_fill(PlpParams* P, long long n_reads, const int* off, int* cur, int* cur_back, int* list, int cap,
                                    cudaStream_t st);
extern "C" int raer_launch_tiles(PlpParams* P, int n_sm, cudaStream_t st, int* launches, cudaEvent_t e0, cudaEvent_t e1);
extern "C" int raer_lane_plan(int n_files, int* stage_bytes);
extern "C" int raer_lane_first(PlpParams* P, int* first, cudaStream_t st);
struct raer_gi;
extern "C" raer_gi* raer_gi_create(int device, int nf, const char* const* paths, const char* const* indexes);
extern "C" void raer_gi_destroy(raer_gi* g);
extern "C" int raer_gi_contig(raer_gi* g, int tid, const int* ftid, int contig_len, const raer::IngestConf* ic,
                              const raer::RegionIndex::Chr* site, int reg_beg, int reg_end, int* n_windows,
                              const uint32_t** site_mask_dev, int64_t* aligned_bases);
extern "C" int raer_gi_window(raer_gi* g, int j, raer_read_batch* b, int64_t* file_off);
extern "C" void raer_gi_stats(raer_gi* g, int64_t* launches, int64_t* inflated, int64_t* compressed, int64_t* records);

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            set_error("CUDA error %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__);  \
            return RAER_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

extern "C" const char* raer_last_error(void) { return raer::last_error(); }
extern "C" const char* raer_version(void) { return "raer_b200 0.1 (sm_100a)"; }
extern "C" int raer_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void* raer_pinned_alloc(size_t n, int* pinned) {
    void* p = nullptr;
    static int have_dev = -1;
    if (have_dev < 0) have_dev = raer_device_count() > 0;
    if (have_dev && cudaHostAlloc(&p, n, cudaHostAllocDefault) == cudaSuccess) { *pinned = 1; return p; }
    if (have_dev) cudaGetLastError();
    *pinned = 0;
    return malloc(n);
}
extern "C" void raer_pinned_free(void* p, int pinned) {
    if (!p) return;
    if (pinned) cudaFreeHost(p);
    else free(p);
}

// ------------------------------------------------------------------------------------- context
// Device buffers of a context come from the device's stream-ordered memory pool, on the context's stream: growing one
// does not synchronise the device, and what a call frees is still there for the next call of the process (cudaMalloc /
// cudaFree cost between 5 and 270 ms per raer_pileup() call, depending on what else the process had allocated).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaStream_t st = nullptr;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFreeAsync(p, st);
        size_t want = n + n / 4 + 256;
        if (cudaMallocAsync(&p, want, st) != cudaSuccess) { p = nullptr; cap = 0; cudaGetLastError(); set_error("cudaMallocAsync(%zu) failed", want); return RAER_ERR_CUDA; }
        cap = want;
        return 0;
    }
    void release() { if (p) cudaFreeAsync(p, st); p = nullptr; cap = 0; }
};

struct raer_ctx {
    int device = 0, n_sm = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6];
    DevBuf hdr, maxend, cigar, seq, qual, pair_b, mask;
    DevBuf phase_clk;                      // RAER_PHASE_PROF: 16 clock sums of k_pileup_fast
    DevBuf aei;                            // aei_mode: [n_files][16] sums of the batch
    std::vector<int64_t> last_aei;
    DevBuf umi_holes;                      // bulk UMI mode: (read, position) pairs that stay counted
    DevBuf ov_scratch;                     // mate overlap: compacted list of the CIGAR-walking pairs
    DevBuf tile_idx, tile_list;            // k_pileup_fast: counts | offsets | cursors, and the per-tile read lists
    DevBuf tile_first, pmask_own, pmask_cin;  // k_pileup_lane: own ranges of the tiles, counted-match masks of phase 1
    DevBuf tile_ranges, tile_rows, small;  // small: tile_counter(4) | pad | row_counter(8) | err(4)
    DevBuf row_pos, row_meta, row_stats, row_counts, row_alt;
    DevBuf ref;
    const char* ref_ptr = nullptr;
    int64_t ref_len = 0;
    int ref_tid = -1;
    // last-run info
    int n_tiles = 0, W = 0;
    long long row_cap = 0;
    float ms_h2d = 0, ms_kernel = 0, ms_d2h = 0;
    int64_t last_launches = 0;
    // pipelined host path (raer_pipe_*): what is in flight on this context
    bool pending = false;
    raer_bulk_conf p_conf;
    raer_read_batch p_dev;            // the batch as it sits in device memory
    std::vector<int64_t> p_file_off;
    const uint32_t* p_mask = nullptr;
    // optional per-launch timing of the dominant kernel (bench roofline)
    bool prof = false;
    std::vector<cudaEvent_t> pev;
    size_t pev_used = 0;
};

static std::vector<DevBuf*> ctx_bufs(raer_ctx* c) {
    return {&c->hdr, &c->maxend, &c->cigar, &c->seq, &c->qual, &c->pair_b, &c->mask, &c->phase_clk, &c->aei, &c->umi_holes,
            &c->ov_scratch, &c->tile_idx, &c->tile_list, &c->tile_first, &c->pmask_own, &c->pmask_cin, &c->tile_ranges,
            &c->tile_rows, &c->small, &c->row_pos, &c->row_meta, &c->row_stats, &c->row_counts, &c->row_alt, &c->ref};
}

extern "C" raer_ctx* raer_ctx_create(int32_t device) {
    int n = raer_device_count();
    if (n <= 0 || device >= n) { set_error("no usable CUDA device (the product has no CPU path)"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    raer_ctx* c = new raer_ctx();
    c->device = device;
    {   // (cudaGetDeviceProperties takes tens of milliseconds: one attribute is all that is needed)
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) c->n_sm = v;
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; set_error("stream"); return nullptr; }
    for (auto& e : c->ev) cudaEventCreate(&e);
    for (DevBuf* b : ctx_bufs(c)) b->st = c->stream;
    {   // keep up to 8 GB of freed buffers in the pool for the next call (raer_trim_device_cache() returns them)
        cudaMemPool_t pool;
        unsigned long long keep = 0;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess && keep < (8ull << 30)) {
            keep = 8ull << 30;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    return c;
}
extern "C" void raer_ctx_destroy(raer_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->phase_clk.p) {
        unsigned long long v[16];
        cudaDeviceSynchronize();
        if (cudaMemcpy(v, c->phase_clk.p, sizeof(v), cudaMemcpyDeviceToHost) == cudaSuccess) {
            double tot = 0;
            for (int k = 0; k < 7; ++k) tot += (double)v[k];
            fprintf(stderr, "[raer phase prof] CTA cycles: fetch %.1f%% clear %.1f%% phase1 %.1f%% coverage %.1f%% decide+rows %.1f%% "
                            "pass2 %.1f%% stats %.1f%% | barrier wait, warp cycles / (16 x CTA cycles): phase1 %.1f%% pass2 %.1f%%\n",
                    100 * v[0] / tot, 100 * v[1] / tot, 100 * v[2] / tot, 100 * v[3] / tot, 100 * v[4] / tot, 100 * v[5] / tot,
                    100 * v[6] / tot, 100 * v[8] / (16 * tot), 100 * v[9] / (16 * tot));
        }
    }
    cudaStreamSynchronize(c->stream);
    for (DevBuf* b : ctx_bufs(c)) b->release();
    cudaStreamSynchronize(c->stream);
    for (auto& e : c->ev) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}
extern "C" void* raer_ctx_stream(raer_ctx* c) { return (void*)c->stream; }
// switch per-launch CUDA-event timing of k_pileup_tiles on/off (resets the accumulated launches)
extern "C" void raer_ctx_profile(raer_ctx* c, int on) { c->prof = on != 0; c->pev_used = 0; }
// after a stream synchronize: total milliseconds and number of k_pileup_tiles launches since raer_ctx_profile(1)
extern "C" int raer_ctx_profile_read(raer_ctx* c, double* ms_sum, int64_t* n_launches) {
    double tot = 0;
    for (size_t i = 0; i + 1 < c->pev_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->pev[i], c->pev[i + 1]) != cudaSuccess) { cudaGetLastError(); return RAER_ERR_CUDA; }
        tot += ms;
    }
    *ms_sum = tot;
    *n_launches = (int64_t)(c->pev_used / 2);
    return RAER_OK;
}

extern "C" int raer_ctx_set_reference(raer_ctx* c, int32_t tid, const char* seq, int64_t len) {
    cudaSetDevice(c->device);
    if (len > 0) {
        int rc = c->ref.ensure((size_t)len);
        if (rc) return rc;
        CK(cudaMemcpyAsync(c->ref.p, seq, (size_t)len, cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    c->ref_ptr = len > 0 ? (const char*)c->ref.p : nullptr;
    c->ref_len = len;
    c->ref_tid = tid;
    return RAER_OK;
}
extern "C" int raer_ctx_set_reference_device(raer_ctx* c, int32_t tid, const void* dseq, int64_t len) {
    c->ref_ptr = (const char*)dseq;
    c->ref_len = len;
    c->ref_tid = tid;
    return RAER_OK;
}

// choose the tile width: counters [file][strand][6][W] x 4 B must fit, leave room for 2 CTAs/SM when cheap
// the small device block of a run: tile_counter | row_counter | err | list total | khash growth positions
struct SmallState { unsigned tc, pad; unsigned long long rows; int err; int pad2[3]; int list_total; int pad3[7];
                    int grow[2 * RAER_MAX_FILES]; };
static_assert(offsetof(SmallState, grow) == 64, "SmallState layout");

static int choose_tile_width(int n_files) {
    // 227 KB per CTA minus the per-warp read stages (raer_tile_smem_bytes) and slack
    int cw, sw;
    int fixed = (int)raer_tile_smem_bytes(n_files, 0, 0, &cw, &sw);
    int budget = 227 * 1024 - fixed - 512;
    int w = budget / (50 * n_files + 14);  // 48 B of counters per position and file, + 1/32 bank-skew padding
    w = std::min(w, n_files <= 2 ? 1024 : 2048);
    w &= ~31;
    return std::max(w, 32);
}

enum { RUN_FORCE_GENERIC = 1, RUN_SKIP_OVERLAP = 2 };

// Enqueue the kernels of one batch on the context's stream. Nothing here waits for the device: conditions only
// the device can see (tile lists larger than their allocation, coverage beyond the fast kernel's 16-bit counters)
// come back through err_flag and finish_run() re-runs the batch accordingly.
static int run_device(raer_ctx* c, const raer_bulk_conf* conf, const raer_read_batch* b, const uint32_t* d_mask,
                      int64_t* launches, int flags = 0, long long list_need = 0, long long row_need = 0) {
    const int nf = b->n_files;
    PlpParams P;
    memset(&P, 0, sizeof(P));
    P.hdr = b->hdr; P.maxend = b->maxend; P.cigar = b->cigar; P.seq = b->seq; P.qual = (uint8_t*)b->qual;
    for (int i = 0; i <= nf; ++i) P.file_off[i] = b->file_off[i];
    P.n_files = nf;
    P.t_beg = b->beg; P.t_end = b->end;
    // k_pileup_fast: every filter, as long as the threshold is within the byte-parallel compare's range.
    // RAER_KERNEL=generic forces k_pileup_tiles (tests run both kernels on the same inputs).
    const bool only_bq = !(conf->trim_5p > 0 || conf->trim_3p > 0 || conf->ftrim_5p > 0 || conf->ftrim_3p > 0 ||
                           conf->splice_dist || conf->min_overhang || conf->indel_dist || conf->n_mm_type > 0 || conf->n_mm > 0);
    const char* kforce = getenv("RAER_KERNEL");
    const int fw = raer_fast_width(nf);
    (void)only_bq;
    P.use_fast = conf->min_bq <= 128 && fw > 0 && !(kforce && !strcmp(kforce, "generic")) &&
                 !(flags & RUN_FORCE_GENERIC);
    P.W = P.use_fast ? fw : choose_tile_width(nf);
    P.fast = !(conf->umi_mode || conf->trim_5p > 0 || conf->trim_3p > 0 || conf->ftrim_5p > 0 || conf->ftrim_3p > 0 || conf->splice_dist ||
               conf->min_overhang || conf->indel_dist || conf->n_mm_type > 0 || conf->n_mm > 0);
    // k_pileup_lane (lane per read, TMA-staged chunks) when only the mapq / base-quality filters are configured;
    // RAER_KERNEL=fast keeps k_pileup_fast for those calls too
    if (P.use_fast && P.fast && !(kforce && !strcmp(kforce, "fast"))) {
        int sb = 0;
        const int lw = raer_lane_plan(nf, &sb);
        if (lw > 0) { P.use_lane = 1; P.carry_only = 1; P.W = lw; P.stage_bytes = sb; }
    }
    if (P.use_fast && getenv("RAER_FAST_W")) P.W = std::max(32, std::min(P.W, atoi(getenv("RAER_FAST_W")) & ~31));
    long long span = (long long)b->end - b->beg;
    P.n_tiles = span > 0 ? (int)((span + P.W - 1) / P.W) : 0;
    P.ref = c->ref_ptr; P.ref_len = (int)std::min<int64_t>(c->ref_len, INT_MAX);
    P.min_depth = conf->min_depth; P.min_bq = conf->min_bq; P.trim_5p = conf->trim_5p; P.trim_3p = conf->trim_3p;
    P.indel_dist = conf->indel_dist; P.splice_dist = conf->splice_dist; P.min_overhang = conf->min_overhang;
    P.nmer = conf->nmer; P.min_var_reads = conf->min_var_reads; P.n_mm_type = conf->n_mm_type; P.n_mm = conf->n_mm;
    P.report_multiallelic = conf->report_multiallelic; P.want_stats = conf->want_stats;
    P.ftrim_5p = conf->ftrim_5p; P.ftrim_3p = conf->ftrim_3p; P.min_af = conf->min_af;
    P.any_okv = 0;
    P.aei_mode = conf->aei_mode;
    if (P.aei_mode) {
        if (!P.use_fast) { set_error("aei_mode needs the fast tile kernel"); return RAER_ERR_LIMIT; }
        int rc2 = c->aei.ensure((size_t)RAER_MAX_FILES * 16 * sizeof(unsigned long long));
        if (rc2) return rc2;
        CK(cudaMemsetAsync(c->aei.p, 0, (size_t)nf * 16 * sizeof(unsigned long long), c->stream));
        P.aei_out = (unsigned long long*)c->aei.p;
    }
    P.umi_mode = conf->umi_mode;
    P.umi_holes = b->umi_holes;
    P.n_umi_holes = (int)b->n_umi_holes;
    if (P.umi_mode && !P.use_fast) {
        set_error("umi_tag in pileup_sites needs the fast tile kernel (min_base_quality <= 128, coverage <= 65535)");
        return RAER_ERR_LIMIT;
    }
    for (int i = 0; i < nf; ++i) {
        P.min_mapq[i] = conf->min_mapq[i];
        P.okv[i] = conf->only_keep_variants[i];
        P.any_okv |= P.okv[i] != 0;
    }
    P.reg_beg = conf->reg_beg; P.reg_end = conf->reg_end;
    P.site_mask = d_mask; P.mask_beg = conf->mask_beg; P.mask_bits = conf->mask_bits;
    c->n_tiles = P.n_tiles; c->W = P.W;
    int nl = 0;
    int rc;
    if ((rc = c->small.ensure(sizeof(SmallState)))) return rc;
    CK(cudaMemsetAsync(c->small.p, 0, sizeof(SmallState), c->stream));
    P.tile_counter = (unsigned int*)c->small.p;
    P.row_counter = (unsigned long long*)((char*)c->small.p + 8);
    P.err_flag = (int*)((char*)c->small.p + 16);
    P.grow = (int*)((char*)c->small.p + offsetof(SmallState, grow));
    P.phase_clk = nullptr;
    if (getenv("RAER_PHASE_PROF")) {
        if (!c->phase_clk.p) {
            if ((rc = c->phase_clk.ensure(16 * 8))) return rc;
            CK(cudaMemsetAsync(c->phase_clk.p, 0, 16 * 8, c->stream));
        }
        P.phase_clk = (unsigned long long*)c->phase_clk.p;
    }
    if (P.n_tiles > 0) {
        if ((rc = c->tile_ranges.ensure((size_t)P.n_tiles * nf * sizeof(int2)))) return rc;
        if ((rc = c->tile_rows.ensure((size_t)P.n_tiles * sizeof(int2)))) return rc;
        // Row buffers: a site writes at most two rows, and only where a read lies, so the packed bases of the batch bound
        // the rows far below 2 x span for a sparse BAM or a site-list call on a long contig (40 B x files + 32 B per row:
        // sizing by span asked for tens of GB there). The bound is loose on purpose and not exact (deletions, min_depth 0
        // inside ref-skips): the device keeps counting past the end of the buffers and the batch is re-run with the count.
        long long cap = 2 * span + 64;
        if (row_need > 0) cap = std::min(cap, row_need + 64);
        else cap = std::min(cap, std::max<long long>(1 << 16, 4 * (long long)b->n_sq_bytes + 64));
        if (getenv("RAER_TEST_ROWCAP") && row_need == 0) cap = std::min<long long>(cap, std::max(1, atoi(getenv("RAER_TEST_ROWCAP"))));
        if ((rc = c->row_pos.ensure((size_t)cap * 4))) return rc;
        if ((rc = c->row_meta.ensure((size_t)cap * 4))) return rc;
        if ((rc = c->row_stats.ensure((size_t)cap * 24))) return rc;
        if ((rc = c->row_counts.ensure((size_t)cap * nf * 32))) return rc;
        if ((rc = c->row_alt.ensure((size_t)cap * nf * 4))) return rc;
        c->row_cap = cap;
        P.row_cap = cap;
        P.tile_ranges = (int2*)c->tile_ranges.p; P.tile_rows = (int2*)c->tile_rows.p;
        P.row_pos = (int*)c->row_pos.p; P.row_meta = (unsigned*)c->row_meta.p; P.row_stats = (double*)c->row_stats.p;
        P.row_counts = (int*)c->row_counts.p; P.row_alt = (unsigned*)c->row_alt.p;
        // second-pass slots: reuse the counter region, at least 32 sites per round
        int cw, sw;
        P.S = 0;
        raer_tile_smem_bytes(nf, P.W, 0, &cw, &sw);
        P.S = std::max(32, cw / sw);
        if (P.S > 4096) P.S = 4096;
        if (conf->remove_overlaps && b->n_pairs > 0 && !(flags & RUN_SKIP_OVERLAP)) {
            if ((rc = c->ov_scratch.ensure(((size_t)b->n_pairs + 1) * sizeof(int)))) return rc;
            int r = raer_launch_overlap(b->hdr, b->cigar, b->seq, (uint8_t*)b->qual, b->pair_b, b->n_pairs,
                                        conf->remove_overlaps, c->stream, (int*)c->ov_scratch.p);
            if (r < 0) { set_error("overlap kernel launch failed"); return RAER_ERR_CUDA; }
            nl += r;
        }
        if (P.use_fast) {
            // per-tile read lists: count, scan, fill. The lists are sized from the read count (a read overlaps
            // 1 + span / W tiles); if that guess is too small the device says so and the batch is re-run.
            const size_t ne = (size_t)P.n_tiles * nf + 1;
            if ((rc = c->tile_idx.ensure(ne * 3 * sizeof(int) + 16))) return rc;
            int* cnt = (int*)c->tile_idx.p;
            int* off = cnt + ne;
            int* cur = off + ne;
            int* d_total = (int*)((char*)c->small.p + 32);
            const long long n_reads = b->file_off[nf];
            long long want = std::max(list_need, (P.carry_only ? 1 : 2) * n_reads + (long long)ne + 1024);
            if (want > INT_MAX) want = INT_MAX;
            if ((rc = c->tile_list.ensure((size_t)want * sizeof(int)))) return rc;
            int cap = (int)std::min<size_t>(c->tile_list.cap / sizeof(int), (size_t)INT_MAX);
            // test hook: pretend the first guess was this small, so that the re-run path is exercised
            if (list_need == 0 && getenv("RAER_TEST_LISTCAP")) cap = std::min(cap, std::max(1, atoi(getenv("RAER_TEST_LISTCAP"))));
            int r = raer_fast_index_count(&P, n_reads, cnt, off, cur, d_total, c->stream);
            if (r < 0) { set_error("tile index kernels failed"); return RAER_ERR_CUDA; }
            nl += r;
            r = raer_fast_index_fill(&P, n_reads, off, cur, cnt, (int*)c->tile_list.p, cap, c->stream);
            if (r < 0) { set_error("tile index kernels failed"); return RAER_ERR_CUDA; }
            nl += r;
            P.tile_off = off;
            P.tile_list = (const int*)c->tile_list.p;
            P.tile_list_cap = cap;
            if (P.use_lane) {
                if ((rc = c->tile_first.ensure(((size_t)P.n_tiles + 1) * nf * sizeof(int)))) return rc;
                if (P.want_stats && !P.aei_mode) {
                    if ((rc = c->pmask_own.ensure((size_t)n_reads * RAER_LANE_MW * 4 + 16))) return rc;
                    if ((rc = c->pmask_cin.ensure(((size_t)cap + 32) * RAER_LANE_MW * 4 + 16))) return rc;
                }
                P.tile_first = (const int*)c->tile_first.p;
                P.pmask_own = (uint32_t*)c->pmask_own.p;
                P.pmask_cin = (uint32_t*)c->pmask_cin.p;
                r = raer_lane_first(&P, (int*)c->tile_first.p, c->stream);
                if (r < 0) { set_error("tile index kernels failed"); return RAER_ERR_CUDA; }
                nl += r;
            }
        }
        int tl = 0;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (c->prof) {
            while (c->pev.size() < c->pev_used + 2) { cudaEvent_t e; cudaEventCreate(&e); c->pev.push_back(e); }
            e0 = c->pev[c->pev_used++];
            e1 = c->pev[c->pev_used++];
        }
        if (raer_launch_tiles(&P, c->n_sm, c->stream, &tl, e0, e1) < 0) {
            set_error("tile kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return RAER_ERR_CUDA;
        }
        nl += tl;
    }
    if (launches) *launches += nl;
    return RAER_OK;
}


// Wait for the batch enqueued by run_device(), re-running it when the device asked for it. b must be the batch
// in device memory (qualities already rewritten by the mate-overlap kernel: re-runs skip it).
static int finish_run(raer_ctx* c, const raer_bulk_conf* conf, const raer_read_batch* b, const uint32_t* d_mask,
                      int64_t* launches, SmallState* out) {
    int flags = RUN_SKIP_OVERLAP;  // what the re-runs have learnt so far
    long long list_need = 0, row_need = 0;
    for (int attempt = 0; attempt < 6; ++attempt) {
        SmallState s;
        CK(cudaMemcpyAsync(&s, c->small.p, sizeof(s), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (s.err == RAER_RETRY_LISTCAP) list_need = (long long)s.list_total + 16;
        else if (s.err == RAER_RETRY_DEEP) flags |= RUN_FORCE_GENERIC;
        else if (s.err == RAER_RETRY_ROWCAP) row_need = (long long)s.rows;
        else if (s.err) {
            set_error(s.err == RAER_ERR_MD ? "inconsistent MD tag with max_mismatch_type" : "device reported error %d", s.err);
            return s.err;
        } else {
            *out = s;
            return RAER_OK;
        }
        int rc = run_device(c, conf, b, d_mask, launches, flags, list_need, row_need);
        if (rc) return rc;
    }
    set_error("device kept asking for a re-run");
    return RAER_ERR;
}

extern "C" int raer_batch_pileup_device(raer_ctx* c, const raer_bulk_conf* conf, const raer_read_batch* b,
                                        int64_t* n_rows, int64_t* launches) {
    cudaSetDevice(c->device);
    if (conf->n_files != b->n_files || b->n_files > RAER_MAX_FILES) { set_error("bad file count"); return RAER_ERR_ARG; }
    int rc = run_device(c, conf, b, conf->site_mask, launches);
    if (rc) return rc;
    SmallState s;
    rc = finish_run(c, conf, b, conf->site_mask, launches, &s);
    if (rc) return rc;
    if (n_rows) *n_rows = (int64_t)s.rows;
    return RAER_OK;
}

static int fetch_rows(raer_ctx* c, int nf, const SmallState& s, raer_rows* out);
// device-resident batch -> rows in host memory (GPU ingest: the batch was packed by kernels, nothing to upload)
extern "C" int raer_batch_pileup_device_rows(raer_ctx* c, const raer_bulk_conf* conf, const raer_read_batch* b, raer_rows* out,
                                             int64_t* launches) {
    cudaSetDevice(c->device);
    memset(out, 0, sizeof(*out));
    out->n_files = b->n_files;
    for (int k = 0; k < 2 * RAER_MAX_FILES; ++k) out->grow_pos[k] = INT32_MAX;
    if (conf->n_files != b->n_files || b->n_files > RAER_MAX_FILES) { set_error("bad file count"); return RAER_ERR_ARG; }
    if (b->n_reads == 0 || b->end <= b->beg) return RAER_OK;
    cudaEventRecord(c->ev[1], c->stream);
    int rc = run_device(c, conf, b, conf->site_mask, launches);
    if (rc) return rc;
    cudaEventRecord(c->ev[2], c->stream);
    SmallState s;
    rc = finish_run(c, conf, b, conf->site_mask, launches, &s);
    if (rc) return rc;
    rc = fetch_rows(c, b->n_files, s, out);
    if (rc == RAER_OK && conf->aei_mode) {
        c->last_aei.assign((size_t)b->n_files * 16, 0);
        CK(cudaMemcpyAsync(c->last_aei.data(), c->aei.p, c->last_aei.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    cudaEventRecord(c->ev[3], c->stream);
    cudaEventSynchronize(c->ev[3]);
    c->ms_h2d = 0;
    cudaEventElapsedTime(&c->ms_kernel, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&c->ms_d2h, c->ev[2], c->ev[3]);
    return rc;
}

extern "C" void raer_rows_free(raer_rows* r) {
    if (!r) return;
    free(r->pos); free(r->meta); free(r->stats); free(r->counts); free(r->altcode);
    memset(r, 0, sizeof(*r));
}

// download the rows of the last run in tile order (host-side ordered concatenation)
static int fetch_rows(raer_ctx* c, int nf, const SmallState& s, raer_rows* out) {
    memset(out, 0, sizeof(*out));
    out->n_files = nf;
    size_t n = (size_t)s.rows;
    out->n_rows = (int64_t)n;
    for (int k = 0; k < 2 * RAER_MAX_FILES; ++k) out->grow_pos[k] = s.grow[k] ? 0x7fffffff - s.grow[k] : INT32_MAX;
    if (n == 0 || c->n_tiles == 0) return RAER_OK;
    std::vector<int2> tr(c->n_tiles);
    std::vector<int> pos(n);
    std::vector<unsigned> meta(n), alt(n * nf);
    std::vector<double> st(n * 3);
    std::vector<int> cnt(n * nf * 8);
    CK(cudaMemcpyAsync(tr.data(), c->tile_rows.p, sizeof(int2) * c->n_tiles, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(pos.data(), c->row_pos.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(meta.data(), c->row_meta.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(st.data(), c->row_stats.p, n * 24, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(cnt.data(), c->row_counts.p, n * nf * 32, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(alt.data(), c->row_alt.p, n * nf * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    out->pos = (int32_t*)malloc(n * 4);
    out->meta = (uint32_t*)malloc(n * 4);
    out->stats = (double*)malloc(n * 24);
    out->counts = (int32_t*)malloc(n * nf * 32);
    out->altcode = (uint32_t*)malloc(n * nf * 4);
    size_t o = 0;
    for (int t = 0; t < c->n_tiles; ++t) {
        int base = tr[t].x, k = tr[t].y;
        if (k <= 0) continue;
        memcpy(out->pos + o, pos.data() + base, (size_t)k * 4);
        memcpy(out->meta + o, meta.data() + base, (size_t)k * 4);
        memcpy(out->stats + o * 3, st.data() + (size_t)base * 3, (size_t)k * 24);
        memcpy(out->counts + o * nf * 8, cnt.data() + (size_t)base * nf * 8, (size_t)k * nf * 32);
        memcpy(out->altcode + o * nf, alt.data() + (size_t)base * nf, (size_t)k * nf * 4);
        o += k;
    }
    if (o != n) { set_error("row bookkeeping mismatch (%zu vs %zu)", o, n); return RAER_ERR; }
    return RAER_OK;
}

// Host batch -> device buffers of the context and the kernels, all enqueued on the context's stream (no wait).
// The host arrays must stay valid until host_collect() returned.
static int host_submit(raer_ctx* c, const raer_bulk_conf* conf, const raer_read_batch* hb) {
    const int nf = hb->n_files;
    int rc;
    size_t nr = (size_t)hb->n_reads;
    if ((rc = c->hdr.ensure(nr * sizeof(raer_read_hdr)))) return rc;
    if ((rc = c->maxend.ensure(nr * 4))) return rc;
    if ((rc = c->cigar.ensure((size_t)hb->n_cigar_words * 4 + 16))) return rc;
    if ((rc = c->seq.ensure((size_t)hb->n_sq_bytes + 16))) return rc;
    if ((rc = c->qual.ensure((size_t)hb->n_sq_bytes * 2 + 16))) return rc;
    if ((rc = c->pair_b.ensure((size_t)hb->n_pairs * 4 + 16))) return rc;
    cudaEventRecord(c->ev[0], c->stream);
    CK(cudaMemcpyAsync(c->hdr.p, hb->hdr, nr * sizeof(raer_read_hdr), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->maxend.p, hb->maxend, nr * 4, cudaMemcpyHostToDevice, c->stream));
    if (hb->n_cigar_words) CK(cudaMemcpyAsync(c->cigar.p, hb->cigar, (size_t)hb->n_cigar_words * 4, cudaMemcpyHostToDevice, c->stream));
    if (hb->n_sq_bytes) {
        CK(cudaMemcpyAsync(c->seq.p, hb->seq, (size_t)hb->n_sq_bytes, cudaMemcpyHostToDevice, c->stream));
        CK(cudaMemcpyAsync(c->qual.p, hb->qual, (size_t)hb->n_sq_bytes * 2, cudaMemcpyHostToDevice, c->stream));
    }
    if (hb->n_pairs) CK(cudaMemcpyAsync(c->pair_b.p, hb->pair_b, (size_t)hb->n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
    if (hb->n_umi_holes) {
        if ((rc = c->umi_holes.ensure((size_t)hb->n_umi_holes * 8))) return rc;
        CK(cudaMemcpyAsync(c->umi_holes.p, hb->umi_holes, (size_t)hb->n_umi_holes * 8, cudaMemcpyHostToDevice, c->stream));
    }
    const uint32_t* d_mask = nullptr;
    if (conf->site_mask) {
        size_t words = ((size_t)conf->mask_bits + 31) / 32;
        if ((rc = c->mask.ensure(words * 4 + 16))) return rc;
        CK(cudaMemcpyAsync(c->mask.p, conf->site_mask, words * 4, cudaMemcpyHostToDevice, c->stream));
        d_mask = (const uint32_t*)c->mask.p;
    }
    cudaEventRecord(c->ev[1], c->stream);
    c->p_conf = *conf;
    c->p_dev = *hb;
    c->p_dev.umi_holes = hb->n_umi_holes ? (const int32_t*)c->umi_holes.p : nullptr;
    c->p_file_off.assign(hb->file_off, hb->file_off + nf + 1);
    c->p_dev.file_off = c->p_file_off.data();
    c->p_dev.hdr = (const raer_read_hdr*)c->hdr.p; c->p_dev.maxend = (const int32_t*)c->maxend.p;
    c->p_dev.cigar = (const uint32_t*)c->cigar.p; c->p_dev.seq = (const uint8_t*)c->seq.p; c->p_dev.qual = (const uint8_t*)c->qual.p;
    c->p_dev.pair_b = (const int32_t*)c->pair_b.p;
    c->p_mask = d_mask;
    int64_t nl = 0;
    rc = run_device(c, &c->p_conf, &c->p_dev, d_mask, &nl);
    if (rc) return rc;
    c->last_launches = nl;
    cudaEventRecord(c->ev[2], c->stream);
    c->pending = true;
    return RAER_OK;
}

static int host_collect(raer_ctx* c, raer_rows* out) {
    c->pending = false;
    SmallState st;
    int64_t nl = 0;
    int rc = finish_run(c, &c->p_conf, &c->p_dev, c->p_mask, &nl, &st);
    c->last_launches += nl;
    if (rc) return rc;
    rc = fetch_rows(c, c->p_dev.n_files, st, out);
    if (rc == RAER_OK && c->p_conf.aei_mode) {
        c->last_aei.assign((size_t)c->p_dev.n_files * 16, 0);
        CK(cudaMemcpyAsync(c->last_aei.data(), c->aei.p, c->last_aei.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    cudaEventRecord(c->ev[3], c->stream);
    cudaEventSynchronize(c->ev[3]);
    cudaEventElapsedTime(&c->ms_h2d, c->ev[0], c->ev[1]);
    cudaEventElapsedTime(&c->ms_kernel, c->ev[1], c->ev[2]);
    cudaEventElapsedTime(&c->ms_d2h, c->ev[2], c->ev[3]);
    return rc;
}

extern "C" int raer_batch_pileup_host(raer_ctx* c, const raer_bulk_conf* conf, const raer_read_batch* hb, raer_rows* out) {
    cudaSetDevice(c->device);
    const int nf = hb->n_files;
    if (conf->n_files != nf || nf > RAER_MAX_FILES || nf < 1) { set_error("bad file count"); return RAER_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    out->n_files = nf;
    if (hb->n_reads == 0 || hb->end <= hb->beg) return RAER_OK;
    int rc = host_submit(c, conf, hb);
    if (rc) return rc;
    return host_collect(c, out);
}

// ------------------------------------------------------------------------------------- pipelined host path
// Two contexts (own stream, own device buffers) used alternately: the H2D copy of batch i + 1 runs while the
// kernels of batch i execute and the rows of batch i - 1 travel back. Batches complete in submission order.
struct raer_pipe {
    raer_ctx* c[2] = {nullptr, nullptr};
    int64_t submitted = 0, collected = 0;
    bool empty[2] = {false, false};  // the slot holds a batch with no reads (nothing was enqueued)
    int nf[2] = {0, 0};
};

extern "C" raer_pipe* raer_pipe_create(int32_t device) {
    raer_pipe* p = new raer_pipe();
    for (int i = 0; i < 2; ++i) {
        p->c[i] = raer_ctx_create(device);
        if (!p->c[i]) { if (i) raer_ctx_destroy(p->c[0]); delete p; return nullptr; }
    }
    return p;
}
extern "C" void raer_pipe_destroy(raer_pipe* p) {
    if (!p) return;
    for (int i = 0; i < 2; ++i) raer_ctx_destroy(p->c[i]);
    delete p;
}
// Enqueue one host batch together with the reference bases of its contig (raw FASTA bytes, host memory). At most
// two batches may be in flight: returns RAER_ERR_LIMIT if raer_pipe_collect() has to be called first.
extern "C" int raer_pipe_submit(raer_pipe* p, const raer_bulk_conf* conf, const raer_read_batch* hb, int32_t tid,
                                const char* ref, int64_t ref_len) {
    if (p->submitted - p->collected >= 2) { set_error("two batches already in flight"); return RAER_ERR_LIMIT; }
    const int k = (int)(p->submitted & 1);
    raer_ctx* c = p->c[k];
    cudaSetDevice(c->device);
    const int nf = hb->n_files;
    if (conf->n_files != nf || nf > RAER_MAX_FILES || nf < 1) { set_error("bad file count"); return RAER_ERR_ARG; }
    p->nf[k] = nf;
    p->empty[k] = hb->n_reads == 0 || hb->end <= hb->beg;
    if (!p->empty[k]) {
        int rc;
        if (ref_len > 0) {
            if ((rc = c->ref.ensure((size_t)ref_len))) return rc;
            CK(cudaMemcpyAsync(c->ref.p, ref, (size_t)ref_len, cudaMemcpyHostToDevice, c->stream));
        }
        c->ref_ptr = ref_len > 0 ? (const char*)c->ref.p : nullptr;
        c->ref_len = ref_len;
        c->ref_tid = tid;
        if ((rc = host_submit(c, conf, hb))) return rc;
    }
    ++p->submitted;
    return RAER_OK;
}
// Rows of the oldest batch in flight.
extern "C" int raer_pipe_collect(raer_pipe* p, raer_rows* out) {
    if (p->submitted == p->collected) { set_error("no batch in flight"); return RAER_ERR_ARG; }
    const int k = (int)(p->collected & 1);
    ++p->collected;
    memset(out, 0, sizeof(*out));
    out->n_files = p->nf[k];
    if (p->empty[k]) return RAER_OK;
    cudaSetDevice(p->c[k]->device);
    return host_collect(p->c[k], out);
}

extern "C" int raer_ctx_last_aei(raer_ctx* c, int64_t* sums, int32_t n_files) {
    if (!c || !sums || (size_t)n_files * 16 != c->last_aei.size()) { set_error("no AEI sums for %d files on this context", n_files); return RAER_ERR_ARG; }
    memcpy(sums, c->last_aei.data(), c->last_aei.size() * sizeof(int64_t));
    return RAER_OK;
}

extern "C" void raer_ctx_last_timing(raer_ctx* c, double* h2d_s, double* kernel_s, double* d2h_s) {
    *h2d_s = c->ms_h2d * 1e-3; *kernel_s = c->ms_kernel * 1e-3; *d2h_s = c->ms_d2h * 1e-3;
}

extern "C" int64_t raer_batch_algorithmic_bytes(const raer_read_batch* b, int64_t n_rows, int64_t ref_bases) {
    // 32 B/record + 4 B/CIGAR op + 0.5 B/base (seq) + 1 B/base (qual)
    // + 1 B/reference base + (40 B x files + 24 B) per output row
    int64_t bases = b->n_sq_bytes * 2;
    return 32 * b->n_reads + 4 * b->n_cigar_words + bases / 2 + bases + ref_bases + n_rows * (40LL * b->n_files + 24);
}

// ------------------------------------------------------------------------------------- ALT strings
// Simulates the 4/8 bucket khash of single-character keys the reference walks to print ALT
// (src/plp.c:193-215; htslib khash.h: home = key & mask, probe i += ++step). State per (file, strand).
extern "C" void raer_altcode_to_string(uint32_t code, int ref_char, int32_t* buckets, char* out) {
    (void)ref_char;
    static const char L[4] = {'A', 'C', 'G', 'T'};
    static const char NT16[17] = "=ACMGRSVTWYHKDBN";
    unsigned mask = code & RAER_ALT_MASK;
    int nins = (code >> RAER_ALT_NINS_SHIFT) & 7;
    char okey = NT16[(code >> RAER_ALT_OCODE_SHIFT) & 0xf];
    char keys[5];
    bool live[5];
    int nk = 0;
    if (nins == 0) {
        // rows without a second pass (want_stats == 0): no order information, use base order
        for (int b = 0; b < 4; ++b)
            if (mask & (1u << b)) { keys[nk] = L[b]; live[nk++] = true; }
    } else {
        for (int k = 0; k < nins && k < 5; ++k) {
            int id = (code >> (RAER_ALT_ORDER_SHIFT + 3 * k)) & 7;
            if (id < 4) { keys[nk] = L[id]; live[nk++] = (mask >> id) & 1u; }
            else { keys[nk] = okey; live[nk++] = (code & RAER_ALT_OTHER) != 0; }
        }
    }
    // khash emulation: home = key & (nb-1), probe i += ++step; a put finding n_occupied >= 0.77*nb resizes
    // (4 -> 8) and re-inserts the live keys in bucket order before going on. The table never shrinks.
    if (*buckets != 8 && *buckets != 16) *buckets = 4;
    int nb = *buckets;
    char slot[16];
    memset(slot, 0, sizeof(slot));
    int occ = 0;
    auto put = [&](char key) {
        unsigned i = (unsigned)(unsigned char)key & (nb - 1), step = 0;
        while (slot[i]) i = (i + (++step)) & (nb - 1);
        slot[i] = key;
        ++occ;
    };
    auto grow = [&]() {
        char old[16];
        int onb = nb;
        memcpy(old, slot, sizeof(old));
        nb *= 2;
        memset(slot, 0, sizeof(slot));
        occ = 0;
        for (int i = 0; i < onb; ++i)
            if (old[i]) put(old[i]);
    };
    for (int k = 0; k < nk; ++k) {
        if (occ >= (int)(nb * 0.77 + 0.5) && nb < 16) grow();
        put(keys[k]);
    }
    // a further put of an already present key with 3 keys occupied also resizes (flag from the device)
    if ((code & RAER_ALT_GROW) && nb == 4) grow();
    *buckets = nb;
    int z = 0;
    for (int i = 0; i < nb; ++i) {
        if (!slot[i]) continue;
        bool alive = false;
        for (int k = 0; k < nk; ++k)
            if (keys[k] == slot[i]) alive = live[k];
        if (!alive) continue;  // pruned by min_allelic_freq after insertion (kh_del keeps the layout)
        if (z && z < 10) out[z++] = ',';
        if (z < 11) out[z++] = slot[i];
    }
    if (z == 0) { out[0] = '-'; z = 1; }
    out[z] = 0;
}

// ------------------------------------------------------------------------------------- .Call(".pileup")
struct ResultSink {
    raer_pileup_result* r;
    int64_t cap = 0;
    std::vector<char> alt;       // [file][cap][12]
    std::vector<int32_t> counts; // [file][cap][8]
};

static void sink_reserve(ResultSink& s, int64_t need) {
    raer_pileup_result* r = s.r;
    if (need <= s.cap) return;
    int64_t ncap = std::max<int64_t>(need + need / 2, 1 << 16);
    r->tid = (int32_t*)realloc(r->tid, ncap * 4);
    r->pos = (int32_t*)realloc(r->pos, ncap * 4);
    r->strand = (char*)realloc(r->strand, ncap);
    r->ref = (char*)realloc(r->ref, ncap);
    r->rpbz = (double*)realloc(r->rpbz, ncap * 8);
    r->vdb = (double*)realloc(r->vdb, ncap * 8);
    r->sor = (double*)realloc(r->sor, ncap * 8);
    char* nalt = (char*)calloc((size_t)r->nbam * ncap, 12);
    int32_t* ncnt = (int32_t*)calloc((size_t)r->nbam * ncap * 8, 4);
    for (int f = 0; f < r->nbam && r->nrows; ++f) {
        memcpy(nalt + (size_t)f * ncap * 12, r->alt + (size_t)f * s.cap * 12, (size_t)r->nrows * 12);
        memcpy(ncnt + (size_t)f * ncap * 8, r->counts + (size_t)f * s.cap * 8, (size_t)r->nrows * 32);
    }
    free(r->alt); free(r->counts);
    r->alt = nalt; r->counts = ncnt;
    s.cap = ncap;
}

extern "C" void raer_pileup_result_free(raer_pileup_result* r) {
    if (!r) return;
    free(r->tid); free(r->pos); free(r->strand); free(r->ref); free(r->rpbz); free(r->vdb); free(r->sor);
    free(r->alt); free(r->counts); free(r->aei);
    if (r->target_names) {
        for (int i = 0; i < r->n_targets; ++i) free(r->target_names[i]);
        free(r->target_names);
    }
    memset(r, 0, sizeof(*r));
}

extern "C" int raer_pileup(const raer_pileup_args* a, raer_pileup_result* out) {
    double t_start = raer::now_seconds();
    memset(out, 0, sizeof(*out));
    // check_plp_args (src/plp.c:998-1055)
    if (!a || a->nbam < 1 || !a->bampaths || !a->fapath || !a->libtype || !a->only_keep_variants || !a->min_mapq) {
        set_error("invalid arguments"); return RAER_ERR_ARG;
    }
    if (a->nbam > RAER_MAX_FILES) { set_error("at most %d BAM files per call", RAER_MAX_FILES); return RAER_ERR_LIMIT; }
    // documented limit, refused before any file or device work: the UMI shadows and the fused AEI reduction live in the
    // fast tile kernels, whose byte-parallel quality compare handles thresholds up to 128
    if ((a->umi_tag || a->aei_mode) && a->int_args[2] > 128) {
        set_error("%s needs min_base_quality <= 128", a->umi_tag ? "umi_tag in pileup_sites" : "the fused AEI reduction");
        return RAER_ERR_LIMIT;
    }
    if (raer_device_count() <= 0) { set_error("no usable CUDA device (the product has no CPU path)"); return RAER_ERR_NO_DEVICE; }
    const int nf = a->nbam;
    out->nbam = nf;

    raer::RegionIndex ridx;
    bool has_ridx = a->n_regions > 0;
    if (has_ridx) {
        std::vector<int32_t> b0(a->n_regions), e0(a->n_regions);
        for (int i = 0; i < a->n_regions; ++i) { b0[i] = a->reg_start[i] - 1; e0[i] = a->reg_end[i] - 1; }
        ridx.build(a->n_regions, a->reg_seqnames, b0.data(), e0.data());
    }
    raer::IngestConf ic;
    ic.n_files = nf;
    ic.keep_flag[0] = (uint32_t)a->int_args[12];
    ic.keep_flag[1] = (uint32_t)a->int_args[13];
    ic.max_depth = a->int_args[0];
    ic.min_global_mq = 0;
    if (a->min_mapq[0] > 0) {  // src/plp.c:1124-1133
        int mq = a->min_mapq[0];
        for (int i = 0; i < nf; ++i) mq = std::min(mq, (int)a->min_mapq[i]);
        ic.min_global_mq = mq;
    }
    ic.rq_pct = a->dbl_args[3];
    ic.rq_minq = (int)a->dbl_args[4];
    ic.remove_overlaps = a->lgl_args[1] ? (a->overlap_mode == 2 ? 2 : 1) : 0;
    ic.n_mm_type = a->int_args[10];
    ic.n_mm = a->int_args[11];
    ic.min_overhang = a->int_args[7];
    ic.libtype.assign(a->libtype, a->libtype + nf);
    for (int lt : ic.libtype)
        if (lt < 0 || lt > 2) { set_error("invert read orientation failure"); return RAER_ERR; }
    if (a->umi_tag) { ic.bulk_umi = true; ic.umi_tag = a->umi_tag; }
    ic.regidx = has_ridx ? &ridx : nullptr;
    if (const char* e = getenv("RAER_BATCH_READS")) {  // reads per batch (tests exercise batch seams with small values)
        long v = atol(e);
        if (v > 0) ic.target_reads = v;
    }

    // ---- device ingest (raer_gingest.cu): whole contigs are inflated, parsed, filtered and packed by kernels; what that
    // path does not cover falls back, contig by contig, to the host stream
    // (a `region` is fine when it names a whole contig -- the per-chromosome calls of R/pileup.R:208-243 -- which is known
    // once the headers are read; a coordinate range goes through the host stream and the index)
    bool use_gi = a->shard_count <= 1 && !a->umi_tag && !getenv("RAER_SERIAL_INGEST") && !getenv("RAER_BATCH_READS");
    if (const char* e = getenv("RAER_GPU_INGEST")) use_gi = use_gi && atoi(e) != 0;
    if (use_gi) use_gi = raer::indexes_usable(a->bampaths, a->indexes, nf);
    // The device ingest reads the byte ranges the index names and nothing else, so a file that lost its tail would go
    // unnoticed where the single stream runs into the cut. A complete BGZF file ends with the 28-byte empty block
    // (htslib's bgzf_check_EOF): without it the call takes the host path, which reads to the end and reports what it finds.
    for (int i = 0; use_gi && i < nf; ++i) {
        static const unsigned char eof_blk[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43,
                                                  0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        unsigned char tail[28];
        FILE* fp = fopen(a->bampaths[i], "rb");
        const bool ok = fp && fseek(fp, -28, SEEK_END) == 0 && fread(tail, 1, 28, fp) == 28 && memcmp(tail, eof_blk, 28) == 0;
        if (fp) fclose(fp);
        if (!ok) use_gi = false;
    }

    double tm_open = raer::now_seconds();
    int gi_only_contig = -1, gi_reg_beg = 0, gi_reg_end = INT_MAX;
    if (use_gi && a->region) {
        raer::Ingest probe(ic);
        if (!probe.open_headers(a->bampaths)) return RAER_ERR_IO;
        const std::vector<std::string>& hn = probe.names();
        for (size_t k = 0; k < hn.size() && gi_only_contig < 0; ++k)
            if (hn[k] == a->region) gi_only_contig = (int)k;
        if (gi_only_contig < 0) {
            // "contig:beg-end" (hts_parse_reg: the name is what stands before the last colon). A wide range -- a tile of a
            // contig cut for several GPUs, R/pileup.R:208-243 with ranges -- is worth the device path; a narrow one reads
            // a handful of blocks through the index on the host.
            int rb = 0, re = INT_MAX;
            const int nl = raer::parse_region(a->region, &rb, &re);
            long min_span = 1 << 18;
            if (const char* e = getenv("RAER_GI_REGION_MIN")) min_span = atol(e);
            if (nl > 0 && (long)re - rb >= min_span) {
                const std::string nm(a->region, (size_t)nl);
                for (size_t k = 0; k < hn.size() && gi_only_contig < 0; ++k)
                    if (hn[k] == nm) gi_only_contig = (int)k;
                if (gi_only_contig >= 0) { gi_reg_beg = rb; gi_reg_end = re; }
            }
        }
        if (gi_only_contig < 0) use_gi = false;  // the host path parses (or refuses) it
    }
    raer::Ingest ing(ic);
    // the device ingest reads the files itself: only the headers are needed here
    if (!(use_gi ? ing.open_headers(a->bampaths) : ing.open(a->bampaths, a->indexes, a->host_threads, a->region))) return RAER_ERR_IO;
    tm_open = raer::now_seconds() - tm_open;

    double tm_ctx = raer::now_seconds();
    raer_ctx* ctx = raer_ctx_create(a->device);
    if (!ctx) return RAER_ERR_NO_DEVICE;
    tm_ctx = raer::now_seconds() - tm_ctx;
    double tm_fasta = 0, tm_batch = 0, tm_gpu = 0, tm_sink = 0;

    raer_bulk_conf bc;
    memset(&bc, 0, sizeof(bc));
    bc.n_files = nf;
    bc.min_depth = a->int_args[1]; bc.min_bq = a->int_args[2]; bc.trim_5p = a->int_args[3]; bc.trim_3p = a->int_args[4];
    bc.indel_dist = a->int_args[5]; bc.splice_dist = a->int_args[6]; bc.min_overhang = a->int_args[7];
    bc.nmer = a->int_args[8]; bc.min_var_reads = a->int_args[9]; bc.n_mm_type = a->int_args[10]; bc.n_mm = a->int_args[11];
    bc.ftrim_5p = a->dbl_args[0]; bc.ftrim_3p = a->dbl_args[1]; bc.min_af = a->dbl_args[2];
    bc.report_multiallelic = a->lgl_args[0];
    bc.remove_overlaps = ic.remove_overlaps;
    bc.want_stats = 1;
    bc.umi_mode = a->umi_tag ? 1 : 0;
    bc.aei_mode = a->aei_mode ? 1 : 0;
    if (bc.aei_mode) out->aei = (int64_t*)calloc((size_t)nf * 16, sizeof(int64_t));
    for (int i = 0; i < nf; ++i) { bc.min_mapq[i] = a->min_mapq[i]; bc.only_keep_variants[i] = a->only_keep_variants[i]; }

    const std::vector<std::string>& names = ing.names();
    out->n_targets = (int)names.size();
    out->target_names = (char**)calloc(names.size() ? names.size() : 1, sizeof(char*));
    for (size_t i = 0; i < names.size(); ++i) out->target_names[i] = strdup(names[i].c_str());

    ResultSink sink;
    sink.r = out;
    std::vector<int32_t> kh(nf * 2, 4);  // khash bucket count of every (file, strand) variant table
    std::vector<uint32_t> mask;
    int rc = RAER_OK;

    // One batch on the device: region / site bitmap, kernels, rows back in host memory.
    auto gpu_batch = [&](const raer::RegionIndex::Chr* rchr, int reg_beg, int reg_end, raer::PackedBatch& pb, raer_rows* rows) -> int {
        bc.reg_beg = reg_beg;
        bc.reg_end = reg_end;
        bc.site_mask = nullptr;
        if (has_ridx) {  // bitmap of eligible positions inside the batch window
            int wb = pb.b.beg, we = pb.b.end;
            size_t words = ((size_t)(we - wb) + 31) / 32;
            mask.assign(words, 0);
            size_t k = std::lower_bound(rchr->maxend.begin(), rchr->maxend.end(), wb) - rchr->maxend.begin();
            for (; k < rchr->beg.size() && rchr->beg[k] < we; ++k) {
                int lo = std::max(rchr->beg[k], wb), hi = std::min(rchr->end[k] + 1, we);
                for (int p = lo; p < hi; ++p) mask[(p - wb) >> 5] |= 1u << ((p - wb) & 31);
            }
            bc.site_mask = mask.data();
            bc.mask_beg = wb;
            bc.mask_bits = we - wb;
        }
        double t0g = raer::now_seconds();
        int r = raer_batch_pileup_host(ctx, &bc, &pb.b, rows);
        tm_gpu += raer::now_seconds() - t0g;
        if (r) return r;
        if (bc.aei_mode) {
            std::vector<int64_t> sums((size_t)nf * 16);
            r = raer_ctx_last_aei(ctx, sums.data(), nf);
            if (r) return r;
            for (size_t i = 0; i < sums.size(); ++i) out->aei[i] += sums[i];
        }
        out->gpu_launches += ctx->last_launches;
        out->t_h2d += ctx->ms_h2d * 1e-3; out->t_kernel += ctx->ms_kernel * 1e-3; out->t_d2h += ctx->ms_d2h * 1e-3;
        return RAER_OK;
    };
    // Rows of one batch into the result columns. Batches must arrive in (contig, position) order: the ALT strings
    // depend on the khash state the earlier rows of the call left behind.
    auto emit_rows = [&](int tid, raer_rows& rows) {
        double t0s = raer::now_seconds();
        sink_reserve(sink, out->nrows + rows.n_rows);
        for (int64_t i = 0; i < rows.n_rows; ++i) {
            int64_t o = out->nrows + i;
            int s = rows.meta[i] & 1;
            // tables grown by the puts of a site that wrote no row for that strand (src/plp.c:487)
            for (int f = 0; f < nf; ++f)
                if (kh[f * 2 + s] == 4 && rows.grow_pos[f * 2 + s] < rows.pos[i]) kh[f * 2 + s] = 8;
            out->tid[o] = tid;
            out->pos[o] = rows.pos[i] + 1;
            out->strand[o] = s ? '-' : '+';
            out->ref[o] = (char)((rows.meta[i] >> 8) & 0xff);
            out->rpbz[o] = rows.stats[i * 3];
            out->vdb[o] = rows.stats[i * 3 + 1];
            out->sor[o] = rows.stats[i * 3 + 2];
            for (int f = 0; f < nf; ++f) {
                memcpy(out->counts + ((size_t)f * sink.cap + o) * 8, rows.counts + ((size_t)i * nf + f) * 8, 32);
                raer_altcode_to_string(rows.altcode[i * nf + f], out->ref[o], &kh[f * 2 + s],
                                       out->alt + ((size_t)f * sink.cap + o) * 12);
            }
        }
        for (int k = 0; k < nf * 2; ++k)
            if (kh[k] == 4 && rows.grow_pos[k] != INT32_MAX) kh[k] = 8;
        out->nrows += rows.n_rows;
        raer_rows_free(&rows);
        tm_sink += raer::now_seconds() - t0s;
    };

    // ---- contig-parallel ingest: the record parser of a file is serial (stream-order decisions of bam_plp_push), but
    // contigs are independent, so several workers each stream one contig through the BAI index and hand their batches
    // to this thread, which owns the device. Rows are kept per contig and turned into columns in contig order.
    int hw = a->host_threads > 0 ? a->host_threads : raer::usable_cpus();
    if (hw <= 0) hw = 4;
    int n_workers = std::min<int>((int)names.size(), std::min(16, hw / (nf + 1)));
    // a site / region list goes contig by contig through the index even with one worker: the streams skip what lies
    // between the intervals (Ingest::set_jumps; src/plp.c:856-882 queries the index per region)
    if (has_ridx && n_workers < 1) n_workers = 1;
    bool parallel = !a->region && a->shard_count <= 1 && (n_workers >= 2 || (has_ridx && n_workers >= 1)) &&
                    !getenv("RAER_SERIAL_INGEST");
    if (parallel) parallel = raer::indexes_usable(a->bampaths, a->indexes, nf);  // single stream: no index needed
    if (use_gi) {
        double t0c = raer::now_seconds();
        raer_gi* gi = raer_gi_create(a->device, nf, a->bampaths, a->indexes);
        if (!gi) rc = RAER_ERR_IO;
        const double tm_gi_create = raer::now_seconds() - t0c;
        const std::vector<int32_t>& lens = ing.lengths();
        std::vector<int> ftid((size_t)nf);
        std::vector<int64_t> file_off((size_t)nf + 1);
        int64_t n_fallback = 0, fb_records = 0, fb_bases = 0, fb_inflated = 0, fb_seeks = 0;
        // the host stream for one contig (same as a worker of the contig-parallel path, on this thread)
        auto host_contig = [&](int c) -> int {
            raer::Ingest wi(ic);
            std::vector<std::pair<int32_t, int32_t>> jl;
            if (has_ridx) jl = raer::jump_intervals(ridx.find(names[c]));
            const bool ranged = gi_reg_beg > 0 || gi_reg_end < INT_MAX;  // a coordinate range: the call's own region string
            if (!wi.open(a->bampaths, a->indexes, a->host_threads, ranged ? a->region : names[c].c_str(), has_ridx ? &jl : nullptr))
                return RAER_ERR_IO;
            raer::PackedBatch pb;
            int r = RAER_OK;
            while (r == RAER_OK && wi.next_contig() >= 0) {
                std::string seq;
                if (!raer::fasta_fetch(a->fapath, names[c], seq)) seq.clear();
                if ((r = raer_ctx_set_reference(ctx, c, seq.data(), (int64_t)seq.size()))) break;
                while (r == RAER_OK && wi.next_batch(pb)) {
                    if (a->poll && a->poll(a->poll_ctx)) { set_error("interrupted"); r = RAER_ERR_INTERRUPT; break; }
                    if (pb.b.n_reads == 0 || pb.b.end <= pb.b.beg) continue;
                    raer_rows rows;
                    r = gpu_batch(has_ridx ? ridx.find(names[c]) : nullptr, wi.reg_tid() >= 0 ? wi.reg_beg() : 0,
                                  wi.reg_tid() >= 0 ? wi.reg_end() : INT_MAX, pb, &rows);
                    if (r == RAER_OK) emit_rows(c, rows);
                }
            }
            if (r == RAER_OK && wi.failed()) { set_error("%s", wi.error().c_str()); r = RAER_ERR_IO; }
            fb_records += wi.records(); fb_bases += wi.aligned_bases(); fb_inflated += wi.inflated_bytes(); fb_seeks += wi.seeks();
            return r;
        };
        for (int c = 0; rc == RAER_OK && c < (int)names.size(); ++c) {
            const raer::RegionIndex::Chr* rchr = has_ridx ? ridx.find(names[c]) : nullptr;
            if (has_ridx && !rchr) continue;  // no site on this contig: nothing can be written
            if (gi_only_contig >= 0 && c != gi_only_contig) continue;
            if (a->poll && a->poll(a->poll_ctx)) { set_error("interrupted"); rc = RAER_ERR_INTERRUPT; break; }
            if (has_ridx) {
                // a sparse site list reads a few blocks per site through the index (Ingest::set_jumps); inflating the
                // whole contig pays only when the intervals to visit cover most of it anyway
                int64_t covered = 0;
                for (const auto& iv : raer::jump_intervals(rchr)) covered += (int64_t)iv.second - iv.first;
                if (covered * 2 < (int64_t)lens[(size_t)c]) {
                    ++out->n_host_contigs;
                    rc = host_contig(c);
                    continue;
                }
            }
            for (int f = 0; f < nf; ++f) {  // each file resolves the name in its own header, like sam_itr_querys
                ftid[(size_t)f] = c;
                const std::vector<std::string>& fn = ing.file_names(f);
                if (f > 0 && !((size_t)c < fn.size() && fn[(size_t)c] == names[c])) {
                    ftid[(size_t)f] = -1;
                    for (size_t k = 0; k < fn.size(); ++k)
                        if (fn[k] == names[c]) { ftid[(size_t)f] = (int)k; break; }
                    if (ftid[(size_t)f] < 0) { set_error("fail to parse region '%s' with %s", names[c].c_str(), a->bampaths[f]); rc = RAER_ERR_IO; }
                }
            }
            if (rc) break;
            const uint32_t* d_sites = nullptr;
            int64_t bases = 0;
            int n_windows = 0;
            double t0b = raer::now_seconds();
            int r = raer_gi_contig(gi, c, ftid.data(), lens[(size_t)c], &ic, rchr, gi_reg_beg, gi_reg_end, &n_windows, &d_sites, &bases);
            tm_batch += raer::now_seconds() - t0b;
            if (r == 1) {  // RAER_GI_FALLBACK
                ++n_fallback;
                ++out->n_host_contigs;
                rc = host_contig(c);
                continue;
            }
            if (r) { rc = r; break; }
            out->n_aligned_bases += bases;
            ++out->n_device_contigs;
            if (n_windows == 0) continue;
            std::string seq;
            double t0f = raer::now_seconds();
            if (!raer::fasta_fetch(a->fapath, names[c], seq)) seq.clear();
            rc = raer_ctx_set_reference(ctx, c, seq.data(), (int64_t)seq.size());
            tm_fasta += raer::now_seconds() - t0f;
            if (rc) break;
            bc.reg_beg = gi_reg_beg;
            bc.reg_end = gi_reg_end;
            bc.site_mask = d_sites;  // device pointer: the bitmap the prefilter used, over the whole contig
            bc.mask_beg = 0;
            bc.mask_bits = lens[(size_t)c];
            // the contig in position windows (one unless its reads exceed what a batch's 32-bit offsets describe)
            for (int j = 0; j < n_windows && rc == RAER_OK; ++j) {
                if (j > 0 && a->poll && a->poll(a->poll_ctx)) { set_error("interrupted"); rc = RAER_ERR_INTERRUPT; break; }
                raer_read_batch b;
                t0b = raer::now_seconds();
                rc = raer_gi_window(gi, j, &b, file_off.data());
                tm_batch += raer::now_seconds() - t0b;
                if (rc) break;
                if (b.n_reads == 0) continue;
                raer_rows rows;
                int64_t launches = 0;
                double t0g = raer::now_seconds();
                rc = raer_batch_pileup_device_rows(ctx, &bc, &b, &rows, &launches);
                tm_gpu += raer::now_seconds() - t0g;
                if (rc) break;
                if (bc.aei_mode) {
                    std::vector<int64_t> sums((size_t)nf * 16);
                    if ((rc = raer_ctx_last_aei(ctx, sums.data(), nf))) { raer_rows_free(&rows); break; }
                    for (size_t i = 0; i < sums.size(); ++i) out->aei[i] += sums[i];
                }
                out->gpu_launches += launches;
                out->t_kernel += ctx->ms_kernel * 1e-3;
                out->t_d2h += ctx->ms_d2h * 1e-3;
                emit_rows(c, rows);
            }
        }
        if (gi) {
            int64_t gl = 0, gin = 0, gco = 0, grec = 0;
            raer_gi_stats(gi, &gl, &gin, &gco, &grec);
            out->gpu_launches += gl;
            out->n_records = grec + fb_records;
            out->n_aligned_bases += fb_bases;
            out->n_inflated_bytes = gin + fb_inflated;
            out->n_seeks = fb_seeks;
            t0c = raer::now_seconds();
            raer_gi_destroy(gi);
            t0c = raer::now_seconds() - t0c;
        }
        raer_ctx_set_reference(ctx, -1, nullptr, 0);
        if (getenv("RAER_TIMING"))
            fprintf(stderr, "[raer_pileup] device ingest: %lld contig(s) fell back to the host stream; create %.3f s, destroy %.3f s\n",
                    (long long)n_fallback, tm_gi_create, t0c);
    } else if (parallel) {
        struct Slot {
            raer::PackedBatch pb;
            int tid = -1;
            std::shared_ptr<std::string> ref;
            bool busy = false;
        };
        std::vector<Slot> slots((size_t)n_workers * 2);
        // plain host memory: page-locking from several threads serialises on the driver and the kernel's mm locks
        // and stalls every other thread of the process; the device thread copies from pageable memory instead
        for (Slot& sl : slots) sl.pb.want_pinned = false;
        std::mutex mu;
        std::condition_variable cv_ready, cv_free;
        std::deque<Slot*> ready;
        int workers_left = n_workers;
        bool stop = false;
        int worker_rc = RAER_OK;
        std::string worker_msg;
        std::atomic<int> next_idx{0};
        // longest contigs first: the wall time is bounded below by the longest contig's serial parse
        std::vector<int> order(names.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
        const std::vector<int32_t>& lens = ing.lengths();
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return lens[x] > lens[y]; });
        raer::IngestConf icw = ic;
        if (!getenv("RAER_BATCH_READS")) {
            // two batches per worker are in flight, each in its own host buffers: size them by the input (about 50
            // compressed bytes per record) so that a small call does not hold more memory than it reads
            int64_t bytes = 0;
            for (int i = 0; i < nf; ++i) {
                FILE* fp = fopen(a->bampaths[i], "rb");
                if (fp) { fseek(fp, 0, SEEK_END); bytes += ftell(fp); fclose(fp); }
            }
            const int64_t est = bytes / 50 / (4 * n_workers);
            icw.target_reads = std::min<int64_t>(1 << 19, std::max<int64_t>(1 << 16, est));
        }
        // a stream's parser is one thread, its inflate costs about three times the parser's CPU time: give the inflate
        // of every stream two threads (the host is slightly oversubscribed, which keeps every core busy)
        const int inflate_threads = std::min(4, std::max(2, (hw + n_workers * nf - 1) / (n_workers * nf)));
        double w_decode = 0, w_pack = 0;
        int64_t w_records = 0, w_bases = 0, w_inflated = 0, w_seeks = 0;
        auto worker = [&](int w) {
            int my_rc = RAER_OK;
            while (my_rc == RAER_OK) {
                const int oi = next_idx.fetch_add(1);
                if (oi >= (int)order.size()) break;
                const int c = order[oi];
                if (has_ridx && !ridx.find(names[c])) continue;  // no site on this contig: nothing can be written
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (stop) break;
                }
                raer::Ingest wi(icw);
                std::vector<std::pair<int32_t, int32_t>> jl;
                if (has_ridx) jl = raer::jump_intervals(ridx.find(names[c]));
                if (!wi.open(a->bampaths, a->indexes, inflate_threads, names[c].c_str(), has_ridx ? &jl : nullptr)) { my_rc = RAER_ERR_IO; break; }
                std::shared_ptr<std::string> ref;
                while (my_rc == RAER_OK && wi.next_contig() >= 0) {
                    if (!ref) {
                        ref = std::make_shared<std::string>();
                        if (!raer::fasta_fetch(a->fapath, names[c], *ref)) ref->clear();  // every site reads as 'N' and is skipped
                    }
                    while (true) {
                        Slot* sl = nullptr;
                        {
                            std::unique_lock<std::mutex> lk(mu);
                            cv_free.wait(lk, [&]() { return stop || !slots[2 * w].busy || !slots[2 * w + 1].busy; });
                            if (stop) break;
                            sl = !slots[2 * w].busy ? &slots[2 * w] : &slots[2 * w + 1];
                        }
                        if (!wi.next_batch(sl->pb)) break;
                        if (sl->pb.b.n_reads == 0 || sl->pb.b.end <= sl->pb.b.beg) continue;
                        sl->tid = c;
                        sl->ref = ref;
                        std::lock_guard<std::mutex> lk(mu);
                        sl->busy = true;
                        ready.push_back(sl);
                        cv_ready.notify_one();
                    }
                    {
                        std::lock_guard<std::mutex> lk(mu);
                        if (stop) break;
                    }
                }
                if (my_rc == RAER_OK && wi.failed()) {  // a stream stopped on an error: not the end of the contig
                    my_rc = RAER_ERR_IO;
                    set_error("%s", wi.error().c_str());
                }
                std::lock_guard<std::mutex> lk(mu);
                w_decode += wi.decode_seconds(); w_pack += wi.pack_seconds();
                w_records += wi.records(); w_bases += wi.aligned_bases();
                w_inflated += wi.inflated_bytes(); w_seeks += wi.seeks();
            }
            std::lock_guard<std::mutex> lk(mu);
            if (my_rc != RAER_OK && worker_rc == RAER_OK) {
                worker_rc = my_rc;
                worker_msg = raer::last_error();  // the message buffer is per thread: hand it to the calling thread
                stop = true;
                cv_free.notify_all();
            }
            --workers_left;
            cv_ready.notify_one();
        };
        std::vector<std::thread> th;
        for (int w = 0; w < n_workers; ++w) th.emplace_back(worker, w);
        std::vector<std::vector<raer_rows>> stash(names.size());
        // batches of different contigs interleave: every contig's reference is uploaded once and stays on the device
        // for the rest of the call (a genome is small next to 180 GB of HBM)
        std::unordered_map<int, void*> dref;
        int ref_tid = -1;
        while (true) {
            Slot* sl = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                double t0b = raer::now_seconds();
                // wake up twice a second to give the interrupt hook a chance while the workers are still decoding
                while (!cv_ready.wait_for(lk, std::chrono::milliseconds(500), [&]() { return !ready.empty() || workers_left == 0; })) {
                    if (rc == RAER_OK && a->poll) {
                        lk.unlock();
                        const bool stop_now = a->poll(a->poll_ctx) != 0;
                        lk.lock();
                        if (stop_now) { set_error("interrupted"); rc = RAER_ERR_INTERRUPT; stop = true; cv_free.notify_all(); }
                    }
                }
                tm_batch += raer::now_seconds() - t0b;
                if (ready.empty()) break;
                sl = ready.front();
                ready.pop_front();
            }
            if (rc == RAER_OK && a->poll && a->poll(a->poll_ctx)) {  // the calling thread polls between batches (src/plp.c:925-931)
                set_error("interrupted");
                rc = RAER_ERR_INTERRUPT;
            }
            if (rc == RAER_OK) {
                if (sl->tid != ref_tid) {
                    double t0f = raer::now_seconds();
                    const int64_t rlen = (int64_t)sl->ref->size();
                    if (rlen == 0) {
                        rc = raer_ctx_set_reference(ctx, sl->tid, nullptr, 0);
                    } else {
                        auto it = dref.find(sl->tid);
                        if (it == dref.end()) {
                            void* dp = nullptr;
                            if (cudaMalloc(&dp, (size_t)rlen) != cudaSuccess ||
                                cudaMemcpy(dp, sl->ref->data(), (size_t)rlen, cudaMemcpyHostToDevice) != cudaSuccess) {
                                if (dp) cudaFree(dp);
                                set_error("reference upload failed");
                                rc = RAER_ERR_CUDA;
                            } else {
                                it = dref.emplace(sl->tid, dp).first;
                            }
                        }
                        if (rc == RAER_OK) rc = raer_ctx_set_reference_device(ctx, sl->tid, it->second, rlen);
                    }
                    tm_fasta += raer::now_seconds() - t0f;
                    ref_tid = sl->tid;
                }
                if (rc == RAER_OK) {
                    raer_rows rows;
                    rc = gpu_batch(has_ridx ? ridx.find(names[sl->tid]) : nullptr, 0, INT_MAX, sl->pb, &rows);
                    if (rc == RAER_OK) stash[sl->tid].push_back(rows);
                }
            }
            std::lock_guard<std::mutex> lk(mu);
            sl->busy = false;
            sl->ref.reset();
            if (rc != RAER_OK) stop = true;
            cv_free.notify_all();
        }
        for (auto& t : th) t.join();
        cudaStreamSynchronize((cudaStream_t)raer_ctx_stream(ctx));
        for (auto& kv : dref) cudaFree(kv.second);
        raer_ctx_set_reference(ctx, -1, nullptr, 0);
        if (rc == RAER_OK && worker_rc != RAER_OK) {
            rc = worker_rc;
            set_error("%s", worker_msg.c_str());
        }
        for (size_t c = 0; c < stash.size(); ++c)
            for (raer_rows& rows : stash[c]) {
                if (rc == RAER_OK) emit_rows((int)c, rows);
                else raer_rows_free(&rows);
            }
        out->t_decode = w_decode;
        out->t_pack = w_pack;
        out->n_records = w_records;
        out->n_aligned_bases = w_bases;
        out->n_inflated_bytes = w_inflated;
        out->n_seeks = w_seeks;
    } else {
        raer::PackedBatch pb;
        int tid;
        int contig_no = 0;
        while (rc == RAER_OK && (tid = ing.next_contig()) >= 0) {
            bool mine = a->shard_count <= 1 || (contig_no % a->shard_count) == a->shard_index;
            ++contig_no;
            std::string seq;
            double t0f = raer::now_seconds();
            bool have_ref = raer::fasta_fetch(a->fapath, names[tid], seq);
            if (!have_ref) seq.clear();  // reference: every site reads as 'N' and is skipped
            if (mine) {
                rc = raer_ctx_set_reference(ctx, tid, seq.data(), (int64_t)seq.size());
                if (rc) break;
            }
            tm_fasta += raer::now_seconds() - t0f;
            const raer::RegionIndex::Chr* rchr = has_ridx ? ridx.find(names[tid]) : nullptr;
            while (rc == RAER_OK) {
                double t0b = raer::now_seconds();
                const bool got = ing.next_batch(pb);
                tm_batch += raer::now_seconds() - t0b;
                if (!got) break;
                if (a->poll && a->poll(a->poll_ctx)) {  // the calling thread polls between batches (src/plp.c:925-931)
                    set_error("interrupted");
                    rc = RAER_ERR_INTERRUPT;
                    break;
                }
                if (!mine || pb.b.n_reads == 0 || pb.b.end <= pb.b.beg) continue;
                if (has_ridx && !rchr) continue;
                raer_rows rows;
                rc = gpu_batch(rchr, ing.reg_tid() >= 0 ? ing.reg_beg() : 0, ing.reg_tid() >= 0 ? ing.reg_end() : INT_MAX, pb, &rows);
                if (rc) break;
                emit_rows(tid, rows);
            }
        }
        // a stream that stopped on a read / inflate / checksum / record / sort-order error looks like the end of the input
        // to the loops above: the rows so far are not the result of the call (src/plp.c:1195 raises an error)
        if (rc == RAER_OK && ing.failed()) {
            set_error("%s", ing.error().c_str());
            rc = RAER_ERR_IO;
        }
        out->t_decode = ing.decode_seconds();
        out->t_pack = ing.pack_seconds();
        out->n_records = ing.records();
        out->n_aligned_bases = ing.aligned_bases();
        out->n_inflated_bytes = ing.inflated_bytes();
        out->n_seeks = ing.seeks();
    }
    // compact the [file][row] stride to nrows
    if (sink.cap != out->nrows && out->nrows > 0) {
        for (int f = 1; f < nf; ++f) {
            memmove(out->alt + (size_t)f * out->nrows * 12, out->alt + (size_t)f * sink.cap * 12, (size_t)out->nrows * 12);
            memmove(out->counts + (size_t)f * out->nrows * 8, out->counts + (size_t)f * sink.cap * 8, (size_t)out->nrows * 32);
        }
    }
    double tm_destroy = raer::now_seconds();
    raer_ctx_destroy(ctx);
    tm_destroy = raer::now_seconds() - tm_destroy;
    out->t_total = raer::now_seconds() - t_start;
    if (getenv("RAER_TIMING")) fprintf(stderr, "[raer_pileup] ctx destroy %.3f s\n", tm_destroy);
    if (getenv("RAER_TIMING"))
        fprintf(stderr, "[raer_pileup] total %.3f s: open %.3f ctx %.3f fasta+ref %.3f next_batch %.3f (decode %.3f pack %.3f) "
                        "gpu call %.3f (h2d %.3f kernel %.3f d2h %.3f) rows->columns %.3f\n",
                out->t_total, tm_open, tm_ctx, tm_fasta, tm_batch, out->t_decode, out->t_pack, tm_gpu, out->t_h2d, out->t_kernel,
                out->t_d2h, tm_sink);
    if (rc != RAER_OK) {
        raer_pileup_result_free(out);
        return rc;
    }
    return RAER_OK;
}

// Host-only probe of the ingest (no device needed): streams the BAMs exactly like raer_pileup() and
// reports what would be handed to the kernels. out[0..7] = kept reads (summed over batches, halo
// reads counted once per batch), mate pairs, CIGAR words, packed sequence bytes, batches, aligned
// bases of all mapped records, records read, distinct kept reads.
extern "C" int raer_ingest_summary(const raer_pileup_args* a, int64_t target_reads, int64_t* out) {
    memset(out, 0, 16 * sizeof(int64_t));
    if (!a || a->nbam < 1 || a->nbam > RAER_MAX_FILES) return RAER_ERR_ARG;
    const int nf = a->nbam;
    raer::RegionIndex ridx;
    bool has_ridx = a->n_regions > 0;
    if (has_ridx) {
        std::vector<int32_t> b0(a->n_regions), e0(a->n_regions);
        for (int i = 0; i < a->n_regions; ++i) { b0[i] = a->reg_start[i] - 1; e0[i] = a->reg_end[i] - 1; }
        ridx.build(a->n_regions, a->reg_seqnames, b0.data(), e0.data());
    }
    raer::IngestConf ic;
    ic.n_files = nf;
    ic.keep_flag[0] = (uint32_t)a->int_args[12];
    ic.keep_flag[1] = (uint32_t)a->int_args[13];
    ic.max_depth = a->int_args[0];
    if (a->min_mapq[0] > 0) {
        int mq = a->min_mapq[0];
        for (int i = 0; i < nf; ++i) mq = std::min(mq, (int)a->min_mapq[i]);
        ic.min_global_mq = mq;
    }
    ic.rq_pct = a->dbl_args[3];
    ic.rq_minq = (int)a->dbl_args[4];
    ic.remove_overlaps = a->lgl_args[1] ? (a->overlap_mode == 2 ? 2 : 1) : 0;
    ic.n_mm_type = a->int_args[10];
    ic.n_mm = a->int_args[11];
    ic.min_overhang = a->int_args[7];
    ic.libtype.assign(a->libtype, a->libtype + nf);
    if (a->umi_tag) { ic.bulk_umi = true; ic.umi_tag = a->umi_tag; }
    ic.regidx = has_ridx ? &ridx : nullptr;
    if (target_reads > 0) ic.target_reads = target_reads;
    // one pass over `ing`: what raer_pileup() does with the batches, minus the device
    auto drain = [&](raer::Ingest& ing) -> int {
        raer::PackedBatch pb;
        int prev_boundary = 0;
        while (ing.next_contig() >= 0) {
            prev_boundary = INT_MIN;
            while (ing.next_batch(pb)) {
                if (a->poll && a->poll(a->poll_ctx)) { set_error("interrupted"); return RAER_ERR_INTERRUPT; }
                out[0] += pb.b.n_reads;
                out[1] += pb.b.n_pairs;
                out[2] += pb.b.n_cigar_words;
                out[3] += pb.b.n_sq_bytes;
                out[4] += 1;
                for (int64_t i = 0; i < pb.b.n_reads; ++i)
                    if (pb.hdr[i].pos >= prev_boundary) out[7] += 1;  // first appearance of the read
                if (pb.b.n_reads) {
                    int mx = INT_MIN;
                    for (int64_t i = 0; i < pb.b.n_reads; ++i) mx = std::max(mx, (int)pb.hdr[i].pos);
                    prev_boundary = mx + 1 > prev_boundary ? mx + 1 : prev_boundary;
                }
            }
        }
        out[5] += ing.aligned_bases();
        out[6] += ing.records();
        out[8] += ing.inflated_bytes();
        out[9] += ing.seeks();
        if (ing.failed()) { set_error("%s", ing.error().c_str()); return RAER_ERR_IO; }
        return RAER_OK;
    };
    if (has_ridx && !a->region && raer::indexes_usable(a->bampaths, a->indexes, nf) && !getenv("RAER_SERIAL_INGEST")) {
        // a site / region list: contig by contig through the index, skipping what lies between the intervals
        raer::BamStream hs;
        if (!hs.open(a->bampaths[0], 1, true)) return RAER_ERR_IO;
        const std::vector<std::string> names = hs.names();
        for (const std::string& nm : names) {
            const raer::RegionIndex::Chr* rc = ridx.find(nm);
            if (!rc) continue;
            raer::Ingest ing(ic);
            const std::vector<std::pair<int32_t, int32_t>> jl = raer::jump_intervals(rc);
            if (!ing.open(a->bampaths, a->indexes, a->host_threads, nm.c_str(), &jl)) return RAER_ERR_IO;
            int rc2 = drain(ing);
            if (rc2) return rc2;
        }
        return RAER_OK;
    }
    raer::Ingest ing(ic);
    if (!ing.open(a->bampaths, a->indexes, a->host_threads, a->region)) return RAER_ERR_IO;
    {
        int rc2 = drain(ing);
        if (rc2) return rc2;
    }
    if (getenv("RAER_TIMING"))
        fprintf(stderr, "[raer ingest] parse sections %.3f s (of which the parsers waited %.3f s, summed over files, for inflate), "
                        "pack sections %.3f s, inflate tasks %.3f s summed over files\n",
                ing.parse_seconds(), ing.wait_seconds(), ing.pack_seconds(), ing.decode_seconds());
    return RAER_OK;
}

// ------------------------------------------------------------------------------------- misc .Call entries
extern "C" int raer_get_region(const char* region, char* chrom, int32_t chrom_cap, int32_t* beg, int32_t* end) {
    if (!region || !chrom || chrom_cap < 1) return RAER_ERR_ARG;
    int b, e;
    int nl = raer::parse_region(region, &b, &e);
    if (nl < 0) { set_error("could not parse region:%s", region); return RAER_ERR_ARG; }
    if (nl >= chrom_cap) nl = chrom_cap - 1;
    memcpy(chrom, region, nl);
    chrom[nl] = 0;
    *beg = b;
    *end = e;
    return RAER_OK;
}

// Fisher exact test, two-sided p-value as htslib kt_fisher_exact reports it (sum of the
// probabilities of all tables no more likely than the observed one; hypergeometric via lgamma).
static double lbinom(int n, int k) { return lgamma(n + 1.0) - lgamma(k + 1.0) - lgamma(n - k + 1.0); }
static double hyper(int n11, int n1_, int n_1, int n) { return exp(lbinom(n1_, n11) + lbinom(n - n1_, n_1 - n11) - lbinom(n, n_1)); }
extern "C" int raer_fisher_exact(const int32_t* mat, int32_t ncol, double* pvalue) {
    for (int j = 0; j < ncol; ++j) {
        int n11 = mat[4 * j], n12 = mat[4 * j + 1], n21 = mat[4 * j + 2], n22 = mat[4 * j + 3];
        int n1_ = n11 + n12, n_1 = n11 + n21, n = n11 + n12 + n21 + n22;
        int lo = std::max(0, n1_ + n_1 - n), hi = std::min(n1_, n_1);
        if (lo == hi) { pvalue[j] = 1.0; continue; }
        double q = hyper(n11, n1_, n_1, n), two = 0;
        for (int i = lo; i <= hi; ++i) {
            double p = hyper(i, n1_, n_1, n);
            if (p < q * (1 + 1e-7)) two += p;  // htslib uses a relative slack when comparing table probabilities
        }
        pvalue[j] = two > 1 ? 1.0 : two;
    }
    return RAER_OK;
}
