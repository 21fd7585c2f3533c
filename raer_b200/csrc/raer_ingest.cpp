// raer_ingest.cpp -- BGZF/BAM streaming (multi-threaded inflate), read prefilter, batch packing.
// See raer_host.h. Reference semantics: readaln src/plp.c:32-78, screadaln src/sc-plp.c:449-511,
// invert_read_orientation src/plp_utils.c:331-371, read_base_quality :310-328,
// parse_mismatches src/ext/htsbox/htsbox-ext.c:17-97; htslib bam_plp_push / overlap_push for the
// stream-order decisions.
#include <ctype.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <thread>

#include "raer_host.h"
#include "raer_inflate.h"

extern "C" void* raer_pinned_alloc(size_t n, int* pinned);
extern "C" void raer_pinned_free(void* p, int pinned);

namespace raer {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

int usable_cpus() {
#ifdef __linux__
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return n;
    }
#endif
    const int n = (int)std::thread::hardware_concurrency();
    return n > 0 ? n : 4;
}

double now_seconds() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t le32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

// ------------------------------------------------------------------------------------- BamStream
BamStream::~BamStream() {
    if (pf_.valid()) pf_.wait();  // the prefetch task reads fp_
    if (fp_) fclose(fp_);
}

struct BlockJob {
    size_t c_off;   // start of deflate payload in cbuf_
    uint32_t c_len; // payload bytes
    size_t u_off;   // destination in ubuf_
    uint32_t u_len;
    uint32_t crc;   // CRC32 of the inflated bytes (BGZF block trailer)
};

int BamStream::fail(const char* fmt, ...) {
    char buf[400];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> g(err_mu_);
    err_ = buf;
    return -1;
}

int BamStream::produce(std::vector<uint8_t>& out, std::vector<BlkPos>& pos) {
    double t0 = now_seconds();
    // small groups: the next one inflates while this one is parsed (light: just the header, a seek follows; jump list:
    // the next seek is never far)
    const bool small = light_ || small_groups_;
    const size_t kChunk = small_groups_ ? (256u << 10) : small ? (1u << 20) : (32u << 20);
    const size_t kMaxOut = small_groups_ ? (192u << 10) : small ? (256u << 10) : (12u << 20);
    out.clear();
    pos.clear();
    std::vector<BlockJob> jobs;
    size_t out_off = 0;
    while (true) {
        // refill the compressed buffer
        if (!file_eof_ && cbuf_.size() - c_off_ < (small_groups_ ? (128u << 10) : (1u << 20))) {
            if (c_off_ > 0) {
                cbuf_.erase(cbuf_.begin(), cbuf_.begin() + c_off_);
                cbuf_base_ += c_off_;
                c_off_ = 0;
            }
            size_t old = cbuf_.size();
            cbuf_.resize(old + kChunk);
            size_t got = fread(cbuf_.data() + old, 1, kChunk, fp_);
            cbuf_.resize(old + got);
            if (got < kChunk) file_eof_ = true;
        }
        size_t o = c_off_;
        bool at_stop = false;
        while (o + 18 <= cbuf_.size() && out_off < kMaxOut) {
            if (stop_coff_ && cbuf_base_ + o > stop_coff_) { at_stop = true; break; }  // behind the region's contig
            const uint8_t* h = cbuf_.data() + o;
            if (h[0] != 31 || h[1] != 139 || h[2] != 8 || !(h[3] & 4)) return fail("corrupt BGZF block header");
            int xlen = le16(h + 10);
            if (o + 12 + xlen > cbuf_.size()) break;
            int bsize = -1;
            for (int e = 0; e + 4 <= xlen;) {
                int slen = le16(h + 12 + e + 2);
                if (h[12 + e] == 'B' && h[13 + e] == 'C' && slen == 2) bsize = le16(h + 12 + e + 4);
                e += 4 + slen;
            }
            if (bsize < 0) return fail("BGZF block without BC field");
            // header (12) + extra field + at least two payload bytes + CRC32 and ISIZE
            if (bsize + 1 < 12 + xlen + 2 + 8) return fail("corrupt BGZF block (BSIZE smaller than its own header)");
            if (o + bsize + 1 > cbuf_.size()) break;
            uint32_t isize = le32(h + bsize + 1 - 4);
            if (isize > 65536u) return fail("corrupt BGZF block (ISIZE above 64 KiB)");
            BlockJob j;
            j.crc = le32(h + bsize + 1 - 8);
            j.c_off = o + 12 + xlen;
            j.c_len = (uint32_t)(bsize + 1 - 12 - xlen - 8);
            j.u_off = out_off;
            j.u_len = isize;
            if (isize) jobs.push_back(j);
            pos.push_back(BlkPos{(int64_t)out_off, cbuf_base_ + o});
            out_off += isize;
            o += bsize + 1;
        }
        const bool progressed = o != c_off_;
        c_off_ = o;
        if (!jobs.empty()) break;
        if (at_stop) { t_decode_ += now_seconds() - t0; return 0; }
        if (file_eof_) {
            if (c_off_ >= cbuf_.size()) { t_decode_ += now_seconds() - t0; return 0; }  // every byte of the file was a whole block
            // bytes that do not make a whole block: the file was cut (htslib: "truncated file")
            if (!progressed) return fail("truncated BGZF file (incomplete block at the end)");
            continue;
        }
        // a block larger than what is buffered cannot happen (blocks are <= 64 KiB); only empty blocks so far: read on
        if (!progressed && cbuf_.size() - c_off_ >= (128u << 10)) return fail("BGZF framing error");
    }
    out.resize(out_off);
    int nt = std::max(1, std::min<int>(threads_, (int)jobs.size() / 4 + 1));
    std::vector<int> status(nt, 0);
    // blocks go through the ingest's own DEFLATE decoder (raer_inflate.cpp, about 1.7x zlib's speed on BAM data); zlib
    // takes the last block of the buffer (the decoder wants 16 readable bytes behind a payload), anything the decoder
    // refuses, and everything when RAER_ZLIB_INFLATE is set
    static const bool own_inflate = getenv("RAER_ZLIB_INFLATE") == nullptr;
    static const bool check_crc = getenv("RAER_SKIP_CRC") == nullptr;  // htslib verifies every block's CRC32 too
    auto work = [&](int t) {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { status[t] = -1; return; }
        for (size_t i = t; i < jobs.size(); i += nt) {
            const BlockJob& j = jobs[i];
            if (!(own_inflate && j.c_off + j.c_len + 16 <= cbuf_.size() &&
                  inflate_raw(cbuf_.data() + j.c_off, j.c_len, out.data() + j.u_off, j.u_len) == 0)) {
                inflateReset(&zs);
                zs.next_in = cbuf_.data() + j.c_off;
                zs.avail_in = j.c_len;
                zs.next_out = out.data() + j.u_off;
                zs.avail_out = j.u_len;
                int zr = inflate(&zs, Z_FINISH);
                if (zr != Z_STREAM_END || zs.avail_out != 0) { status[t] = -1; break; }
            }
            if (check_crc && (uint32_t)crc32(0L, out.data() + j.u_off, j.u_len) != j.crc) { status[t] = -2; break; }
        }
        inflateEnd(&zs);
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    for (int s : status)
        if (s) return fail(s == -2 ? "BGZF block CRC32 mismatch" : "inflate failed (corrupt BGZF block)");
    n_inflated_ += (int64_t)out_off;
    t_decode_ += now_seconds() - t0;
    return 1;
}

bool BamStream::fill() {
    if (drained_) return false;
    int rc;
    std::vector<uint8_t>& grp = grp_;
    if (pf_.valid()) {
        const double tw = now_seconds();
        rc = pf_.get();
        t_wait_ += now_seconds() - tw;
        grp.swap(pf_buf_);
        grp_blk_.swap(pf_blk_);
    } else {
        rc = produce(grp, grp_blk_);
    }
    if (rc < 0) { failed_ = true; set_error("%s", error().c_str()); }
    if (rc <= 0) { drained_ = true; return false; }
    // keep the unread tail of the previous group (a record can straddle two groups) in front of the new bytes
    const size_t leftover = ubuf_.size() - u_off_;
    {
        // block positions relative to the new ubuf_: the blocks the unread tail still lies in keep their (now possibly
        // negative) start, the new group's blocks follow behind the tail
        std::vector<BlkPos> keep;
        for (size_t i = 0; i < blk_.size(); ++i) {
            const int64_t nxt = i + 1 < blk_.size() ? blk_[i + 1].u_off : (int64_t)ubuf_.size();
            if (leftover > 0 && nxt > (int64_t)u_off_) keep.push_back(BlkPos{blk_[i].u_off - (int64_t)u_off_, blk_[i].c_off});
        }
        for (const BlkPos& b : grp_blk_) keep.push_back(BlkPos{b.u_off + (int64_t)leftover, b.c_off});
        blk_.swap(keep);
    }
    if (leftover == 0) {
        ubuf_.swap(grp);
    } else {
        if (u_off_ > 0) memmove(ubuf_.data(), ubuf_.data() + u_off_, leftover);
        ubuf_.resize(leftover);
        ubuf_.insert(ubuf_.end(), grp.begin(), grp.end());
    }
    u_off_ = 0;
    // the next group is inflated while the caller parses this one
    // (not with a jump list: the next seek would throw the work away)
    if (!light_ && !small_groups_) pf_ = std::async(std::launch::async, [this]() { return produce(pf_buf_, pf_blk_); });
    return true;
}

uint64_t BamStream::tell() const {
    // bgzf_tell: compressed offset of the block the byte lies in << 16 | offset inside the inflated block
    const int64_t u = (int64_t)u_off_;
    size_t lo = 0, hi = blk_.size();
    while (lo < hi) {  // last block with u_off <= u
        const size_t mid = (lo + hi) >> 1;
        if (blk_[mid].u_off <= u) lo = mid + 1; else hi = mid;
    }
    if (lo == 0) return 0;
    return (blk_[lo - 1].c_off << 16) | (uint64_t)(u - blk_[lo - 1].u_off);
}

bool BamStream::seek_voffset(uint64_t v) {
    if (pf_.valid()) pf_.get();  // the prefetch task owns fp_ and the compressed buffer
    pf_buf_.clear();
    pf_blk_.clear();
    drained_ = false;
    if (fseek(fp_, (long)(v >> 16), SEEK_SET) != 0) { failed_ = true; fail("seek failed"); return false; }
    cbuf_.clear(); c_off_ = 0; file_eof_ = false;
    cbuf_base_ = v >> 16;
    ubuf_.clear(); u_off_ = 0;
    blk_.clear();
    ++n_seeks_;
    if (!fill()) return false;
    u_off_ = (size_t)(v & 0xffff);
    return true;
}

void BamStream::set_jumps(const std::vector<std::pair<int32_t, int32_t>>& intervals) {
    jumps_ = intervals;
    jk_ = 0;
    small_groups_ = jumps_.size() > 1;
}

bool BamStream::need(size_t n) {
    while (ubuf_.size() - u_off_ < n)
        if (!fill()) return false;
    return true;
}

bool BamStream::read_header() {
    if (!need(12)) return false;
    const uint8_t* p = ubuf_.data() + u_off_;
    if (memcmp(p, "BAM\1", 4)) { set_error("not a BAM file"); return false; }
    uint32_t l_text = le32(p + 4);
    if (!need(12 + (size_t)l_text)) return false;
    p = ubuf_.data() + u_off_;
    int n_ref = (int)le32(p + 8 + l_text);
    u_off_ += 12 + l_text;
    for (int i = 0; i < n_ref; ++i) {
        if (!need(4)) return false;
        uint32_t ln = le32(ubuf_.data() + u_off_);
        if (!need(8 + (size_t)ln)) return false;
        p = ubuf_.data() + u_off_;
        names_.emplace_back((const char*)p + 4, strnlen((const char*)p + 4, ln));
        lens_.push_back((int32_t)le32(p + 4 + ln));
        u_off_ += 8 + ln;
    }
    return true;
}

bool BamStream::open(const char* path, int threads, bool light) {
    fp_ = fopen(path, "rb");
    if (!fp_) { set_error("failed to open %s", path); return false; }
    threads_ = threads > 0 ? threads : 1;
    light_ = light;
    return read_header();
}

int BamStream::name2tid(const std::string& n) const {
    for (size_t i = 0; i < names_.size(); ++i)
        if (names_[i] == n) return (int)i;
    return -1;
}

bool BamStream::set_region(const char* bai_path, int tid, int beg, int end) {
    has_region_ = true;
    r_tid_ = tid; r_beg_ = beg; r_end_ = end;
    FILE* fp = fopen(bai_path, "rb");
    if (!fp) { set_error("fail to load bamfile index %s", bai_path); return false; }
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    std::vector<uint8_t> d(sz);
    if (fread(d.data(), 1, sz, fp) != (size_t)sz) { fclose(fp); return false; }
    fclose(fp);
    if (sz < 8 || memcmp(d.data(), "BAI\1", 4)) { set_error("bad BAI magic"); return false; }
    int n_ref = (int)le32(d.data() + 4);
    long o = 8;
    uint64_t voff = 0;
    bool found = false;
    // every count and offset comes from the file: nothing is read beyond its size
    auto room = [&](long bytes) { return bytes >= 0 && o >= 0 && o <= sz && bytes <= sz - o; };
    bool bad = n_ref < 0;
    for (int t = 0; !bad && t < n_ref && o + 4 <= sz; ++t) {
        int n_bin = (int)le32(d.data() + o);
        o += 4;
        if (n_bin < 0) { bad = true; break; }
        uint64_t min_chunk = UINT64_MAX, max_chunk_end = 0;
        for (int b = 0; b < n_bin; ++b) {
            if (!room(8)) { bad = true; break; }
            uint32_t bin = le32(d.data() + o);
            int n_chunk = (int)le32(d.data() + o + 4);
            o += 8;
            if (n_chunk < 0 || !room(16L * n_chunk)) { bad = true; break; }
            for (int c = 0; c < n_chunk; ++c) {
                uint64_t cb = le64(d.data() + o);
                if (bin != 37450 && cb < min_chunk) min_chunk = cb;
                if (bin != 37450) max_chunk_end = std::max<uint64_t>(max_chunk_end, le64(d.data() + o + 8));
                o += 16;
            }
        }
        if (bad || !room(4)) { bad = true; break; }
        int n_intv = (int)le32(d.data() + o);
        o += 4;
        if (n_intv < 0 || !room(8L * n_intv)) { bad = true; break; }
        if (t == tid) {
            int w = beg >> 14;
            uint64_t v = 0;
            if (n_intv > 0) {
                if (w >= n_intv) w = n_intv - 1;
                v = le64(d.data() + o + 8L * w);
                while (v == 0 && w > 0) { --w; v = le64(d.data() + o + 8L * w); }
            }
            if (v == 0 && min_chunk != UINT64_MAX) v = min_chunk;
            voff = v;
            found = v != 0;
            lin_.resize((size_t)n_intv);
            for (int k = 0; k < n_intv; ++k) lin_[k] = le64(d.data() + o + 8L * k);
            // no record of this contig lies in a block behind the end of its last chunk: the stream stops reading there
            // instead of inflating into the next contig
            stop_coff_ = max_chunk_end ? (max_chunk_end >> 16) : 0;
            break;
        }
        o += 8L * n_intv;
    }
    if (bad) { set_error("corrupt BAI index %s", bai_path); return false; }
    light_ = false;
    if (!found) { region_done_ = true; return true; }  // no alignments on that contig
    // (a prefetch started after the header holds bytes from the wrong place: seek_voffset drops it)
    if (!seek_voffset(voff)) {
        if (failed_) return false;
        region_done_ = true;
    }
    --n_seeks_;  // the positioning of a region is not a jump
    return true;
}

int BamStream::next(BamRecView& r) {
    while (true) {
        if (region_done_) return -1;
        if (!need(4)) {
            if (failed_) return -2;
            if (ubuf_.size() != u_off_) { failed_ = true; fail("truncated BAM record"); set_error("%s", error().c_str()); return -2; }
            return -1;
        }
        uint32_t bs = le32(ubuf_.data() + u_off_);
        if (bs < 32) { failed_ = true; fail("corrupt BAM record"); set_error("%s", error().c_str()); return -2; }
        if (!need(4 + (size_t)bs)) {
            if (!failed_) { failed_ = true; fail("truncated BAM record"); }
            set_error("%s", error().c_str());
            return -2;
        }
        const uint8_t* p = ubuf_.data() + u_off_ + 4;
        u_off_ += 4 + bs;
        ++n_records_;
        r.tid = (int32_t)le32(p);
        r.pos = (int32_t)le32(p + 4);
        r.l_qname = p[8];
        r.mapq = p[9];
        r.n_cigar = le16(p + 12);
        r.flag = le16(p + 14);
        r.l_qseq = (int)le32(p + 16);
        r.mtid = (int32_t)le32(p + 20);
        r.mpos = (int32_t)le32(p + 24);
        r.isize = (int32_t)le32(p + 28);
        if (r.l_qseq < 0) { failed_ = true; fail("corrupt BAM record (negative l_seq)"); set_error("%s", error().c_str()); return -2; }
        size_t fixed = (size_t)r.l_qname + 4u * r.n_cigar + ((size_t)r.l_qseq + 1) / 2 + (size_t)r.l_qseq;
        if (32 + fixed > bs) { failed_ = true; fail("corrupt BAM record layout"); set_error("%s", error().c_str()); return -2; }
        r.qname = (const char*)p + 32;
        r.cigar_bytes = p + 32 + r.l_qname;
        r.seq = r.cigar_bytes + 4u * r.n_cigar;
        r.qual = r.seq + (r.l_qseq + 1) / 2;
        r.aux = r.qual + r.l_qseq;
        r.l_aux = (int)(bs - 32 - fixed);
        if (has_region_) {
            if (r.tid != r_tid_ || r.pos >= r_end_) {
                if (r.tid < 0 || r.tid > r_tid_ || (r.tid == r_tid_ && r.pos >= r_end_)) { region_done_ = true; return -1; }
                continue;
            }
            int rlen = 0;
            for (int k = 0; k < r.n_cigar; ++k) {
                uint32_t w = le32(r.cigar_bytes + 4 * k);
                if ((0x3C1A7 >> ((w & 0xf) << 1)) & 2) rlen += (int)(w >> 4);
            }
            if ((r.flag & 4) || r.n_cigar == 0 || rlen == 0) rlen = 1;
            if (r.pos + rlen <= r_beg_) continue;
            if (!jumps_.empty()) {
                while (jk_ < jumps_.size() && jumps_[jk_].second <= r.pos) ++jk_;  // intervals this record starts behind
                if (jk_ == jumps_.size()) { region_done_ = true; return -1; }
                if (r.pos + rlen <= jumps_[jk_].first) {
                    // The record lies in the gap in front of the next interval. The linear index gives the offset of
                    // the first record that reaches the 16 kb window the interval starts in; go there if that is well
                    // ahead of where the stream stands (never backwards: what lies between is unread, nothing repeats).
                    const size_t w = (size_t)(jumps_[jk_].first >> 14);
                    uint64_t v = w < lin_.size() ? lin_[w] : 0;
                    for (size_t k = w; v == 0 && k + 1 < lin_.size(); ) v = lin_[++k];  // windows without records
                    const uint64_t here = tell();
                    if (v != 0 && (v >> 16) > (here >> 16) + (192u << 10)) {
                        if (!seek_voffset(v)) { if (failed_) return -2; region_done_ = true; return -1; }
                    }
                    continue;
                }
            }
        }
        return 0;
    }
}

// ------------------------------------------------------------------------------------- regions
void RegionIndex::build(int n, const char* const* names, const int32_t* beg0, const int32_t* end0) {
    struct Iv { int b, e, p; };
    std::unordered_map<std::string, std::vector<Iv>> tmp;
    for (int i = 0; i < n; ++i) tmp[names[i]].push_back({beg0[i], end0[i], i});
    for (auto& kv : tmp) {
        std::stable_sort(kv.second.begin(), kv.second.end(), [](const Iv& a, const Iv& b) {
            return a.b != b.b ? a.b < b.b : a.e < b.e;
        });
        Chr& c = chr[kv.first];
        int me = -1;
        for (auto& iv : kv.second) {
            c.beg.push_back(iv.b);
            c.end.push_back(iv.e);
            c.payload.push_back(iv.p);
            me = std::max(me, iv.e);
            c.maxend.push_back(me);
        }
    }
}
const RegionIndex::Chr* RegionIndex::find(const std::string& name) const {
    auto it = chr.find(name);
    return it == chr.end() ? nullptr : &it->second;
}
bool RegionIndex::overlap(const Chr* c, int from, int to) {
    if (!c || c->beg.empty()) return false;
    size_t k = std::upper_bound(c->beg.begin(), c->beg.end(), to) - c->beg.begin();
    if (k == 0) return false;
    return c->maxend[k - 1] >= from;
}

// ------------------------------------------------------------------------------------- helpers
static const uint8_t* aux_get(const BamRecView& r, const char tag[2]) {
    const uint8_t* p = r.aux;
    const uint8_t* end = r.aux + r.l_aux;
    auto tsz = [](int t) {
        switch (t) {
        case 'A': case 'c': case 'C': return 1;
        case 's': case 'S': return 2;
        case 'i': case 'I': case 'f': return 4;
        case 'd': return 8;
        default: return 0;
        }
    };
    while (p + 3 <= end) {
        bool hit = p[0] == (uint8_t)tag[0] && p[1] == (uint8_t)tag[1];
        int t = p[2];
        const uint8_t* v = p + 2;
        p += 3;
        if (t == 'Z' || t == 'H') {
            while (p < end && *p) ++p;
            ++p;
        } else if (t == 'B') {
            if (p + 5 > end) return nullptr;
            int sz = tsz(p[0]);
            uint32_t cnt = le32(p + 1);
            p += 5 + (size_t)sz * cnt;
        } else {
            int sz = tsz(t);
            if (!sz) return nullptr;
            p += sz;
        }
        if (p > end) return nullptr;
        if (hit) return v;
    }
    return nullptr;
}
static const char* aux_z(const BamRecView& r, const std::string& tag) {
    if (tag.size() != 2) return nullptr;
    const uint8_t* v = aux_get(r, tag.c_str());
    if (!v || (*v != 'Z' && *v != 'H')) return nullptr;
    return (const char*)v + 1;
}
static inline int seqi(const uint8_t* s, int i) { return (s[i >> 1] >> ((~i & 1) << 2)) & 0xf; }

// htsbox-ext.c:17-97 -> 0 pass, >0 fail, -1 inconsistent MD
static int md_verdict(const BamRecView& r, const uint32_t* cigar, int n_types, int n_mis) {
    if (n_types < 1 && n_mis < 1) return 0;
    const uint8_t* md = aux_get(r, "MD");
    if (!md || (*md != 'Z' && *md != 'H')) return 0;
    int x = 0;
    for (int k = 0; k < r.n_cigar; ++k) {
        int op = cigar[k] & 0xf;
        if (op == 0 || op == 7 || op == 8) x += (int)(cigar[k] >> 4);
    }
    const char* p = (const char*)md + 1;
    int y = 0, nm = 0, ntypes = 0;
    unsigned seen = 0;
    while (isdigit((unsigned char)*p)) {
        char* q;
        y += (int)strtol(p, &q, 10);
        p = q;
        if (*p == 0) break;
        if (*p == '^') {
            ++p;
            while (isalpha((unsigned char)*p)) ++p;
        } else {
            while (isalpha((unsigned char)*p)) {
                if (y >= x) { y = -1; break; }
                int code = y < r.l_qseq ? seqi(r.seq, y) : 15;
                ++nm;
                if (!(seen & (1u << code))) { seen |= 1u << code; ++ntypes; }
                ++y;
                ++p;
            }
            if (y == -1) break;
        }
    }
    if (x != y) return -1;
    return (ntypes >= n_types && nm >= n_mis) ? (ntypes > 0 ? ntypes : 0) : 0;
}

static unsigned x31(const char* s) {
    unsigned h = (unsigned)(int)*s;
    if (h)
        for (++s; *s; ++s) h = (h << 5) - h + (unsigned)(int)*s;
    return h;
}
static unsigned wang(unsigned key) {
    key += ~(key << 15); key ^= (key >> 10); key += (key << 3);
    key ^= (key >> 6); key += ~(key << 11); key ^= (key >> 16);
    return key;
}

bool fasta_fetch(const char* fa, const std::string& name, std::string& out) {
    std::string fai = std::string(fa) + ".fai";
    FILE* fp = fopen(fai.c_str(), "r");
    if (!fp) return false;
    char line[4096];
    long long slen = -1, off = 0, lb = 0, lw = 0;
    while (fgets(line, sizeof(line), fp)) {
        char* tab = strchr(line, '\t');
        if (!tab) continue;
        *tab = 0;
        if (name != line) continue;
        if (sscanf(tab + 1, "%lld\t%lld\t%lld\t%lld", &slen, &off, &lb, &lw) != 4) slen = -1;
        break;
    }
    fclose(fp);
    if (slen < 0) return false;
    fp = fopen(fa, "rb");
    if (!fp) return false;
    fseek(fp, (long)off, SEEK_SET);
    out.clear();
    out.reserve(slen);
    // whole lines at once: linebases graph characters then the line terminator
    size_t raw = (size_t)(lb > 0 ? (slen / lb + 1) * lw : slen) + 16;
    std::vector<char> buf(raw);
    size_t got = fread(buf.data(), 1, raw, fp);
    fclose(fp);
    for (size_t i = 0; i < got && (long long)out.size() < slen; ++i)
        if (isgraph((unsigned char)buf[i])) out.push_back(buf[i]);
    return true;
}

int parse_region(const char* s, int* beg, int* end) {
    const char* colon = strrchr(s, ':');
    *beg = 0;
    *end = INT_MAX;
    if (!colon) return (int)strlen(s);
    char tmp[128];
    int k = 0;
    for (const char* p = colon + 1; *p && k < 127; ++p)
        if (*p != ',') tmp[k++] = *p;
    tmp[k] = 0;
    char* q;
    long b = strtol(tmp, &q, 10);
    if (q == tmp) return -1;
    *beg = (int)(b - 1);
    if (*beg < 0) *beg = 0;
    if (*q == '-') {
        char* q2;
        long e = strtol(q + 1, &q2, 10);
        *end = q2 == q + 1 ? INT_MAX : (int)e;
    } else if (*q) {
        return -1;
    }
    if (*beg > *end) return -1;
    return (int)(colon - s);
}

// ------------------------------------------------------------------------------------- batches
void PackedBatch::release() {
    raer_pinned_free(hdr, pinned); raer_pinned_free(maxend, pinned); raer_pinned_free(cigar, pinned);
    raer_pinned_free(seq, pinned); raer_pinned_free(qual, pinned); raer_pinned_free(pair_b, pinned);
    raer_pinned_free(cb, pinned); raer_pinned_free(umi, pinned);
    hdr = nullptr; maxend = nullptr; cigar = nullptr; seq = nullptr; qual = nullptr; pair_b = nullptr;
    cb = nullptr; umi = nullptr;
    cap_reads = cap_cigar = cap_sq = cap_pairs = 0;
}

void PackedBatch::reserve(size_t reads, size_t cig, size_t sq, bool sc) {
    int pin = 0;
    auto grow = [&](void** p, size_t bytes) {
        raer_pinned_free(*p, pinned);
        if (want_pinned) *p = raer_pinned_alloc(bytes ? bytes : 16, &pin);
        else *p = malloc(bytes ? bytes : 16);
    };
    if (reads > cap_reads || !hdr) {
        size_t n = reads + reads / 4 + 1024;
        grow((void**)&hdr, n * sizeof(raer_read_hdr));
        grow((void**)&maxend, n * 4);
        grow((void**)&pair_b, n * 4);
        if (sc) { grow((void**)&cb, n * 4); grow((void**)&umi, n * 4); }
        cap_reads = n;
        pinned = pin;
    }
    if (cig > cap_cigar || !cigar) {
        size_t n = cig + cig / 4 + 1024;
        grow((void**)&cigar, n * 4);
        cap_cigar = n;
    }
    if (sq > cap_sq || !seq) {
        size_t n = sq + sq / 4 + 4096;
        grow((void**)&seq, n);
        grow((void**)&qual, 2 * n);
        cap_sq = n;
    }
}

Ingest::~Ingest() {
    for (auto* f : files_) delete f;
}

bool Ingest::open(const char* const* paths, const char* const* indexes, int threads, const char* region,
                  const std::vector<std::pair<int32_t, int32_t>>* jumps) {
    if (threads <= 0) threads = usable_cpus();
    if (threads <= 0) threads = 4;
    for (int i = 0; i < conf_.n_files; ++i) {
        FilePending* f = new FilePending();
        files_.push_back(f);
        if (!f->bs.open(paths[i], threads, region != nullptr)) return false;
    }
    if (region) {
        int rb, re;
        int nl = parse_region(region, &rb, &re);
        if (nl < 0) { set_error("fail to parse region '%s'", region); return false; }
        int tid = files_[0]->bs.name2tid(region);
        if (tid >= 0) { rb = 0; re = INT_MAX; }
        else tid = files_[0]->bs.name2tid(std::string(region, nl));
        if (tid < 0) { set_error("fail to parse region '%s'", region); return false; }
        has_region_ = true;
        reg_tid_ = tid; reg_beg_ = rb; reg_end_ = re;
        int seek_beg = rb, seek_end = re;
        if (jumps && !jumps->empty()) {
            // a site / region list on this contig: small inflate groups from the start, and the stream is positioned at
            // the first interval, not at the start of the contig (rows still come from [reg_beg_, reg_end_) only)
            for (auto* f : files_) f->bs.set_jumps(*jumps);
            seek_beg = std::max(rb, (int)jumps->front().first);
            seek_end = std::min(re, (int)jumps->back().second);
        }
        for (int i = 0; i < conf_.n_files; ++i) {
            std::string bai = (indexes && indexes[i]) ? indexes[i] : std::string(paths[i]) + ".bai";
            // each file resolves the name in its own header, like sam_itr_querys
            int ftid = files_[i]->bs.name2tid(tid < (int)names().size() ? names()[tid] : "");
            if (ftid < 0) { set_error("fail to parse region '%s' with %s", region, paths[i]); return false; }
            if (!files_[i]->bs.set_region(bai.c_str(), ftid, seek_beg, seek_end)) return false;
        }
    }
    return true;
}

bool Ingest::open_headers(const char* const* paths) {
    for (int i = 0; i < conf_.n_files; ++i) {
        FilePending* f = new FilePending();
        files_.push_back(f);
        if (!f->bs.open(paths[i], 1, true)) return false;
    }
    return true;
}

double Ingest::wait_seconds() const {
    double t = 0;
    for (auto* f : files_) t += f->bs.wait_seconds();
    return t;
}
double Ingest::decode_seconds() const {
    double t = 0;
    for (auto* f : files_) t += f->bs.decode_seconds();
    return t;
}
bool Ingest::failed() const {
    for (auto* f : files_)
        if (f->failed) return true;
    return false;
}
std::string Ingest::error() const {
    for (auto* f : files_)
        if (f->failed) return f->err.empty() ? std::string("BAM read error") : f->err;
    return std::string();
}
int64_t Ingest::records() const {
    int64_t n = 0;
    for (auto* f : files_) n += f->bs.records_read();
    return n;
}
int64_t Ingest::aligned_bases() const {
    int64_t n = 0;
    for (auto* f : files_) n += f->n_aligned_bases;
    return n;
}
int64_t Ingest::inflated_bytes() const {
    int64_t n = 0;
    for (auto* f : files_) n += f->bs.inflated_bytes();
    return n;
}
int64_t Ingest::seeks() const {
    int64_t n = 0;
    for (auto* f : files_) n += f->bs.seeks();
    return n;
}
void Ingest::set_jumps(const std::vector<std::pair<int32_t, int32_t>>& intervals) {
    if (!has_region_) return;  // the linear index of the region's contig is what the jumps go through
    for (auto* f : files_) f->bs.set_jumps(intervals);
}

bool indexes_usable(const char* const* paths, const char* const* indexes, int n) {
    for (int i = 0; i < n; ++i) {
        std::string bai = (indexes && indexes[i]) ? indexes[i] : std::string(paths[i]) + ".bai";
        FILE* fp = fopen(bai.c_str(), "rb");
        char magic[4] = {0, 0, 0, 0};
        const bool ok = fp && fread(magic, 1, 4, fp) == 4 && !memcmp(magic, "BAI\1", 4);
        if (fp) fclose(fp);
        if (!ok) return false;
    }
    return true;
}

// The intervals of one contig of a region index as a jump list: padded by one position on either side (the prefilter's
// overlap test counts a read that ends where an interval starts, src/plp.c:56-60), neighbours closer than `gap` merged
// (a seek only pays across a few index windows). Empty when the intervals cover so much of their span that streaming
// is the better plan.
std::vector<std::pair<int32_t, int32_t>> jump_intervals(const RegionIndex::Chr* c, int32_t gap) {
    std::vector<std::pair<int32_t, int32_t>> out;
    if (!c || c->beg.empty()) return out;
    int64_t covered = 0;
    for (size_t i = 0; i < c->beg.size(); ++i) {
        const int32_t b = c->beg[i] > 0 ? c->beg[i] - 1 : 0, e = c->end[i] < INT32_MAX - 2 ? c->end[i] + 2 : INT32_MAX;
        if (!out.empty() && b <= out.back().second + gap) out.back().second = std::max(out.back().second, e);
        else out.emplace_back(b, e);
    }
    for (auto& iv : out) covered += (int64_t)iv.second - iv.first;
    const int64_t span = (int64_t)out.back().second - out.front().first;
    if (out.size() == 1 || covered * 2 > span) {
        // dense list: one interval from the first to the last site still saves the two ends of the contig
        const std::pair<int32_t, int32_t> all(out.front().first, out.back().second);
        out.assign(1, all);
    }
    return out;
}

// Reads the next record that survives the prefilter into fp.look. Returns false at EOF / error.
// nt16 code of the base a pileup entry of this read shows at reference position p (pos <= p < end), the way htslib's
// resolve_cigar2 assigns qpos: inside an aligned block the aligned base, inside a deletion / ref-skip the next
// query base; 15 (N) when that offset lies beyond the read.
static int umi_code_at(const std::vector<uint32_t>& cig, const std::vector<uint8_t>& seq, int l_qseq, int pos, int p) {
    int x = pos, y = 0;
    for (uint32_t w : cig) {
        const int op = (int)(w & 0xf), len = (int)(w >> 4);
        if (op == 0 || op == 7 || op == 8) {
            if (p < x + len) { const int q = y + (p - x); return q < l_qseq ? (seq[q >> 1] >> ((~q & 1) << 2)) & 0xf : 15; }
            x += len; y += len;
        } else if (op == 2 || op == 3) {
            if (p < x + len) return y < l_qseq ? (seq[y >> 1] >> ((~y & 1) << 2)) & 0xf : 15;
            x += len;
        } else if (op == 1 || op == 4) {
            y += len;
        }
    }
    return 15;
}

bool Ingest::parse_one(int fi) {
    FilePending& F = *files_[fi];
    BamRecView r;
    std::vector<uint32_t>& cig = F.look_cigar;
    while (true) {
        int rc = F.bs.next(r);
        if (rc < 0) {
            // -1 is the end of the file (or of the region); anything else -- or a stream that stopped on a read /
            // inflate / checksum error -- fails the call: the reference raises an error here (src/plp.c:1195)
            if (rc != -1 || F.bs.failed()) { F.failed = true; F.err = F.bs.error(); }
            F.eof = true;
            F.have_look = false;
            return false;
        }
        if (r.tid < 0 || (r.flag & 4)) continue;
        int rlen = 0;
        int64_t nal = 0;
        cig.resize(r.n_cigar);
        for (int k = 0; k < r.n_cigar; ++k) {
            uint32_t w = le32(r.cigar_bytes + 4 * k);
            cig[k] = w;
            int op = w & 0xf;
            if ((0x3C1A7 >> (op << 1)) & 2) rlen += (int)(w >> 4);
            if (op == 0 || op == 7 || op == 8) nal += (w >> 4);
        }
        F.n_aligned_bases += nal;  // unit of the bench metric: every mapped input record, before filters
        int cbcol = 0;
        if (conf_.sc && conf_.cb_index && !conf_.cb_tag.empty()) {  // src/sc-plp.c:464-477
            const char* cb = aux_z(r, conf_.cb_tag);
            if (!cb) continue;
            auto it = conf_.cb_index->find(cb);
            if (it == conf_.cb_index->end()) continue;
            cbcol = it->second;
        }
        if (conf_.sc && conf_.fixed_cb) cbcol = conf_.fixed_cb;
        uint32_t test = (conf_.keep_flag[0] & ~(uint32_t)r.flag) | (conf_.keep_flag[1] & (uint32_t)r.flag);
        if (~test & 2047u) continue;
        int endpos = r.pos + ((r.n_cigar == 0 || rlen == 0) ? 1 : rlen);
        if (conf_.regidx) {
            const std::string& nm = names()[r.tid < (int)names().size() ? r.tid : 0];
            if (!RegionIndex::overlap(conf_.regidx->find(nm), r.pos, endpos)) continue;  // inclusive end, src/plp.c:56-60
        }
        if (r.mapq < conf_.min_global_mq) continue;
        if ((r.flag & 1) && !(r.flag & 2)) continue;
        if (conf_.rq_pct && conf_.rq_minq) {  // read_base_quality, float threshold
            if (!r.l_qseq) continue;
            int lq = 0;
            for (int i = 0; i < r.l_qseq; ++i) lq += r.qual[i] < conf_.rq_minq;
            double frac = (double)lq / r.l_qseq;
            if (!(frac <= (float)conf_.rq_pct)) continue;
        }
        if (r.l_qseq > 65535 || r.n_cigar > 65535) {
            F.failed = true;
            F.err = "read longer than 65535 bases is not supported";
            set_error("%s", F.err.c_str());
            F.eof = true;
            F.have_look = false;
            return false;
        }

        // ---- stream-order decisions of bam_plp_push
        if (F.live_tid != r.tid) {
            F.live_ends = decltype(F.live_ends)();
            F.live_tid = r.tid;
        }
        // the mempool count of bam_plp_push: reads still in the buffer (end >= position of the last pushed read) + 1.
        // It only matters for a read that starts where the previous one did, and only once that many were pushed.
        if (F.prev_tid == r.tid && F.prev_pos == r.pos && (int64_t)F.live_ends.size() + 1 > conf_.max_depth) {
            const int32_t lim = F.prev_pos;
            while (!F.live_ends.empty() && F.live_ends.top() < lim) F.live_ends.pop();
            if ((int64_t)F.live_ends.size() + 1 > conf_.max_depth) {
                if (conf_.remove_overlaps) {
                    const size_t qn = strlen(r.qname);
                    if (QnameTable::Ent* e = F.olap.find(r.qname, qn, QnameTable::hash(r.qname, qn))) F.olap.erase(e);
                }
                continue;  // dropped by maxcnt
            }
        }
        // htslib bam_plp_push refuses input that is not coordinate sorted ("the input is not sorted"): everything below
        // -- batch windows, maxend prefixes, mate pairing, the max_depth count -- relies on the order
        if (r.tid < F.sort_tid || (r.tid == F.sort_tid && r.pos < F.sort_pos)) {
            F.failed = true;
            F.err = r.tid < F.sort_tid ? "the input is not sorted (chromosomes out of order)" : "the input is not sorted (reads out of order)";
            set_error("%s", F.err.c_str());
            F.eof = true;
            F.have_look = false;
            return false;
        }
        F.sort_tid = r.tid;
        F.sort_pos = r.pos;
        HostRead& H = F.look;
        memset(&H, 0, sizeof(H));
        H.h.pos = r.pos;
        H.h.end = r.pos + rlen;
        H.h.flag = r.flag;
        H.h.mapq = r.mapq;
        H.h.l_qseq = (uint16_t)r.l_qseq;
        H.h.n_cigar = (uint16_t)r.n_cigar;
        H.h.mate = -1;
        H.mate_ord = -1;
        H.keep_end = H.h.end;
        H.cb = cbcol;
        uint8_t bits = 0;
        {  // invert_read_orientation
            int lt = conf_.libtype[fi], rev = (r.flag & 16) != 0, inv = 0;
            if (lt == 1 || lt == 2) {
                int r1 = -1;
                if (!(r.flag & 1)) r1 = 1;
                else if (r.flag & 64) r1 = 1;
                else if (r.flag & 128) r1 = 0;
                if (r1 >= 0) {
                    int when_rev = (lt == 2);
                    inv = r1 ? (rev == when_rev) : (rev != when_rev);
                }
            }
            if (inv) bits |= RAER_RB_INVERT;
        }
        if (conf_.n_mm_type > 0 || conf_.n_mm > 0) {
            int m = md_verdict(r, cig.data(), conf_.n_mm_type, conf_.n_mm);
            if (m > 0) bits |= RAER_RB_MMFAIL;
            else if (m < 0) bits |= RAER_RB_MMERR;
        }
        if (conf_.min_overhang && r.l_qseq >= 2 && (r.seq[0] & 0xf) == 3) bits |= RAER_RB_OHANG_N;
        // bam_plp_push keeps a read only if it can still produce a column (end > iterator position)
        bool kept = (r.tid != F.prev_tid) || (H.h.end > F.prev_pos);
        if (!kept) {
            F.prev_tid = r.tid;
            F.prev_pos = r.pos;
            continue;
        }
        H.ordinal = F.next_ordinal++;
        if (conf_.remove_overlaps) {
            // overlap_push
            bool eligible = !(r.flag & 8) && (r.flag & 2);
            long long isz = r.isize < 0 ? -(long long)r.isize : r.isize;
            if ((r.mtid >= 0 && r.tid != r.mtid) || (isz >= 2LL * r.l_qseq && r.mpos >= H.h.end)) eligible = false;
            if (eligible) {
                const size_t qn = strlen(r.qname);
                const uint64_t qh = QnameTable::hash(r.qname, qn);
                // an entry is gone once its read left the pileup buffer (end < position of the last pushed read)
                const int cur_tid = r.tid, cur_prev = F.prev_pos;
                auto dead = [cur_tid, cur_prev](const QnameTable::Ent& e) { return e.tid != cur_tid || e.end < cur_prev; };
                QnameTable::Ent* it = F.olap.find(r.qname, qn, qh);
                if (it && dead(*it)) {
                    F.olap.erase(it);
                    it = nullptr;
                }
                if (!it) {
                    if (r.mpos >= r.pos || ((r.flag & 1) && r.mpos == -1)) {
                        bool keep_a = (wang(x31(r.qname)) & 1) != 0;
                        F.olap.insert(r.qname, qn, qh, H.ordinal, r.tid, H.h.end, keep_a, dead);
                    }
                } else {
                    H.mate_ord = it->ordinal;
                    if (it->keep_a) bits |= RAER_RB_KEEP_A;
                    F.olap.erase(it);
                }
            }
        }
        if (conf_.bulk_umi) {
            // check_umi (src/plp.c:622-646, called at :738 before the filter verdict is used): at every site the first
            // entry (BAM order) of each UMI per strand and file goes on, later ones are skipped; an entry takes part if
            // its base character is not N, whatever the filters say, deletions and ref-skips included. The reads are
            // coordinate sorted, so for this read the positions taken by earlier reads of its group form a prefix of
            // its span (hdr.aux), except where every earlier read shows an N (umi_holes).
            const char* u = aux_z(r, conf_.umi_tag);
            if (F.umi_tid != r.tid) { F.umi_groups.clear(); F.umi_tid = r.tid; }
            if (!u) {
                H.h.aux = 0xffffffffu;
            } else {
                std::string key(1, (bits & RAER_RB_INVERT) ? '-' : '+');
                key += u;
                auto& G = F.umi_groups[key];
                size_t w = 0;
                int max_end = r.pos;
                bool any_n = false;
                for (size_t k = 0; k < G.size(); ++k)
                    if (G[k].end > r.pos) {
                        max_end = std::max(max_end, G[k].end);
                        any_n |= G[k].has_n;
                        if (w != k) G[w] = std::move(G[k]);
                        ++w;
                    }
                G.resize(w);
                const int s_end = std::min(max_end, (int)H.h.end);
                if (s_end > r.pos) {
                    if (!any_n) {
                        H.h.aux = (uint32_t)s_end;
                    } else {
                        std::vector<char> taken((size_t)(s_end - r.pos), 0);
                        for (const auto& e : G) {
                            const int hi = std::min(e.end, s_end);
                            for (int p = r.pos; p < hi; ++p)
                                if (!e.has_n || umi_code_at(e.cigar, e.seq, e.l_qseq, e.pos, p) != 15) taken[p - r.pos] = 1;
                        }
                        int last = -1;
                        for (int p = s_end - 1; p >= r.pos; --p) if (taken[p - r.pos]) { last = p; break; }
                        if (last >= 0) {
                            H.h.aux = (uint32_t)(last + 1);
                            std::vector<int32_t> holes;
                            for (int p = r.pos; p <= last; ++p) if (!taken[p - r.pos]) holes.push_back(p);
                            if (!holes.empty()) { bits |= RAER_RB_UMI_HOLES; F.umi_holes[H.ordinal] = std::move(holes); }
                        }
                    }
                }
                FilePending::UmiLive me;
                me.pos = r.pos; me.end = H.h.end; me.l_qseq = r.l_qseq;
                me.has_n = false;
                for (int q = 0; q < r.l_qseq && !me.has_n; ++q) me.has_n = ((r.seq[q >> 1] >> ((~q & 1) << 2)) & 0xf) == 15;
                // a read that ends in a deletion / ref-skip shows N there; treat it like a read with an N
                if (!me.has_n && r.n_cigar) { const int lop = (int)(cig[r.n_cigar - 1] & 0xf); me.has_n = lop == 2 || lop == 3; }
                if (me.has_n) { me.cigar = cig; me.seq.assign(r.seq, r.seq + (r.l_qseq + 1) / 2); }
                G.push_back(std::move(me));
                if (F.umi_groups.size() > 2000000) {  // drop the groups whose reads have all left the pileup
                    for (auto it = F.umi_groups.begin(); it != F.umi_groups.end();) {
                        bool live = false;
                        for (const auto& e : it->second) live |= e.end > r.pos;
                        it = live ? std::next(it) : F.umi_groups.erase(it);
                    }
                }
            }
        }
        H.h.bits = bits;
        F.prev_tid = r.tid;
        F.prev_pos = r.pos;
        F.live_ends.push(H.h.end);
        if (F.live_ends.size() > F.live_compact_at) {  // every later check asks for ends >= a position >= this one
            const int32_t lim = r.pos;
            while (!F.live_ends.empty() && F.live_ends.top() < lim) F.live_ends.pop();
            F.live_compact_at = std::max<size_t>(4096, 2 * F.live_ends.size());
        }
        if (conf_.sc) {
            if (!conf_.umi_tag.empty()) {
                const char* u = aux_z(r, conf_.umi_tag);
                if (u) {
                    auto ins = umi_dict_.emplace(u, (uint32_t)umi_dict_.size() + 1);
                    H.umi = ins.first->second;
                }
            }
        }
        F.look_seq = r.seq;  // the stream buffer keeps the record until the next one is asked for
        F.look_qual = r.qual;
        F.look_tid = r.tid;
        F.have_look = true;
        return true;
    }
}

bool Ingest::parse_ahead(int fi, size_t want) {
    FilePending& F = *files_[fi];
    while (true) {
        if (!F.have_look) {
            if (F.eof) return false;
            if (!parse_one(fi)) return false;
        }
        if (F.look_tid != cur_tid_) return false;  // next contig: stays in look
        if (F.reads.size() - F.front >= want) return true;  // enough pending; look holds the next read
        HostRead H = F.look;
        H.h.cigar_off = (uint32_t)(F.cigar.size());
        H.h.sq_off = (uint32_t)(F.seq.size());
        F.cigar.insert(F.cigar.end(), F.look_cigar.begin(), F.look_cigar.end());
        F.seq.insert(F.seq.end(), F.look_seq, F.look_seq + (H.h.l_qseq + 1) / 2);
        // every read starts on a 4-byte (sequence) / 8-byte (quality) boundary: the kernels fetch 8 bases
        // per load; qualities live at 2 * sq_off
        while (F.seq.size() & 3) F.seq.push_back(0xff);
        F.qual.insert(F.qual.end(), F.look_qual, F.look_qual + H.h.l_qseq);
        F.qual.resize(2 * F.seq.size(), 0);
        if (H.mate_ord >= 0 && F.front < F.reads.size()) {
            // htslib's overlap pass can rewrite qualities of the later mate far beyond the earlier mate's
            // end (deletion catch-up across a ref-skip), so the earlier mate must stay packed with it
            int64_t k = (int64_t)F.front + (H.mate_ord - F.reads[F.front].ordinal);
            if (k >= (int64_t)F.front && k < (int64_t)F.reads.size())
                F.reads[k].keep_end = std::max(F.reads[k].keep_end, (int32_t)H.h.end);
        }
        F.reads.push_back(H);
        F.have_look = false;
    }
}

int Ingest::next_contig() {
    int best = -1;
    for (int i = 0; i < conf_.n_files; ++i) {
        FilePending& F = *files_[i];
        // drop whatever is left of the previous contig
        F.reads.clear(); F.front = 0; F.cigar.clear(); F.seq.clear(); F.qual.clear();
        if (!F.have_look && !F.eof) parse_one(i);
        if (F.have_look && (best < 0 || F.look_tid < best)) best = F.look_tid;
    }
    cur_tid_ = best;
    win_beg_ = 0;
    if (has_region_ && best >= 0) win_beg_ = reg_beg_;
    return best;
}

bool Ingest::next_batch(PackedBatch& out) {
    if (cur_tid_ < 0) return false;
    const int nf = conf_.n_files;
    size_t want = (size_t)std::max<int64_t>(16, conf_.target_reads / nf);
    int boundary;
    const double t0p = now_seconds();
    while (true) {
        boundary = INT_MAX;
        bool all_done = true;
        // The files are independent until the window boundary is taken: parse them side by side (the UMI
        // dictionary of the single-cell path is shared, that path stays serial).
        if (nf > 1 && !conf_.sc) {
            std::vector<std::thread> th;
            th.reserve(nf);
            for (int i = 0; i < nf; ++i) th.emplace_back([this, i, want]() { parse_ahead(i, want); });
            for (auto& t : th) t.join();
        } else {
            for (int i = 0; i < nf; ++i) parse_ahead(i, want);
        }
        for (int i = 0; i < nf; ++i) {
            FilePending& F = *files_[i];
            bool done = !F.have_look || F.look_tid != cur_tid_;
            if (!done) {
                all_done = false;
                boundary = std::min(boundary, (int)F.look.h.pos);
            }
        }
        if (all_done || boundary > win_beg_) break;
        want *= 2;  // every pending read starts at the window start: look further
    }
    double t0 = now_seconds();
    t_parse_ += t0 - t0p;
    // reads of each file starting before the boundary
    std::vector<size_t> upto(nf);
    size_t n_reads = 0, n_cig = 0, n_sq = 0;
    int min_pos = INT_MAX, max_end = INT_MIN;
    for (int i = 0; i < nf; ++i) {
        FilePending& F = *files_[i];
        size_t lo = F.front, hi = F.reads.size();
        while (lo < hi) {
            size_t mid = (lo + hi) >> 1;
            if (F.reads[mid].h.pos >= boundary) hi = mid; else lo = mid + 1;
        }
        upto[i] = lo;
        if (lo > F.front) {
            n_reads += lo - F.front;
            size_t c0 = F.reads[F.front].h.cigar_off, s0 = F.reads[F.front].h.sq_off;
            size_t c1 = lo < F.reads.size() ? F.reads[lo].h.cigar_off : F.cigar.size();
            size_t s1 = lo < F.reads.size() ? F.reads[lo].h.sq_off : F.seq.size();
            n_cig += c1 - c0;
            n_sq += s1 - s0;
            min_pos = std::min(min_pos, (int)F.reads[F.front].h.pos);
        }
    }
    if (n_reads == 0) {
        bool any_left = false;
        for (int i = 0; i < nf; ++i) any_left |= files_[i]->have_look && files_[i]->look_tid == cur_tid_;
        if (!any_left) return false;
    }
    out.reserve(n_reads, n_cig, n_sq, conf_.sc);
    out.umi_holes.clear();
    out.file_off.assign(nf + 1, 0);
    // where each file's share of the batch starts, then the files are copied side by side
    std::vector<size_t> ro_of(nf + 1, 0), co_of(nf + 1, 0), so_of(nf + 1, 0);
    for (int i = 0; i < nf; ++i) {
        FilePending& F = *files_[i];
        size_t nr_i = 0, nc_i = 0, ns_i = 0;
        if (upto[i] > F.front) {
            size_t c0 = F.reads[F.front].h.cigar_off, s0 = F.reads[F.front].h.sq_off;
            size_t c1 = upto[i] < F.reads.size() ? F.reads[upto[i]].h.cigar_off : F.cigar.size();
            size_t s1 = upto[i] < F.reads.size() ? F.reads[upto[i]].h.sq_off : F.seq.size();
            nr_i = upto[i] - F.front; nc_i = c1 - c0; ns_i = s1 - s0;
        }
        ro_of[i + 1] = ro_of[i] + nr_i; co_of[i + 1] = co_of[i] + nc_i; so_of[i + 1] = so_of[i] + ns_i;
        out.file_off[i] = (int64_t)ro_of[i];
    }
    std::vector<std::vector<int32_t>> pairs_of(nf), holes_of(nf);
    std::vector<int> max_end_of(nf, INT_MIN);
    auto pack_file = [&](int i) {
        FilePending& F = *files_[i];
        if (upto[i] <= F.front) return;
        const size_t ro = ro_of[i], co = co_of[i], so = so_of[i];
        size_t c0 = F.reads[F.front].h.cigar_off, s0 = F.reads[F.front].h.sq_off;
        size_t c1 = upto[i] < F.reads.size() ? F.reads[upto[i]].h.cigar_off : F.cigar.size();
        size_t s1 = upto[i] < F.reads.size() ? F.reads[upto[i]].h.sq_off : F.seq.size();
        memcpy(out.cigar + co, F.cigar.data() + c0, (c1 - c0) * 4);
        memcpy(out.seq + so, F.seq.data() + s0, s1 - s0);
        memcpy(out.qual + 2 * so, F.qual.data() + 2 * s0, 2 * (s1 - s0));
        int64_t front_ord = F.reads[F.front].ordinal;
        int me = INT_MIN, mx = INT_MIN;
        for (size_t k = F.front; k < upto[i]; ++k) {
            const HostRead& H = F.reads[k];
            raer_read_hdr h = H.h;
            h.cigar_off = (uint32_t)(co + (H.h.cigar_off - c0));
            h.sq_off = (uint32_t)(so + (H.h.sq_off - s0));
            h.mate = (H.mate_ord >= front_ord) ? (int32_t)(ro + (size_t)(H.mate_ord - front_ord)) : -1;
            if (!conf_.bulk_umi) h.aux = (uint32_t)(ro + (k - F.front));  // bulk UMI mode keeps its verdict there
            size_t bi = ro + (k - F.front);
            out.hdr[bi] = h;
            me = std::max(me, (int)h.end);
            out.maxend[bi] = me;
            if (h.mate >= 0) pairs_of[i].push_back((int32_t)bi);
            if (conf_.sc) { out.cb[bi] = H.cb; out.umi[bi] = H.umi; }
            if (h.bits & RAER_RB_UMI_HOLES) {
                auto it = F.umi_holes.find(H.ordinal);
                if (it != F.umi_holes.end())
                    for (int32_t hp : it->second) { holes_of[i].push_back((int32_t)bi); holes_of[i].push_back(hp); }
            }
            if (h.end > win_beg_) mx = std::max(mx, (int)h.end);
        }
        max_end_of[i] = mx;
    };
    if (nf > 1) {
        std::vector<std::thread> th;
        th.reserve(nf);
        for (int i = 0; i < nf; ++i) th.emplace_back(pack_file, i);
        for (auto& t : th) t.join();
    } else {
        pack_file(0);
    }
    size_t np = 0;
    for (int i = 0; i < nf; ++i) {
        for (int32_t v : pairs_of[i]) out.pair_b[np++] = v;
        out.umi_holes.insert(out.umi_holes.end(), holes_of[i].begin(), holes_of[i].end());
        max_end = std::max(max_end, max_end_of[i]);
    }
    const size_t ro = ro_of[nf], co = co_of[nf], so = so_of[nf];
    int64_t nal = 0;
    out.file_off[nf] = (int64_t)ro;
    raer_read_batch& b = out.b;
    memset(&b, 0, sizeof(b));
    b.n_files = nf;
    b.tid = cur_tid_;
    b.beg = std::max(win_beg_, min_pos == INT_MAX ? win_beg_ : std::min(min_pos, boundary));
    b.beg = std::max(b.beg, win_beg_);
    b.end = std::min(boundary, max_end == INT_MIN ? win_beg_ : max_end);
    if (has_region_) { b.beg = std::max(b.beg, reg_beg_); b.end = std::min(b.end, reg_end_); }
    if (b.end < b.beg) b.end = b.beg;
    b.n_reads = (int64_t)ro;
    b.file_off = out.file_off.data();
    b.hdr = out.hdr; b.maxend = out.maxend; b.cigar = out.cigar; b.seq = out.seq; b.qual = out.qual;
    b.n_cigar_words = (int64_t)co;
    b.n_sq_bytes = (int64_t)so;
    b.pair_b = out.pair_b;
    b.n_pairs = (int64_t)np;
    b.cb = conf_.sc ? out.cb : nullptr;
    b.umi = conf_.sc ? out.umi : nullptr;
    b.n_aligned_bases = nal;
    b.umi_holes = out.umi_holes.empty() ? nullptr : out.umi_holes.data();
    b.n_umi_holes = (int64_t)(out.umi_holes.size() / 2);
    // retire reads that end before the boundary (front only, so ordinals stay contiguous) and compact pools
    win_beg_ = boundary;
    for (int i = 0; i < nf; ++i) {
        FilePending& F = *files_[i];
        while (F.front < upto[i] && F.reads[F.front].keep_end <= boundary) {
            if (F.reads[F.front].h.bits & RAER_RB_UMI_HOLES) F.umi_holes.erase(F.reads[F.front].ordinal);
            ++F.front;
        }
        if (F.front == F.reads.size()) {
            F.reads.clear(); F.front = 0; F.cigar.clear(); F.seq.clear(); F.qual.clear();
        } else if (F.front > (1u << 16) && F.front * 2 > F.reads.size()) {
            uint32_t c0 = F.reads[F.front].h.cigar_off, s0 = F.reads[F.front].h.sq_off;
            F.cigar.erase(F.cigar.begin(), F.cigar.begin() + c0);
            F.seq.erase(F.seq.begin(), F.seq.begin() + s0);
            F.qual.erase(F.qual.begin(), F.qual.begin() + 2 * (size_t)s0);
            F.reads.erase(F.reads.begin(), F.reads.begin() + F.front);
            F.front = 0;
            for (auto& H : F.reads) { H.h.cigar_off -= c0; H.h.sq_off -= s0; }
        }
    }
    t_pack_ += now_seconds() - t0;
    if (boundary == INT_MAX) {
        // the contig is finished after this batch
        for (int i = 0; i < nf; ++i) { files_[i]->reads.clear(); files_[i]->front = 0; files_[i]->cigar.clear(); files_[i]->seq.clear(); files_[i]->qual.clear(); }
    }
    return true;
}

}  // namespace raer
