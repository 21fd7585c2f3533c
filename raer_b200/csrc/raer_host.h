// raer_host.h -- host-side ingest: BGZF/BAM streaming, read-level prefilter, batch packing.
//
// Replaces what the reference gets from htslib + readaln()/screadaln() (src/plp.c:32-78,
// src/sc-plp.c:449-511) and the buffer-state dependent parts of bam_plp_push / overlap_push
// (mate pairing, max_depth) that must be decided in stream order. Everything positional is left
// to the device.
#ifndef RAER_HOST_H
#define RAER_HOST_H

#include <stdint.h>
#include <stdio.h>

#include <future>
#include <mutex>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>

#include "raer_b200.h"

namespace raer {

void set_error(const char* fmt, ...);

struct BamRecView {  // points into the stream's inflate buffer; valid until the next call
    int32_t tid, pos, mtid, mpos, isize;
    uint16_t flag;
    uint8_t mapq;
    int l_qname, n_cigar, l_qseq, l_aux;
    const char* qname;
    const uint8_t* cigar_bytes;  // unaligned little-endian words
    const uint8_t* seq;
    const uint8_t* qual;
    const uint8_t* aux;
};

class BamStream {
public:
    ~BamStream();
    // light: the caller will seek with set_region(): read the header in small groups, no read-ahead before the seek
    bool open(const char* path, int threads, bool light = false);
    // restrict to [beg, end) of tid using the BAI linear index; returns false on a missing/bad index
    bool set_region(const char* bai_path, int tid, int beg, int end);
    // After set_region(): only the records that overlap one of these intervals (0-based, half open, sorted, disjoint)
    // are wanted. The stream skips the gaps between them through the linear index -- forward only, so every record
    // comes out at most once -- whenever the next interval starts far enough ahead in the file to pay for a seek
    // (site lists and sparse region lists: src/plp.c:856-882 queries the index per region through sam_itr_querys).
    void set_jumps(const std::vector<std::pair<int32_t, int32_t>>& intervals);
    int64_t inflated_bytes() const { return n_inflated_; }
    int64_t seeks() const { return n_seeks_; }
    // 0 ok, -1 EOF (or past the region), -2 error
    int next(BamRecView& r);
    int n_targets() const { return (int)names_.size(); }
    const std::vector<std::string>& names() const { return names_; }
    const std::vector<int32_t>& lengths() const { return lens_; }
    int name2tid(const std::string& n) const;
    double decode_seconds() const { return t_decode_; }
    double wait_seconds() const { return t_wait_; }  // time the parser waited for an inflated group
    int64_t records_read() const { return n_records_; }
    // a read, framing, inflate, checksum or record-layout error stopped the stream (sticky; distinct from the end of the file)
    bool failed() const { return failed_; }
    std::string error() const { std::lock_guard<std::mutex> g(err_mu_); return err_; }

private:
    int fail(const char* fmt, ...);  // records the message in the object (the prefetch task runs on another thread); returns -1
    bool failed_ = false;
    std::string err_;
    mutable std::mutex err_mu_;
    bool fill();  // next group of inflated BGZF blocks into ubuf_; the group after it is inflated in the background
    bool read_header();
    bool need(size_t n);
    FILE* fp_ = nullptr;
    int threads_ = 1;
    std::vector<uint8_t> cbuf_;  // compressed bytes not yet inflated
    size_t c_off_ = 0;
    bool file_eof_ = false;
    std::vector<uint8_t> ubuf_;  // inflated bytes
    size_t u_off_ = 0;
    std::vector<std::string> names_;
    std::vector<int32_t> lens_;
    bool has_region_ = false, region_done_ = false;
    int r_tid_ = -1, r_beg_ = 0, r_end_ = 0;
    // jump list (set_jumps) and what it needs: the linear index of r_tid_, and the file offset of every inflated
    // block still in ubuf_, so that the virtual offset of the record being parsed is known exactly (bgzf_tell)
    std::vector<std::pair<int32_t, int32_t>> jumps_;
    size_t jk_ = 0;
    std::vector<uint64_t> lin_;
    struct BlkPos { int64_t u_off; uint64_t c_off; };  // offset in ubuf_ (negative: the block began before the kept tail)
    std::vector<BlkPos> blk_, grp_blk_, pf_blk_;
    uint64_t cbuf_base_ = 0;  // file offset of cbuf_[0]
    uint64_t stop_coff_ = 0;  // set_region: file offset of the last block that can hold a record of the contig (0 = unknown)
    bool small_groups_ = false;
    int64_t n_inflated_ = 0, n_seeks_ = 0;
    uint64_t tell() const;     // virtual offset of ubuf_[u_off_]
    bool seek_voffset(uint64_t v);
    int produce(std::vector<uint8_t>& out, std::vector<BlkPos>& pos);
    double t_decode_ = 0;
    double t_wait_ = 0;
    int64_t n_records_ = 0;
    // prefetch: while the parser walks ubuf_, a task inflates the following group (it owns fp_, cbuf_, c_off_,
    // file_eof_ and t_decode_ until fill() collects it)
    std::vector<uint8_t> pf_buf_;
    std::vector<uint8_t> grp_;  // the group fill() hands over; a member so that its storage is recycled
    std::future<int> pf_;
    bool drained_ = false;  // produce() reported the end of the file
    bool light_ = false;
};

// sorted, merged-on-query interval index (htslib regidx semantics: 0-based inclusive)
struct RegionIndex {
    struct Chr {
        std::vector<int32_t> beg, end, maxend, payload;
    };
    std::unordered_map<std::string, Chr> chr;
    void build(int n, const char* const* names, const int32_t* beg0, const int32_t* end0);
    const Chr* find(const std::string& name) const;
    static bool overlap(const Chr* c, int from, int to);
};

struct IngestConf {
    int n_files = 0;
    uint32_t keep_flag[2] = {0, 0};
    int min_global_mq = 0;
    int rq_minq = 0;
    double rq_pct = 0;
    int max_depth = 10000;
    int remove_overlaps = 1;  // 0 none, 1 htslib >= 1.13, 2 <= 1.12
    int n_mm_type = 0, n_mm = 0;
    int min_overhang = 0;
    std::vector<int> libtype;
    const RegionIndex* regidx = nullptr;
    // single cell
    bool sc = false;
    const std::unordered_map<std::string, int>* cb_index = nullptr;  // barcode -> 1-based column
    std::string cb_tag, umi_tag;                                     // empty = not configured
    bool bulk_umi = false;  // pileup_sites(umi_tag = ...): first read per UMI, site, strand and file (src/plp.c:622-646)
    int fixed_cb = 0;  // Smart-seq2 mode: every read of this BAM belongs to this whitelist column (src/sc-plp.c:245-246)
    int64_t target_reads = 1 << 21;                                  // reads per batch over all files
};

struct HostRead {
    raer_read_hdr h;   // cigar_off / sq_off relative to the owning file pool
    int64_t ordinal;   // arrival index among kept reads of the file
    int64_t mate_ord;  // ordinal of the earlier mate or -1
    int32_t keep_end;  // the read stays in the pending window until this position (its own end or its mate's)
    int32_t cb;
    uint32_t umi;
};

// qname -> the earlier mate waiting for its partner (htslib overlap_push's khash). Open addressing with the names in
// an arena: no allocation per read. Entries whose read has left the pileup buffer are dead; find() callers treat them
// as absent and rebuild() drops them.
struct QnameTable {
    struct Ent {
        uint64_t h;
        uint32_t off;
        uint16_t len;
        uint8_t state;  // 0 empty, 1 live, 2 deleted
        bool keep_a;
        int64_t ordinal;
        int32_t tid, end;
    };
    std::vector<Ent> tab;
    std::vector<char> arena;
    size_t live = 0, used = 0;  // used = live + deleted slots
    static uint64_t hash(const char* s, size_t n) {
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
        return h ^ (h >> 29);
    }
    size_t size() const { return live; }
    void clear() { tab.clear(); arena.clear(); live = used = 0; }
    Ent* find(const char* q, size_t n, uint64_t h) {
        if (tab.empty()) return nullptr;
        const size_t mask = tab.size() - 1;
        for (size_t i = h & mask; tab[i].state; i = (i + 1) & mask)
            if (tab[i].state == 1 && tab[i].h == h && tab[i].len == n && !memcmp(arena.data() + tab[i].off, q, n)) return &tab[i];
        return nullptr;
    }
    void erase(Ent* e) { e->state = 2; --live; }
    template <class Dead>
    void rebuild(Dead dead) {
        std::vector<Ent> keep;
        keep.reserve(live);
        for (const Ent& e : tab)
            if (e.state == 1 && !dead(e)) keep.push_back(e);
        size_t cap = 4096;
        while (cap < 4 * keep.size()) cap <<= 1;
        std::vector<char> na;
        na.reserve(arena.size() / 2 + 64);
        tab.assign(cap, Ent{});
        const size_t mask = cap - 1;
        for (Ent e : keep) {
            const uint32_t off = (uint32_t)na.size();
            na.insert(na.end(), arena.data() + e.off, arena.data() + e.off + e.len);
            e.off = off;
            size_t i = e.h & mask;
            while (tab[i].state) i = (i + 1) & mask;
            tab[i] = e;
        }
        arena.swap(na);
        live = used = keep.size();
    }
    // the key must be absent (find() returned nullptr)
    template <class Dead>
    void insert(const char* q, size_t n, uint64_t h, int64_t ordinal, int tid, int end, bool keep_a, Dead dead) {
        if (tab.empty() || (used + 1) * 10 > tab.size() * 7 || arena.size() > 64 * tab.size()) rebuild(dead);
        const size_t mask = tab.size() - 1;
        size_t i = h & mask;
        while (tab[i].state == 1) i = (i + 1) & mask;
        if (tab[i].state == 0) ++used;
        Ent& e = tab[i];
        e.h = h; e.off = (uint32_t)arena.size(); e.len = (uint16_t)n; e.state = 1; e.keep_a = keep_a;
        e.ordinal = ordinal; e.tid = tid; e.end = end;
        arena.insert(arena.end(), q, q + n);
        ++live;
    }
};

struct FilePending {
    BamStream bs;
    std::vector<HostRead> reads;
    size_t front = 0;
    std::vector<uint32_t> cigar;
    std::vector<uint8_t> seq, qual;
    size_t cig_base = 0, sq_base = 0;  // pool offsets already compacted away
    bool eof = false;
    bool failed = false;     // the stream or a record was corrupt / the input is not coordinate sorted: the call must fail
    std::string err;
    int sort_tid = -1, sort_pos = -1;  // last read handed to the pileup (htslib bam_plp_push: max_tid, max_pos)
    bool have_look = false;  // a parsed record waiting in `look`
    HostRead look;
    std::vector<uint32_t> look_cigar;
    const uint8_t* look_seq = nullptr;   // into the stream buffer: valid until the next record is read
    const uint8_t* look_qual = nullptr;
    int look_tid = -1;
    // stream-order state (bam_plp_push emulation)
    int64_t next_ordinal = 0;
    int prev_tid = 0, prev_pos = 0;
    QnameTable olap;
    // ends of the reads pushed on the current contig; entries below the iterator position are dead and are
    // dropped lazily (the count is only needed when a read starts where the previous one did)
    // ends of the pushed reads that can still be in htslib's pileup buffer, smallest on top: the max_depth rule asks how
    // many have not fallen behind a position that never moves backwards, so stale ends are popped when it is asked
    // (a vector that was compacted at every check cost max_depth steps per read at a locus deeper than max_depth)
    std::priority_queue<int32_t, std::vector<int32_t>, std::greater<int32_t>> live_ends;
    size_t live_compact_at = 4096;
    int live_tid = -1;
    int64_t n_aligned_bases = 0;
    // bulk UMI mode: earlier reads of each (strand, UMI) that can still be in the pileup, and per read the
    // positions inside its shadow that no earlier read of its group takes (rare: needs an N in a duplicate)
    struct UmiLive { int pos, end; bool has_n; std::vector<uint32_t> cigar; std::vector<uint8_t> seq; int l_qseq; };
    std::unordered_map<std::string, std::vector<UmiLive>> umi_groups;
    int umi_tid = -1;
    std::unordered_map<int64_t, std::vector<int32_t>> umi_holes;  // ordinal -> positions
};

// Owned, host-resident batch (pinned when a device is present)
struct PackedBatch {
    raer_read_batch b{};
    std::vector<int64_t> file_off;
    raer_read_hdr* hdr = nullptr;
    int32_t* maxend = nullptr;
    uint32_t* cigar = nullptr;
    uint8_t* seq = nullptr;
    uint8_t* qual = nullptr;
    int32_t* pair_b = nullptr;
    int32_t* cb = nullptr;
    uint32_t* umi = nullptr;
    std::vector<int32_t> umi_holes;  // bulk UMI mode: (read index, position) pairs, sorted by read index
    size_t cap_reads = 0, cap_cigar = 0, cap_sq = 0, cap_pairs = 0;
    bool pinned = false;
    bool want_pinned = true;  // false: plain host memory (workers that must not enter the CUDA driver)
    void reserve(size_t reads, size_t cig, size_t sq, bool sc);
    void release();
    ~PackedBatch() { release(); }
};

class Ingest {
public:
    Ingest(const IngestConf& c) : conf_(c) {}
    // jumps: the sorted, disjoint intervals wanted on the region's contig (site / region lists, see BamStream::set_jumps)
    bool open(const char* const* paths, const char* const* indexes, int threads, const char* region,
              const std::vector<std::pair<int32_t, int32_t>>* jumps = nullptr);
    // headers only (names(), lengths(), file_names()): nothing is inflated beyond the header blocks, no batches can be read
    bool open_headers(const char* const* paths);
    const std::vector<std::string>& file_names(int i) const { return files_[i]->bs.names(); }
    // next contig with reads in any file (ascending tid of file 0's header); -1 when done
    int next_contig();
    // pack the next batch of the current contig; false when the contig is exhausted
    bool next_batch(PackedBatch& out);
    const std::vector<std::string>& names() const { return files_[0]->bs.names(); }
    const std::vector<int32_t>& lengths() const { return files_[0]->bs.lengths(); }
    int reg_tid() const { return reg_tid_; }
    int reg_beg() const { return reg_beg_; }
    int reg_end() const { return reg_end_; }
    // a file could not be read to its end, a record was corrupt, or the input is not coordinate sorted: the rows
    // produced so far are not the result of the call (the reference raises Rf_error, src/plp.c:1195)
    bool failed() const;
    std::string error() const;
    double decode_seconds() const;
    double pack_seconds() const { return t_pack_; }
    double parse_seconds() const { return t_parse_; }  // wall time of the parse section of next_batch (all files side by side)
    double wait_seconds() const;
    int64_t records() const;
    int64_t aligned_bases() const;
    int64_t inflated_bytes() const;
    int64_t seeks() const;
    // site / region lists: let every file skip what lies between the intervals of the open region's contig
    void set_jumps(const std::vector<std::pair<int32_t, int32_t>>& intervals);
    std::unordered_map<std::string, uint32_t>& umi_dict() { return umi_dict_; }
    ~Ingest();

private:
    // parse records of file f until it holds `want` pending reads of contig cur_tid_ past the window start
    bool parse_ahead(int f, size_t want);
    bool parse_one(int f);  // parse + prefilter one record into look / pending
    IngestConf conf_;
    std::vector<FilePending*> files_;
    int cur_tid_ = -1;
    int win_beg_ = 0;
    int reg_tid_ = -1, reg_beg_ = 0, reg_end_ = INT32_MAX;
    bool has_region_ = false;
    double t_pack_ = 0;
    double t_parse_ = 0;
    std::unordered_map<std::string, uint32_t> umi_dict_;
};

// true iff every file has a readable .bai (explicit path or "<bam>.bai")
bool indexes_usable(const char* const* paths, const char* const* indexes, int n);
// jump list of one contig of a region index (see Ingest::set_jumps)
std::vector<std::pair<int32_t, int32_t>> jump_intervals(const RegionIndex::Chr* c, int32_t gap = 1 << 16);

// FASTA via .fai, whole contig, case preserved (faidx_fetch_seq64 semantics)
bool fasta_fetch(const char* fa, const std::string& name, std::string& out);
int parse_region(const char* s, int* beg, int* end);  // hts_parse_reg; returns length of the name part or -1
double now_seconds();
int usable_cpus();  // CPUs this process may run on (affinity mask), not the machine's count

}  // namespace raer
#endif
