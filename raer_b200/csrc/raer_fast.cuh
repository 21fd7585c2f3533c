// raer_fast.cuh -- k_pileup_fast: the tile kernel. k_pileup_fast<false>: only the mapq and base-quality filters of check_read_filters are
// active (src/plp.c:568-586, the default FilterParam). k_pileup_fast<true>: also the trim, splice-distance,
// splice-overhang and indel-distance filters (:588-612), which all turn a base into nX and depend on nothing
// but the read and the query offset (trims are two bounds per read, the distance filters are evaluated per base
// only for reads that have a ref-skip / insertion / deletion), and the mismatch filter (:744-760), a per-read MD
// verdict that drops the bases consulting it. Included from raer_kernels.cu; base-quality thresholds above 128
// and tiles deeper than 65535 run k_pileup_tiles.
//
// The reference touches a counter for every aligned base. Here a base that equals the reference and
// passes the quality filter touches nothing:
//   * coverage per (file, strand) comes from +1/-1 events at the ends of each aligned block, written once
//     per read while its record is staged, and one prefix sum per tile;
//   * the 8 query bases of an item (one 32-bit sequence word, one 64-bit quality word) are compared with
//     8 packed reference bases by XOR, and with the quality threshold by byte-parallel arithmetic;
//   * only exceptional bases -- mismatch, quality below the threshold, N -- run scalar code and issue a
//     shared-memory RED; runs of quality 0 (what htslib's mate-overlap rule leaves behind) are taken out
//     of the coverage with one -1/+1 event pair per run;
//   * nRef of a site = coverage - everything explicit (CntPacked::get).
// Counters are 16-bit pairs in 32-bit words (A|C, G|T, other|X): a tile whose coverage exceeds 65535 is
// reported through err_flag (RAER_ERR_LIMIT).
//
// Reads reach a tile through per-tile index lists built by k_tile_count / k_tile_scan / k_tile_fill (every
// read that overlaps the tile, nothing else). Warps work on their own: each pulls chunks of 32 listed reads
// from a per-tile queue, stages their records in its slice of shared memory and spreads the chunk's items over
// its lanes, so there is no CTA barrier inside a phase and no idle lanes for reads that only pass by. The second pass (read-position histograms, strand tallies and first-seen
// order at the sites that carry a variant) is pair parallel: a prefix count of flagged positions gives every
// read its number of (read, flagged position) pairs in O(1), warps pull chunks of 32 listed reads and their
// lanes take one pair each.
// Two CTAs of 512 threads per SM.
#ifndef RAER_FAST_CUH
#define RAER_FAST_CUH

#define FB_THREADS 512
#define FSTAGE_WORDS (32 * FREC + 64)  // per warp: 32 read records, pair prefix sums, read ordinals
#define FREC 13   // words per record (odd: lanes reading one field of different records spread over banks)
// FR_CIG0..+3: CIGAR words, or [0] = offset if n > 4. Second pass: FR_FIRST = f0 | f1 << 16, FR_B2 = f2 | c0 << 16,
// FR_B3 = c1: first flagged rank and number of flagged sites of the first, second and remaining aligned blocks.
enum { FR_POS = 0, FR_LEN, FR_SQ, FR_FLAGS, FR_QSQE, FR_FIRST, FR_RCP, FR_CIG0, FR_B2 = FR_CIG0 + 4, FR_B3 };
#define STF_SIMPLE 0x400u  // one aligned block, nothing but clips around it: query q sits on pos + q - qs
#define STF_SLOWFILT 0x800u  // EXT: the read has a ref-skip / indel (or the overhang over-read): distance filters per base
#define FWMAX 1024         // widest tile; the fixed part of the shared-memory layout is sized for it
#define FRF_FILE_SHIFT 24
#define FWS(W) ((W) + 8)  // words per counter array: W positions + the -1 event one past the end

struct FastSmem {
    uint32_t* cnt;      // [file][strand]{coverage events, A|C<<16, G|T<<16, other|X<<16}[WS]; second-pass slots later
    uint32_t* refn;     // packed reference: nibble i of word k = nt16 code of position t0 - 8 + 8k + i
    uint32_t* covered;  // min_depth <= 0 only: a pileup column exists here
    int* flag_row;      // per flagged site: its first output row
    uint32_t* stage;    // FSTAGE_WORDS per warp
    uint16_t* fpre;     // per position (W + 1): number of flagged sites before it (their ranks are position sorted)
    uint16_t* flag_pidx;  // per flagged site: its tile position
    uint8_t* flag_d;    // per flagged site: decision bits
    uint8_t* dec;       // per position: decision bits
    int WS;
};

__device__ __forceinline__ void red_add(uint32_t* p, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_s(unsigned saddr, uint32_t v) {  // saddr: shared-window byte address
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
// per byte: bit 7 set iff byte < m (m4 = m in every byte, m <= 128: no borrow crosses a byte)
__device__ __forceinline__ uint32_t byte_lt(uint32_t v, uint32_t m4) { return ~(((v | 0x80808080u) - m4) | v) & 0x80808080u; }
__device__ __forceinline__ uint32_t byte_zero(uint32_t v) { return ~(((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u; }
// low nibble of byte j of lo -> nibble j, of hi -> nibble 4 + j
__device__ __forceinline__ uint32_t bytes_to_nibbles(uint32_t lo, uint32_t hi) {
    lo = (lo | (lo >> 4)) & 0x00ff00ffu;
    lo = (lo | (lo >> 8)) & 0xffffu;
    hi = (hi | (hi >> 4)) & 0x00ff00ffu;
    hi = ((hi << 16) | (hi << 8)) & 0xffff0000u;
    return lo | hi;
}
// nt16 code of the upper-cased reference letter; letters outside the nt16 alphabet -> 0 (never equal to a
// read base except the never-produced '=' code), beyond the contig -> N
__device__ __forceinline__ int ref_nibble(const PlpParams& P, int p) {
    if (!P.ref || p < 0 || p >= P.ref_len) return 15;
    const int u = d_upper((unsigned char)__ldg(P.ref + p));
    if (u < 'A' || u > 'Y') return 0;
    //                                   A  B  C  D  E F G  H  I J K   L M N   O P Q R S T U V W X Y
    const unsigned char lut[25] = {1, 14, 2, 13, 0, 0, 4, 11, 0, 0, 12, 0, 3, 15, 0, 0, 0, 5, 6, 8, 0, 7, 9, 0, 10};
    return lut[u - 'A'];
}
__device__ __forceinline__ void cover_range(uint32_t* cov, int lo, int hi) {  // tile-relative [lo, hi), lo < hi
    const int w0 = lo >> 5, w1 = (hi - 1) >> 5;
    for (int w = w0; w <= w1; ++w) {
        uint32_t m = 0xffffffffu;
        if (w == w0) m &= 0xffffffffu << (lo & 31);
        if (w == w1) m &= 0xffffffffu >> (31 - ((hi - 1) & 31));
        atomicOr(&cov[w], m);
    }
}

// bulk UMI mode: true iff position p of read r stays counted although it lies below the read's hdr.aux
__device__ __forceinline__ bool umi_is_hole(const PlpParams& P, int r, int p) {
    int lo = 0, hi = P.n_umi_holes;
    while (lo < hi) {  // first pair with (read, pos) >= (r, p)
        const int mid = (lo + hi) >> 1;
        const int mr = __ldg(P.umi_holes + 2 * mid), mp = __ldg(P.umi_holes + 2 * mid + 1);
        if (mr < r || (mr == r && mp < p)) lo = mid + 1; else hi = mid;
    }
    return lo < P.n_umi_holes && __ldg(P.umi_holes + 2 * lo) == r && __ldg(P.umi_holes + 2 * lo + 1) == p;
}

struct FastCtx {  // per-kernel constants of the item walk
    int t0, tw, minbq;
    uint32_t mb4, zall;
    int round0, ns, HW, SW;  // second pass: flagged ranks [round0, round0 + ns) own slots of SW words
};

// ---- phase 1: one aligned block's share of an item. Bases j in [jlo, jhi) sit on tile positions rel + j.
// NM/LT/Z: nibble-bit masks (bit 4j) of read base N, quality below the threshold, quality 0.
// mmf: RAER_RB_MMFAIL / RAER_RB_MMERR bits of the read when the mismatch filter is configured, else 0
__device__ __forceinline__ void fast_block_count(const FastSmem& S, const FastCtx& X, unsigned cbs, uint32_t rw, uint32_t NM,
                                                 uint32_t LT, uint32_t Z, bool mapq_fail, int cxor, int rel, int jlo, int jhi,
                                                 int mmf = 0, int* err_flag = nullptr, uint32_t skip = 0) {
    // bulk UMI mode: skip = bases of a read with UMI holes that an earlier read takes (in the coverage: dropped here)
    const int lo = max(max(jlo, -rel), 0), hi = min(min(jhi, X.tw - rel), 8);
    if (lo >= hi) return;
    const uint32_t V = (0x11111111u >> ((8 - (hi - lo)) << 2)) << (lo << 2);
    const int a = rel + 8;  // >= 1 here: index of the item's first base in the packed reference
    const uint32_t refw = __funnelshift_r(S.refn[a >> 3], S.refn[(a >> 3) + 1], (a & 7) << 2);
    uint32_t D, XM, MM;
    if (mapq_fail) {  // every base of the read is nX (src/plp.c:574); read base N is counted nowhere (:718)
        D = (NM | skip) & V;
        XM = V & ~D;
        MM = 0;
    } else {
        D = (Z | NM | skip) & V;   // quality 0 below the threshold is dropped silently (src/plp.c:580-586), so is N
        XM = LT & V & ~D;   // nX
        MM = nib_nonzero(rw ^ refw) & V & ~(LT | D);
        if (mmf) {
            // src/plp.c:750: a base that passed every filter consults the read's MD verdict if the read is inverted or
            // the base differs from the reference; a failing read drops such bases silently, an inconsistent MD tag is
            // a hard error
            const uint32_t consulted = cxor ? (V & ~(D | XM)) : MM;
            if (consulted) {
                if (mmf & RAER_RB_MMERR) atomicExch(err_flag, RAER_ERR_MD);
                if (mmf & RAER_RB_MMFAIL) { D |= consulted; MM &= ~consulted; }
            }
        }
    }
    const unsigned base = cbs + ((unsigned)rel << 2);  // shared address of this item's first coverage word
    // runs of dropped bases leave the coverage: -1 where a run starts, +1 after its end
    uint32_t rs = D & ~(D << 4), re = D & ~(D >> 4);
    while (rs) {
        const int b1 = __ffs(rs) - 1, b2 = __ffs(re) - 1;
        rs &= rs - 1;
        re &= re - 1;
        red_add_s(base + b1, 0xffffffffu);  // b = 4j: byte offset of word j
        red_add_s(base + b2 + 4, 1u);
    }
    const unsigned xb = base + (unsigned)(3 * S.WS) * 4;
    while (XM) {
        const int b = __ffs(XM) - 1;
        XM &= XM - 1;
        red_add_s(xb + b, 0x10000u);
    }
    while (MM) {
        const int b = __ffs(MM) - 1;
        MM &= MM - 1;
        const int code = (rw >> b) & 15;
        // A=1 C=2 G=4 T=8 -> 0..3; complement on the minus strand = xor 3; any other code -> other
        int cls = ((code >> 1) - (code >> 3)) ^ cxor;
        if (__popc(code) != 1) cls = RAER_CLS_OTHER;
        red_add_s(base + (unsigned)((1 + (cls >> 1)) * S.WS) * 4 + b, 1u << ((cls & 1) << 4));
    }
}

__device__ __noinline__ void fast_block_count_call(const FastSmem& S, const FastCtx& X, unsigned cbs, uint32_t rw, uint32_t NM,
                                                   uint32_t LT, uint32_t Z, bool mapq_fail, int cxor, int rel, int jlo, int jhi,
                                                   int mmf, int* err_flag, uint32_t skip) {
    fast_block_count(S, X, cbs, rw, NM, LT, Z, mapq_fail, cxor, rel, jlo, jhi, mmf, err_flag, skip);
}

// EXT: nibble-bit mask of the item's bases that the read-geometry filters turn into nX (check_read_filters after the
// base-quality test, src/plp.c:588-612)
__device__ __forceinline__ uint32_t fast_filter_mask(const PlpParams& P, const uint32_t* rec, int q0, int q1) {
    // trims: nX iff q < qL or q >= qR (bounds staged per read)
    const int qL = (int)(rec[FR_RCP] & 0xffff), qR = (int)(rec[FR_RCP] >> 16);
    const int nl = min(max(qL - q0, 0), 8), nr = min(max(qR - q0, 0), 8);
    uint32_t tx = (nl ? (0x11111111u >> ((8 - nl) << 2)) : 0u) | (nr < 8 ? (0x11111111u << (nr << 2)) : 0u);
    const unsigned fl = rec[FR_FLAGS];
    if (fl & STF_SLOWFILT) {
        ReadCtx c;
        c.n_cigar = (int)(rec[FR_LEN] >> 16);
        c.cig = c.n_cigar <= 4 ? rec + FR_CIG0 : P.cigar + rec[FR_CIG0];
        c.bits = (int)(fl & 0xff);
        // One pass over the CIGAR decides whether any base of the item can be hit at all: a ref-skip, insertion
        // or deletion within the configured distance of [q0, q1), or an aligned block shorter than the overhang
        // that touches the item. Most items of a spliced read are far from its junctions.
        bool near = (c.bits & RAER_RB_OHANG_N) && P.min_overhang;
        const int dmax = max(P.splice_dist, P.indel_dist);
        for (int k = 0, ip = 0; k < c.n_cigar && !near; ++k) {
            const uint32_t w = c.cig[k];
            const int op = cg_op(w), cl = cg_len(w);
            if (op == OP_N || op == OP_D) near = ip >= q0 - dmax && ip <= q1 - 1 + dmax;
            else if (op == OP_I) near = ip + cl >= q0 - dmax && ip <= q1 - 1 + dmax;
            else if (op == OP_M && P.min_overhang && cl < P.min_overhang) near = ip + cl >= q0 && ip <= q1 - 1;
            if (cg_type(op) & 1) ip += cl;
        }
        for (int q = q0; near && q < q1; ++q) {
            bool hit = P.splice_dist && d_dist_to_splice(c, q, P.splice_dist) >= 0;
            if (!hit && P.min_overhang) hit = d_check_splice_overhang(c, q, P.min_overhang) > 0;
            if (!hit && P.indel_dist) hit = d_dist_to_indel(c, q, P.indel_dist) >= 0;
            if (hit) tx |= 1u << ((q - q0) << 2);
        }
    }
    return tx;
}

template <bool EXT>
__device__ __forceinline__ void fast_item(const PlpParams& P, const FastSmem& S, const FastCtx& X, const uint32_t* rec, int q0,
                                          uint32_t sw, uint2 qw) {
    const int pos = (int)rec[FR_POS];
    const int lq = (int)(rec[FR_LEN] & 0xffff), n = (int)(rec[FR_LEN] >> 16);
    const unsigned fl = rec[FR_FLAGS];
    const int q1 = min(q0 + 8, lq);
    const bool mapq_fail = (fl & STF_MAPQ_FAIL) != 0;
    const int s = fl & RAER_RB_INVERT;
    const unsigned cbs = (unsigned)__cvta_generic_to_shared(S.cnt + (size_t)((((fl >> FRF_FILE_SHIFT) & 0x3f) * 2 + s) * 4) * S.WS);
    const uint32_t rw = nib_swap(sw);
    const uint32_t NM = rw & (rw >> 1) & (rw >> 2) & (rw >> 3) & 0x11111111u;
    // nibble j: bit 0 = quality below the threshold, bit 1 = quality 0
    const uint32_t Q = bytes_to_nibbles((byte_lt(qw.x, X.mb4) >> 7) | (byte_zero(qw.x) >> 6),
                                        (byte_lt(qw.y, X.mb4) >> 7) | (byte_zero(qw.y) >> 6));
    uint32_t LT = Q & 0x11111111u;
    const uint32_t Z = (Q >> 1) & 0x11111111u & X.zall;
    // a base that fails a later filter is nX exactly like one below the quality threshold (and quality 0 below the
    // threshold has been dropped before those filters are consulted): same mask from here on
    if (EXT) LT |= fast_filter_mask(P, rec, q0, q1) & ~Z;
    const uint32_t* cig = n <= 4 ? rec + FR_CIG0 : P.cigar + rec[FR_CIG0];
    const bool simple = (fl & STF_SIMPLE) != 0;
    int x = pos, y = 0, k = 0;
    bool has = false;
    int rel = 0, jlo = 0, jhi = 0;
    // next aligned block that intersects the item (reads that walk their CIGAR)
    auto next_block = [&]() {
        has = false;
        while (k < n && y < q1 && !has) {
            const uint32_t w = cig[k++];
            const int op = cg_op(w), len = cg_len(w);
            if ((0x181 >> op) & 1) {  // M = X
                if (y + len > q0) {
                    has = true;
                    rel = x + (q0 - y) - X.t0;
                    jlo = y - q0;
                    jhi = min(y + len, q1) - q0;
                }
                x += len;
                y += len;
            } else if ((0xC >> op) & 1) {  // D N
                x += len;
            } else if ((0x12 >> op) & 1) {  // I S
                y += len;
            }
        }
    };
    if (simple) {  // the aligned block is query [qs, qe)
        const int qs = (int)(rec[FR_QSQE] & 0xffff), qe = (int)(rec[FR_QSQE] >> 16);
        has = qe > q0 && qs < q1;
        rel = pos + (q0 - qs) - X.t0;
        jlo = qs - q0;
        jhi = min(qe, q1) - q0;
    } else {
        next_block();
    }
    // one inlined copy of the block code for every kind of read ...
    const int mmf = (EXT && (P.n_mm_type > 0 || P.n_mm > 0)) ? (int)(fl & (RAER_RB_MMFAIL | RAER_RB_MMERR)) : 0;
    const bool umi = EXT && P.umi_mode;
    const bool holes = umi && (fl & RAER_RB_UMI_HOLES);
    const int shadow = umi ? (int)rec[FR_B2] : 0;  // positions below belong to an earlier read of the same UMI
    auto umi_skip = [&]() -> uint32_t {            // reads with holes: per base, minus the positions that stay theirs
        uint32_t m = 0;
        for (int j = max(jlo, 0); j < min(jhi, 8); ++j) {
            const int p = X.t0 + rel + j;
            if (p < shadow && !umi_is_hole(P, (int)rec[FR_B3], p)) m |= 1u << (j << 2);
        }
        return m;
    };
    const int plo = (umi && !holes) ? min(max(shadow - X.t0, 0), X.tw) : 0;
    // a read without holes is simply absent below plo (its coverage events were clipped there too)
    if (has) fast_block_count(S, X, cbs, rw, NM, LT, Z, mapq_fail, s ? 3 : 0, rel, max(jlo, plo - rel), jhi, mmf, P.err_flag, holes ? umi_skip() : 0u);
    // ... and a call for the rare item that straddles two aligned blocks
    while (!simple && has && k < n && y < q1) {
        next_block();
        if (has)
            fast_block_count_call(S, X, cbs, rw, NM, LT, Z, mapq_fail, s ? 3 : 0, rel, max(jlo, plo - rel), jhi, mmf, P.err_flag,
                                  holes ? umi_skip() : 0u);
    }
}

// Phase 1: stage one read record (thread per read) and write the read's coverage events.
template <bool NEEDCOV, bool EXT>
__device__ __forceinline__ int fast_stage(const PlpParams& P, const FastSmem& S, const FastCtx& X, int r, int f, int t1,
                                          uint32_t* rec) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(P.hdr + r));
    const int4 b = __ldg(reinterpret_cast<const int4*>(P.hdr + r) + 1);
    const int flag = a.z & 0xffff, mapq = (a.z >> 16) & 0xff, bits = (a.z >> 24) & 0xff;
    const int lq = a.w & 0xffff, n = (a.w >> 16) & 0xffff;
    const int s = bits & RAER_RB_INVERT;
    unsigned fl = (unsigned)bits | ((unsigned)((flag >> 4) & 1) << 16) | ((unsigned)f << FRF_FILE_SHIFT);
    if (mapq < P.min_mapq[f]) fl |= STF_MAPQ_FAIL;
    const uint32_t* cg = P.cigar + (unsigned)b.x;
    uint32_t* dl = S.cnt + (size_t)((f * 2 + s) * 4) * S.WS;
    // bulk UMI mode: positions below `shadow` belong to an earlier read of the same UMI (src/plp.c:738): this read is
    // absent there. Without holes the read is simply clipped; with holes the items sort it out base by base.
    int shadow = 0;
    if (EXT && P.umi_mode) {
        shadow = (unsigned)b.w == 0xffffffffu ? 0x7fffffff : b.w;
        rec[FR_B2] = (uint32_t)shadow;
        rec[FR_B3] = (uint32_t)r;
        if (bits & RAER_RB_UMI_HOLES) shadow = 0;
    }
    int x = a.x, y = 0;
    int qs = 0, tail = 0, nblk = 0, other = 0;
    int qlo = 0x7fffffff, qhi = 0;  // query bases whose positions fall inside the tile
    bool lead = true;
    // One operation that consumes as many reference as query bases is an aligned block (M, = or X): when only the
    // mapq / base-quality filters are active nothing asks which of the three, so the CIGAR word is not fetched.
    const bool one_block = !EXT && n == 1 && a.y - a.x == lq && lq > 0;
    for (int k = 0; k < n; ++k) {
        const uint32_t w = one_block ? ((uint32_t)lq << 4) : __ldg(cg + k);
        if (k < 4) rec[FR_CIG0 + k] = w;
        const int op = cg_op(w), len = cg_len(w);
        // query_start / query_end (src/plp_utils.c:52-104): soft clips at either end, hard clips skipped
        if (op == OP_S) { if (lead) qs += len; tail += len; }
        else if (op != OP_H) { lead = false; tail = 0; if ((0x181 >> op) & 1) ++nblk; else other = 1; }
        if (NEEDCOV && (cg_type(op) & 2)) {
            const int lo = max(x, X.t0), h2 = min(x + len, t1);
            if (lo < h2) cover_range(S.covered, lo - X.t0, h2 - X.t0);
        }
        if ((0x181 >> op) & 1) {
            // coverage events of the aligned block, clipped to the tile (and to the bases the read really has)
            const int lo = max(max(x, X.t0), shadow), h2 = min(x + min(len, lq - y), t1);
            if (lo < h2) {
                red_add(dl + (lo - X.t0), 1u);
                red_add(dl + (h2 - X.t0), 0xffffffffu);
                qlo = min(qlo, y + (lo - x));
                qhi = max(qhi, y + (h2 - x));
            }
            x += len;
            y += len;
        } else if ((0xC >> op) & 1) {
            x += len;
        } else if ((0x12 >> op) & 1) {
            y += len;
        }
    }
    if (nblk == 1 && !other && qs + (x - a.x) == lq - tail) fl |= STF_SIMPLE;
    if (EXT) {
        // check_variant_fpos / check_variant_pos (src/plp_utils.c:137-193): nX iff q < qL or q >= qR
        const int qe = lq - tail, rev = (flag >> 4) & 1;
        int qL = 0, qR = 65535;
        if (P.ftrim_5p > 0 || P.ftrim_3p > 0) {
            const int ql = qe - qs;
            if (ql <= 0) qL = 65535;
            const int d5 = (int)floor(P.ftrim_5p * ql), d3 = (int)ceil(P.ftrim_3p * ql);
            qL = max(qL, (rev ? d3 : d5) + qs);
            qR = min(qR, qe - (rev ? d5 : d3));
        }
        if (P.trim_5p > 0 || P.trim_3p > 0) {
            qL = max(qL, (rev ? P.trim_3p : P.trim_5p) + qs);
            qR = min(qR, qe - (rev ? P.trim_5p : P.trim_3p));
        }
        rec[FR_RCP] = (uint32_t)min(max(qL, 0), 65535) | ((uint32_t)min(max(qR, 0), 65535) << 16);
        // the distance filters can only fire on reads with a ref-skip / insertion / deletion, or through the
        // over-read of check_splice_overhang (RAER_RB_OHANG_N)
        if ((P.splice_dist || P.min_overhang || P.indel_dist) && (!(fl & STF_SIMPLE) || (bits & RAER_RB_OHANG_N))) fl |= STF_SLOWFILT;
    }
    if (n > 4) rec[FR_CIG0] = (uint32_t)b.x;
    rec[FR_POS] = (uint32_t)a.x;
    rec[FR_LEN] = (uint32_t)lq | ((uint32_t)n << 16);
    rec[FR_SQ] = (uint32_t)b.y;
    rec[FR_FLAGS] = fl;
    rec[FR_QSQE] = (uint32_t)qs | ((uint32_t)(lq - tail) << 16);
    // items (8 query bases each) that can touch the tile: a read that overlaps two tiles, or a spliced read that
    // only passes through, does not fetch the rest of its bases here
    const int ilo = qlo >> 3;
    rec[FR_FIRST] = (uint32_t)ilo;
    return qhi > qlo ? ((qhi - 1) >> 3) - ilo + 1 : 0;
}

// Phase 1 for one chunk of 32 listed reads, handled by one warp on its own (no CTA barrier): stage the records,
// then the lanes take the chunk's items, two in flight per lane.
template <bool EXT>
__device__ __forceinline__ void fast_count_chunk(const PlpParams& P, const FastSmem& S, const FastCtx& X, int tile, int li0,
                                                 int l1, int t1, int lane, uint32_t* st, bool need_cov) {
    const int F = P.n_files;
    const int li = li0 + lane;
    uint32_t* rec = st + lane * FREC;
    int nq = 0;
    if (li < l1) {
        int f = 0;
        while (f + 1 < F && li >= P.tile_off[tile * F + f + 1]) ++f;
        const int r = __ldg(P.tile_list + li);
        nq = need_cov ? fast_stage<true, EXT>(P, S, X, r, f, t1, rec) : fast_stage<false, EXT>(P, S, X, r, f, t1, rec);
    }
    // items of the chunk in read order: inclusive prefix of the per-read item counts
    uint32_t* pref = st + 32 * FREC;
    int inc = nq;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += v;
    }
    pref[lane] = (uint32_t)inc;
    const int tot = __shfl_sync(FULL, inc, 31);
    __syncwarp();
    // Software pipeline: the loads of the next item are issued before the current one is processed, so one item per
    // lane is always in flight while another is being worked on.
    const uint32_t* recN = st;
    int qN = 0;
    uint32_t swN = 0;
    uint2 qwN = make_uint2(0, 0);
    bool okN = false;
    auto fetch = [&](int it) {
        okN = it < tot;
        if (okN) {
            int lo = 0, hh = 31;  // smallest i with pref[i] > it
#pragma unroll
            for (int step = 0; step < 5; ++step) {
                const int mid = (lo + hh) >> 1;
                if ((int)pref[mid] > it) hh = mid; else lo = mid + 1;
            }
            recN = st + lo * FREC;
            qN = ((int)recN[FR_FIRST] + it - (lo ? (int)pref[lo - 1] : 0)) << 3;
            const uint32_t sq = recN[FR_SQ];
            swN = __ldg(reinterpret_cast<const uint32_t*>(P.seq + sq + (qN >> 1)));
            qwN = __ldg(reinterpret_cast<const uint2*>(P.qual + 2ull * sq + qN));
        }
    };
    fetch(lane);
    for (int itb = 0; itb < tot; itb += 32) {
        // The trip count is warp-uniform, so the warp can be re-joined here: without it the lanes that took
        // different paths through the previous item keep running as separate sub-warps.
        __syncwarp();
        const uint32_t* rec = recN;
        const int q0 = qN;
        const uint32_t sw = swN;
        const uint2 qw = qwN;
        const bool ok = okN;
        fetch(itb + 32 + lane);
        if (ok) fast_item<EXT>(P, S, X, rec, q0, sw, qw);
    }
    __syncwarp();
}

// ---- second pass: one (read, flagged position) pair per lane. Called by all 32 lanes (active = this lane has a
// pair); the lanes are re-joined after the position lookup and after the loads, so single-block reads and reads
// that walk a CIGAR share the rest of the code.
template <bool EXT>
__device__ __forceinline__ void fast_pair(const PlpParams& P, const FastSmem& S, const FastCtx& X, const TileSmem& T2,
                                          const uint32_t* rec, int rank, int ordinal, bool active) {
    const int pos = (int)rec[FR_POS];
    const int lq = (int)(rec[FR_LEN] & 0xffff), n = (int)(rec[FR_LEN] >> 16);
    const unsigned fl = rec[FR_FLAGS];
    const uint32_t sq = rec[FR_SQ];
    const int qs = (int)(rec[FR_QSQE] & 0xffff), qe = (int)(rec[FR_QSQE] >> 16);
    const int pidx = active ? (int)S.flag_pidx[rank] : 0;
    const int p = X.t0 + pidx;
    bool valid = active;
    int q = 0;
    if (valid) {
        if (fl & STF_SIMPLE) {
            q = qs + (p - pos);
            valid = q < qe;
        } else {
            const uint32_t* cig = n <= 4 ? rec + FR_CIG0 : P.cigar + rec[FR_CIG0];
            int x = pos, y = 0, k = 0, op = -1, len = 0;
            for (; k < n; ++k) {
                const uint32_t w = cig[k];
                op = cg_op(w);
                len = cg_len(w);
                if ((0x18D >> op) & 1) {  // M D N = X
                    if (p < x + len) break;
                    x += len;
                    if ((0x181 >> op) & 1) y += len;
                } else if ((0x12 >> op) & 1) {
                    y += len;
                }
            }
            // past the end of the read, or a deletion / ref-skip at p (filtered silently, src/plp.c:571)
            valid = k < n && !((0xC >> op) & 1);
            q = y + (p - x);
        }
        valid = valid && q < lq;
    }
    __syncwarp();
    int code = 15, bq = 0;
    if (valid) {
        code = seq_code(P.seq + sq, q);
        bq = P.qual[2ull * sq + q];
    }
    // read base N (src/plp.c:718); nX or dropped bases are not part of the statistics
    valid = valid && code != 15 && bq >= X.minbq;
    if (EXT && P.umi_mode && valid) {  // an earlier read of the same UMI takes this site (src/plp.c:738)
        const int r = ordinal + (int)P.file_off[(fl >> FRF_FILE_SHIFT) & 0x3f];
        const unsigned sh = __ldg(&P.hdr[r].aux);
        const int shadow = sh == 0xffffffffu ? 0x7fffffff : (int)sh;
        if (p < shadow && !((fl & RAER_RB_UMI_HOLES) && umi_is_hole(P, r, p))) valid = false;
    }
    if (EXT && valid) {  // the remaining filters of check_read_filters, as k_pileup_tiles evaluates them
        ReadCtx c;
        c.n_cigar = n;
        c.cig = n <= 4 ? rec + FR_CIG0 : P.cigar + rec[FR_CIG0];
        c.bits = (int)(fl & 0xff);
        c.l_qseq = lq;
        c.qs = qs;
        c.qe = qe;
        c.rev = (fl >> 16) & 1;
        c.d5f = c.d3f = 0;
        c.ftrim_all = 0;
        if (P.ftrim_5p > 0 || P.ftrim_3p > 0) {
            const int ql = qe - qs;
            if (ql <= 0) c.ftrim_all = 1;
            c.d5f = (int)floor(P.ftrim_5p * ql);
            c.d3f = (int)ceil(P.ftrim_3p * ql);
        }
        valid = base_verdict(P, c, 0, q, bq) == 0;
    }
    __syncwarp();
    if (valid) {
        const int a = pidx + 8;
        const int rcode = (S.refn[a >> 3] >> ((a & 7) << 2)) & 15;
        if (EXT && (P.n_mm_type > 0 || P.n_mm > 0) && (fl & (RAER_RB_MMFAIL | RAER_RB_MMERR)) &&
            ((fl & RAER_RB_INVERT) || !((code == rcode) & (code != 0)))) {  // src/plp.c:750, as in phase 1
            if (fl & RAER_RB_MMERR) atomicExch(P.err_flag, RAER_ERR_MD);
            return;
        }
        ReadCtx c;
        c.l_qseq = lq;
        c.qs = qs;
        c.qe = qe;
        c.rev = (fl >> 16) & 1;
        c.invert = fl & RAER_RB_INVERT;
        pass2_accumulate_core(P, T2, c, (int)((fl >> FRF_FILE_SHIFT) & 0x3f), c.invert ? 1 : 0, ordinal, rank - X.round0, q, code,
                              (code == rcode) & (code != 0), rec[FR_RCP]);
    }
}

// One chunk of 32 listed reads, handled by one warp: stage the reads that cover a flagged site of this round,
// prefix-sum their pair counts, lanes take pairs.
template <bool EXT>
__device__ __forceinline__ void fast_pair_chunk(const PlpParams& P, const FastSmem& S, const FastCtx& X, const TileSmem& T2,
                                                int tile, int li0, int l1, int lane, uint32_t* st) {
    uint32_t* pref = st + 32 * FREC;  // inclusive prefix of pairs per read
    const int F = P.n_files;
    const int li = li0 + lane;
    uint32_t* rec = st + lane * FREC;
    int np = 0, r = 0, f = 0;
    if (li < l1) {
        while (f + 1 < F && li >= P.tile_off[tile * F + f + 1]) ++f;
        r = __ldg(P.tile_list + li);
        const int4 a = __ldg(reinterpret_cast<const int4*>(P.hdr + r));
        // a read below the mapq threshold only produces nX: nothing to accumulate (src/plp.c:574)
        if (((a.z >> 16) & 0xff) >= P.min_mapq[f]) {
            const int4 b = __ldg(reinterpret_cast<const int4*>(P.hdr + r) + 1);
            const int flag = a.z & 0xffff, bits = (a.z >> 24) & 0xff;
            const int lq = a.w & 0xffff, n = (a.w >> 16) & 0xffff;
            const uint32_t* cg = P.cigar + (unsigned)b.x;
            const int r0 = X.round0, r1 = X.round0 + X.ns;
            // Flagged sites are counted per aligned block (prefix counts, O(1) each): the first two blocks exactly,
            // everything from the third block to the end of the read as one span. A spliced read does not pay for
            // the flagged sites inside its introns.
            int qs = 0, tail = 0, nblk = 0, other = 0, rl = 0, x = a.x;
            int fb[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
            bool lead = true;
            const bool one_block = !EXT && n == 1 && a.y - a.x == lq && lq > 0;  // see fast_stage
            for (int k = 0; k < n; ++k) {
                const uint32_t w = one_block ? ((uint32_t)lq << 4) : __ldg(cg + k);
                if (k < 4) rec[FR_CIG0 + k] = w;
                const int op = cg_op(w), len = cg_len(w);
                // query_start / query_end (src/plp_utils.c:52-104): soft clips at either end, hard clips skipped
                if (op == OP_S) { if (lead) qs += len; tail += len; }
                else if (op != OP_H) { lead = false; tail = 0; if (!((0x181 >> op) & 1)) other = 1; }
                if ((0x181 >> op) & 1) {
                    if (nblk < 3) {
                        const int lo = min(max(x - X.t0, 0), X.tw);
                        const int hi = min(max((nblk < 2 ? x + len : a.y) - X.t0, 0), X.tw);
                        fb[nblk] = max((int)S.fpre[lo], r0);
                        cb[nblk] = max(min((int)S.fpre[hi], r1) - fb[nblk], 0);
                    }
                    ++nblk;
                    rl += len;
                    x += len;
                } else if ((0xC >> op) & 1) {
                    x += len;
                }
            }
            np = cb[0] + cb[1] + cb[2];
            if (np > 0) {
                const int qe = lq - tail;
                unsigned fl = (unsigned)bits | ((unsigned)((flag >> 4) & 1) << 16) | ((unsigned)f << FRF_FILE_SHIFT);
                if (nblk == 1 && !other && qs + rl == qe) fl |= STF_SIMPLE;
                if (n > 4) rec[FR_CIG0] = (uint32_t)b.x;
                rec[FR_POS] = (uint32_t)a.x;
                rec[FR_LEN] = (uint32_t)lq | ((uint32_t)n << 16);
                rec[FR_SQ] = (uint32_t)b.y;
                rec[FR_FLAGS] = fl;
                rec[FR_QSQE] = (uint32_t)qs | ((uint32_t)qe << 16);
                rec[FR_FIRST] = (uint32_t)fb[0] | ((uint32_t)fb[1] << 16);
                rec[FR_B2] = (uint32_t)fb[2] | ((uint32_t)cb[0] << 16);
                rec[FR_B3] = (uint32_t)cb[1];
                // get_relative_position divides by alen + 1 = qe - qs + 1 (src/plp_utils.c:401-412): reciprocal once per read
                const int rb = qe - qs + 1;
                rec[FR_RCP] = rb > 1 ? 0xffffffffu / (uint32_t)rb + 1u : 0u;  // ceil(2^32 / rb) without a 64-bit divide
            }
        }
    }
    int inc = np;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += v;
    }
    pref[lane] = (uint32_t)inc;
    // ordinal of the read inside its file: what "first seen" means in BAM order
    pref[32 + lane] = (uint32_t)(r - (int)P.file_off[f]);
    const int tot = __shfl_sync(FULL, inc, 31);
    __syncwarp();
    for (int itb = 0; itb < tot; itb += 32) {
        __syncwarp();
        const int it = itb + lane;
        int lo = 0, hh = 31;  // smallest i with pref[i] > it
#pragma unroll
        for (int step = 0; step < 5; ++step) {
            const int mid = (lo + hh) >> 1;
            if ((int)pref[mid] > it) hh = mid; else lo = mid + 1;
        }
        const int i = lo;
        const int firstp = i ? (int)pref[i - 1] : 0;
        const uint32_t* ri = st + i * FREC;
        // pair j of the read -> flagged rank, through the per-block counts
        const int j = it - firstp, c0 = (int)(ri[FR_B2] >> 16), c1 = (int)ri[FR_B3];
        const int rank = j < c0 ? (int)(ri[FR_FIRST] & 0xffff) + j
                                : (j < c0 + c1 ? (int)(ri[FR_FIRST] >> 16) + (j - c0) : (int)(ri[FR_B2] & 0xffff) + (j - c0 - c1));
        fast_pair<EXT>(P, S, X, T2, ri, rank, (int)pref[32 + i], it < tot);
    }
    __syncwarp();
}

// rpbz / vdb / sor and the ALT order codes of one second-pass slot (shared by both tile kernels)
__device__ void slot_stats(const PlpParams& P, const uint32_t* slot, int HW, int rowp, int rowm, int growmask, int pos) {
    const double sor = d_calc_sor((int)slot[HW], (int)slot[HW + 1], (int)slot[HW + 2], (int)slot[HW + 3]);
    for (int s = 0; s < 2; ++s) {
        const long long row = s == 0 ? rowp : rowm;
        if (row < 0 && !((growmask >> s) & 1)) continue;
        if (row >= 0) {
            const HistView hr = {slot, (s * 2) * RAER_NHIST, HW != 4 * RAER_NHIST};
            const HistView ha = {slot, (s * 2 + 1) * RAER_NHIST, HW != 4 * RAER_NHIST};
            P.row_stats[row * 3 + 0] = d_calc_mwu_z(hr, ha);
            P.row_stats[row * 3 + 1] = d_calc_vdb(ha);
            P.row_stats[row * 3 + 2] = sor;
        }
        for (int f = 0; f < P.n_files; ++f) {
            const uint32_t* fs = slot + HW + 4 + (f * 2 + s) * RAER_SLOT_FS;
            // order the (at most 5) variant keys by first-seen ordinal: 0..3 = ACGT, 4 = IUPAC key
            unsigned ord[5] = {fs[0], fs[1], fs[2], fs[3], fs[5]};
            unsigned code = 0;
            int nins = 0;
            unsigned third = 0;
            for (int k = 0; k < 5; ++k) {
                int best = -1;
                for (int b = 0; b < 5; ++b)
                    if (ord[b] != 0xffffffffu && (best < 0 || ord[b] < ord[best])) best = b;
                if (best < 0) break;
                code |= (unsigned)best << (3 * nins);
                if (nins == 2) third = ord[best];
                ord[best] = 0xffffffffu;
                ++nins;
            }
            // kh_put resizes 4 -> 8 buckets when a put arrives with 3 keys occupied (khash.h)
            const bool grows = nins >= 3 && fs[4] > third + 1u;
            if (row < 0) {
                // no row carries the flag: the host gets the first such position of the (file, strand)
                if (grows) atomicMax(&P.grow[f * 2 + s], 0x7fffffff - pos);
                continue;
            }
            unsigned alt = P.row_alt[row * P.n_files + f] & (RAER_ALT_MASK | RAER_ALT_OTHER);
            alt |= code << RAER_ALT_ORDER_SHIFT;
            alt |= (unsigned)nins << RAER_ALT_NINS_SHIFT;
            if (grows) alt |= RAER_ALT_GROW;
            if (fs[5] != 0xffffffffu) {
                alt |= (fs[6] & 0xfu) << RAER_ALT_OCODE_SHIFT;
                if (fs[6] != fs[7]) alt |= RAER_ALT_OAMBIG;
            }
            P.row_alt[row * P.n_files + f] = alt;
        }
    }
}

// development aid: thread 0 adds the cycles since the previous mark to phase_clk[k]
#define PHASE_MARK(k)                                                                   \
    do {                                                                                \
        if (P.phase_clk && threadIdx.x == 0) {                                          \
            const long long now_ = clock64();                                           \
            atomicAdd(&P.phase_clk[k], (unsigned long long)(now_ - ph_t));              \
            ph_t = now_;                                                                \
        }                                                                               \
    } while (0)
// ... and lane 0 of every warp the cycles it waited at the barrier that ends a warp-autonomous phase
#define PHASE_BARRIER(k)                                                                \
    do {                                                                                \
        if (P.phase_clk) {                                                              \
            const long long a_ = clock64();                                             \
            __syncthreads();                                                            \
            if (lane == 0) atomicAdd(&P.phase_clk[k], (unsigned long long)(clock64() - a_)); \
        } else {                                                                        \
            __syncthreads();                                                            \
        }                                                                               \
    } while (0)

template <bool EXT>
__global__ void __launch_bounds__(FB_THREADS, 2) k_pileup_fast(PlpParams P) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int s_scan[40];
    __shared__ int s_tile, s_nflag, s_rowbase, s_deep, s_chunk;
    const int W = P.W, F = P.n_files;
    FastSmem S;
    S.WS = FWS(W);
    const int cnt_words = F * 8 * S.WS;
    // everything but the counters sits at constant offsets (no address arithmetic on W or the file count in the
    // item loops); the counters, whose size depends on both, come last
    S.stage = smem;
    S.refn = S.stage + FB_THREADS / 32 * FSTAGE_WORDS;
    S.covered = S.refn + (FWMAX >> 3) + 4;
    S.flag_row = (int*)(S.covered + (FWMAX >> 5) + 4);
    S.fpre = (uint16_t*)(S.flag_row + FWMAX);
    S.flag_pidx = S.fpre + FWMAX + 8;
    S.flag_d = (uint8_t*)(S.flag_pidx + FWMAX);
    S.dec = S.flag_d + FWMAX;
    S.cnt = (uint32_t*)(S.dec + FWMAX);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const CntPacked CN = {S.cnt, S.WS};
    TileSmem T2;
    T2.cnt = S.cnt;
    T2.hist_words = 2 * RAER_NHIST;  // coverage <= 65535 in this kernel: two 16-bit bins per word
    T2.slot_words = T2.hist_words + 4 + 2 * RAER_SLOT_FS * F;
    FastCtx X;
    X.minbq = P.min_bq;
    X.mb4 = (uint32_t)max(P.min_bq, 0) * 0x01010101u;
    X.zall = P.min_bq > 0 ? 0xffffffffu : 0u;
    X.HW = T2.hist_words;
    X.SW = T2.slot_words;
    X.round0 = 0;
    X.ns = 0;
    long long ph_t = clock64();

    while (true) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(P.tile_counter, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= P.n_tiles) break;
        PHASE_MARK(0);  // tile fetch (and what follows the last mark of the previous tile)
        const int t0 = P.t_beg + tile * W;
        const int t1 = min(t0 + W, P.t_end);
        X.t0 = t0;
        X.tw = t1 - t0;
        {
            uint4* z = reinterpret_cast<uint4*>(S.cnt);
            for (int i = threadIdx.x; i < (cnt_words >> 2); i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        }
        for (int i = threadIdx.x; i < (W >> 5) + 1; i += blockDim.x) S.covered[i] = 0;
        for (int i = threadIdx.x; i < (W >> 3) + 3; i += blockDim.x) {
            uint32_t w = 0;
            const int p0 = t0 - 8 + 8 * i;
#pragma unroll
            for (int m = 0; m < 8; ++m) w |= (uint32_t)ref_nibble(P, p0 + m) << (4 * m);
            S.refn[i] = w;
        }
        if (threadIdx.x == 0) { s_nflag = 0; s_deep = 0; s_chunk = 0; }
        __syncthreads();
        PHASE_MARK(1);  // clear + reference packing
        if (P.tile_off[tile * F + F] > P.tile_list_cap) {  // lists overflowed their allocation: the host re-runs the batch
            if (threadIdx.x == 0) {
                atomicCAS(P.err_flag, 0, RAER_RETRY_LISTCAP);
                P.tile_rows[tile] = make_int2(0, 0);
            }
            continue;
        }

        // ---- phase 1: coverage events + exceptional bases
        {
            const int l0 = P.tile_off[tile * F], l1 = P.tile_off[tile * F + F];
            const int nchunks = (l1 - l0 + 31) >> 5;
            const bool need_cov = P.min_depth <= 0;
            uint32_t* st = S.stage + wid * FSTAGE_WORDS;
            while (true) {  // warps pull chunks of 32 listed reads
                int c = 0;
                if (lane == 0) c = atomicAdd(&s_chunk, 1);
                c = __shfl_sync(FULL, c, 0);
                if (c >= nchunks) break;
                // from the back of the list: the chunks of CIGAR-walking reads are the slow ones, they go first
                fast_count_chunk<EXT>(P, S, X, tile, l0 + ((nchunks - 1 - c) << 5), l1, t1, lane, st, need_cov);
            }
        }
        PHASE_BARRIER(8);
        PHASE_MARK(2);  // phase 1

        // ---- coverage: prefix sum of the events, one warp per (file, strand)
        for (int a = wid; a < 2 * F; a += FB_THREADS / 32) {
            uint32_t* dl = S.cnt + (size_t)(a * 4) * S.WS;
            uint32_t carry = 0;
            for (int e0 = 0; e0 < W; e0 += 128) {
                const int e = e0 + lane * 4;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (e < W) v = *reinterpret_cast<uint4*>(dl + e);
                v.y += v.x; v.z += v.y; v.w += v.z;
                uint32_t inc = v.w;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(FULL, inc, o);
                    if (lane >= o) inc += u;
                }
                const uint32_t ex = inc - v.w + carry;
                v.x += ex; v.y += ex; v.z += ex; v.w += ex;
                if (e < W) {
                    *reinterpret_cast<uint4*>(dl + e) = v;
                    if (max(max(v.x, v.y), max(v.z, v.w)) > 65535u) s_deep = 1;
                }
                carry += __shfl_sync(FULL, inc, 31);
            }
        }
        __syncthreads();
        PHASE_MARK(3);  // coverage prefix sums
        if (s_deep && threadIdx.x == 0) atomicCAS(P.err_flag, 0, RAER_RETRY_DEEP);  // 16-bit class counters cannot hold this tile

        // ---- phase 2: decide, reserve rows, write them in (pos, strand) order
        int my_rows = 0;
        for (int pidx = threadIdx.x; pidx < W; pidx += blockDim.x) {
            int d = 0;
            if (t0 + pidx < t1) d = decide_site(P, CN, S.covered, pidx, t0 + pidx);
            S.dec[pidx] = (uint8_t)d;
            my_rows += (d & 1) + ((d >> 1) & 1) + ((d & 48) ? 0x10000 : 0);
        }
        if (P.aei_mode) {
            // calc_AEI (reference R/calc_AEI.R:274-323) sums, per file, n<base> over the rows whose REF is a given base:
            // a 4 x 4 matrix per file is all that leaves the tile. flag_row (unused here) holds the tile's sums.
            uint32_t* acc_s = reinterpret_cast<uint32_t*>(S.flag_row);
            for (int i = threadIdx.x; i < F * 16; i += blockDim.x) acc_s[i] = 0;
            __syncthreads();
            for (int f = 0; f < F; ++f) {
                int acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0;
                for (int pidx = threadIdx.x; pidx < W; pidx += blockDim.x) {
                    const int d = S.dec[pidx];
                    if (!(d & 3)) continue;
                    const int pc = cls_of_ref(d_upper((unsigned char)P.ref[t0 + pidx]));
                    if (pc >= 4) continue;  // REF outside A C G T takes part in no (REF, ALT) pair
                    for (int st2 = 0; st2 < 2; ++st2) {
                        if (!((d >> st2) & 1)) continue;
                        int cc[RAER_NCLS];
                        CN.get(f, st2, pidx, pc, cc);
                        const int rc = st2 == 0 ? pc : 3 - pc;  // REF of the row on that strand
#pragma unroll
                        for (int r4 = 0; r4 < 4; ++r4)
#pragma unroll
                            for (int b4 = 0; b4 < 4; ++b4) acc[r4 * 4 + b4] += (r4 == rc) ? cc[b4] : 0;
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int v = __reduce_add_sync(FULL, acc[i]);
                    if (lane == 0 && v) atomicAdd(&acc_s[f * 16 + i], (uint32_t)v);
                }
            }
            __syncthreads();
            for (int i = threadIdx.x; i < F * 16; i += blockDim.x)
                if (acc_s[i]) atomicAdd(&P.aei_out[i], (unsigned long long)acc_s[i]);
            if (threadIdx.x == 0) P.tile_rows[tile] = make_int2(0, 0);
            __syncthreads();
            continue;
        }
        // rows in the low half; sites that only need the second pass for the khash state in the high half
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
                const int pidx = chunk + threadIdx.x;
                const int d = pidx < W ? S.dec[pidx] : 0;
                const int n = (d & 1) + ((d >> 1) & 1);
                const bool fl = P.want_stats && (((d & 1) && (d & 4)) || ((d & 2) && (d & 8)) || (d & 48));
                // one scan for both: rows in the low half, flagged sites in the high half (a tile has < 2^15 of either)
                int ptot;
                const int pex = block_exclusive_scan(n | ((fl ? 1 : 0) << 16), &ptot, s_scan);
                const int off = (pex & 0xffff) + carried;
                const int rank = (pex >> 16) + fcarried;  // position-sorted
                carried += ptot & 0xffff;
                fcarried += ptot >> 16;
                if (pidx < W) S.fpre[pidx] = (uint16_t)rank;
                if (n || fl) {
                    const long long row = rowbase + off;
                    if (d & 1) write_row(P, CN, pidx, t0 + pidx, 0, row);
                    if (d & 2) write_row(P, CN, pidx, t0 + pidx, 1, row + (d & 1));
                    if (fl) {
                        S.flag_pidx[rank] = (uint16_t)pidx;
                        S.flag_row[rank] = (int)row;
                        S.flag_d[rank] = (uint8_t)d;
                    }
                }
            }
            if (threadIdx.x == 0) { s_nflag = fcarried; S.fpre[W] = (uint16_t)fcarried; }
        }
        __syncthreads();
        PHASE_MARK(4);  // decide + rows

        // ---- phases 3+4: second pass over the sites that carry a variant, `spr` sites per round
        const int nflag = s_nflag;
        const int HW = X.HW, SW = X.SW;
        const int spr = min(cnt_words / SW, 4096);
        for (int round0 = 0; round0 < nflag; round0 += spr) {
            const int ns = min(spr, nflag - round0);
            X.round0 = round0;
            X.ns = ns;
            for (int i = threadIdx.x; i < ns * SW; i += blockDim.x) {
                const int wsl = i % SW;
                // first-seen ordinals (and the minimum IUPAC code) start at +inf, everything else at 0
                const int fw = (wsl - (HW + 4)) % RAER_SLOT_FS;
                const bool is_fs = wsl >= HW + 4 && (fw < 4 || fw == 5 || fw == 6);
                S.cnt[i] = is_fs ? 0xffffffffu : 0u;
            }
            if (threadIdx.x == 0) s_chunk = 0;
            __syncthreads();
            {
                const int l0 = P.tile_off[tile * F], l1 = P.tile_off[tile * F + F];
                const int nchunks = (l1 - l0 + 31) >> 5;
                uint32_t* st = S.stage + wid * FSTAGE_WORDS;
                while (true) {  // warps pull chunks of 32 listed reads
                    int c = 0;
                    if (lane == 0) c = atomicAdd(&s_chunk, 1);
                    c = __shfl_sync(FULL, c, 0);
                    if (c >= nchunks) break;
                    fast_pair_chunk<EXT>(P, S, X, T2, tile, l0 + ((nchunks - 1 - c) << 5), l1, lane, st);
                }
            }
            PHASE_BARRIER(9);
            PHASE_MARK(5);  // second pass (slot clear + pairs)
            for (int i = threadIdx.x; i < ns; i += blockDim.x) {
                const int d = S.flag_d[round0 + i], row = S.flag_row[round0 + i];
                slot_stats(P, S.cnt + (size_t)i * SW, HW, ((d & 1) && (d & 4)) ? row : -1,
                           ((d & 2) && (d & 8)) ? row + (d & 1) : -1, d >> 4, t0 + (int)S.flag_pidx[round0 + i]);
            }
            __syncthreads();
            PHASE_MARK(6);  // statistics
        }
        __syncthreads();
    }
}

// ---- per-tile read lists: every read that overlaps the tile [t_beg + t*W, +W), per (tile, file)
__device__ __forceinline__ bool tile_span(const PlpParams& P, long long r, int& f, int& tf, int& tl, int* cplx = nullptr) {
    f = 0;
    while (f + 1 < P.n_files && r >= P.file_off[f + 1]) ++f;
    const int4 pe = __ldg(reinterpret_cast<const int4*>(P.hdr + r));
    if (cplx) *cplx = ((pe.w >> 16) & 0xffff) != 1;  // more than one CIGAR operation
    if (pe.y <= P.t_beg || pe.x >= P.t_end || pe.y <= pe.x) return false;
    tf = (max(pe.x, P.t_beg) - P.t_beg) / P.W;
    tl = (min(pe.y, P.t_end) - 1 - P.t_beg) / P.W;
    return true;
}
// Reads are coordinate sorted, so the lanes of a warp mostly share their (tile, file): one atomic per distinct key
// and warp step instead of one per read (the tile counters would serialise them).
#define TILE_AGG_STEPS 8  // tiles of a long read handled warp-synchronously; anything beyond takes plain atomics
__global__ void k_tile_count(PlpParams P, long long n_reads, int* cnt) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int f = 0, tf = 0, tl = -1;
    bool ok = r < n_reads && tile_span(P, r, f, tf, tl);
    // k_pileup_lane: a tile finds the reads that start inside it by range (tile_first); only the tiles behind a read's
    // first one list it
    if (ok && P.carry_only && __ldg(&P.hdr[r].pos) >= P.t_beg) { ++tf; ok = tf <= tl; }
    const int more = __reduce_max_sync(FULL, ok ? tl - tf : -1);
    for (int k = 0; k <= min(more, TILE_AGG_STEPS); ++k) {
        const int key = (ok && tf + k <= tl) ? (tf + k) * P.n_files + f : -1;
        const unsigned peers = __match_any_sync(FULL, key);
        if (key >= 0 && (int)(__ffs(peers) - 1) == lane) atomicAdd(&cnt[key], __popc(peers));
    }
    for (int t = tf + TILE_AGG_STEPS + 1; ok && t <= tl; ++t) atomicAdd(&cnt[t * P.n_files + f], 1);
}
// exclusive scan of n counts into off[0..n] (one CTA), cursors cleared
__global__ void k_tile_scan(const int* cnt, int* off, int* cur, int n, int* total) {
    __shared__ int s_w[33];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n ? cnt[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int z = lane < (int)(blockDim.x >> 5) ? s_w[lane] : 0;
            const int own = z;
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FULL, z, o);
                if (lane >= o) z += y;
            }
            s_w[lane] = z - own;
            if (lane == 31) s_w[32] = z;
        }
        __syncthreads();
        if (i < n) {
            off[i] = s_carry + s_w[wid] + x - v;
            cur[i] = 0;
        }
        // (the counts themselves are cleared by the caller once the scan is done: k_tile_fill uses them as back cursors)
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) { off[n] = s_carry; *total = s_carry; }
}
// A (tile, file) list is filled from both ends: reads with a single CIGAR operation from the front, the others from
// the back, so that the chunks of 32 reads the tile kernel takes are (but for one per list) all of one kind and
// the lanes of a warp do not wait for each other's CIGAR walks.
__global__ void k_tile_fill(PlpParams P, long long n_reads, const int* off, int* cur, int* cur_back, int* list, int cap) {
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int f = 0, tf = 0, tl = -1, cplx = 0;
    bool ok = r < n_reads && tile_span(P, r, f, tf, tl, &cplx);
    if (ok && P.carry_only && __ldg(&P.hdr[r].pos) >= P.t_beg) { ++tf; ok = tf <= tl; }
    const int more = __reduce_max_sync(FULL, ok ? tl - tf : -1);
    for (int k = 0; k <= min(more, TILE_AGG_STEPS); ++k) {
        const int key = (ok && tf + k <= tl) ? (tf + k) * P.n_files + f : -1;
        const unsigned peers = __match_any_sync(FULL, key < 0 ? -1 : key * 2 + cplx);
        const int leader = __ffs(peers) - 1;
        int first = 0;
        if (key >= 0 && leader == lane) first = atomicAdd(cplx ? &cur_back[key] : &cur[key], __popc(peers));
        first = __shfl_sync(FULL, first, leader);
        if (key >= 0) {
            const int rank = first + __popc(peers & ((1u << lane) - 1u));
            const int slot = cplx ? off[key + 1] - 1 - rank : off[key] + rank;
            if (slot >= 0 && slot < cap) list[slot] = (int)r;
        }
    }
    for (int t = tf + TILE_AGG_STEPS + 1; ok && t <= tl; ++t) {
        const int k2 = t * P.n_files + f;
        const int rank = atomicAdd(cplx ? &cur_back[k2] : &cur[k2], 1);
        const int slot = cplx ? off[k2 + 1] - 1 - rank : off[k2] + rank;
        if (slot >= 0 && slot < cap) list[slot] = (int)r;
    }
}

#define FFIXED_BYTES ((FB_THREADS / 32 * FSTAGE_WORDS + (FWMAX >> 3) + 4 + (FWMAX >> 5) + 4 + FWMAX) * 4 + (2 * FWMAX + 8) * 2 + 2 * FWMAX)
extern "C" size_t raer_fast_smem_bytes(int n_files, int W) {
    return (size_t)FFIXED_BYTES + (size_t)n_files * 8 * FWS(W) * 4 + 16;
}

// widest tile (multiple of 128 positions, at most 1024) that lets two CTAs share one SM; 0 = does not fit
extern "C" int raer_fast_width(int n_files) {
    const size_t budget = (227 * 1024 - 2 * 1024) / 2 - 768;  // per CTA, minus the static part
    int best = 0;
    for (int w = 128; w <= 1024; w += 128)
        if (raer_fast_smem_bytes(n_files, w) <= budget) best = w;
    if (best == 0)
        for (int w = 32; w < 128; w += 32)
            if (raer_fast_smem_bytes(n_files, w) <= budget) best = w;
    return best;
}

// count + scan of the tile lists; the caller reads *total back, sizes the list and calls raer_fast_fill
extern "C" int raer_fast_index_count(PlpParams* P, long long n_reads, int* cnt, int* off, int* cur, int* total,
                                     cudaStream_t st) {
    const int n = P->n_tiles * P->n_files;
    cudaMemsetAsync(cnt, 0, (size_t)n * sizeof(int), st);
    if (n_reads > 0) k_tile_count<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(*P, n_reads, cnt);
    k_tile_scan<<<1, 1024, 0, st>>>(cnt, off, cur, n, total);
    return cudaGetLastError() == cudaSuccess ? 2 : -1;
}
extern "C" int raer_fast_index_fill(PlpParams* P, long long n_reads, const int* off, int* cur, int* cur_back, int* list, int cap,
                                    cudaStream_t st) {
    cudaMemsetAsync(cur_back, 0, (size_t)P->n_tiles * P->n_files * sizeof(int), st);
    if (n_reads > 0) k_tile_fill<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(*P, n_reads, off, cur, cur_back, list, cap);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

#endif
