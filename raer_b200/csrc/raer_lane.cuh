// raer_lane.cuh -- k_pileup_lane: the tile kernel of the default FilterParam (only the mapq and base-quality filters of
// check_read_filters are active, src/plp.c:568-586; no UMI mode). Same tile phases, counter layout, site decision, rows,
// second-pass slots and statistics as k_pileup_fast (raer_fast.cuh); what differs is how the reads get to the counters:
//
//   * A tile owns the reads that START inside it: a contiguous range per file (tile_first, coordinate-sorted input).
//     Warps pull chunks of 32 consecutive reads. Their headers are one coalesced 1 KB load, and their packed sequences
//     and qualities are two contiguous spans of the batch arrays: one lane issues two cp.async.bulk copies (TMA unit,
//     completion on the warp's mbarrier) that land them in the warp's slice of shared memory. No per-lane global load
//     touches sequence or quality bytes; no sector is fetched twice.
//   * ONE LANE PER READ. A lane walks its read's CIGAR, writes the coverage events of every aligned block and runs
//     through the block in chunks of 16 query bases (two 32-bit sequence words, four 32-bit quality words from shared
//     memory): XOR with 16 packed reference bases, byte-parallel quality classes, everything reduced to 16-bit masks
//     with multiply-gathers. Only exceptional bases (mismatch, below the quality threshold, N, quality-0 runs) issue a
//     shared-memory RED. No per-item binary search, no staged records, no warp-wide prefix sums.
//   * Reads that started in an earlier tile reach the tile through the carry-in lists (k_tile_count / k_tile_fill with
//     carry_only): the same lane code with pointers into global memory.
//   * Phase 1 leaves, per read, a bit mask of the query bases that were counted as reference matches (5 words, written
//     coalesced). The second pass -- one (read, flagged site) pair per lane -- tests one bit of it instead of fetching a
//     sequence and a quality sector per pair; only reads that carry a counted mismatch (flag in the mask) look at
//     their bases again. This removes the 2x DRAM traffic of k_pileup_fast's second pass.
//   * Statistics: one warp per slot (integer Mann-Whitney sums over 25 lanes, vdb over the non-empty bins in order).
//   * One CTA of LN_THREADS per SM; the second-pass slots use the counter region plus the idle part of the staging area.
#ifndef RAER_LANE_CUH
#define RAER_LANE_CUH

#define LN_THREADS 768
#define LN_WARPS (LN_THREADS / 32)
#define LN_MW RAER_LANE_MW           // mask words per read: query bases 0 .. LN_MAXQ - 1; two flag bits on top
#define LN_MAXQ (32 * LN_MW - 2)
#define LN_HASMM 0x40000000u         // last mask word: the read has a counted mismatch inside this tile
#define LN_NOMASK 0x80000000u        //                 the read is longer than the mask: the second pass looks at its bases
#define LN_MASK_BYTES (32 * LN_MW * 4)
#define LN_REC2_WORDS (32 * LN_MW + 32)  // second pass, per warp: the masks of its 32 reads
#define LN_DEFER 768                 // reads with several aligned blocks a tile can set aside for dense chunks of their own
#define LN_PAD 32                    // reference bases packed in front of the tile

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const unsigned addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    } while (!ok);
}
// TMA unit, non-tensor form: `bytes` (multiple of 16) from global to shared memory, both 16-byte aligned; completion is
// signalled on the mbarrier as transferred bytes
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// bits at 4j (j = 0..7) -> bits j: one multiply gathers both half words (all partial products land on distinct bits)
__device__ __forceinline__ uint32_t nibbits_to_c8(uint32_t x) {
    const uint32_t y = x * 0x1248u;
    return ((y >> 12) & 0xfu) | ((y >> 24) & 0xf0u);
}
// bits at 8j + 7 (j = 0..3) -> bits j
__device__ __forceinline__ uint32_t msb4_to_c4(uint32_t t) { return (t * 0x00204081u) >> 28; }

struct LaneCtx {
    int t0, t1, tw, minbq;
    uint32_t mb4;
    bool zall, want_mask, need_cov;
    int round0, ns;
};

// bits 8j + 7 of `hi` and of `lo` (j = 0..3) -> bits 4 + j and j: one multiply gathers both words
__device__ __forceinline__ uint32_t msb8_to_c8(uint32_t lo, uint32_t hi) { return ((hi | (lo >> 4)) * 0x00204081u) >> 24; }

// One step of 32 query bases of one aligned block. Bases j in [jlo, jhi) are part of the block and sit on tile
// positions rel0 + j. sw[i]: packed bases 8i .. 8i + 7 as they lie in the BAM record; qw[i]: their qualities.
// Returns the mask of the bases that were counted as reference matches.
__device__ __forceinline__ uint32_t lane_step(const FastSmem& S, const LaneCtx& X, unsigned cbs, const uint32_t (&sw)[4],
                                              const uint2 (&qw)[4], const uint8_t* sb, int rel0, int jlo, int jhi, bool mapq_fail,
                                              int cxor, uint32_t& hasmm) {
    // sb: the packed bases of the step once more, as bytes: the rare per-base lookups read them from there instead of
    // indexing the register copies with a run-time index
    const uint32_t V = (0xffffffffu >> (32 - jhi)) & (0xffffffffu << jlo);
    const int a = rel0 + LN_PAD;  // >= 1: index of the step's first base in the packed reference
    const int wi = a >> 3, sh = (a & 7) << 2;
    uint32_t rf[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) rf[i] = S.refn[wi + i];
    uint32_t E = 0, LTc = 0, D = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        E |= nibbits_to_c8(nib_nonzero(nib_swap(sw[i]) ^ __funnelshift_r(rf[i], rf[i + 1], sh))) << (8 * i);
    E &= V;  // differs from the reference (read base N among them: the reference is never N where a row is written)
    if (!mapq_fail) {
#pragma unroll
        for (int i = 0; i < 4; ++i) LTc |= msb8_to_c8(byte_lt(qw[i].x, X.mb4), byte_lt(qw[i].y, X.mb4)) << (8 * i);
        // quality 0 below the threshold is dropped silently (src/plp.c:580-586)
        if (X.zall) {
#pragma unroll
            for (int i = 0; i < 4; ++i) D |= msb8_to_c8(byte_zero(qw[i].x), byte_zero(qw[i].y)) << (8 * i);
            D &= V;
        }
    }
    const unsigned base = cbs + ((unsigned)rel0 << 2);  // shared address of the coverage word of base 0
    // a base that differs from the reference and passes the filters is a variant base (rare); read base N is counted
    // nowhere (src/plp.c:718): it leaves the coverage like a dropped base
    for (uint32_t e = mapq_fail ? 0u : (E & ~(LTc | D)); e;) {
        const int j = __ffs(e) - 1;
        e &= e - 1;
        const int code = (sb[j >> 1] >> ((~j & 1) << 2)) & 15;
        if (code == 15) {
            red_add_s(base + (j << 2), 0xffffffffu);
            red_add_s(base + (j << 2) + 4, 1u);
            continue;
        }
        // A=1 C=2 G=4 T=8 -> 0..3; complement on the minus strand = xor 3; any other code -> other
        int cls = ((code >> 1) - (code >> 3)) ^ cxor;
        if (__popc(code) != 1) cls = RAER_CLS_OTHER;
        red_add_s(base + (unsigned)((1 + (cls >> 1)) * S.WS) * 4 + (j << 2), 1u << ((cls & 1) << 4));
        hasmm = 1;
    }
    // every base of a read below the mapq threshold is nX (src/plp.c:574)
    const uint32_t XM = (mapq_fail ? V : LTc & V) & ~D;
    const unsigned xb = base + (unsigned)(3 * S.WS) * 4;
    // One loop for both kinds of events (their bits are disjoint): nX bases, and the starts of the runs of dropped
    // bases, which leave the coverage with -1 where the run starts and +1 behind its end
    for (uint32_t u = XM | (D & ~(D << 1)); u;) {
        const int b1 = __ffs(u) - 1;
        u &= u - 1;
        if ((XM >> b1) & 1u) {
            // an nX base that differs from the reference may be an N: dropped, whatever its quality
            if (((E >> b1) & 1u) && ((sb[b1 >> 1] >> ((~b1 & 1) << 2)) & 15) == 15) {
                red_add_s(base + (b1 << 2), 0xffffffffu);
                red_add_s(base + (b1 << 2) + 4, 1u);
            } else {
                red_add_s(xb + (b1 << 2), 0x10000u);
            }
        } else {
            const uint32_t above = ~(D >> b1);  // first zero above the start = length of the run; none: all 32 bases dropped
            const int run = above ? __ffs(above) - 1 : 32;
            red_add_s(base + (b1 << 2), 0xffffffffu);
            red_add_s(base + ((b1 + run) << 2), 1u);
        }
    }
    return mapq_fail ? 0u : (V & ~(LTc | D | E));
}

// Phase 1 for up to 32 reads, one per lane, called by the whole warp (have: this lane has a read): coverage events and
// steps of 32 query bases of every aligned block that reaches into the tile. The lanes walk their CIGARs on their own,
// but they take their steps in lock step -- the ballot at the top of the loop re-joins the warp before every step, so
// the long straight-line step code runs with all lanes that still have work (without it lanes that once diverged would
// keep running as separate sub-warps). sp / qp: the read's packed sequence / qualities (shared memory when STAGED,
// else global); mk: LN_MW mask words.
template <bool STAGED>
__device__ __forceinline__ void warp_reads(const PlpParams& P, const FastSmem& S, const LaneCtx& X, bool have, const int4 a, const int4 b,
                                           int f, const uint8_t* sp, const uint8_t* qp, uint32_t* mk) {
    const int mapq = (a.z >> 16) & 0xff, bits = (a.z >> 24) & 0xff;
    const int lq = a.w & 0xffff, n = have ? (a.w >> 16) & 0xffff : 0;
    const int s = bits & RAER_RB_INVERT;
    const bool mapq_fail = mapq < P.min_mapq[f];
    uint32_t* dl = S.cnt + (size_t)((f * 2 + s) * 4) * S.WS;
    const unsigned cbs = smem_u32(dl);
    const uint32_t* cg = P.cigar + (unsigned)b.x;
    // one operation that consumes as many reference as query bases is an aligned block: no CIGAR fetch (see fast_stage)
    const bool one_block = n == 1 && a.y - a.x == lq && lq > 0;
    const bool masked = X.want_mask && lq <= LN_MAXQ;
    uint32_t hasmm = 0;
    if (X.want_mask && have) {
#pragma unroll
        for (int i = 0; i < LN_MW; ++i) mk[i] = 0;
    }
    int x = a.x, y = 0, k = 0;
    int q0 = 0, qa = 0, qb = 0, rel_a = 0;  // current aligned block inside the tile: query bases [qa, qb), next step at q0
    while (true) {
        while (q0 >= qb && k < n) {  // next aligned block that reaches into the tile
            const uint32_t w = one_block ? ((uint32_t)lq << 4) : __ldg(cg + k);
            ++k;
            const int op = cg_op(w), len = cg_len(w);
            if (X.need_cov && (cg_type(op) & 2)) {
                const int lo = max(x, X.t0), h2 = min(x + len, X.t1);
                if (lo < h2) cover_range(S.covered, lo - X.t0, h2 - X.t0);
            }
            if ((0x181 >> op) & 1) {  // M = X
                const int lo = max(x, X.t0), hi = min(x + min(len, lq - y), X.t1);
                if (lo < hi) {
                    red_add(dl + (lo - X.t0), 1u);
                    red_add(dl + (hi - X.t0), 0xffffffffu);
                    qa = y + (lo - x);  // query bases [qa, qb) sit on tile positions rel_a ...
                    qb = y + (hi - x);
                    rel_a = lo - X.t0;
                    q0 = qa & ~31;
                }
                x += len;
                y += len;
            } else if ((0xC >> op) & 1) {  // D N
                x += len;
            } else if ((0x12 >> op) & 1) {  // I S
                y += len;
            }
        }
        const bool act = q0 < qb;
        if (!__ballot_sync(FULL, act)) break;
        if (act) {
            uint32_t sw[4];
            uint2 qw[4];
            if (STAGED) {
                // the staged spans end with 32 readable bytes of slack: no guard for the words behind the read's last base
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    sw[i] = *reinterpret_cast<const uint32_t*>(sp + (q0 >> 1) + 4 * i);
                    qw[i] = *reinterpret_cast<const uint2*>(qp + q0 + 8 * i);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool in = q0 + 8 * i < lq;
                    sw[i] = in ? __ldg(reinterpret_cast<const uint32_t*>(sp + (q0 >> 1) + 4 * i)) : 0xffffffffu;
                    qw[i] = in ? __ldg(reinterpret_cast<const uint2*>(qp + q0 + 8 * i)) : make_uint2(0, 0);
                }
            }
            const uint32_t cm = lane_step(S, X, cbs, sw, qw, sp + (q0 >> 1), rel_a + (q0 - qa), max(qa - q0, 0), min(qb - q0, 32),
                                          mapq_fail, s ? 3 : 0, hasmm);
            if (masked && cm) atomicOr(&mk[q0 >> 5], cm);  // a word can be visited from two aligned blocks
            q0 += 32;
        }
    }
    if (X.want_mask && have) {
        if (!masked) mk[LN_MW - 1] = LN_NOMASK;
        else if (hasmm) mk[LN_MW - 1] |= LN_HASMM;
    }
}

// ---- second pass for up to 32 reads, one per lane and called by the whole warp (r < 0: this lane has no read): every
// flagged site of the round that the read covers, answered from the read's mask. A lane walks its CIGAR; the prefix
// count of flagged positions gives each aligned block its range of flagged ranks in O(1) (a spliced read does not pay
// for the sites inside its introns); the lanes then take their sites in lock step (ballot at the top of the loop, see
// warp_reads). mk: the lane's LN_MW words in the warp's shared-memory slice.
__device__ __forceinline__ void warp_pairs(const PlpParams& P, const FastSmem& S, const LaneCtx& X, const TileSmem& T2, int r, int f,
                                           const uint32_t* maskp, uint32_t* mk) {
    bool have = r >= 0;
    int4 a = make_int4(0, 0, 0, 0), b = make_int4(0, 0, 0, 0);
    if (have) {
        a = __ldg(reinterpret_cast<const int4*>(P.hdr + r));
        // a read below the mapq threshold only produces nX: nothing to accumulate (src/plp.c:574)
        have = ((a.z >> 16) & 0xff) >= P.min_mapq[f];
    }
    const int lq = a.w & 0xffff, n = have ? (a.w >> 16) & 0xffff : 0;
    const bool one_block = n == 1 && a.y - a.x == lq && lq > 0;
    uint32_t top = 0;
    ReadCtx c;
    c.l_qseq = lq;
    c.qs = 0;
    c.qe = lq;
    const uint32_t* cg = P.cigar;
    uint32_t sq = 0;
    if (have) {
        b = __ldg(reinterpret_cast<const int4*>(P.hdr + r) + 1);
        cg = P.cigar + (unsigned)b.x;
        sq = (uint32_t)b.y;
#pragma unroll
        for (int i = 0; i < LN_MW; ++i) mk[i] = __ldcg(maskp + i);  // written by this kernel: not through the read-only path
        top = mk[LN_MW - 1];
        if (!one_block) {
            // query_start / query_end (src/plp_utils.c:52-104): soft clips at either end, hard clips skipped
            int qs = 0, tail = 0;
            bool lead = true;
            for (int k = 0; k < n; ++k) {
                const uint32_t w = __ldg(cg + k);
                const int op = cg_op(w);
                if (op == OP_S) { if (lead) qs += cg_len(w); tail += cg_len(w); }
                else if (op != OP_H) { lead = false; tail = 0; }
            }
            c.qs = qs;
            c.qe = lq - tail;
        }
    }
    const int flag = a.z & 0xffff, bits = (a.z >> 24) & 0xff;
    c.rev = (flag >> 4) & 1;
    c.invert = bits & RAER_RB_INVERT;
    // get_relative_position divides by alen + 1 = qe - qs + 1 (src/plp_utils.c:401-412): reciprocal once per read
    const int rb = c.qe - c.qs + 1;
    const uint32_t rcp = rb > 1 ? 0xffffffffu / (uint32_t)rb + 1u : 0u;  // ceil(2^32 / rb) without a 64-bit divide
    const int ordinal = r - (int)P.file_off[f];  // position of the read inside its file: what "first seen" means in BAM order
    const int r0 = X.round0, r1 = X.round0 + X.ns;
    int x = a.x, y = 0, k = 0, cur = 0, end = 0, qoff = 0;
    while (true) {
        while (cur >= end && k < n) {  // next aligned block with flagged sites of this round
            const uint32_t w = one_block ? ((uint32_t)lq << 4) : __ldg(cg + k);
            ++k;
            const int op = cg_op(w), len = cg_len(w);
            if ((0x181 >> op) & 1) {  // M = X
                const int lo = min(max(x - X.t0, 0), X.tw), hi = min(max(x + min(len, lq - y) - X.t0, 0), X.tw);
                if (lo < hi) {
                    const int c0 = max((int)S.fpre[lo], r0), c1 = min((int)S.fpre[hi], r1);
                    if (c0 < c1) { cur = c0; end = c1; qoff = y - x; }
                }
                x += len;
                y += len;
            } else if ((0xC >> op) & 1) {  // D N
                x += len;
            } else if ((0x12 >> op) & 1) {  // I S
                y += len;
            }
        }
        const bool act = cur < end;
        if (!__ballot_sync(FULL, act)) break;
        if (act) {
            const int rank = cur++;
            const int pidx = (int)S.flag_pidx[rank];
            const int q = X.t0 + pidx + qoff;
            int code = 0, isref = 1;
            bool valid = true;
            if (!(top & LN_NOMASK) && ((mk[q >> 5] >> (q & 31)) & 1u)) {
                // counted in phase 1 as a base that equals the reference
            } else if (top & (LN_NOMASK | LN_HASMM)) {
                // the read carries a variant base somewhere in the tile (or has no mask): look at this base
                code = seq_code(P.seq + sq, q);
                const int bq = P.qual[2ull * sq + q];
                const int ai = pidx + LN_PAD;
                const int rcode = (S.refn[ai >> 3] >> ((ai & 7) << 2)) & 15;
                // read base N (src/plp.c:718); nX or dropped bases are not part of the statistics
                valid = code != 15 && bq >= X.minbq;
                isref = (code == rcode) & (code != 0);
            } else {
                valid = false;
            }
            if (valid) pass2_accumulate_core(P, T2, c, f, c.invert ? 1 : 0, ordinal, rank - r0, q, code, isref, rcp);
        }
    }
}

// ---- slot_stats with one warp per slot. calc_mwu_biasZ (bcftools-ext.c:144-212): every sum is an integer sum (int
// arithmetic wraps the same way in any order), so the 100 bins are spread over 25 lanes, 4 consecutive bins each.
struct MwuSums { int e, l, na, nb; long long t; };
__device__ __forceinline__ double finish_mwu(const MwuSums m) {
    const int na = m.na, nb = m.nb;
    if (!na || !nb) return CUDART_INF;  // covers the empty-b branch of the reference as well
    const double U = m.l + m.e * 0.5;
    const double mm = na * nb / 2.0;
    const double var2 = (na * nb) / 12.0 * ((na + nb + 1) - m.t / (double)((na + nb) * (na + nb - 1)));
    if (var2 <= 0) return 0;
    return (U - mm) / sqrt(var2);
}
__device__ __forceinline__ MwuSums warp_mwu_sums(const HistView a, const HistView b, int lane) {
    int av[4] = {0, 0, 0, 0}, bv[4] = {0, 0, 0, 0};
    if (lane < RAER_NHIST / 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { av[k] = a[4 * lane + k]; bv[k] = b[4 * lane + k]; }
    }
    const int bsum = bv[0] + bv[1] + bv[2] + bv[3];
    // number of b entries in the bins above this lane's (the serial loop walks the bins downwards)
    int above = bsum;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_down_sync(FULL, above, o);
        if (lane + o < 32) above += v;
    }
    above -= bsum;
    unsigned e = 0, l = 0;
    long long t = 0;
    int nb_run = above;
#pragma unroll
    for (int k = 3; k >= 0; --k) {
        e += (unsigned)(av[k] * bv[k]);
        l += (unsigned)(av[k] * nb_run);
        nb_run += bv[k];
        const int p = av[k] + bv[k];
        t += (long long)((p * p - 1) * p);
    }
    e = __reduce_add_sync(FULL, e);
    l = __reduce_add_sync(FULL, l);
    const int na = (int)__reduce_add_sync(FULL, (unsigned)(av[0] + av[1] + av[2] + av[3]));
    const int nb = (int)__reduce_add_sync(FULL, (unsigned)bsum);
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL, t, o);
    MwuSums m;
    m.e = (int)e; m.l = (int)l; m.na = na; m.nb = nb; m.t = t;
    return m;
}

// calc_vdb (bcftools-ext.c:51-111): the float accumulations depend on the order of the bins, so every lane walks the
// non-empty bins in ascending order (found by ballots, values broadcast) and computes the same value.
// the part of calc_vdb behind the two histogram loops: dp reads, mean_diff already divided by dp
__device__ __forceinline__ double finish_vdb(int dp, float mean_diff) {
    const float prm[15][3] = {{3, 0.079f, 18},    {4, 0.09f, 19.8f},  {5, 0.1f, 20.5f},   {6, 0.11f, 21.5f},
                              {7, 0.125f, 21.6f}, {8, 0.135f, 22},    {9, 0.14f, 22.2f},  {10, 0.153f, 22.3f},
                              {15, 0.19f, 22.8f}, {20, 0.22f, 23.2f}, {30, 0.26f, 23.4f}, {40, 0.29f, 23.5f},
                              {50, 0.35f, 23.65f}, {100, 0.5f, 23.7f}, {200, 0.7f, 23.7f}};
    if (dp < 2) return CUDART_INF;
    const int ipos = (int)mean_diff;
    if (dp == 2) return (double)((2 * 100 - 2 * (ipos + 1) - 1) * (ipos + 1) / 99) / (100 * 0.5);
    int i;
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
__device__ __forceinline__ void warp_vdb_sums(const HistView pos, int lane, int& dp_out, float& md_out) {
    int v[4];
    unsigned nz[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = lane + 32 * k;
        v[k] = i < RAER_NHIST ? (int)pos[i] : 0;
        nz[k] = __ballot_sync(FULL, v[k] != 0);
    }
    int dp = 0;
    float mean_pos = 0, mean_diff = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        for (unsigned m = nz[k]; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const int val = __shfl_sync(FULL, v[k], src);
            dp += val;
            mean_pos += (float)(val * (src + 32 * k));
        }
    dp_out = dp;
    md_out = 0;
    if (dp < 2) return;
    mean_pos /= (float)dp;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        for (unsigned m = nz[k]; m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const int val = __shfl_sync(FULL, v[k], src);
            // float += int * fabs(float): evaluated in double, rounded back to float
            mean_diff = (float)((double)mean_diff + (double)val * fabs((double)((float)(src + 32 * k) - mean_pos)));
        }
    mean_diff /= (float)dp;
    md_out = mean_diff;
}

// Statistics of one slot in two steps. slot_sums_warp (all 32 lanes, one warp per slot): the integer sums of
// calc_mwu_biasZ and the two float accumulations of calc_vdb for both strands, parked in the slot's own first words
// (LN_SUM_WORDS per strand; the histogram of strand 1 starts at word 100, so strand 0's summary does not touch what is
// still to be read), and the ALT order codes (lane f takes file f). slot_finish (one thread per slot and strand, all
// slots of the round side by side): the fp64 tails -- square root, erfc, logarithms -- whose latency a warp per slot
// would pay once per slot instead of once per round.
#define LN_SUM_WORDS 8
__device__ __forceinline__ void slot_sums_warp(const PlpParams& P, uint32_t* slot, int HW, int rowp, int rowm, int growmask, int pos,
                                               int lane) {
    for (int s = 0; s < 2; ++s) {
        const long long row = s == 0 ? rowp : rowm;
        if (row < 0 && !((growmask >> s) & 1)) continue;
        if (row >= 0) {
            const HistView hr = {slot, (s * 2) * RAER_NHIST, HW != 4 * RAER_NHIST};
            const HistView ha = {slot, (s * 2 + 1) * RAER_NHIST, HW != 4 * RAER_NHIST};
            const MwuSums m = warp_mwu_sums(hr, ha, lane);
            int dp;
            float md;
            warp_vdb_sums(ha, lane, dp, md);
            __syncwarp();  // every lane has read its bins of this strand
            if (lane == 0) {
                uint32_t* o = slot + s * LN_SUM_WORDS;
                o[0] = (uint32_t)m.e; o[1] = (uint32_t)m.l; o[2] = (uint32_t)m.na; o[3] = (uint32_t)m.nb;
                o[4] = (uint32_t)(unsigned long long)m.t; o[5] = (uint32_t)((unsigned long long)m.t >> 32);
                o[6] = (uint32_t)dp; o[7] = __float_as_uint(md);
            }
        }
        for (int f = lane; f < P.n_files; f += 32) {
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
__device__ __forceinline__ void slot_finish(const PlpParams& P, const uint32_t* slot, int HW, int s, long long row) {
    const uint32_t* o = slot + s * LN_SUM_WORDS;
    MwuSums m;
    m.e = (int)o[0]; m.l = (int)o[1]; m.na = (int)o[2]; m.nb = (int)o[3];
    m.t = (long long)((unsigned long long)o[4] | ((unsigned long long)o[5] << 32));
    P.row_stats[row * 3 + 0] = finish_mwu(m);
    P.row_stats[row * 3 + 1] = finish_vdb((int)o[6], __uint_as_float(o[7]));
    P.row_stats[row * 3 + 2] = d_calc_sor((int)slot[HW], (int)slot[HW + 1], (int)slot[HW + 2], (int)slot[HW + 3]);
}

// file of own-range chunk c (chunks of the files are numbered one file after the other) and its first read
__device__ __forceinline__ int own_chunk(const int* s_first, const int* s_coff, int F, int c, int& f, int& n) {
    f = 0;
    while (f + 1 < F && c >= s_coff[f + 1]) ++f;
    const int r0 = s_first[2 * f] + ((c - s_coff[f]) << 5);
    n = min(32, s_first[2 * f + 1] - r0);
    return r0;
}

__global__ void __launch_bounds__(LN_THREADS, 1) k_pileup_lane(PlpParams P) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int s_scan[40];
    __shared__ int s_first[2 * RAER_MAX_FILES], s_coff[RAER_MAX_FILES + 1];
    __shared__ int s_tile, s_nflag, s_rowbase, s_deep, s_chunk, s_ndefer, s_own_seen;
    const int W = P.W, F = P.n_files;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    FastSmem S;
    S.WS = FWS(W);
    const int cnt_words = F * 8 * S.WS;
    // staging slices first, the counters right behind them: the second pass keeps its slots in the counter region and in
    // the part of the staging area its records do not use
    const int stage_words = P.stage_bytes >> 2;
    S.stage = smem;
    S.cnt = smem + LN_WARPS * stage_words;
    S.refn = S.cnt + cnt_words;
    S.covered = S.refn + (FWMAX >> 3) + 12;
    S.flag_row = (int*)(S.covered + (FWMAX >> 5) + 4);
    S.fpre = (uint16_t*)(S.flag_row + FWMAX);
    S.flag_pidx = S.fpre + FWMAX + 8;
    S.flag_d = (uint8_t*)(S.flag_pidx + FWMAX);
    S.dec = S.flag_d + FWMAX;
    uint64_t* mbars = (uint64_t*)(S.dec + FWMAX);
    int* const s_defer = (int*)(mbars + LN_WARPS);
    uint8_t* const s_defer_f = (uint8_t*)(s_defer + LN_DEFER);
    const CntPacked CN = {S.cnt, S.WS};
    uint32_t* const slot0 = smem + LN_WARPS * LN_REC2_WORDS;
    const int slot_region_words = LN_WARPS * (stage_words - LN_REC2_WORDS) + cnt_words;
    TileSmem T2;
    T2.cnt = slot0;
    T2.hist_words = 2 * RAER_NHIST;  // coverage <= 65535 in this kernel: two 16-bit bins per word
    T2.slot_words = T2.hist_words + 4 + 2 * RAER_SLOT_FS * F;
    LaneCtx X;
    X.minbq = P.min_bq;
    X.mb4 = (uint32_t)max(P.min_bq, 0) * 0x01010101u;
    X.zall = P.min_bq > 0;
    X.want_mask = P.want_stats && !P.aei_mode;
    X.need_cov = P.min_depth <= 0;
    X.round0 = 0;
    X.ns = 0;
    // this warp's staging slice: sequence span | quality span | masks
    const int SB = ((P.stage_bytes - LN_MASK_BYTES) / 3) & ~15;
    uint8_t* const w_seq = (uint8_t*)(smem + wid * stage_words);
    uint8_t* const w_qual = w_seq + SB;
    uint32_t* const w_mask = (uint32_t*)(w_seq + 3 * SB);
    uint64_t* const w_bar = mbars + wid;
    uint32_t w_par = 0;
    const bool can_stage = SB >= 256 && (((uintptr_t)P.seq | (uintptr_t)P.qual) & 15) == 0;
    if (lane == 0) mbar_init(w_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    long long ph_t = clock64();

    while (true) {
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(P.tile_counter, 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= P.n_tiles) break;
        PHASE_MARK(0);  // tile fetch (and what follows the last mark of the previous tile)
        {
            // a tile no read reaches writes no row (a sparse BAM or a site-list call on a long contig is mostly such tiles)
            bool any = P.tile_off[tile * F + F] > P.tile_off[tile * F];
            for (int f = 0; f < F && !any; ++f) any = __ldg(P.tile_first + (tile + 1) * F + f) > __ldg(P.tile_first + tile * F + f);
            if (!any) {
                if (threadIdx.x == 0) P.tile_rows[tile] = make_int2(0, 0);
                __syncthreads();  // everybody has read s_tile before the next tile overwrites it
                continue;
            }
        }
        const int t0 = P.t_beg + tile * W;
        const int t1 = min(t0 + W, P.t_end);
        X.t0 = t0;
        X.t1 = t1;
        X.tw = t1 - t0;
        {
            uint4* z = reinterpret_cast<uint4*>(S.cnt);
            for (int i = threadIdx.x; i < (cnt_words >> 2); i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
        }
        for (int i = threadIdx.x; i < (W >> 5) + 1; i += blockDim.x) S.covered[i] = 0;
        for (int i = threadIdx.x; i < (W >> 3) + 9; i += blockDim.x) {
            uint32_t w = 0;
            const int p0 = t0 - LN_PAD + 8 * i;
#pragma unroll
            for (int m = 0; m < 8; ++m) w |= (uint32_t)ref_nibble(P, p0 + m) << (4 * m);
            S.refn[i] = w;
        }
        if (threadIdx.x < F) {
            // own range of the tile in this file: the reads that start inside it
            const int f = threadIdx.x;
            const int r0 = __ldg(P.tile_first + tile * F + f), r1 = __ldg(P.tile_first + (tile + 1) * F + f);
            s_first[2 * f] = r0;
            s_first[2 * f + 1] = r1;
        }
        if (threadIdx.x == 0) { s_nflag = 0; s_deep = 0; s_chunk = 0; s_ndefer = 0; s_own_seen = 0; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int c = 0;
            for (int f = 0; f < F; ++f) { s_coff[f] = c; c += (s_first[2 * f + 1] - s_first[2 * f] + 31) >> 5; }
            s_coff[F] = c;
        }
        __syncthreads();
        PHASE_MARK(1);  // clear + reference packing
        const int l0 = P.tile_off[tile * F], l1 = P.tile_off[tile * F + F];
        if (l1 > P.tile_list_cap) {  // carry-in lists overflowed their allocation: the host re-runs the batch
            if (threadIdx.x == 0) {
                atomicCAS(P.err_flag, 0, RAER_RETRY_LISTCAP);
                P.tile_rows[tile] = make_int2(0, 0);
            }
            continue;
        }
        const int n_own = s_coff[F];
        const int n_cin = (l1 - l0 + 31) >> 5;

        // ---- phase 1: coverage events + exceptional bases, one lane per read. Chunk ids: the tile's own reads, the
        // carry-in lists, then the reads set aside by the own chunks. A warp's steps last as long as its slowest lane:
        // reads with several aligned blocks (spliced, indels: n_cigar >= 3) would hold up the single-block reads they sit
        // between, so an own chunk puts them on a list and they are walked together, 32 of a kind, at the end.
        while (true) {
            int c = 0;
            if (lane == 0) c = atomicAdd(&s_chunk, 1);
            c = __shfl_sync(FULL, c, 0);
            bool have, staged = false;
            int4 a = make_int4(0, 0, 0, 0), b = make_int4(0, 0, 0, 0);
            int f = 0, n;
            unsigned dmask = 0;  // own chunk: lanes whose read was set aside
            const uint8_t *sp, *qp;
            uint32_t* mdst;
            bool scattered = false;  // masks go to their reads one by one
            if (c < n_own) {
                const int r0 = own_chunk(s_first, s_coff, F, c, f, n);
                have = lane < n;
                if (have) {
                    a = __ldg(reinterpret_cast<const int4*>(P.hdr + r0 + lane));
                    b = __ldg(reinterpret_cast<const int4*>(P.hdr + r0 + lane) + 1);
                }
                const bool dfr = have && ((a.w >> 16) & 0xffff) >= 3;
                dmask = __ballot_sync(FULL, dfr);
                if (dmask) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_ndefer, __popc(dmask));
                    base = __shfl_sync(FULL, base, 0);
                    const int slot = base + __popc(dmask & ((1u << lane) - 1u));
                    if (dfr && slot < LN_DEFER) {
                        s_defer[slot] = r0 + lane;
                        s_defer_f[slot] = (uint8_t)f;
                        have = false;
                    }
                    dmask = __ballot_sync(FULL, dfr && !have);  // a full list leaves the read with this chunk
                    __threadfence_block();
                }
                __syncwarp();
                if (lane == 0) atomicAdd(&s_own_seen, 1);
                // the chunk's packed sequences are one span of the batch's sequence array (every read starts on a
                // 4-byte boundary and occupies a multiple of 4 bytes; qualities live at twice the offsets)
                const unsigned sq = (unsigned)b.y;
                const unsigned sqe = sq + (((unsigned)(((a.w & 0xffff) + 1) >> 1) + 3u) & ~3u);
                const unsigned lo = __reduce_min_sync(FULL, have ? sq : 0xffffffffu) & ~15u;
                const unsigned hi = __reduce_max_sync(FULL, have ? sqe : 0u);
                const unsigned len = (hi - lo + 15) & ~15u;
                staged = can_stage && hi > lo && len + 32 <= (unsigned)SB;  // 32 bytes of slack behind the span (see warp_reads)
                if (staged) {
                    if (lane == 0) {
                        fence_async_smem();
                        mbar_expect_tx(w_bar, 3 * len);
                        bulk_g2s(w_seq, P.seq + lo, len, w_bar);
                        bulk_g2s(w_qual, P.qual + 2ull * lo, 2 * len, w_bar);
                    }
                    mbar_wait(w_bar, w_par);
                    w_par ^= 1;
                    sp = w_seq + (sq - lo);
                    qp = w_qual + 2 * (sq - lo);
                } else {
                    sp = P.seq + sq;
                    qp = P.qual + 2ull * sq;
                }
                mdst = P.pmask_own + (size_t)r0 * LN_MW;
            } else if (c < n_own + n_cin) {
                const int li0 = l0 + ((c - n_own) << 5);
                const int li = li0 + lane;
                n = min(32, l1 - li0);
                have = li < l1;
                if (have) {
                    while (f + 1 < F && li >= P.tile_off[tile * F + f + 1]) ++f;
                    const int r = __ldg(P.tile_list + li);
                    a = __ldg(reinterpret_cast<const int4*>(P.hdr + r));
                    b = __ldg(reinterpret_cast<const int4*>(P.hdr + r) + 1);
                }
                sp = P.seq + (unsigned)b.y;
                qp = P.qual + 2ull * (unsigned)b.y;
                mdst = P.pmask_cin + (size_t)li0 * LN_MW;
            } else {
                // the list is complete once every own chunk has been through its hand-over (which comes first thing in a
                // chunk: nobody waits here for long)
                if (lane == 0)
                    while (*(volatile int*)&s_own_seen < n_own) {}
                __syncwarp();
                const int nd = min(*(volatile int*)&s_ndefer, LN_DEFER);
                const int i0 = (c - n_own - n_cin) << 5;
                if (i0 >= nd) break;
                n = min(32, nd - i0);
                have = lane < n;
                int r = 0;
                if (have) {
                    r = s_defer[i0 + lane];
                    f = s_defer_f[i0 + lane];
                    a = __ldg(reinterpret_cast<const int4*>(P.hdr + r));
                    b = __ldg(reinterpret_cast<const int4*>(P.hdr + r) + 1);
                }
                sp = P.seq + (unsigned)b.y;
                qp = P.qual + 2ull * (unsigned)b.y;
                mdst = P.pmask_own + (size_t)r * LN_MW;
                scattered = true;
            }
            if (staged) warp_reads<true>(P, S, X, have, a, b, f, sp, qp, w_mask + lane * LN_MW);
            else warp_reads<false>(P, S, X, have, a, b, f, sp, qp, w_mask + lane * LN_MW);
            __syncwarp();
            if (X.want_mask) {
                if (scattered) {
                    if (have)
                        for (int i = 0; i < LN_MW; ++i) mdst[i] = w_mask[lane * LN_MW + i];
                } else {
                    for (int i = lane; i < n * LN_MW; i += 32)
                        if (!((dmask >> (i / LN_MW)) & 1u)) mdst[i] = w_mask[i];
                }
            }
            __syncwarp();
        }
        PHASE_BARRIER(8);
        PHASE_MARK(2);  // phase 1

        // ---- coverage: prefix sum of the events, one warp per (file, strand)
        for (int a = wid; a < 2 * F; a += LN_WARPS) {
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
        // every thread takes a run of consecutive positions, so one block scan places rows and flagged sites in position order
        const int KP = (W + (int)blockDim.x - 1) / (int)blockDim.x;
        const int p_lo = min((int)threadIdx.x * KP, W), p_hi = min(p_lo + KP, W);
        int my_rows = 0;  // rows in the low half, flagged sites in the high half (a tile has < 2^15 of either)
        for (int pidx = p_lo; pidx < p_hi; ++pidx) {
            int d = 0;
            if (t0 + pidx < t1) d = decide_site(P, CN, S.covered, pidx, t0 + pidx);
            S.dec[pidx] = (uint8_t)d;
            const bool fl = P.want_stats && (((d & 1) && (d & 4)) || ((d & 2) && (d & 8)) || (d & 48));
            my_rows += (d & 1) + ((d >> 1) & 1) + (fl ? 0x10000 : 0);
        }
        if (P.aei_mode) {
            // calc_AEI (reference R/calc_AEI.R:274-323): per file the 4 x 4 (REF of the row, counted base) sums of the rows
            // the tile would write; flag_row (unused here) holds the tile's sums
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
        int tile_tot;
        const int my_ex = block_exclusive_scan(my_rows, &tile_tot, s_scan);
        const int tile_rows_total = tile_tot & 0xffff, tile_flagged = tile_tot >> 16;
        if (threadIdx.x == 0) {
            long long base = (long long)atomicAdd(P.row_counter, (unsigned long long)tile_rows_total);
            if (base + tile_rows_total > P.row_cap) { atomicCAS(P.err_flag, 0, RAER_RETRY_ROWCAP); base = -1; }  // row_counter keeps the total: the host re-runs with that many
            s_rowbase = (int)base;
            P.tile_rows[tile] = make_int2((int)base, base < 0 ? 0 : tile_rows_total);
        }
        __syncthreads();
        const long long rowbase = s_rowbase;
        if (rowbase >= 0 && (tile_rows_total > 0 || tile_flagged > 0)) {
            int off = my_ex & 0xffff, rank = my_ex >> 16;  // position-sorted
            for (int pidx = p_lo; pidx < p_hi; ++pidx) {
                const int d = S.dec[pidx];
                const int n = (d & 1) + ((d >> 1) & 1);
                const bool fl = P.want_stats && (((d & 1) && (d & 4)) || ((d & 2) && (d & 8)) || (d & 48));
                S.fpre[pidx] = (uint16_t)rank;
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
                off += n;
                rank += fl ? 1 : 0;
            }
            if (threadIdx.x == 0) { s_nflag = tile_flagged; S.fpre[W] = (uint16_t)tile_flagged; }
        }
        __syncthreads();
        PHASE_MARK(4);  // decide + rows

        // ---- phases 3+4: second pass over the sites that carry a variant, `spr` sites per round. The dense counters are
        // not needed any more: the slots take their place (and the idle part of the staging area).
        const int nflag = s_nflag;
        const int HW = T2.hist_words, SW = T2.slot_words;
        const int spr = min(slot_region_words / SW, 4096);
        for (int round0 = 0; round0 < nflag; round0 += spr) {
            const int ns = min(spr, nflag - round0);
            X.round0 = round0;
            X.ns = ns;
            for (int i = wid; i < ns; i += LN_WARPS) {  // a warp per slot
                uint32_t* sl = slot0 + (size_t)i * SW;
                for (int w = lane; w < SW; w += 32) {
                    // first-seen ordinals (and the minimum IUPAC code) start at +inf, everything else at 0
                    const int fw = (w - (HW + 4)) & (RAER_SLOT_FS - 1);
                    sl[w] = (w >= HW + 4 && (fw < 4 || fw == 5 || fw == 6)) ? 0xffffffffu : 0u;
                }
            }
            if (threadIdx.x == 0) s_chunk = 0;
            __syncthreads();
            {
                uint32_t* st = smem + wid * LN_REC2_WORDS;
                while (true) {  // warps pull chunks of 32 reads: the tile's own reads, then the carry-in lists
                    int c = 0;
                    if (lane == 0) c = atomicAdd(&s_chunk, 1);
                    c = __shfl_sync(FULL, c, 0);
                    if (c >= n_own + n_cin) break;
                    if (c < n_own) {
                        int f, n;
                        const int r0 = own_chunk(s_first, s_coff, F, c, f, n);
                        warp_pairs(P, S, X, T2, lane < n ? r0 + lane : -1, f, P.pmask_own + (size_t)(r0 + lane) * LN_MW, st + lane * LN_MW);
                    } else {
                        const int li = l0 + ((c - n_own) << 5) + lane;
                        int f = 0, r = -1;
                        if (li < l1) {
                            while (f + 1 < F && li >= P.tile_off[tile * F + f + 1]) ++f;
                            r = __ldg(P.tile_list + li);
                        }
                        warp_pairs(P, S, X, T2, r, f, P.pmask_cin + (size_t)li * LN_MW, st + lane * LN_MW);
                    }
                }
            }
            PHASE_BARRIER(9);
            PHASE_MARK(5);  // second pass (slot clear + pairs)
            for (int i = wid; i < ns; i += LN_WARPS) {  // one warp per slot: integer sums, ALT order codes
                const int d = S.flag_d[round0 + i], row = S.flag_row[round0 + i];
                slot_sums_warp(P, slot0 + (size_t)i * SW, HW, ((d & 1) && (d & 4)) ? row : -1,
                               ((d & 2) && (d & 8)) ? row + (d & 1) : -1, d >> 4, t0 + (int)S.flag_pidx[round0 + i], lane);
            }
            __syncthreads();
            for (int i = threadIdx.x; i < 2 * ns; i += blockDim.x) {  // one thread per slot and strand: the fp64 tails
                const int sl = i >> 1, s2 = i & 1;
                const int d = S.flag_d[round0 + sl], row = S.flag_row[round0 + sl];
                const long long rw = s2 == 0 ? (((d & 1) && (d & 4)) ? row : -1) : (((d & 2) && (d & 8)) ? row + (d & 1) : -1);
                if (rw >= 0) slot_finish(P, slot0 + (size_t)sl * SW, HW, s2, rw);
            }
            __syncthreads();
            PHASE_MARK(6);  // statistics
        }
        __syncthreads();
    }
}

// first read of every (tile, file) that starts at or behind the tile's first position: [(n_tiles + 1) * n_files]
__global__ void k_tile_first(PlpParams P, int* first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (P.n_tiles + 1) * P.n_files) return;
    const int t = i / P.n_files, f = i % P.n_files;
    const long long target = (long long)P.t_beg + (long long)t * P.W;
    long long lo = P.file_off[f], hi = P.file_off[f + 1];
    if (t == P.n_tiles) lo = hi;  // reads that start behind the batch window belong to no tile
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if ((long long)__ldg(&P.hdr[mid].pos) < target) lo = mid + 1; else hi = mid;
    }
    first[i] = (int)lo;
}

// per-CTA shared memory of k_pileup_lane without the staging slices
extern "C" size_t raer_lane_fixed_bytes(int n_files, int W) {
    return (size_t)n_files * 8 * FWS(W) * 4 + ((FWMAX >> 3) + 12 + (FWMAX >> 5) + 4 + FWMAX) * 4 + (2 * FWMAX + 8) * 2 + 2 * FWMAX +
           LN_WARPS * 8 + LN_DEFER * 5 + 128;
}
// widest tile (multiple of 128, at most 1024) that leaves every warp a staging slice of at least `min_stage` bytes;
// *stage_bytes = the slice the remaining shared memory allows (capped)
extern "C" int raer_lane_plan(int n_files, int* stage_bytes) {
    const size_t total = 227 * 1024 - 1024;  // per CTA, minus the static part
    const size_t min_stage = (size_t)LN_REC2_WORDS * 4 + 256;
    int best = 0;
    for (int w = 128; w <= 1024; w += 128)
        if (raer_lane_fixed_bytes(n_files, w) + LN_WARPS * min_stage <= total) best = w;
    if (best == 0)
        for (int w = 32; w < 128; w += 32)
            if (raer_lane_fixed_bytes(n_files, w) + LN_WARPS * min_stage <= total) best = w;
    if (best == 0) return 0;
    size_t sb = (total - raer_lane_fixed_bytes(n_files, best)) / LN_WARPS;
    sb &= ~(size_t)127;
    if (sb > 12 * 1024) sb = 12 * 1024;
    *stage_bytes = (int)sb;
    return best;
}

#endif
