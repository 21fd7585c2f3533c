// raer_walk.cuh -- the read walker shared by phase 1 (counting) and phase 3 (second pass over flagged
// sites) of k_pileup_tiles. Included from raer_kernels.cu after the helpers.
//
// A warp fetches read headers 32 at a time (one 32-byte header per lane, coalesced). The owning lane
// derives every per-read quantity (soft clips, fractional trim distances, filter bits) and appends a
// record to the warp's shared-memory stage; only reads overlapping the target span are kept. Once at
// least ST_FLUSH reads are staged the warp-wide work list is processed with full lanes:
//   phase 1: items of 8 consecutive QUERY bases (one 32-bit sequence load + one 64-bit quality load per
//            item, unrolled predicated body per aligned block);
//   phase 3: (read, flagged position) pairs, so the second pass costs (flagged sites x depth) instead of
//            a second walk over every base.
// No warp-uniform per-read work, no per-lane loops of data-dependent length.
#ifndef RAER_WALK_CUH
#define RAER_WALK_CUH

#define ST_CAP 48    // staged reads per warp (>= ST_FLUSH + 31)
#define ST_FLUSH 16  // process the stage once it holds this many reads
#define ST_STRIDE (ST_CAP * ST_WORDS + 2 * ST_CAP)  // records, work prefix sums, first-flag index

template <bool PASS2, bool FAST>
__device__ __forceinline__ void walk_item(const PlpParams& P, const TileSmem& S, const uint32_t* rec, int q0, int f,
                                          int t0, int tw, bool mm_cfg) {
    const int pos = (int)rec[ST_POS];
    const int lq = (int)(rec[ST_LEN] & 0xffff), n = (int)(rec[ST_LEN] >> 16);
    const unsigned fl = rec[ST_FLAGS];
    const uint32_t sq = rec[ST_SQ];
    const uint32_t sw = __ldg(reinterpret_cast<const uint32_t*>(P.seq + sq + (q0 >> 1)));
    const uint2 qw = __ldg(reinterpret_cast<const uint2*>(P.qual + 2ull * sq + q0));
    const int q1 = min(q0 + 8, lq);
    const bool mapq_fail = (fl & STF_MAPQ_FAIL) != 0;
    const int invert = fl & RAER_RB_INVERT;
    const int s = invert ? 1 : 0;
    const uint32_t* cig = n <= 4 ? rec + ST_CIG0 : P.cigar + rec[ST_CIGOFF];
    const int W = P.W, WP = WPAD(P.W), minbq = P.min_bq;
    const int cxor = invert ? 3 : 0;
    const unsigned cbase_s = (unsigned)__cvta_generic_to_shared(S.cnt + (size_t)((f * 2 + s) * RAER_NCLS) * WP);
    ReadCtx c;  // only the fields a path really reads survive dead-code elimination
    if (!FAST) stage_ctx(P, rec, c);
    const bool mm_on = !FAST && mm_cfg && (fl & (RAER_RB_MMFAIL | RAER_RB_MMERR));
    int x = pos, y = 0;
    for (int k = 0; k < n && y < q1; ++k) {
        const uint32_t w = cig[k];
        const int op = cg_op(w), len = cg_len(w);
        if (op == OP_M || op == OP_EQ || op == OP_X) {
            if (y + len > q0) {
                const int rel = x + (q0 - y) - t0;  // tile index query base q0 would have under this block
                const int jlo = y - q0, jhi = min(y + len, q1) - q0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int pidx = rel + j;
                    const int code = (sw >> (((j >> 1) << 3) + ((~j & 1) << 2))) & 0xf;
                    const int bq = (int)(((j < 4 ? qw.x : qw.y) >> ((j & 3) << 3)) & 0xff);
                    // inside this aligned block, inside the tile, and not a read base N (src/plp.c:718)
                    bool go = (j >= jlo) & (j < jhi) & ((unsigned)pidx < (unsigned)tw) & (code != 15);
                    if (FAST) {
                        const bool lowq = bq < minbq;
                        // src/plp.c:574 (mapq first), :580-586 (quality 0 below the threshold is dropped silently)
                        go = go & (mapq_fail | !(lowq & (bq == 0)));
                        if (go) {
                            // A=1 C=2 G=4 T=8 -> 0..3 via (code>>1)-(code>>3); complement = xor 3; other codes -> 4
                            int cls = ((code >> 1) - (code >> 3)) ^ cxor;
                            if (__popc(code) != 1) cls = RAER_CLS_OTHER;
                            const int k2 = (mapq_fail | lowq) ? RAER_CLS_X : cls;
                            // no return value needed: RED on the shared window
                            asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(cbase_s + (unsigned)((k2 * WP + CIDX(pidx, W)) << 2)) : "memory");
                        }
                    } else if (go) {
                        process_base<false>(P, S, c, f, s, mapq_fail, mm_on, 0, t0 + pidx, pidx, q0 + j, code, bq);
                    }
                }
            }
            x += len;
            y += len;
        } else if (op == OP_D || op == OP_N) {
            x += len;
        } else if (op == OP_I || op == OP_S) {
            y += len;
        }
    }
}

// One (read, flagged position) pair per lane: second-pass slot j <-> tile index flist[j].
template <bool FAST>
__device__ __forceinline__ void walk_pair(const PlpParams& P, const TileSmem& S, const uint32_t* rec, int f, int t0,
                                          const uint16_t* flist, int j, bool mm_cfg) {
    const unsigned fl = rec[ST_FLAGS];
    if (fl & STF_MAPQ_FAIL) return;  // every base of the read is nX: nothing to accumulate (src/plp.c:574)
    ReadCtx c;
    stage_ctx(P, rec, c);
    const int s = c.invert ? 1 : 0;
    const bool mm_on = !FAST && mm_cfg && (fl & (RAER_RB_MMFAIL | RAER_RB_MMERR));
    const int ordinal = (int)rec[ST_ORD];
    const int p = t0 + (int)flist[j];
    // the CIGAR operation that covers p
    int x = c.pos, y = 0, k = 0, op = -1, len = 0;
    for (; k < c.n_cigar; ++k) {
        const uint32_t w = c.cig[k];
        op = cg_op(w);
        len = cg_len(w);
        if (op == OP_M || op == OP_EQ || op == OP_X || op == OP_D || op == OP_N) {
            if (p < x + len) break;
            x += len;
            if (op != OP_D && op != OP_N) y += len;
        } else if (op == OP_I || op == OP_S) {
            y += len;
        }
    }
    if (k >= c.n_cigar) return;              // past the end of the read
    if (op == OP_D || op == OP_N) return;    // deletion / ref-skip at p: filtered silently (src/plp.c:571)
    const int q = y + (p - x);
    if (q >= c.l_qseq) return;
    const int code = seq_code(c.seq, q);
    if (code == 15) return;
    const int bq = c.qual[q];
    if (FAST) {
        if (bq < P.min_bq) return;
    } else {
        if (base_verdict(P, c, 0, q, bq) != 0) return;
        if (mm_on) {
            int pref = (p < P.ref_len) ? d_upper((unsigned char)P.ref[p]) : 'N';
            int letter = "=ACMGRSVTWYHKDBN"[code];
            if (c.invert || letter != pref) {
                if (c.bits & RAER_RB_MMERR) { atomicExch(P.err_flag, RAER_ERR_MD); return; }
                if (c.bits & RAER_RB_MMFAIL) return;
            }
        }
    }
    pass2_accumulate_slot(P, S, c, f, s, ordinal, p, j, q, code);
}

__device__ __forceinline__ int lower_bound_u16(const uint16_t* a, int n, int v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if ((int)a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <bool PASS2, bool FAST>
__device__ void scan_reads(const PlpParams& P, const TileSmem& S, int f, int2 rg, int t0, int t1, int span_lo,
                           int span_hi, int lane, int wid, int nwarp, const uint16_t* flist = nullptr, int nfl = 0) {
    uint32_t* st = S.cigs + wid * ST_STRIDE;
    uint32_t* pref = st + ST_CAP * ST_WORDS;  // inclusive prefix of work units (items / pairs) per staged read
    uint32_t* plo = pref + ST_CAP;            // phase 3: index of the read's first flagged position
    const bool need_cov = !PASS2 && P.min_depth <= 0;
    const bool mm_cfg = (P.n_mm_type > 0 || P.n_mm > 0);
    const int tw = t1 - t0;
    int nst = 0, tot = 0, nq0 = 0;
    bool uni = true;
    long long base = rg.x + (long long)wid * 32;
    while (true) {
        const bool more = base < rg.y;
        if (more) {
            long long r = base + lane;
            int nq = 0, first_flag = 0;
            int4 a = make_int4(0, 0, 0, 0);
            bool ov = false;
            if (r < rg.y) {
                a = __ldg(reinterpret_cast<const int4*>(P.hdr + r));
                ov = a.y > span_lo && a.x < span_hi;
            }
            if (PASS2 && ov) {
                // flagged positions inside [pos, end): the read's share of the pair list
                first_flag = lower_bound_u16(flist, nfl, a.x - t0);
                nq = lower_bound_u16(flist, nfl, a.y - t0) - first_flag;
                ov = nq > 0;
            }
            const unsigned m = __ballot_sync(FULL, ov);
            const int slot = nst + __popc(m & ((1u << lane) - 1u));
            if (ov) {
                const int4 b = __ldg(reinterpret_cast<const int4*>(P.hdr + r) + 1);
                const int flag = a.z & 0xffff, mapq = (a.z >> 16) & 0xff, bits = (a.z >> 24) & 0xff;
                const int lq = a.w & 0xffff, n = (a.w >> 16) & 0xffff;
                const uint32_t* cg = P.cigar + (unsigned)b.x;
                uint32_t* rec = st + slot * ST_WORDS;
                // query_start / query_end (src/plp_utils.c:52-104): soft clips at either end, hard clips skipped
                int qs = 0, qe = lq, k;
                for (k = 0; k < n; ++k) {
                    uint32_t w = __ldg(cg + k);
                    if (k < 4) rec[ST_CIG0 + k] = w;
                    int op = cg_op(w);
                    if (op == OP_H) continue;
                    if (op == OP_S) qs += cg_len(w);
                    else break;
                }
                for (++k; k < n && k < 4; ++k) rec[ST_CIG0 + k] = __ldg(cg + k);
                for (k = n - 1; k >= 0; --k) {
                    uint32_t w = __ldg(cg + k);
                    int op = cg_op(w);
                    if (op == OP_H) continue;
                    if (op == OP_S) qe -= cg_len(w);
                    else break;
                }
                unsigned fl = (unsigned)bits | ((unsigned)((flag >> 4) & 1) << 16);
                if (mapq < P.min_mapq[f]) fl |= STF_MAPQ_FAIL;
                int d5f = 0, d3f = 0;
                if (!FAST && (P.ftrim_5p > 0 || P.ftrim_3p > 0)) {  // check_variant_fpos, src/plp_utils.c:163-193
                    int ql = qe - qs;
                    if (ql <= 0) fl |= STF_FTRIM_ALL;
                    d5f = (int)floor(P.ftrim_5p * ql);
                    d3f = (int)ceil(P.ftrim_3p * ql);
                }
                rec[ST_POS] = (uint32_t)a.x;
                rec[ST_LEN] = (uint32_t)lq | ((uint32_t)n << 16);
                rec[ST_SQ] = (uint32_t)b.y;
                rec[ST_FLAGS] = fl;
                rec[ST_QSQE] = (uint32_t)qs | ((uint32_t)qe << 16);
                rec[ST_FTRIM] = (uint32_t)(d5f & 0xffff) | ((uint32_t)d3f << 16);
                rec[ST_CIGOFF] = (uint32_t)b.x;
                rec[ST_ORD] = (uint32_t)(r - P.file_off[f]);
                if (!PASS2) nq = (lq + 7) >> 3;
                plo[slot] = (uint32_t)first_flag;
            }
            int inc = ov ? nq : 0;
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += v;
            }
            if (ov) pref[slot] = (uint32_t)(tot + inc);
            if (m) {
                if (nst == 0) nq0 = __shfl_sync(FULL, nq, __ffs(m) - 1);
                uni = uni && __all_sync(FULL, !ov || nq == nq0);
            }
            nst += __popc(m);
            tot += __shfl_sync(FULL, inc, 31);
            base += (long long)nwarp * 32;
        }
        if (nst >= ST_FLUSH || (!more && nst > 0)) {
            __syncwarp();
            if (need_cov) {
                // a pileup column exists wherever a read spans the position, deletions and ref-skips included
                for (int i = 0; i < nst; ++i) {
                    const uint32_t* rec = st + i * ST_WORDS;
                    int n = (int)(rec[ST_LEN] >> 16);
                    const uint32_t* cg = n <= 4 ? rec + ST_CIG0 : P.cigar + rec[ST_CIGOFF];
                    int x = (int)rec[ST_POS];
                    for (int k = 0; k < n; ++k) {
                        uint32_t w = cg[k];
                        if (cg_type(cg_op(w)) & 2) {
                            int lo = max(x, t0), hi = min(x + cg_len(w), t1);
                            for (int p = lo + lane; p < hi; p += 32) atomicOr(&S.covered[(p - t0) >> 5], 1u << ((p - t0) & 31));
                            x += cg_len(w);
                        }
                    }
                }
            }
            const float inv_nq = uni && nq0 > 0 ? 1.0f / (float)nq0 : 0.f;
            for (int itb = 0; itb < tot; itb += 32) {
                // warp-uniform trip count: re-join the lanes that took different paths through the previous item
                __syncwarp();
                const int it = itb + lane;
                if (it >= tot) continue;
                int i, first;
                if (uni) {
                    i = __float2int_rz(((float)it + 0.5f) * inv_nq);
                    first = i * nq0;
                } else {
                    int lo = 0, hi = nst - 1;  // smallest i with pref[i] > it
                    while (lo < hi) {
                        int mid = (lo + hi) >> 1;
                        if ((int)pref[mid] > it) hi = mid; else lo = mid + 1;
                    }
                    i = lo;
                    first = i ? (int)pref[i - 1] : 0;
                }
                if (PASS2) walk_pair<FAST>(P, S, st + i * ST_WORDS, f, t0, flist, (int)plo[i] + (it - first), mm_cfg);
                else walk_item<false, FAST>(P, S, st + i * ST_WORDS, (it - first) << 3, f, t0, tw, mm_cfg);
            }
            __syncwarp();
            nst = 0;
            tot = 0;
            nq0 = 0;
            uni = true;
        }
        if (!more) break;
    }
}

#endif
