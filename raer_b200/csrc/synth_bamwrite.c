/*
 * synth_bamwrite.c -- writes a coordinate-sorted BAM (+ .bai) from struct-of-arrays records.
 * Used by the synthetic-data generator (raer_b200/synth.py) for tests and bench inputs; it is not
 * part of the pileup path. Format per SAMv1 sections 4.1 (BGZF), 4.2 (BAM), 5.2 (BAI).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

typedef struct {
    FILE* fp;
    uint8_t buf[0xff00];
    int used;
    int level;
    int64_t coff; /* file offset of the block being filled */
} bgzf_w;

static int flush_block(bgzf_w* w) {
    if (w->used == 0) return 0;
    uint8_t out[0x10000 + 64];
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, w->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -1;
    zs.next_in = w->buf;
    zs.avail_in = w->used;
    zs.next_out = out + 18;
    zs.avail_out = sizeof(out) - 18 - 8;
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); return -1; }
    int clen = (int)zs.total_out;
    deflateEnd(&zs);
    int bsize = clen + 25;
    static const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
    memcpy(out, hdr, 16);
    out[16] = bsize & 0xff;
    out[17] = (bsize >> 8) & 0xff;
    uint32_t crc = crc32(crc32(0L, NULL, 0), w->buf, w->used);
    uint32_t isz = w->used;
    uint8_t* t = out + 18 + clen;
    for (int i = 0; i < 4; ++i) { t[i] = (crc >> (8 * i)) & 0xff; t[4 + i] = (isz >> (8 * i)) & 0xff; }
    if (fwrite(out, 1, bsize + 1, w->fp) != (size_t)(bsize + 1)) return -1;
    w->coff += bsize + 1;
    w->used = 0;
    return 0;
}

static int bw(bgzf_w* w, const void* data, size_t n) {
    const uint8_t* d = (const uint8_t*)data;
    while (n) {
        size_t k = sizeof(w->buf) - w->used;
        if (k > n) k = n;
        memcpy(w->buf + w->used, d, k);
        w->used += (int)k;
        d += k;
        n -= k;
        if (w->used == (int)sizeof(w->buf) && flush_block(w) < 0) return -1;
    }
    return 0;
}

static void p32(uint8_t* p, uint32_t v) { p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; p[2] = (v >> 16) & 0xff; p[3] = v >> 24; }
static void p16(uint8_t* p, uint32_t v) { p[0] = v & 0xff; p[1] = (v >> 8) & 0xff; }

static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (int)(beg >> 14);
    if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (int)(beg >> 17);
    if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (int)(beg >> 20);
    if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (int)(beg >> 23);
    if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (int)(beg >> 26);
    return 0;
}

typedef struct { uint32_t bin; uint64_t beg, end; } chunk_t;
static int chunk_cmp(const void* a, const void* b) {
    const chunk_t *x = (const chunk_t*)a, *y = (const chunk_t*)b;
    if (x->bin != y->bin) return x->bin < y->bin ? -1 : 1;
    return x->beg < y->beg ? -1 : (x->beg > y->beg);
}

int synth_write_bam(const char* path, const char* header_text, int n_targets, const char** names, const int32_t* lens,
                    int64_t n, const int32_t* tid, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq,
                    const int32_t* mtid, const int32_t* mpos, const int32_t* isize, const int64_t* cig_off,
                    const uint32_t* cigar, const int64_t* sq_off, const uint8_t* seq_codes, const uint8_t* qual,
                    const int64_t* qn_off, const char* qnames, const int64_t* aux_off, const uint8_t* aux, int level) {
    bgzf_w* w = (bgzf_w*)calloc(1, sizeof(bgzf_w));
    w->fp = fopen(path, "wb");
    if (!w->fp) { free(w); return -1; }
    w->level = level;
    uint8_t b4[8];
    uint32_t l_text = (uint32_t)strlen(header_text);
    bw(w, "BAM\1", 4);
    p32(b4, l_text); bw(w, b4, 4);
    bw(w, header_text, l_text);
    p32(b4, (uint32_t)n_targets); bw(w, b4, 4);
    for (int i = 0; i < n_targets; ++i) {
        uint32_t ln = (uint32_t)strlen(names[i]) + 1;
        p32(b4, ln); bw(w, b4, 4);
        bw(w, names[i], ln);
        p32(b4, (uint32_t)lens[i]); bw(w, b4, 4);
    }
    if (flush_block(w) < 0) return -1;

    /* index accumulators */
    chunk_t* chunks = (chunk_t*)malloc(sizeof(chunk_t) * (n ? n : 1));
    int64_t* chunk_tid_start = (int64_t*)calloc(n_targets + 1, sizeof(int64_t));
    int64_t n_chunks = 0;
    uint64_t** lin = (uint64_t**)calloc(n_targets ? n_targets : 1, sizeof(uint64_t*));
    int* n_lin = (int*)calloc(n_targets ? n_targets : 1, sizeof(int));
    for (int t = 0; t < n_targets; ++t) {
        n_lin[t] = (lens[t] >> 14) + 1;
        lin[t] = (uint64_t*)calloc(n_lin[t], sizeof(uint64_t));
    }
    int cur_t = -1;
    uint8_t* rec = (uint8_t*)malloc(1 << 20);
    for (int64_t i = 0; i < n; ++i) {
        int nc = (int)(cig_off[i + 1] - cig_off[i]);
        int lq = (int)(sq_off[i + 1] - sq_off[i]);
        int lqn = (int)(qn_off[i + 1] - qn_off[i]) + 1;
        int laux = aux_off ? (int)(aux_off[i + 1] - aux_off[i]) : 0;
        int rlen = 0;
        for (int k = 0; k < nc; ++k) {
            uint32_t c = cigar[cig_off[i] + k];
            if ((0x3C1A7 >> ((c & 0xf) << 1)) & 2) rlen += (int)(c >> 4);
        }
        int64_t end = (int64_t)pos[i] + (rlen ? rlen : 1);
        int bsz = 32 + lqn + 4 * nc + (lq + 1) / 2 + lq + laux;
        if (bsz + 4 > (1 << 20)) return -1;
        uint8_t* p = rec;
        p32(p, (uint32_t)bsz); p += 4;
        p32(p, (uint32_t)tid[i]); p32(p + 4, (uint32_t)pos[i]);
        p[8] = (uint8_t)lqn; p[9] = mapq[i];
        p16(p + 10, (uint32_t)reg2bin(pos[i], end));
        p16(p + 12, (uint32_t)nc); p16(p + 14, flag[i]);
        p32(p + 16, (uint32_t)lq); p32(p + 20, (uint32_t)mtid[i]); p32(p + 24, (uint32_t)mpos[i]); p32(p + 28, (uint32_t)isize[i]);
        p += 32;
        memcpy(p, qnames + qn_off[i], lqn - 1); p[lqn - 1] = 0; p += lqn;
        for (int k = 0; k < nc; ++k) { p32(p, cigar[cig_off[i] + k]); p += 4; }
        memset(p, 0, (lq + 1) / 2);
        for (int k = 0; k < lq; ++k) p[k >> 1] |= (seq_codes[sq_off[i] + k] & 0xf) << ((~k & 1) << 2);
        p += (lq + 1) / 2;
        memcpy(p, qual + sq_off[i], lq); p += lq;
        if (laux) { memcpy(p, aux + aux_off[i], laux); p += laux; }
        /* keep a record inside one block when it fits so that virtual offsets stay simple */
        if (w->used + bsz + 4 > (int)sizeof(w->buf) && flush_block(w) < 0) return -1;
        uint64_t vbeg = ((uint64_t)w->coff << 16) | (uint64_t)w->used;
        if (bw(w, rec, bsz + 4) < 0) return -1;
        uint64_t vend = ((uint64_t)w->coff << 16) | (uint64_t)w->used;
        if (tid[i] >= 0 && tid[i] < n_targets) {
            int t = tid[i];
            if (t != cur_t) {
                for (int tt = cur_t + 1; tt <= t; ++tt) chunk_tid_start[tt] = n_chunks;
                cur_t = t;
            }
            chunks[n_chunks].bin = (uint32_t)reg2bin(pos[i], end);
            chunks[n_chunks].beg = vbeg;
            chunks[n_chunks].end = vend;
            ++n_chunks;
            for (int64_t wdw = pos[i] >> 14; wdw <= (end - 1) >> 14 && wdw < n_lin[t]; ++wdw)
                if (lin[t][wdw] == 0 || vbeg < lin[t][wdw]) lin[t][wdw] = vbeg;
        }
    }
    for (int tt = cur_t + 1; tt <= n_targets; ++tt) chunk_tid_start[tt] = n_chunks;
    if (flush_block(w) < 0) return -1;
    static const uint8_t eof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0,
                                    0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, w->fp);
    fclose(w->fp);
    free(rec);

    /* BAI */
    char* bai = (char*)malloc(strlen(path) + 5);
    sprintf(bai, "%s.bai", path);
    FILE* fi = fopen(bai, "wb");
    free(bai);
    if (!fi) return -1;
    fwrite("BAI\1", 1, 4, fi);
    p32(b4, (uint32_t)n_targets); fwrite(b4, 1, 4, fi);
    for (int t = 0; t < n_targets; ++t) {
        int64_t c0 = chunk_tid_start[t], c1 = chunk_tid_start[t + 1];
        qsort(chunks + c0, (size_t)(c1 - c0), sizeof(chunk_t), chunk_cmp);
        /* merge adjacent chunks of a bin, count bins */
        int64_t m = c0;
        for (int64_t k = c0; k < c1; ++k) {
            if (m > c0 && chunks[m - 1].bin == chunks[k].bin && (chunks[m - 1].end >> 16) == (chunks[k].beg >> 16)) {
                chunks[m - 1].end = chunks[k].end;
            } else {
                chunks[m++] = chunks[k];
            }
        }
        int n_bin = 0;
        for (int64_t k = c0; k < m; ++k)
            if (k == c0 || chunks[k].bin != chunks[k - 1].bin) ++n_bin;
        p32(b4, (uint32_t)n_bin); fwrite(b4, 1, 4, fi);
        for (int64_t k = c0; k < m;) {
            int64_t e = k;
            while (e < m && chunks[e].bin == chunks[k].bin) ++e;
            p32(b4, chunks[k].bin); fwrite(b4, 1, 4, fi);
            p32(b4, (uint32_t)(e - k)); fwrite(b4, 1, 4, fi);
            for (int64_t q = k; q < e; ++q) {
                uint8_t b16[16];
                for (int z = 0; z < 8; ++z) { b16[z] = (chunks[q].beg >> (8 * z)) & 0xff; b16[8 + z] = (chunks[q].end >> (8 * z)) & 0xff; }
                fwrite(b16, 1, 16, fi);
            }
            k = e;
        }
        int nl = 0;
        for (int k = 0; k < n_lin[t]; ++k)
            if (lin[t][k]) nl = k + 1;
        p32(b4, (uint32_t)nl); fwrite(b4, 1, 4, fi);
        uint64_t last = 0;
        for (int k = 0; k < nl; ++k) {
            if (lin[t][k]) last = lin[t][k];
            uint8_t b8[8];
            for (int z = 0; z < 8; ++z) b8[z] = (last >> (8 * z)) & 0xff;
            fwrite(b8, 1, 8, fi);
        }
        free(lin[t]);
    }
    fclose(fi);
    free(lin); free(n_lin); free(chunks); free(chunk_tid_start); free(w);
    return 0;
}
