/*
 * oracle/hts_min.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 * BGZF/BAM/BAI/FASTA readers restated from the SAM specification; see hts_min.h.
 */
#include "hts_min.h"

#include <ctype.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

const char nt16_str[17] = "=ACMGRSVTWYHKDBN";

static inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t le32(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
static inline uint64_t le64(const uint8_t* p) { return (uint64_t)le32(p) | ((uint64_t)le32(p + 4) << 32); }

/* ---- BGZF (SAMv1 4.1) ---- */
static int bgzf_load_block(bam_file_t* f) {
    uint8_t hdr[18];
    f->blk_addr = ftell(f->fp);
    size_t n = fread(hdr, 1, 18, f->fp);
    if (n == 0) { f->eof = 1; f->ulen = f->uoff = 0; return -1; }
    if (n != 18 || hdr[0] != 31 || hdr[1] != 139 || hdr[2] != 8 || !(hdr[3] & 4)) return -2;
    int xlen = le16(hdr + 10);
    if (xlen < 6) return -2;
    /* locate the BC subfield; the common layout has it first */
    int bsize = -1;
    uint8_t* extra = (uint8_t*)malloc(xlen);
    memcpy(extra, hdr + 12, xlen < 6 ? xlen : 6);
    if (xlen > 6 && fread(extra + 6, 1, xlen - 6, f->fp) != (size_t)(xlen - 6)) { free(extra); return -2; }
    for (int o = 0; o + 4 <= xlen;) {
        int slen = le16(extra + o + 2);
        if (extra[o] == 'B' && extra[o + 1] == 'C' && slen == 2) bsize = le16(extra + o + 4);
        o += 4 + slen;
    }
    free(extra);
    if (bsize < 0) return -2;
    int clen = bsize + 1 - 12 - xlen; /* deflate payload + crc32 + isize */
    if (clen < 8 || fread(f->cbuf, 1, clen, f->fp) != (size_t)clen) return -2;
    uint32_t isize = le32(f->cbuf + clen - 4);
    if (isize > 65536) return -2;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    zs.next_in = f->cbuf;
    zs.avail_in = clen - 8;
    zs.next_out = f->ubuf;
    zs.avail_out = 65536;
    if (inflateInit2(&zs, -15) != Z_OK) return -2;
    int zr = inflate(&zs, Z_FINISH);
    inflateEnd(&zs);
    if (zr != Z_STREAM_END || zs.total_out != isize) return -2;
    f->ulen = (int)isize;
    f->uoff = 0;
    return 0;
}

static int bgzf_read(bam_file_t* f, void* dst, int n) {
    uint8_t* d = (uint8_t*)dst;
    int got = 0;
    while (got < n) {
        if (f->uoff >= f->ulen) {
            int r = bgzf_load_block(f);
            if (r == -1) break;
            if (r < 0) return r;
            continue;
        }
        int k = f->ulen - f->uoff;
        if (k > n - got) k = n - got;
        memcpy(d + got, f->ubuf + f->uoff, k);
        f->uoff += k;
        got += k;
    }
    return got;
}

int bamf_seek(bam_file_t* f, uint64_t voffset) {
    if (fseek(f->fp, (long)(voffset >> 16), SEEK_SET) != 0) return -1;
    f->eof = 0;
    f->ulen = f->uoff = 0;
    int r = bgzf_load_block(f);
    if (r == -1) return 0;
    if (r < 0) return -1;
    f->uoff = (int)(voffset & 0xffff);
    return 0;
}

/* ---- BAM (SAMv1 4.2) ---- */
bam_file_t* bamf_open(const char* path) {
    FILE* fp = fopen(path, "rb");
    if (!fp) return NULL;
    bam_file_t* f = (bam_file_t*)calloc(1, sizeof(*f));
    f->fp = fp;
    f->cbuf = (uint8_t*)malloc(65536 + 64);
    f->ubuf = (uint8_t*)malloc(65536);
    uint8_t b[8];
    if (bgzf_read(f, b, 8) != 8 || memcmp(b, "BAM\1", 4)) { bamf_close(f); return NULL; }
    int l_text = (int)le32(b + 4);
    char* text = (char*)malloc(l_text + 1);
    if (bgzf_read(f, text, l_text) != l_text) { free(text); bamf_close(f); return NULL; }
    free(text);
    if (bgzf_read(f, b, 4) != 4) { bamf_close(f); return NULL; }
    f->n_targets = (int)le32(b);
    f->target_name = (char**)calloc(f->n_targets ? f->n_targets : 1, sizeof(char*));
    f->target_len = (int32_t*)calloc(f->n_targets ? f->n_targets : 1, sizeof(int32_t));
    for (int i = 0; i < f->n_targets; ++i) {
        if (bgzf_read(f, b, 4) != 4) { bamf_close(f); return NULL; }
        int ln = (int)le32(b);
        f->target_name[i] = (char*)malloc(ln + 1);
        if (bgzf_read(f, f->target_name[i], ln) != ln) { bamf_close(f); return NULL; }
        f->target_name[i][ln] = 0;
        if (bgzf_read(f, b, 4) != 4) { bamf_close(f); return NULL; }
        f->target_len[i] = (int32_t)le32(b);
    }
    return f;
}

void bamf_close(bam_file_t* f) {
    if (!f) return;
    if (f->fp) fclose(f->fp);
    free(f->cbuf);
    free(f->ubuf);
    if (f->target_name) {
        for (int i = 0; i < f->n_targets; ++i) free(f->target_name[i]);
        free(f->target_name);
    }
    free(f->target_len);
    free(f);
}

int bamf_name2tid(const bam_file_t* f, const char* name) {
    for (int i = 0; i < f->n_targets; ++i)
        if (!strcmp(f->target_name[i], name)) return i;
    return -1;
}

static void rec_set_ptrs(rec_t* r) {
    uint8_t* p = r->data;
    r->qname = (char*)p;
    p += r->l_qname;
    /* keep cigar 4-byte aligned by copying into an aligned slot at the end of data */
    r->cigar = (uint32_t*)(r->data + r->m_data - 4 * (r->n_cigar ? r->n_cigar : 1));
    p += 4 * r->n_cigar;
    r->seq = p;
    p += (r->l_qseq + 1) / 2;
    r->qual = p;
    p += r->l_qseq;
    r->aux = p;
}

int bamf_read1(bam_file_t* f, rec_t* r) {
    uint8_t b[36];
    int n = bgzf_read(f, b, 4);
    if (n == 0) return -1;
    if (n != 4) return -2;
    int block_size = (int)le32(b);
    if (block_size < 32) return -2;
    if (bgzf_read(f, b, 32) != 32) return -2;
    r->tid = (int32_t)le32(b);
    r->pos = (int32_t)le32(b + 4);
    r->l_qname = b[8];
    r->mapq = b[9];
    r->n_cigar = le16(b + 12);
    r->flag = le16(b + 14);
    r->l_qseq = (int)le32(b + 16);
    r->mtid = (int32_t)le32(b + 20);
    r->mpos = (int32_t)le32(b + 24);
    r->isize = (int32_t)le32(b + 28);
    int l_data = block_size - 32;
    int fixed = r->l_qname + 4 * r->n_cigar + (r->l_qseq + 1) / 2 + r->l_qseq;
    if (fixed > l_data) return -2;
    r->l_aux = l_data - fixed;
    int need = l_data + 4 * (r->n_cigar ? r->n_cigar : 1) + 8;
    need = (need + 7) & ~7;
    if (need > r->m_data) {
        r->data = (uint8_t*)realloc(r->data, need);
    }
    r->m_data = need;
    if (bgzf_read(f, r->data, l_data) != l_data) return -2;
    rec_set_ptrs(r);
    memcpy(r->cigar, r->data + r->l_qname, 4 * (size_t)r->n_cigar);
    return 0;
}

void rec_free(rec_t* r) {
    free(r->data);
    memset(r, 0, sizeof(*r));
}

void rec_copy(rec_t* dst, const rec_t* src) {
    uint8_t* d = dst->data;
    int m = dst->m_data;
    if (m < src->m_data) {
        d = (uint8_t*)realloc(d, src->m_data);
    }
    *dst = *src;
    dst->data = d;
    dst->m_data = src->m_data;
    memcpy(dst->data, src->data, src->m_data);
    rec_set_ptrs(dst);
}

int rec_rlen(const rec_t* r) {
    int l = 0;
    for (int k = 0; k < r->n_cigar; ++k)
        if (cig_type(cig_op(r->cigar[k])) & 2) l += cig_len(r->cigar[k]);
    return l;
}

int rec_endpos(const rec_t* r) {
    int rlen = (r->flag & BAM_FUNMAP) || r->n_cigar == 0 ? 1 : rec_rlen(r);
    if (rlen == 0) rlen = 1;
    return r->pos + rlen;
}

static int aux_type_size(int t) {
    switch (t) {
    case 'A': case 'c': case 'C': return 1;
    case 's': case 'S': return 2;
    case 'i': case 'I': case 'f': return 4;
    case 'd': return 8;
    default: return 0;
    }
}

const uint8_t* rec_aux_get(const rec_t* r, const char tag[2]) {
    const uint8_t* p = r->aux;
    const uint8_t* end = r->aux + r->l_aux;
    while (p + 3 <= end) {
        int hit = (p[0] == (uint8_t)tag[0] && p[1] == (uint8_t)tag[1]);
        int t = p[2];
        const uint8_t* v = p + 2;
        p += 3;
        if (t == 'Z' || t == 'H') {
            while (p < end && *p) ++p;
            ++p;
        } else if (t == 'B') {
            if (p + 5 > end) return NULL;
            int sz = aux_type_size(p[0]);
            uint32_t cnt = le32(p + 1);
            p += 5 + (size_t)sz * cnt;
        } else {
            int sz = aux_type_size(t);
            if (!sz) return NULL;
            p += sz;
        }
        if (p > end) return NULL;
        if (hit) return v;
    }
    return NULL;
}

const char* rec_aux_z(const rec_t* r, const char tag[2]) {
    const uint8_t* v = rec_aux_get(r, tag);
    if (!v) return NULL;
    if (*v != 'Z' && *v != 'H') return NULL;
    return (const char*)(v + 1);
}

/* ---- BAI (SAMv1 5.2): only the linear index is used ---- */
int bai_query_start(const char* bai_path, int tid, int beg, uint64_t* voff) {
    FILE* fp = fopen(bai_path, "rb");
    if (!fp) return -1;
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    uint8_t* d = (uint8_t*)malloc(sz);
    if (fread(d, 1, sz, fp) != (size_t)sz) { fclose(fp); free(d); return -1; }
    fclose(fp);
    int ret = -1;
    if (sz < 8 || memcmp(d, "BAI\1", 4)) goto done;
    {
        int n_ref = (int)le32(d + 4);
        long o = 8;
        for (int t = 0; t < n_ref && o + 4 <= sz; ++t) {
            int n_bin = (int)le32(d + o);
            o += 4;
            uint64_t min_chunk = UINT64_MAX;
            for (int b = 0; b < n_bin; ++b) {
                uint32_t bin = le32(d + o);
                int n_chunk = (int)le32(d + o + 4);
                o += 8;
                for (int c = 0; c < n_chunk; ++c) {
                    uint64_t cb = le64(d + o);
                    if (bin != 37450 && cb < min_chunk) min_chunk = cb;
                    o += 16;
                }
            }
            int n_intv = (int)le32(d + o);
            o += 4;
            if (t == tid) {
                int w = beg >> 14;
                uint64_t v = 0;
                if (n_intv > 0) {
                    if (w >= n_intv) w = n_intv - 1;
                    v = le64(d + o + 8L * w);
                    /* windows without alignments hold 0 in some writers: walk back */
                    while (v == 0 && w > 0) { --w; v = le64(d + o + 8L * w); }
                }
                if (v == 0 && min_chunk != UINT64_MAX) v = min_chunk;
                *voff = v;
                ret = (v != 0) ? 0 : 1; /* 1: index readable, no alignments on tid */
                goto done;
            }
            o += 8L * n_intv;
        }
    }
done:
    free(d);
    return ret;
}

/* ---- FASTA + .fai ---- */
char* fasta_fetch(const char* fa_path, const char* name, int64_t* len) {
    size_t pl = strlen(fa_path);
    char* fai = (char*)malloc(pl + 5);
    sprintf(fai, "%s.fai", fa_path);
    FILE* fp = fopen(fai, "r");
    free(fai);
    if (!fp) return NULL;
    char line[4096];
    long long slen = -1, off = 0, lb = 0, lw = 0;
    while (fgets(line, sizeof(line), fp)) {
        char* tab = strchr(line, '\t');
        if (!tab) continue;
        *tab = 0;
        if (strcmp(line, name)) continue;
        if (sscanf(tab + 1, "%lld\t%lld\t%lld\t%lld", &slen, &off, &lb, &lw) != 4) slen = -1;
        break;
    }
    fclose(fp);
    if (slen < 0) return NULL;
    fp = fopen(fa_path, "rb");
    if (!fp) return NULL;
    char* seq = (char*)malloc(slen + 1);
    fseek(fp, (long)off, SEEK_SET);
    int64_t n = 0;
    int c;
    while (n < slen && (c = fgetc(fp)) != EOF) {
        if (isgraph(c)) seq[n++] = (char)c;
    }
    fclose(fp);
    seq[n] = 0;
    *len = n;
    return seq;
}

/* hts_parse_reg semantics: last ':' splits name and range, commas ignored,
 * "name" -> [0, INT_MAX), "name:b" -> [b-1, INT_MAX), "name:b-e" -> [b-1, e) */
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
        if (q2 == q + 1) *end = INT_MAX;
        else *end = (int)e;
    } else if (*q) {
        return -1;
    }
    if (*beg > *end) return -1;
    return (int)(colon - s);
}
