/*
 * oracle/oracle_utils.c -- TEST INFRASTRUCTURE ONLY (CPU oracle). See oracle_utils.h.
 */
#include "oracle_utils.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* complement table: reference src/plp_utils.c:438-455 (IUPAC aware, case preserving, U->A) */
static unsigned char comp_tab[256];
static int comp_ready = 0;
static void comp_init(void) {
    const char* from = "ABCDGHKMRTUVY";
    const char* to = "TVGHCDMKYAABR";
    for (int i = 0; i < 256; ++i) comp_tab[i] = (unsigned char)i;
    for (int i = 0; from[i]; ++i) {
        comp_tab[(unsigned char)from[i]] = (unsigned char)to[i];
        comp_tab[(unsigned char)tolower(from[i])] = (unsigned char)tolower(to[i]);
    }
    comp_ready = 1;
}
unsigned char o_comp(unsigned char c) {
    if (!comp_ready) comp_init();
    return comp_tab[c];
}

/* src/plp_utils.c:52-73 */
int o_query_start(const rec_t* b) {
    int off = 0;
    for (int i = 0; i < b->n_cigar; ++i) {
        int op = cig_op(b->cigar[i]);
        if (op == OP_H) {
            if (off != 0 && off != b->l_qseq) return -1;
        } else if (op == OP_S) {
            off += cig_len(b->cigar[i]);
        } else {
            break;
        }
    }
    return off;
}

/* src/plp_utils.c:75-104 */
int o_query_end(const rec_t* b) {
    int off = b->l_qseq;
    if (off == 0) return -1;
    for (int i = b->n_cigar - 1; i >= 0; --i) {
        int op = cig_op(b->cigar[i]);
        if (op == OP_H) {
            if (off != 0 && off != b->l_qseq) return -1;
        } else if (op == OP_S) {
            off -= cig_len(b->cigar[i]);
        } else {
            break;
        }
    }
    return off;
}

/* src/plp_utils.c:106-133: any run >= nmer inside [pos-nmer+1, pos+nmer-1] clipped to the contig */
int o_check_simple_repeat(const char* ref, int64_t ref_len, int pos, int nmer) {
    int n_pos = nmer * 2 - 1;
    if (n_pos < 1 || nmer < 1 || ref_len < nmer || pos < 0) return 0;
    int start = pos - nmer + 1;
    if (start < 0) { n_pos += start; start = 0; }
    if ((int64_t)n_pos + start > ref_len) n_pos = (int)(ref_len - start);
    int run = 0;
    for (int i = 0; i < n_pos; ++i) {
        if (i > 0 && ref[start + i] == ref[start + i - 1]) ++run;
        else run = 1;
        if (run >= nmer) return 1;
    }
    return 0;
}

/* src/plp_utils.c:137-159 */
int o_check_variant_pos(const rec_t* b, int qpos, int d5, int d3) {
    int qs = o_query_start(b), qe = o_query_end(b);
    if (qs < 0 || qe < 0) return -1;
    if (!(b->flag & BAM_FREVERSE)) return (qpos < d5 + qs || qe - qpos <= d3) ? 1 : 0;
    return (qe - qpos <= d5 || qpos < d3 + qs) ? 1 : 0;
}

/* src/plp_utils.c:163-193 */
int o_check_variant_fpos(const rec_t* b, int qpos, double f5, double f3) {
    int qs = o_query_start(b), qe = o_query_end(b);
    if (qs < 0 || qe < 0) return -1;
    int ql = qe - qs;
    if (ql <= 0) return 1;
    int d5 = (int)floor(f5 * ql);
    int d3 = (int)ceil(f3 * ql);
    if (!(b->flag & BAM_FREVERSE)) return (qpos < d5 + qs || qe - qpos <= d3) ? 1 : 0;
    return (qe - qpos <= d5 || qpos < d3 + qs) ? 1 : 0;
}

/* src/plp_utils.c:196-221: distance to the first ref-skip whose query offset lies in [qpos-d, qpos+d] */
int o_dist_to_splice(const rec_t* b, int qpos, int dist) {
    int ip = 0;
    for (int i = 0; i < b->n_cigar; ++i) {
        int op = cig_op(b->cigar[i]);
        if (cig_type(op) & 1) { ip += cig_len(b->cigar[i]); continue; }
        if (op == OP_N && ip >= qpos - dist && ip <= qpos + dist) return abs(ip - qpos);
    }
    return -1;
}

/* src/plp_utils.c:228-266; note the inclusive upper bound, BAM_CMATCH only, and the read of
 * cigar[i+1] which may be one past the end (there the reference reads the first seq bytes) */
int o_check_splice_overhang(const rec_t* b, int qpos, int dist) {
    int ip = 0, p_op = -1;
    for (int i = 0; i < b->n_cigar; ++i) {
        int op = cig_op(b->cigar[i]);
        int cl = cig_len(b->cigar[i]);
        if (op == OP_M && qpos >= ip && qpos <= cl + ip) {
            if (i > 0 && p_op == OP_N) return cl >= dist ? -1 : cl;
            int nop;
            if (i + 1 < b->n_cigar) {
                nop = cig_op(b->cigar[i + 1]);
            } else {
                /* one past the CIGAR: in a BAM record the packed sequence follows */
                uint32_t w = 0;
                int nb = (b->l_qseq + 1) / 2;
                for (int k = 0; k < 4 && k < nb; ++k) w |= (uint32_t)b->seq[k] << (8 * k);
                if (nb < 4) {
                    for (int k = nb; k < 4 && k - nb < b->l_qseq; ++k) w |= (uint32_t)b->qual[k - nb] << (8 * k);
                }
                nop = cig_op(w);
            }
            if (nop == OP_N) return cl >= dist ? -1 : cl;
            return 0;
        }
        p_op = op;
        if (cig_type(op) & 1) ip += cl;
    }
    return -2;
}

/* src/plp_utils.c:269-306 */
int o_dist_to_indel(const rec_t* b, int qpos, int dist) {
    int rp = 0, sp = qpos - dist, ep = qpos + dist;
    for (int i = 0; i < b->n_cigar; ++i) {
        int op = cig_op(b->cigar[i]);
        int cl = cig_len(b->cigar[i]);
        if (cig_type(op) & 1) {
            if (op == OP_I) {
                int is = rp, ie = rp + cl;
                if (sp <= ie && is <= ep) {
                    int l = qpos - ie, r = is - qpos;
                    return l > r ? l : r;
                }
            }
            rp += cl;
        }
        if (op == OP_D && rp >= sp && rp <= ep) return abs(rp - qpos);
    }
    return -1;
}

/* src/plp_utils.c:310-328: note the float threshold */
int o_read_base_quality(const rec_t* b, float pc, int mq) {
    if (!b->l_qseq) return 0;
    int lq = 0;
    for (int i = 0; i < b->l_qseq; ++i)
        if (b->qual[i] < mq) ++lq;
    double frac = (double)lq / b->l_qseq;
    return frac <= pc ? 1 : 0;
}

/* src/plp_utils.c:331-371 */
int o_invert_read_orientation(const rec_t* b, int libtype) {
    int rev = (b->flag & BAM_FREVERSE) != 0;
    if (libtype == 0) return 0;
    if (libtype != 1 && libtype != 2) return -1;
    int as_read1;
    if (!(b->flag & BAM_FPAIRED)) as_read1 = 1;
    else if (b->flag & BAM_FREAD1) as_read1 = 1;
    else if (b->flag & BAM_FREAD2) as_read1 = 0;
    else return 0;
    /* fr-first-strand: read1 forward => minus; fr-second-strand: read1 reverse => minus */
    int r1_inverts_when_rev = (libtype == 2);
    if (as_read1) return rev == r1_inverts_when_rev;
    return rev != r1_inverts_when_rev;
}

/* src/plp_utils.c:401-412 */
int o_relative_position(const rec_t* b, int qpos) {
    int qs = o_query_start(b);
    if (qs < 0) return -1;
    int qe = b->l_qseq - o_query_end(b);
    int pos = qpos + 1 - qs;
    int alen = b->l_qseq - qs - qe;
    int rpos = (int)((double)pos / (alen + 1) * 99);
    if (rpos < 0 || rpos >= 100) return -1;
    return rpos;
}

/* src/ext/htsbox/htsbox-ext.c:17-97. Returns 0 pass / no MD, >0 = distinct mismatch bases when the
 * read fails, -1 when the MD tag is inconsistent with the CIGAR. */
int o_parse_mismatches(const rec_t* b, int n_types, int n_mis) {
    if (n_types < 1 && n_mis < 1) return 0;
    const uint8_t* md = rec_aux_get(b, "MD");
    if (!md) return 0;
    int x = 0;
    for (int k = 0; k < b->n_cigar; ++k) {
        int op = cig_op(b->cigar[k]);
        if (op == OP_M || op == OP_EQ || op == OP_X) x += cig_len(b->cigar[k]);
    }
    const char* p = (*md == 'Z' || *md == 'H') ? (const char*)md + 1 : NULL;
    if (!p) return 0; /* bam_aux2Z on a non-string tag gives NULL; isdigit(NULL deref) is UB upstream */
    int y = 0, nm = 0;
    unsigned seen = 0; /* bit per 4-bit base code */
    int ntypes = 0;
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
                int code = y < b->l_qseq ? rec_seqi(b, y) : 15;
                ++nm;
                if (!(seen & (1u << code))) { seen |= 1u << code; ++ntypes; }
                ++y;
                ++p;
            }
            if (y == -1) break;
        }
    }
    if (x != y) return -1;
    if (ntypes >= n_types && nm >= n_mis) return ntypes;
    return 0;
}

/* src/plp.c:568-615 and src/sc-plp.c:171-216 */
int o_check_read_filters(const rec_t* b, int qpos, int is_del, int is_refskip, const ofilter_t* ef,
                         int min_bq, int min_mq) {
    if (is_del || is_refskip) return 2;
    if (b->mapq < min_mq) return 1;
    int bq = qpos < b->l_qseq ? b->qual[qpos] : 0;
    if (bq < min_bq) return bq == 0 ? 2 : 1;
    int res;
    if (ef->trim_f5p > 0 || ef->trim_f3p > 0) {
        res = o_check_variant_fpos(b, qpos, ef->trim_f5p, ef->trim_f3p);
        if (res < 0) return -1;
        if (res > 0) return 1;
    }
    if (ef->trim_i5p > 0 || ef->trim_i3p > 0) {
        res = o_check_variant_pos(b, qpos, ef->trim_i5p, ef->trim_i3p);
        if (res < 0) return -1;
        if (res > 0) return 1;
    }
    if (ef->splice_dist && o_dist_to_splice(b, qpos, ef->splice_dist) >= 0) return 1;
    if (ef->min_overhang) {
        res = o_check_splice_overhang(b, qpos, ef->min_overhang);
        if (res == -2) return -1;
        if (res > 0) return 1;
    }
    if (ef->indel_dist && o_dist_to_indel(b, qpos, ef->indel_dist) >= 0) return 1;
    return 0;
}

/* ---- khash emulation (htslib khash.h: 2-bit flags, quadratic-ish probe i += ++step, load 0.77) ---- */
void okh_init(okh_t* h) { memset(h, 0, sizeof(*h)); }
void okh_free(okh_t* h) {
    free(h->flags);
    free(h->keys);
    free(h->vals);
    memset(h, 0, sizeof(*h));
}
void okh_clear(okh_t* h) {
    if (h->flags) {
        memset(h->flags, 2, h->n_buckets);
        h->size = h->n_occupied = 0;
    }
}
static unsigned roundup32(unsigned x) {
    --x;
    x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;
    return ++x;
}
static void okh_resize(okh_t* h, unsigned new_n) {
    new_n = roundup32(new_n);
    if (new_n < 4) new_n = 4;
    if (h->size >= (unsigned)(new_n * 0.77 + 0.5)) return; /* requested size too small */
    unsigned char* nflags = (unsigned char*)malloc(new_n);
    memset(nflags, 2, new_n);
    if (h->n_buckets < new_n) {
        h->keys = (unsigned char*)realloc(h->keys, new_n);
        h->vals = (int*)realloc(h->vals, new_n * sizeof(int));
    }
    unsigned mask = new_n - 1;
    for (unsigned j = 0; j != h->n_buckets; ++j) {
        if (h->flags[j] != 0) continue;
        unsigned char key = h->keys[j];
        int val = h->vals[j];
        h->flags[j] = 1;
        while (1) { /* kick-out process */
            unsigned i = (unsigned)key & mask, step = 0;
            while (nflags[i] != 2) i = (i + (++step)) & mask;
            nflags[i] = 0;
            if (i < h->n_buckets && h->flags[i] == 0) {
                unsigned char tk = h->keys[i]; h->keys[i] = key; key = tk;
                int tv = h->vals[i]; h->vals[i] = val; val = tv;
                h->flags[i] = 1;
            } else {
                h->keys[i] = key;
                h->vals[i] = val;
                break;
            }
        }
    }
    if (h->n_buckets > new_n) {
        h->keys = (unsigned char*)realloc(h->keys, new_n);
        h->vals = (int*)realloc(h->vals, new_n * sizeof(int));
    }
    free(h->flags);
    h->flags = nflags;
    h->n_buckets = new_n;
    h->n_occupied = h->size;
    h->upper_bound = (unsigned)(new_n * 0.77 + 0.5);
}
unsigned okh_put(okh_t* h, unsigned char key, int* ret) {
    if (h->n_occupied >= h->upper_bound) {
        if (h->n_buckets > (h->size << 1)) okh_resize(h, h->n_buckets - 1);
        else okh_resize(h, h->n_buckets + 1);
    }
    unsigned mask = h->n_buckets - 1, step = 0;
    unsigned x, site, i, last;
    x = site = h->n_buckets;
    i = (unsigned)key & mask;
    if (h->flags[i] == 2) {
        x = i;
    } else {
        last = i;
        while (h->flags[i] != 2 && (h->flags[i] == 1 || h->keys[i] != key)) {
            if (h->flags[i] == 1) site = i;
            i = (i + (++step)) & mask;
            if (i == last) { x = site; break; }
        }
        if (x == h->n_buckets) {
            if (h->flags[i] == 2 && site != h->n_buckets) x = site;
            else x = i;
        }
    }
    if (h->flags[x] == 2) {
        h->keys[x] = key; h->flags[x] = 0; ++h->size; ++h->n_occupied; *ret = 1;
    } else if (h->flags[x] == 1) {
        h->keys[x] = key; h->flags[x] = 0; ++h->size; *ret = 2;
    } else {
        *ret = 0;
    }
    return x;
}
void okh_del(okh_t* h, unsigned x) {
    if (x != h->n_buckets && h->flags[x] == 0) { h->flags[x] = 1; --h->size; }
}

/* ---- region index ---- */
typedef struct { int b, e, p; } ivl_t;
static int ivl_cmp(const void* a, const void* b) {
    const ivl_t *x = (const ivl_t*)a, *y = (const ivl_t*)b;
    if (x->b != y->b) return x->b < y->b ? -1 : 1;
    if (x->e != y->e) return x->e < y->e ? -1 : 1;
    return x->p < y->p ? -1 : (x->p > y->p);
}
oregidx_t* oreg_build(int n, const char** seqnames, const int* beg0, const int* end0) {
    oregidx_t* r = (oregidx_t*)calloc(1, sizeof(*r));
    int* chr_of = (int*)malloc(sizeof(int) * (n ? n : 1));
    for (int i = 0; i < n; ++i) {
        int c = -1;
        /* consecutive entries usually share the contig */
        if (r->n_chr && !strcmp(r->names[r->n_chr - 1], seqnames[i])) c = r->n_chr - 1;
        for (int k = 0; c < 0 && k < r->n_chr; ++k)
            if (!strcmp(r->names[k], seqnames[i])) c = k;
        if (c < 0) {
            r->names = (char**)realloc(r->names, sizeof(char*) * (r->n_chr + 1));
            r->chr = (oreg_chr_t*)realloc(r->chr, sizeof(oreg_chr_t) * (r->n_chr + 1));
            memset(&r->chr[r->n_chr], 0, sizeof(oreg_chr_t));
            r->names[r->n_chr] = strdup(seqnames[i]);
            c = r->n_chr++;
        }
        chr_of[i] = c;
        r->chr[c].n++;
    }
    ivl_t** tmp = (ivl_t**)calloc(r->n_chr ? r->n_chr : 1, sizeof(ivl_t*));
    int* fill = (int*)calloc(r->n_chr ? r->n_chr : 1, sizeof(int));
    for (int c = 0; c < r->n_chr; ++c) tmp[c] = (ivl_t*)malloc(sizeof(ivl_t) * r->chr[c].n);
    for (int i = 0; i < n; ++i) {
        int c = chr_of[i];
        tmp[c][fill[c]].b = beg0[i];
        tmp[c][fill[c]].e = end0[i];
        tmp[c][fill[c]].p = i;
        fill[c]++;
    }
    for (int c = 0; c < r->n_chr; ++c) {
        oreg_chr_t* ch = &r->chr[c];
        qsort(tmp[c], ch->n, sizeof(ivl_t), ivl_cmp);
        ch->beg = (int*)malloc(sizeof(int) * ch->n);
        ch->end = (int*)malloc(sizeof(int) * ch->n);
        ch->payload = (int*)malloc(sizeof(int) * ch->n);
        ch->maxend = (int*)malloc(sizeof(int) * ch->n);
        int me = -1;
        for (int i = 0; i < ch->n; ++i) {
            ch->beg[i] = tmp[c][i].b;
            ch->end[i] = tmp[c][i].e;
            ch->payload[i] = tmp[c][i].p;
            if (ch->end[i] > me) me = ch->end[i];
            ch->maxend[i] = me;
        }
        free(tmp[c]);
    }
    free(tmp);
    free(fill);
    free(chr_of);
    return r;
}
void oreg_free(oregidx_t* r) {
    if (!r) return;
    for (int c = 0; c < r->n_chr; ++c) {
        free(r->names[c]);
        free(r->chr[c].beg); free(r->chr[c].end); free(r->chr[c].payload); free(r->chr[c].maxend);
    }
    free(r->names);
    free(r->chr);
    free(r);
}
const oreg_chr_t* oreg_chr(const oregidx_t* r, const char* name) {
    for (int c = 0; c < r->n_chr; ++c)
        if (!strcmp(r->names[c], name)) return &r->chr[c];
    return NULL;
}
/* number of intervals with beg <= to */
static int upper_beg(const oreg_chr_t* c, int to) {
    int lo = 0, hi = c->n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (c->beg[mid] <= to) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
int oreg_overlap(const oreg_chr_t* c, int from, int to) {
    if (!c || c->n == 0) return 0;
    int k = upper_beg(c, to);
    if (k == 0) return 0;
    return c->maxend[k - 1] >= from;
}
int oreg_first(const oreg_chr_t* c, int from, int to) {
    if (!c) return 0;
    /* first index whose prefix max end reaches from */
    int lo = 0, hi = c->n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (c->maxend[mid] >= from) hi = mid;
        else lo = mid + 1;
    }
    (void)to;
    return lo;
}
