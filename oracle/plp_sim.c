/*
 * oracle/plp_sim.c -- TEST INFRASTRUCTURE ONLY (CPU oracle). See plp_sim.h for provenance.
 */
#include "plp_sim.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>

typedef struct node {
    rec_t b;
    int beg, end;
    int k, x, y; /* CIGAR cursor: op index, ref coordinate and query offset at its start */
    int ordinal;
    struct node* next;
} node_t;

/* ---- qname -> node map (htslib keeps a khash keyed by qname; only membership matters) ---- */
#define OL_NB 16384
typedef struct olent {
    node_t* node;
    struct olent* next;
} olent_t;

struct plp_iter {
    node_t *head, *tail;
    node_t* free_list;
    int cnt; /* live nodes including the tail sentinel (htslib mempool cnt) */
    int tid, pos, max_tid, max_pos;
    int is_eof, error, maxcnt;
    int overlap_mode;
    olent_t** ol;
    plp_read_f func;
    void* data;
    rec_t tmp;
    plp1_t* plp;
    int max_plp;
    int next_ordinal;
};

static unsigned x31(const char* s) {
    unsigned h = (unsigned)(int)*s;
    if (h)
        for (++s; *s; ++s) h = (h << 5) - h + (unsigned)(int)*s;
    return h;
}
static unsigned wang(unsigned key) {
    key += ~(key << 15);
    key ^= (key >> 10);
    key += (key << 3);
    key ^= (key >> 6);
    key += ~(key << 11);
    key ^= (key >> 16);
    return key;
}

static node_t* ol_get(plp_iter_t* it, const char* q, int del) {
    olent_t** pp = &it->ol[x31(q) & (OL_NB - 1)];
    for (; *pp; pp = &(*pp)->next) {
        if (!strcmp((*pp)->node->b.qname, q)) {
            node_t* n = (*pp)->node;
            if (del) {
                olent_t* e = *pp;
                *pp = e->next;
                free(e);
            }
            return n;
        }
    }
    return NULL;
}
static void ol_put(plp_iter_t* it, node_t* n) {
    olent_t* e = (olent_t*)malloc(sizeof(*e));
    unsigned h = x31(n->b.qname) & (OL_NB - 1);
    e->node = n;
    e->next = it->ol[h];
    it->ol[h] = e;
}
static void overlap_remove(plp_iter_t* it, const rec_t* b) {
    if (!it->ol) return;
    ol_get(it, b->qname, 1);
}

static node_t* mp_alloc(plp_iter_t* it) {
    node_t* n;
    ++it->cnt;
    if (it->free_list) {
        n = it->free_list;
        it->free_list = n->next;
        n->next = NULL;
    } else {
        n = (node_t*)calloc(1, sizeof(*n));
    }
    return n;
}
static void mp_free(plp_iter_t* it, node_t* n) {
    --it->cnt;
    n->next = it->free_list;
    it->free_list = n;
}

/* ---- mate overlap (htslib tweak_overlap_quality) ---- */
/* position the cursor on the M/=/X base covering reference offset *iref (relative to read start) */
static int iref2iseq_set(const uint32_t** cigar, const uint32_t* cmax, int* icig, int* iseq, int* iref) {
    int pos = *iref;
    if (pos < 0) return -1;
    *icig = 0;
    *iseq = 0;
    *iref = 0;
    while (*cigar < cmax) {
        int op = cig_op(**cigar), n = cig_len(**cigar);
        if (op == OP_S) { (*cigar)++; *iseq += n; *icig = 0; continue; }
        if (op == OP_H || op == OP_P) { (*cigar)++; *icig = 0; continue; }
        if (op == OP_M || op == OP_EQ || op == OP_X) {
            pos -= n;
            if (pos < 0) { *icig = n + pos; *iseq += *icig; *iref += *icig; return OP_M; }
            (*cigar)++; *iseq += n; *icig = 0; *iref += n;
            continue;
        }
        if (op == OP_I) { (*cigar)++; *iseq += n; *icig = 0; continue; }
        if (op == OP_D || op == OP_N) {
            pos -= n;
            if (pos < 0) pos = 0;
            (*cigar)++; *icig = 0; *iref += n;
            continue;
        }
        return -2;
    }
    *iseq = -1;
    return -1;
}
static int iref2iseq_next(const uint32_t** cigar, const uint32_t* cmax, int* icig, int* iseq, int* iref) {
    while (*cigar < cmax) {
        int op = cig_op(**cigar), n = cig_len(**cigar);
        if (op == OP_M || op == OP_EQ || op == OP_X) {
            if (*icig >= n - 1) { *icig = -1; (*cigar)++; continue; }
            (*iseq)++; (*icig)++; (*iref)++;
            return OP_M;
        }
        if (op == OP_D || op == OP_N) { (*cigar)++; (*iref) += n; *icig = -1; continue; }
        if (op == OP_I) { (*cigar)++; *iseq += n; *icig = -1; continue; }
        if (op == OP_S) { (*cigar)++; *iseq += n; *icig = -1; continue; }
        if (op == OP_H || op == OP_P) { (*cigar)++; *icig = -1; continue; }
        return -2;
    }
    *iseq = -1;
    *iref = -1;
    return -1;
}

static int tweak_overlap_quality(rec_t* a, rec_t* b, int mode) {
    const uint32_t *a_cigar = a->cigar, *a_cmax = a->cigar + a->n_cigar;
    const uint32_t *b_cigar = b->cigar, *b_cmax = b->cigar + b->n_cigar;
    int a_icig = 0, a_iseq = 0, b_icig = 0, b_iseq = 0;
    uint8_t *a_qual = a->qual, *b_qual = b->qual;
    int iref = b->pos;
    int a_iref = iref - a->pos;
    int b_iref = iref - b->pos;

    int a_ret = iref2iseq_set(&a_cigar, a_cmax, &a_icig, &a_iseq, &a_iref);
    if (a_ret < 0) return a_ret < -1 ? -1 : 0;
    int b_ret = iref2iseq_set(&b_cigar, b_cmax, &b_icig, &b_iseq, &b_iref);
    if (b_ret < 0) return b_ret < -1 ? -1 : 0;

    int amul, bmul;
    if (mode == 2) { amul = 1; bmul = 0; }
    else if (wang(x31(a->qname)) & 1) { amul = 1; bmul = 0; }
    else { amul = 0; bmul = 1; }

    int err = 0;
    while (1) {
        while (a_ret >= 0 && a_iref >= 0 && a_iref < iref - a->pos)
            a_ret = iref2iseq_next(&a_cigar, a_cmax, &a_icig, &a_iseq, &a_iref);
        if (a_ret < 0) { err = a_ret < -1 ? -1 : 0; break; }
        if (iref < a_iref + a->pos) iref = a_iref + a->pos;

        while (b_ret >= 0 && b_iref >= 0 && b_iref < iref - b->pos)
            b_ret = iref2iseq_next(&b_cigar, b_cmax, &b_icig, &b_iseq, &b_iref);
        if (b_ret < 0) { err = b_ret < -1 ? -1 : 0; break; }
        if (iref < b_iref + b->pos) iref = b_iref + b->pos;

        iref++;

        if (a_iref + a->pos != b_iref + b->pos) {
            if (mode == 2) continue; /* htslib <= 1.12: only positions matched in both mates */
            if (a_iref + a->pos < b_iref + b->pos && b_cigar > b->cigar && cig_op(b_cigar[-1]) == OP_D) {
                do {
                    a_qual[a_iseq] = amul ? (uint8_t)(a_qual[a_iseq] * 0.8) : 0;
                    a_ret = iref2iseq_next(&a_cigar, a_cmax, &a_icig, &a_iseq, &a_iref);
                    if (a_ret < 0) return -(a_ret < -1);
                } while (a_iref + a->pos < b_iref + b->pos);
            } else if (a_cigar > a->cigar && cig_op(a_cigar[-1]) == OP_D) {
                do {
                    b_qual[b_iseq] = bmul ? (uint8_t)(b_qual[b_iseq] * 0.8) : 0;
                    b_ret = iref2iseq_next(&b_cigar, b_cmax, &b_icig, &b_iseq, &b_iref);
                    if (b_ret < 0) return -(b_ret < -1);
                } while (b_iref + b->pos < a_iref + a->pos);
            } else {
                continue;
            }
        }

        if (a_iseq > a->l_qseq || b_iseq > b->l_qseq) return -1;

        if (rec_seqi(a, a_iseq) == rec_seqi(b, b_iseq)) {
            int qual = a_qual[a_iseq] + b_qual[b_iseq];
            if (qual > 200) qual = 200;
            if (mode == 2) {
                a_qual[a_iseq] = (uint8_t)qual;
                b_qual[b_iseq] = 0;
            } else {
                a_qual[a_iseq] = (uint8_t)(amul * qual);
                b_qual[b_iseq] = (uint8_t)(bmul * qual);
            }
        } else if (mode == 2) {
            if (a_qual[a_iseq] >= b_qual[b_iseq]) {
                a_qual[a_iseq] = (uint8_t)(0.8 * a_qual[a_iseq]);
                b_qual[b_iseq] = 0;
            } else {
                b_qual[b_iseq] = (uint8_t)(0.8 * b_qual[b_iseq]);
                a_qual[a_iseq] = 0;
            }
        } else {
            if (a_qual[a_iseq] > b_qual[b_iseq]) {
                a_qual[a_iseq] = (uint8_t)(0.8 * a_qual[a_iseq]);
                b_qual[b_iseq] = 0;
            } else if (a_qual[a_iseq] < b_qual[b_iseq]) {
                b_qual[b_iseq] = (uint8_t)(0.8 * b_qual[b_iseq]);
                a_qual[a_iseq] = 0;
            } else {
                a_qual[a_iseq] = (uint8_t)(amul * 0.8 * a_qual[a_iseq]);
                b_qual[b_iseq] = (uint8_t)(bmul * 0.8 * b_qual[b_iseq]);
            }
        }
    }
    return err;
}

static int overlap_push(plp_iter_t* it, node_t* node) {
    if (!it->ol) return 0;
    const rec_t* b = &node->b;
    if ((b->flag & BAM_FMUNMAP) || !(b->flag & BAM_FPROPER_PAIR)) return 0;
    long long isz = b->isize < 0 ? -(long long)b->isize : b->isize;
    if ((b->mtid >= 0 && b->tid != b->mtid) || (isz >= 2LL * b->l_qseq && b->mpos >= node->end)) return 0;

    node_t* a = ol_get(it, b->qname, 0);
    if (!a) {
        if (b->mpos >= b->pos || ((b->flag & BAM_FPAIRED) && b->mpos == -1)) ol_put(it, node);
    } else {
        int err = tweak_overlap_quality(&a->b, &node->b, it->overlap_mode);
        ol_get(it, b->qname, 1);
        return err;
    }
    return 0;
}

/* ---- single-file iterator (bam_plp_push / bam_plp64_next / bam_plp64_auto) ---- */
static plp_iter_t* plp_init(plp_read_f func, void* data, int overlap_mode, int maxcnt) {
    plp_iter_t* it = (plp_iter_t*)calloc(1, sizeof(*it));
    it->head = it->tail = mp_alloc(it);
    it->max_tid = it->max_pos = -1;
    it->maxcnt = maxcnt;
    it->func = func;
    it->data = data;
    it->overlap_mode = overlap_mode;
    if (overlap_mode) it->ol = (olent_t**)calloc(OL_NB, sizeof(olent_t*));
    return it;
}

static void plp_destroy(plp_iter_t* it) {
    node_t *p, *q;
    for (p = it->head; p; p = q) { q = p->next; rec_free(&p->b); free(p); }
    for (p = it->free_list; p; p = q) { q = p->next; rec_free(&p->b); free(p); }
    if (it->ol) {
        for (int i = 0; i < OL_NB; ++i)
            for (olent_t *e = it->ol[i], *n; e; e = n) { n = e->next; free(e); }
        free(it->ol);
    }
    rec_free(&it->tmp);
    free(it->plp);
    free(it);
}

static int plp_push(plp_iter_t* it, const rec_t* b) {
    if (it->error) return -1;
    if (!b) { it->is_eof = 1; return 0; }
    if (b->tid < 0 || (b->flag & BAM_FUNMAP)) { overlap_remove(it, b); return 0; }
    if (it->tid == b->tid && it->pos == b->pos && it->cnt > it->maxcnt) { overlap_remove(it, b); return 0; }
    rec_copy(&it->tail->b, b);
    it->tail->ordinal = it->next_ordinal++;
    it->tail->beg = b->pos;
    it->tail->end = b->pos + rec_rlen(b);
    it->tail->k = -1;
    it->tail->x = it->tail->y = 0;
    if (b->tid < it->max_tid) { it->error = 1; return -1; }
    if (b->tid == it->max_tid && it->tail->beg < it->max_pos) { it->error = 1; return -1; }
    it->max_tid = b->tid;
    it->max_pos = it->tail->beg;
    if (it->tail->end > it->pos || it->tail->b.tid > it->tid) {
        node_t* next = mp_alloc(it);
        if (overlap_push(it, it->tail) < 0) { mp_free(it, next); it->error = 1; return -1; }
        it->tail->next = next;
        it->tail = next;
    }
    return 0;
}

/* resolve_cigar2: locate pos inside the read, advancing the per-read cursor */
static void resolve_cigar(plp1_t* p, node_t* n, int pos) {
    const rec_t* b = &n->b;
    const uint32_t* cigar = b->cigar;
    int k;
    if (n->k == -1) {
        if (b->n_cigar == 1) {
            int op = cig_op(cigar[0]);
            if (op == OP_M || op == OP_EQ || op == OP_X) { n->k = 0; n->x = b->pos; n->y = 0; }
        } else {
            n->x = b->pos;
            n->y = 0;
            for (k = 0; k < b->n_cigar; ++k) {
                int op = cig_op(cigar[k]), l = cig_len(cigar[k]);
                if (op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X) break;
                else if (op == OP_I || op == OP_S) n->y += l;
            }
            n->k = k;
        }
    } else {
        int l = cig_len(cigar[n->k]);
        if (pos - n->x >= l) {
            int cur = cig_op(cigar[n->k]);
            if (cur == OP_M || cur == OP_EQ || cur == OP_X) n->y += l;
            n->x += l;
            for (k = n->k + 1; k < b->n_cigar; ++k) {
                int op = cig_op(cigar[k]);
                l = cig_len(cigar[k]);
                if (op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X) break;
                else if (op == OP_I || op == OP_S) n->y += l;
            }
            n->k = k;
        }
    }
    p->is_del = p->is_refskip = 0;
    p->qpos = 0;
    if (n->k >= 0 && n->k < b->n_cigar) {
        int op = cig_op(cigar[n->k]);
        if (op == OP_M || op == OP_EQ || op == OP_X) {
            p->qpos = n->y + (pos - n->x);
        } else if (op == OP_D || op == OP_N) {
            p->is_del = 1;
            p->qpos = n->y;
            p->is_refskip = (op == OP_N);
        }
    }
}

static const plp1_t* plp_next(plp_iter_t* it, int* _tid, int* _pos, int* _n_plp) {
    if (it->error) { *_n_plp = -1; return NULL; }
    *_n_plp = 0;
    if (it->is_eof && it->head == it->tail) return NULL;
    while (it->is_eof || it->max_tid > it->tid || (it->max_tid == it->tid && it->max_pos > it->pos)) {
        int n_plp = 0;
        node_t** pptr = &it->head;
        while (*pptr != it->tail) {
            node_t* p = *pptr;
            if (p->b.tid < it->tid || (p->b.tid == it->tid && p->end <= it->pos)) {
                overlap_remove(it, &p->b);
                *pptr = p->next;
                mp_free(it, p);
            } else {
                if (p->b.tid == it->tid && p->beg <= it->pos) {
                    if (n_plp == it->max_plp) {
                        it->max_plp = it->max_plp ? it->max_plp << 1 : 256;
                        it->plp = (plp1_t*)realloc(it->plp, sizeof(plp1_t) * it->max_plp);
                    }
                    it->plp[n_plp].b = &p->b;
                    it->plp[n_plp].ordinal = p->ordinal;
                    resolve_cigar(&it->plp[n_plp], p, it->pos);
                    ++n_plp;
                }
                pptr = &(*pptr)->next;
            }
        }
        *_n_plp = n_plp;
        *_tid = it->tid;
        *_pos = it->pos;
        if (it->head != it->tail) {
            if (it->tid > it->head->b.tid) { it->error = 1; *_n_plp = -1; return NULL; }
        }
        if (it->tid < it->head->b.tid) {
            it->tid = it->head->b.tid;
            it->pos = it->head->beg;
        } else if (it->pos < it->head->beg) {
            it->pos = it->head->beg;
        } else {
            ++it->pos;
        }
        if (n_plp) return it->plp;
        if (it->is_eof && it->head == it->tail) break;
    }
    return NULL;
}

static const plp1_t* plp_auto(plp_iter_t* it, int* _tid, int* _pos, int* _n_plp) {
    const plp1_t* plp;
    if (it->error) { *_n_plp = -1; return NULL; }
    if ((plp = plp_next(it, _tid, _pos, _n_plp)) != NULL) return plp;
    *_n_plp = 0;
    if (it->is_eof) return NULL;
    int ret;
    while ((ret = it->func(it->data, &it->tmp)) >= 0) {
        if (plp_push(it, &it->tmp) < 0) { *_n_plp = -1; return NULL; }
        if ((plp = plp_next(it, _tid, _pos, _n_plp)) != NULL) return plp;
    }
    if (ret < -1) { it->error = ret; *_n_plp = -1; return NULL; }
    if (plp_push(it, NULL) < 0) { *_n_plp = -1; return NULL; }
    if ((plp = plp_next(it, _tid, _pos, _n_plp)) != NULL) return plp;
    return NULL;
}

/* ---- multi-file merge (bam_mplp64_auto) ---- */
struct mplp {
    int n;
    int* tid;
    int* pos;
    int min_tid, min_pos; /* unsigned compare for tid like htslib's uint32_t */
    plp_iter_t** iter;
    int* n_plp;
    const plp1_t** plp;
};

mplp_t* mplp_init(int n, plp_read_f func, void** data, int overlap_mode, int maxcnt) {
    mplp_t* m = (mplp_t*)calloc(1, sizeof(*m));
    m->n = n;
    m->tid = (int*)calloc(n, sizeof(int));
    m->pos = (int*)calloc(n, sizeof(int));
    m->n_plp = (int*)calloc(n, sizeof(int));
    m->plp = (const plp1_t**)calloc(n, sizeof(plp1_t*));
    m->iter = (plp_iter_t**)calloc(n, sizeof(plp_iter_t*));
    m->min_tid = -1; /* (uint32_t)-1 */
    m->min_pos = INT_MAX;
    for (int i = 0; i < n; ++i) {
        m->pos[i] = INT_MAX;
        m->tid[i] = -1;
        m->iter[i] = plp_init(func, data[i], overlap_mode, maxcnt);
    }
    return m;
}

void mplp_destroy(mplp_t* m) {
    if (!m) return;
    for (int i = 0; i < m->n; ++i) plp_destroy(m->iter[i]);
    free(m->iter);
    free(m->tid);
    free(m->pos);
    free(m->n_plp);
    free(m->plp);
    free(m);
}

int mplp_auto(mplp_t* m, int* _tid, int* _pos, int* n_plp, const plp1_t** plp) {
    int ret = 0;
    int new_min_pos = INT_MAX;
    unsigned new_min_tid = (unsigned)-1;
    for (int i = 0; i < m->n; ++i) {
        if (m->pos[i] == m->min_pos && m->tid[i] == m->min_tid) {
            int tid, pos;
            m->plp[i] = plp_auto(m->iter[i], &tid, &pos, &m->n_plp[i]);
            if (m->iter[i]->error) return -1;
            if (m->plp[i]) { m->tid[i] = tid; m->pos[i] = pos; }
            else { m->tid[i] = 0; m->pos[i] = 0; }
        }
        if (m->plp[i]) {
            if ((unsigned)m->tid[i] < new_min_tid) {
                new_min_tid = (unsigned)m->tid[i];
                new_min_pos = m->pos[i];
            } else if ((unsigned)m->tid[i] == new_min_tid && m->pos[i] < new_min_pos) {
                new_min_pos = m->pos[i];
            }
        }
    }
    m->min_pos = new_min_pos;
    m->min_tid = (int)new_min_tid;
    if (new_min_pos == INT_MAX) return 0;
    *_tid = (int)new_min_tid;
    *_pos = new_min_pos;
    for (int i = 0; i < m->n; ++i) {
        if (m->pos[i] == m->min_pos && m->tid[i] == m->min_tid) {
            n_plp[i] = m->n_plp[i];
            plp[i] = m->plp[i];
            ++ret;
        } else {
            n_plp[i] = 0;
            plp[i] = NULL;
        }
    }
    return ret;
}
