/*
 * oracle/oracle_sc.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restatement of the single-cell pileup engine behind pileup_cells():
 *   scpileup()                 src/sc-plp.c:945-1031 (set_sc_mplp_conf :755-844)
 *   run_scpileup()             src/sc-plp.c:612-750
 *   screadaln()                src/sc-plp.c:449-511
 *   process_one_site()         src/sc-plp.c:528-597
 *   count_record()             src/sc-plp.c:235-309
 *   count_consensus_base()     src/sc-plp.c:321-348, resolve_consensus_bases :361-387
 *   write_counts()             src/sc-plp.c:410-437, write_all_sites :852-869, write_barcodes :846-850
 * Output line order inside the matrix file follows first-touch order of the barcodes at a site
 * instead of khash order; R builds a sparse matrix from the triplets so the order carries no meaning.
 */
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "oracle_utils.h"
#include "plp_sim.h"

typedef struct {
    char* umi;
    int ref, alt, oth;
} umi_ent_t;

typedef struct {
    int ref, alt;
    umi_ent_t* u;
    int n_u, m_u;
    int touched;
} cb_ent_t;

/* string -> int open-addressing map for the barcode whitelist (cbidx) */
typedef struct {
    int nb;
    const char** keys;
    int* vals;
} smap_t;
static unsigned shash(const char* s) {
    unsigned h = 2166136261u;
    for (; *s; ++s) h = (h ^ (unsigned char)*s) * 16777619u;
    return h;
}
static void smap_init(smap_t* m, int n) {
    int nb = 16;
    while (nb < 2 * n + 1) nb <<= 1;
    m->nb = nb;
    m->keys = (const char**)calloc(nb, sizeof(char*));
    m->vals = (int*)calloc(nb, sizeof(int));
}
static void smap_put_first(smap_t* m, const char* k, int v) {
    unsigned i = shash(k) & (m->nb - 1);
    while (m->keys[i]) {
        if (!strcmp(m->keys[i], k)) return; /* duplicate barcode: first index wins, src/sc-plp.c:820 */
        i = (i + 1) & (m->nb - 1);
    }
    m->keys[i] = k;
    m->vals[i] = v;
}
static int smap_get(const smap_t* m, const char* k) {
    unsigned i = shash(k) & (m->nb - 1);
    while (m->keys[i]) {
        if (!strcmp(m->keys[i], k)) return m->vals[i];
        i = (i + 1) & (m->nb - 1);
    }
    return 0;
}

typedef struct {
    int min_mq, libtype, min_bq, max_depth, rq_minq;
    double rq_pct;
    unsigned keep_flag[2];
    ofilter_t ef;
    const oregidx_t* reg_idx;
    smap_t cbidx;
    int has_cbidx, has_umi, remove_overlaps, min_counts, is_ss2;
    const char *umi_tag, *cb_tag;
    FILE* fps[3];
    cb_ent_t* cb; /* indexed by 1-based column */
    int* touched;
    int n_touched;
} sconf_t;

typedef struct {
    bam_file_t* fp;
    const sconf_t* conf;
    int has_iter, it_tid, it_beg, it_end, it_done;
} saux_t;

static int screadaln(void* data, rec_t* b) {
    saux_t* g = (saux_t*)data;
    int ret;
    while (1) {
        if (g->has_iter) {
            if (g->it_done) return -1;
            while (1) {
                ret = bamf_read1(g->fp, b);
                if (ret < 0) break;
                if (b->tid != g->it_tid || b->pos >= g->it_end) {
                    if (b->tid < 0 || b->tid > g->it_tid || (b->tid == g->it_tid && b->pos >= g->it_end)) { ret = -1; g->it_done = 1; break; }
                    continue;
                }
                if (rec_endpos(b) > g->it_beg) break;
            }
        } else {
            ret = bamf_read1(g->fp, b);
        }
        if (ret < 0) return ret;
        if (b->tid < 0 || (b->flag & BAM_FUNMAP)) continue;
        if (g->conf->has_cbidx && g->conf->cb_tag) {
            const char* cb = rec_aux_z(b, g->conf->cb_tag);
            if (!cb) continue;
            if (!smap_get(&g->conf->cbidx, cb)) continue;
        }
        unsigned test = (g->conf->keep_flag[0] & ~(unsigned)b->flag) | (g->conf->keep_flag[1] & (unsigned)b->flag);
        if (~test & 2047u) continue;
        if (g->conf->reg_idx) {
            const char* nm = b->tid < g->fp->n_targets ? g->fp->target_name[b->tid] : "";
            if (!oreg_overlap(oreg_chr(g->conf->reg_idx, nm), b->pos, rec_endpos(b))) continue;
        }
        if (b->mapq < g->conf->min_mq) continue;
        if ((b->flag & BAM_FPAIRED) && !(b->flag & BAM_FPROPER_PAIR)) continue;
        if (g->conf->rq_pct && g->conf->rq_minq && !o_read_base_quality(b, (float)g->conf->rq_pct, g->conf->rq_minq)) continue;
        return ret;
    }
}

static void clear_site(sconf_t* c) {
    for (int t = 0; t < c->n_touched; ++t) {
        cb_ent_t* e = &c->cb[c->touched[t]];
        for (int k = 0; k < e->n_u; ++k) free(e->u[k].umi);
        e->n_u = 0;
        e->ref = e->alt = 0;
        e->touched = 0;
    }
    c->n_touched = 0;
}

/* count_record: 0 missing cb/umi, 1 other, 2 ref, 3 alt */
static int count_record(sconf_t* c, const plp1_t* p, int pld_strand, const char* pld_ref, const char* pld_alt,
                        unsigned char base, int strand, const char* bamid) {
    const rec_t* b = p->b;
    const char* cb;
    if (c->is_ss2) cb = bamid;
    else {
        cb = c->cb_tag ? rec_aux_z(b, c->cb_tag) : NULL;
        if (!cb) return 0;
    }
    if (pld_strand != strand) return 1;
    int col = c->has_cbidx ? smap_get(&c->cbidx, cb) : 0;
    if (!col) return 0;
    cb_ent_t* e = &c->cb[col];
    if (!e->touched) { e->touched = 1; c->touched[c->n_touched++] = col; }
    if (c->has_umi) {
        const char* umi = rec_aux_z(b, c->umi_tag);
        if (!umi) return 0;
        umi_ent_t* u = NULL;
        for (int k = 0; k < e->n_u; ++k)
            if (!strcmp(e->u[k].umi, umi)) { u = &e->u[k]; break; }
        if (!u) {
            if (e->n_u == e->m_u) {
                e->m_u = e->m_u ? e->m_u * 2 : 4;
                e->u = (umi_ent_t*)realloc(e->u, sizeof(umi_ent_t) * e->m_u);
            }
            u = &e->u[e->n_u++];
            u->umi = strdup(umi);
            u->ref = u->alt = u->oth = 0;
        }
        int bq = p->qpos < b->l_qseq ? b->qual[p->qpos] : 0;
        if ((unsigned char)pld_ref[0] == base) { u->ref += bq; return 2; }
        if ((unsigned char)pld_alt[0] == base) { u->alt += bq; return 3; }
        u->oth += bq;
        return 1;
    }
    if ((unsigned char)pld_ref[0] == base) { e->ref++; return 2; }
    if ((unsigned char)pld_alt[0] == base) { e->alt++; return 3; }
    return 1;
}

static int run_one(sconf_t* conf, const oracle_sc_args* a, const char* bamfn, const char* index, const char* bamid) {
    int ret = 0;
    saux_t aux;
    memset(&aux, 0, sizeof(aux));
    aux.fp = bamf_open(bamfn);
    if (!aux.fp) return -1;
    aux.conf = conf;
    int beg0 = 0, end0 = INT_MAX;
    if (a->region) {
        int rb, re;
        int nl = parse_region(a->region, &rb, &re);
        if (nl < 0) { bamf_close(aux.fp); return -1; }
        char nm[256];
        int tid = bamf_name2tid(aux.fp, a->region);
        if (tid >= 0) { rb = 0; re = INT_MAX; }
        else {
            if (nl > 255) nl = 255;
            memcpy(nm, a->region, nl);
            nm[nl] = 0;
            tid = bamf_name2tid(aux.fp, nm);
        }
        if (tid < 0) { bamf_close(aux.fp); return -1; }
        aux.has_iter = 1; aux.it_tid = tid; aux.it_beg = rb; aux.it_end = re;
        beg0 = rb; end0 = re;
        uint64_t voff;
        char bai[4096];
        if (index) snprintf(bai, sizeof(bai), "%s", index);
        else snprintf(bai, sizeof(bai), "%s.bai", bamfn);
        int br = bai_query_start(bai, tid, rb, &voff);
        if (br == 0) { if (bamf_seek(aux.fp, voff) < 0) { bamf_close(aux.fp); return -1; } }
        else if (br == 1) aux.it_done = 1;
        else { bamf_close(aux.fp); return -1; }
    }
    void* dv = &aux;
    mplp_t* iter = mplp_init(1, screadaln, &dv, conf->remove_overlaps ? (a->overlap_mode ? a->overlap_mode : 1) : 0,
                             conf->max_depth);
    int tid, pos, n_plp;
    const plp1_t* plp;
    while ((ret = mplp_auto(iter, &tid, &pos, &n_plp, &plp)) > 0) {
        if (a->region && (pos < beg0 || pos >= end0)) continue;
        if (tid < 0) break;
        if (!conf->reg_idx) continue;
        const oreg_chr_t* ch = oreg_chr(conf->reg_idx, aux.fp->target_name[tid]);
        if (!oreg_overlap(ch, pos, pos)) continue;
        for (int si = oreg_first(ch, pos, pos); si < ch->n && ch->beg[si] <= pos; ++si) {
            if (ch->end[si] < pos) continue;
            int s = ch->payload[si];
            /* ---- process_one_site ---- */
            int all_ref = 0, all_alt = 0;
            clear_site(conf);
            for (int j = 0; j < n_plp; ++j) {
                const plp1_t* p = plp + j;
                const rec_t* b = p->b;
                int c = p->qpos < b->l_qseq ? nt16_str[rec_seqi(b, p->qpos)] : 'N';
                int invert = o_invert_read_orientation(b, conf->libtype);
                if (invert < 0) { ret = -1; goto done; }
                int rret = o_check_read_filters(b, p->qpos, p->is_del, p->is_refskip, &conf->ef, conf->min_bq, conf->min_mq);
                if (rret > 0) continue;
                if (invert) c = o_comp((unsigned char)c);
                int cret = count_record(conf, p, a->site_strand[s], a->site_ref[s], a->site_alt[s], (unsigned char)c,
                                        invert + 1, bamid);
                if (cret == 2) all_ref++;
                else if (cret == 3) all_alt++;
            }
            if (conf->has_umi) {
                all_ref = all_alt = 0;
                for (int t = 0; t < conf->n_touched; ++t) {
                    cb_ent_t* e = &conf->cb[conf->touched[t]];
                    for (int k = 0; k < e->n_u; ++k) {
                        umi_ent_t* u = &e->u[k];
                        if (u->oth > u->alt && u->oth > u->ref) continue;
                        if (u->alt == 0 && u->ref == 0) continue;
                        if (u->alt >= u->ref) { e->alt++; all_alt++; }
                        else { e->ref++; all_ref++; }
                    }
                }
            }
            int n_counted = all_ref + all_alt;
            if (conf->min_counts == 0 || (n_counted >= conf->min_counts && all_alt >= conf->ef.min_var_reads)) {
                for (int t = 0; t < conf->n_touched; ++t) {
                    int col = conf->touched[t];
                    cb_ent_t* e = &conf->cb[col];
                    if (e->ref == 0 && e->alt == 0) continue;
                    fprintf(conf->fps[0], "%d %d %d %d\n", a->site_idx[s], col, e->ref, e->alt);
                }
            }
        }
    }
    if (ret > 0) ret = 0;
done:
    clear_site(conf);
    mplp_destroy(iter);
    bamf_close(aux.fp);
    return ret;
}

int oracle_scpileup(const oracle_sc_args* a) {
    if (a->nbam < 1 || a->n_sites < 1 || a->n_barcodes < 1) return -1;
    sconf_t conf;
    memset(&conf, 0, sizeof(conf));
    conf.is_ss2 = a->nbam > 1;
    for (int i = 0; i < 3; ++i) conf.fps[i] = fopen(a->outfns[i], "w");
    if (!conf.fps[0] || !conf.fps[1] || !conf.fps[2]) {
        for (int i = 0; i < 3; ++i) if (conf.fps[i]) fclose(conf.fps[i]);
        return -1;
    }
    conf.max_depth = a->int_args[0];
    conf.min_bq = a->int_args[2];
    conf.ef.trim_i5p = a->int_args[3];
    conf.ef.trim_i3p = a->int_args[4];
    conf.ef.indel_dist = a->int_args[5];
    conf.ef.splice_dist = a->int_args[6];
    conf.ef.min_overhang = a->int_args[7];
    conf.ef.nmer = a->int_args[8];
    conf.ef.min_var_reads = a->int_args[9];
    conf.ef.n_mm_type = a->int_args[10];
    conf.ef.n_mm = a->int_args[11];
    conf.keep_flag[0] = (unsigned)a->int_args[12];
    conf.keep_flag[1] = (unsigned)a->int_args[13];
    conf.ef.trim_f5p = a->dbl_args[0];
    conf.ef.trim_f3p = a->dbl_args[1];
    conf.rq_pct = a->dbl_args[3];
    conf.rq_minq = (int)a->dbl_args[4];
    if (conf.min_bq < 0) conf.min_bq = 0;
    if (!conf.max_depth) conf.max_depth = 10000;
    conf.libtype = a->libtype;
    smap_init(&conf.cbidx, a->n_barcodes);
    for (int i = 0; i < a->n_barcodes; ++i) smap_put_first(&conf.cbidx, a->barcodes[i], i + 1);
    conf.has_cbidx = 1;
    conf.cb = (cb_ent_t*)calloc(a->n_barcodes + 1, sizeof(cb_ent_t));
    conf.touched = (int*)calloc(a->n_barcodes + 1, sizeof(int));
    conf.has_umi = a->umi_tag != NULL;
    conf.umi_tag = a->umi_tag;
    conf.cb_tag = a->cb_tag;
    conf.remove_overlaps = a->remove_overlaps;
    conf.min_mq = a->min_mapq < 0 ? 0 : a->min_mapq;
    conf.min_counts = a->min_counts < 0 ? 0 : a->min_counts;

    int* b0 = (int*)malloc(sizeof(int) * a->n_sites);
    for (int i = 0; i < a->n_sites; ++i) b0[i] = a->site_pos[i] - 1;
    oregidx_t* ridx = oreg_build(a->n_sites, a->site_seqnames, b0, b0);
    free(b0);
    conf.reg_idx = ridx;

    for (int i = 0; i < a->n_barcodes; ++i) fprintf(conf.fps[2], "%s\n", a->barcodes[i]);
    for (int c = 0; c < ridx->n_chr; ++c) {
        const oreg_chr_t* ch = &ridx->chr[c];
        for (int k = 0; k < ch->n; ++k) {
            int s = ch->payload[k];
            fprintf(conf.fps[1], "%d\t%s\t%d\t%d\t%s\t%s\n", a->site_idx[s], ridx->names[c], ch->beg[k] + 1,
                    a->site_strand[s], a->site_ref[s], a->site_alt[s]);
        }
    }
    int res = 0;
    if (conf.is_ss2) {
        for (int i = 0; i < a->nbam; ++i) {
            res = run_one(&conf, a, a->bampaths[i], a->indexes ? a->indexes[i] : NULL, a->barcodes[i]);
            if (res < 0) break;
        }
    } else {
        res = run_one(&conf, a, a->bampaths[0], a->indexes ? a->indexes[0] : NULL, NULL);
    }
    for (int i = 0; i < 3; ++i) fclose(conf.fps[i]);
    for (int i = 0; i <= a->n_barcodes; ++i) free(conf.cb[i].u);
    free(conf.cb);
    free(conf.touched);
    free((void*)conf.cbidx.keys);
    free(conf.cbidx.vals);
    oreg_free(ridx);
    return res;
}
