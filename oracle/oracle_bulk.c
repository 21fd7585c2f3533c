/*
 * oracle/oracle_bulk.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).
 *
 * Restatement of the bulk pileup engine behind pileup_sites()/calc_AEI():
 *   pileup()            src/plp.c:1143-1198   (argument unpacking: set_mplp_conf :1057-1141)
 *   run_pileup()        src/plp.c:804-995
 *   readaln()           src/plp.c:32-78
 *   process_one_site()  src/plp.c:671-802
 *   count_one_record()  src/plp.c:463-561
 *   check_umi()         src/plp.c:622-646
 *   calc_biases()       src/plp.c:335-348
 *   store_counts()      src/plp.c:358-461, add_counts :230-307, get_var_string :193-215
 *   mplp_get_ref()      src/ext/samtools/samtools-ext.c:8-65
 */
#include <ctype.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"
#include "oracle_utils.h"
#include "plp_sim.h"

#define NBASE_POS 100

typedef struct {
    int total, nr, nv, na, nt, ng, nc, nn, nx;
    okh_t var;        /* variant base -> count, khash order drives the ALT string */
    char** umi;       /* UMI strings seen at this site (strset) */
    int n_umi, m_umi;
} counts_t;

typedef struct {
    int min_global_mq, min_bq, min_depth, max_depth, report_multiallelics, remove_overlaps, nbam;
    double min_af;
    int umi;
    const char* umi_tag;
    const int *min_mqs, *libtype, *only_keep_variants;
    ofilter_t ef;
    int rq_minq;
    double rq_pct;
    unsigned keep_flag[2];
    const oregidx_t* reg_idx;
} conf_t;

typedef struct {
    bam_file_t* fp;
    bam_file_t* h; /* header of the first file (src/plp.c:884-889) */
    const conf_t* conf;
    int has_iter, it_tid, it_beg, it_end, it_done;
} aux_t;

/* readaln, src/plp.c:32-78; the region iterator follows hts_itr semantics: a record is returned
 * when endpos > beg and pos < end on the query contig */
static int readaln(void* data, rec_t* b) {
    aux_t* g = (aux_t*)data;
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
        unsigned test = (g->conf->keep_flag[0] & ~(unsigned)b->flag) | (g->conf->keep_flag[1] & (unsigned)b->flag);
        if (~test & 2047u) continue;
        if (g->conf->reg_idx) {
            const char* nm = b->tid < g->h->n_targets ? g->h->target_name[b->tid] : "";
            /* inclusive end as in the reference (src/plp.c:56-60) */
            if (!oreg_overlap(oreg_chr(g->conf->reg_idx, nm), b->pos, rec_endpos(b))) continue;
        }
        if (b->mapq < g->conf->min_global_mq) continue;
        if ((b->flag & BAM_FPAIRED) && !(b->flag & BAM_FPROPER_PAIR)) continue;
        if (g->conf->rq_pct && g->conf->rq_minq && !o_read_base_quality(b, (float)g->conf->rq_pct, g->conf->rq_minq)) continue;
        return ret;
    }
}

static void counts_clear(counts_t* c) {
    c->total = c->nr = c->nv = c->na = c->nt = c->ng = c->nc = c->nn = c->nx = 0;
    okh_clear(&c->var);
    for (int i = 0; i < c->n_umi; ++i) free(c->umi[i]);
    c->n_umi = 0;
}

/* check_umi: -1 no tag, 0 seen, 1 new */
static int umi_check(counts_t* c, const rec_t* b, const char* tag) {
    const char* u = rec_aux_z(b, tag);
    if (!u) return -1;
    for (int i = 0; i < c->n_umi; ++i)
        if (!strcmp(c->umi[i], u)) return 0;
    if (c->n_umi == c->m_umi) {
        c->m_umi = c->m_umi ? c->m_umi * 2 : 16;
        c->umi = (char**)realloc(c->umi, sizeof(char*) * c->m_umi);
    }
    c->umi[c->n_umi++] = strdup(u);
    return 1;
}

static void var_string(const okh_t* h, char* out) {
    int z = 0;
    out[0] = 0;
    if (h->size > 0) {
        for (unsigned k = 0; k < h->n_buckets; ++k) {
            if (!okh_exist(h, k)) continue;
            if (z && z < 10) out[z++] = ',';
            if (z < 11) out[z++] = (char)h->keys[k];
            out[z] = 0;
        }
    } else {
        strcpy(out, "-");
    }
}

typedef struct {
    oracle_plp_result* r;
    int64_t cap;
} sink_t;

static void sink_grow(sink_t* s) {
    oracle_plp_result* r = s->r;
    int64_t ncap = s->cap ? s->cap * 2 : 4096;
    int nb = r->nbam;
    r->tid = (int32_t*)realloc(r->tid, sizeof(int32_t) * ncap);
    r->pos = (int32_t*)realloc(r->pos, sizeof(int32_t) * ncap);
    r->strand = (char*)realloc(r->strand, ncap);
    r->ref = (char*)realloc(r->ref, ncap);
    r->rpbz = (double*)realloc(r->rpbz, sizeof(double) * ncap);
    r->vdb = (double*)realloc(r->vdb, sizeof(double) * ncap);
    r->sor = (double*)realloc(r->sor, sizeof(double) * ncap);
    /* per-file blocks are row-major inside [file][row]; regrow by re-striding */
    char* nalt = (char*)calloc((size_t)nb * ncap, 12);
    int32_t* ncnt = (int32_t*)calloc((size_t)nb * ncap * 8, sizeof(int32_t));
    for (int f = 0; f < nb; ++f) {
        if (r->nrows) {
            memcpy(nalt + (size_t)f * ncap * 12, r->alt + (size_t)f * s->cap * 12, (size_t)r->nrows * 12);
            memcpy(ncnt + (size_t)f * ncap * 8, r->counts + (size_t)f * s->cap * 8, (size_t)r->nrows * 8 * sizeof(int32_t));
        }
    }
    free(r->alt);
    free(r->counts);
    r->alt = nalt;
    r->counts = ncnt;
    s->cap = ncap;
}

static void sink_add(sink_t* s, counts_t** per_file, int tid, int gpos, int ref, int strand, const double* st) {
    oracle_plp_result* r = s->r;
    if (r->nrows == s->cap) sink_grow(s);
    int64_t i = r->nrows;
    r->tid[i] = tid;
    r->pos[i] = gpos + 1;
    r->strand[i] = (char)strand;
    r->ref[i] = (char)ref;
    r->rpbz[i] = st[0];
    r->vdb[i] = st[1];
    r->sor[i] = st[2];
    for (int f = 0; f < r->nbam; ++f) {
        counts_t* c = per_file[f];
        var_string(&c->var, r->alt + ((size_t)f * s->cap + i) * 12);
        int32_t* o = r->counts + ((size_t)f * s->cap + i) * 8;
        o[0] = c->nr; o[1] = c->nv; o[2] = c->na; o[3] = c->nt;
        o[4] = c->nc; o[5] = c->ng; o[6] = c->nn; o[7] = c->nx;
    }
    r->nrows++;
}

static void sink_compact(sink_t* s) {
    /* make the [file][row] stride equal to nrows */
    oracle_plp_result* r = s->r;
    if (s->cap == r->nrows || r->nrows == 0) return;
    for (int f = 1; f < r->nbam; ++f) {
        memmove(r->alt + (size_t)f * r->nrows * 12, r->alt + (size_t)f * s->cap * 12, (size_t)r->nrows * 12);
        memmove(r->counts + (size_t)f * r->nrows * 8, r->counts + (size_t)f * s->cap * 8, (size_t)r->nrows * 8 * sizeof(int32_t));
    }
}

/* min_allelic_freq pruning of the variant set, src/plp.c:392-420 */
static void prune_af(counts_t* c, double min_af) {
    if (c->var.size == 0) return;
    for (unsigned k = 0; k < c->var.n_buckets; ++k) {
        if (!okh_exist(&c->var, k)) continue;
        double vf = (double)c->var.vals[k] / c->total;
        if (vf < min_af) okh_del(&c->var, k);
    }
}

int oracle_pileup(const oracle_plp_args* a, oracle_plp_result* out) {
    int ret = 0, i;
    memset(out, 0, sizeof(*out));
    if (a->nbam <= 0) return -1;

    conf_t conf;
    memset(&conf, 0, sizeof(conf));
    conf.nbam = a->nbam;
    conf.max_depth = a->int_args[0];
    conf.min_depth = a->int_args[1];
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
    conf.min_af = a->dbl_args[2];
    conf.rq_pct = a->dbl_args[3];
    conf.rq_minq = (int)a->dbl_args[4];
    conf.report_multiallelics = a->lgl_args[0];
    conf.remove_overlaps = a->lgl_args[1];
    conf.libtype = a->libtype;
    conf.only_keep_variants = a->only_keep_variants;
    conf.min_mqs = a->min_mapq;
    if (a->min_mapq[0] > 0) { /* src/plp.c:1124-1133 */
        int mq = a->min_mapq[0];
        for (i = 0; i < a->nbam; ++i)
            if (a->min_mapq[i] < mq) mq = a->min_mapq[i];
        conf.min_global_mq = mq;
    }
    if (a->umi_tag) { conf.umi = 1; conf.umi_tag = a->umi_tag; }

    oregidx_t* ridx = NULL;
    if (a->n_regions > 0) {
        int* b0 = (int*)malloc(sizeof(int) * a->n_regions);
        int* e0 = (int*)malloc(sizeof(int) * a->n_regions);
        for (i = 0; i < a->n_regions; ++i) { b0[i] = a->reg_start[i] - 1; e0[i] = a->reg_end[i] - 1; }
        ridx = oreg_build(a->n_regions, a->reg_seqnames, b0, e0);
        free(b0);
        free(e0);
        conf.reg_idx = ridx;
    }

    aux_t** data = (aux_t**)calloc(a->nbam, sizeof(aux_t*));
    counts_t* pc = (counts_t*)calloc(a->nbam, sizeof(counts_t));
    counts_t* mc = (counts_t*)calloc(a->nbam, sizeof(counts_t));
    counts_t** prow = (counts_t**)calloc(a->nbam, sizeof(counts_t*));
    int* n_plp = (int*)calloc(a->nbam, sizeof(int));
    const plp1_t** plp = (const plp1_t**)calloc(a->nbam, sizeof(plp1_t*));
    mplp_t* iter = NULL;
    bam_file_t* h = NULL;
    int beg0 = 0, end0 = INT_MAX;
    char* ref[2] = {NULL, NULL};
    int ref_id[2] = {-1, -1};
    int64_t ref_len[2] = {0, 0};
    sink_t sink;
    sink.r = out;
    sink.cap = 0;
    out->nbam = a->nbam;

    for (i = 0; i < a->nbam; ++i) {
        data[i] = (aux_t*)calloc(1, sizeof(aux_t));
        data[i]->fp = bamf_open(a->bampaths[i]);
        if (!data[i]->fp) { ret = -1; goto fail; }
        data[i]->conf = &conf;
        if (i == 0) h = data[i]->fp;
        data[i]->h = h;
        if (a->region) {
            int rb, re;
            int nl = parse_region(a->region, &rb, &re);
            if (nl < 0) { ret = -1; goto fail; }
            char nm[256];
            int tid = bamf_name2tid(data[i]->fp, a->region); /* whole string may be a contig name */
            if (tid >= 0) { rb = 0; re = INT_MAX; }
            else {
                if (nl > 255) nl = 255;
                memcpy(nm, a->region, nl);
                nm[nl] = 0;
                tid = bamf_name2tid(data[i]->fp, nm);
            }
            if (tid < 0) { ret = -1; goto fail; }
            data[i]->has_iter = 1;
            data[i]->it_tid = tid;
            data[i]->it_beg = rb;
            data[i]->it_end = re;
            uint64_t voff;
            char bai[4096];
            if (a->indexes && a->indexes[i]) snprintf(bai, sizeof(bai), "%s", a->indexes[i]);
            else snprintf(bai, sizeof(bai), "%s.bai", a->bampaths[i]);
            int br = bai_query_start(bai, tid, rb, &voff);
            if (br == 0) {
                if (bamf_seek(data[i]->fp, voff) < 0) { ret = -1; goto fail; }
            } else if (br == 1) {
                data[i]->it_done = 1; /* no alignments on this contig */
            } else {
                ret = -1; /* reference: "fail to load bamfile index" */
                goto fail;
            }
            if (i == 0) { beg0 = rb; end0 = re; }
        }
    }
    out->n_targets = h->n_targets;
    out->target_names = (char**)calloc(h->n_targets ? h->n_targets : 1, sizeof(char*));
    for (i = 0; i < h->n_targets; ++i) out->target_names[i] = strdup(h->target_name[i]);

    for (i = 0; i < a->nbam; ++i) { okh_init(&pc[i].var); okh_init(&mc[i].var); }

    iter = mplp_init(a->nbam, readaln, (void**)data, conf.remove_overlaps ? (a->overlap_mode ? a->overlap_mode : 1) : 0,
                     conf.max_depth);

    int tid, pos;
    int p_ref_pos[NBASE_POS], p_alt_pos[NBASE_POS], m_ref_pos[NBASE_POS], m_alt_pos[NBASE_POS];
    while ((ret = mplp_auto(iter, &tid, &pos, n_plp, plp)) > 0) {
        if (a->region && (pos < beg0 || pos >= end0)) continue;

        /* mplp_get_ref: two-entry whole-contig cache */
        char* rseq;
        int64_t rlen;
        if (tid == ref_id[0]) { rseq = ref[0]; rlen = ref_len[0]; }
        else if (tid == ref_id[1]) {
            char* t = ref[0]; ref[0] = ref[1]; ref[1] = t;
            int ti = ref_id[0]; ref_id[0] = ref_id[1]; ref_id[1] = ti;
            int64_t tl = ref_len[0]; ref_len[0] = ref_len[1]; ref_len[1] = tl;
            rseq = ref[0]; rlen = ref_len[0];
        } else {
            free(ref[1]);
            ref[1] = ref[0]; ref_id[1] = ref_id[0]; ref_len[1] = ref_len[0];
            ref_id[0] = tid;
            ref[0] = a->fapath ? fasta_fetch(a->fapath, h->target_name[tid], &ref_len[0]) : NULL;
            if (!ref[0]) { ref_id[0] = -1; ref_len[0] = 0; }
            rseq = ref[0]; rlen = ref_len[0];
        }
        if (tid < 0) break;

        const oreg_chr_t* rchr = NULL;
        if (conf.reg_idx) {
            rchr = oreg_chr(conf.reg_idx, h->target_name[tid]);
            if (!oreg_overlap(rchr, pos, pos)) continue;
        }

        /* ---- process_one_site ---- */
        int pref_b = (rseq && pos < rlen) ? rseq[pos] : 'N';
        pref_b = toupper(pref_b);
        if (pref_b == 'N') continue;
        int mref_b = o_comp((unsigned char)pref_b);

        memset(p_ref_pos, 0, sizeof(p_ref_pos)); memset(p_alt_pos, 0, sizeof(p_alt_pos));
        memset(m_ref_pos, 0, sizeof(m_ref_pos)); memset(m_alt_pos, 0, sizeof(m_alt_pos));
        int ref_fwd = 0, ref_rev = 0, alt_fwd = 0, alt_rev = 0, p_has_var = 0, m_has_var = 0;
        for (i = 0; i < a->nbam; ++i) { counts_clear(&pc[i]); counts_clear(&mc[i]); }
        double st[6] = {0, 0, 0, 0, 0, 0};

        if (conf.ef.nmer > 0 && o_check_simple_repeat(rseq, rlen, pos, conf.ef.nmer)) continue;

        int pass_reads = 0;
        for (i = 0; i < a->nbam; ++i)
            if (n_plp[i] >= conf.min_depth) pass_reads = 1;
        if (!pass_reads) continue;

        for (i = 0; i < a->nbam; ++i) {
            for (int j = 0; j < n_plp[i]; ++j) {
                const plp1_t* p = plp[i] + j;
                const rec_t* b = p->b;
                int c = p->qpos < b->l_qseq ? nt16_str[rec_seqi(b, p->qpos)] : 'N';
                if (c == 'N') continue;
                int invert = o_invert_read_orientation(b, conf.libtype[i]);
                if (invert < 0) { ret = -1; goto fail; }
                int rret = o_check_read_filters(b, p->qpos, p->is_del, p->is_refskip, &conf.ef, conf.min_bq, conf.min_mqs[i]);
                if (conf.umi) {
                    int u = umi_check(invert ? &mc[i] : &pc[i], b, conf.umi_tag);
                    if (u != 1) continue;
                }
                if (rret > 0) {
                    if (rret == 1) { if (invert) mc[i].nx++; else pc[i].nx++; }
                    continue;
                }
                int ci = invert ? o_comp((unsigned char)c) : c;
                if (conf.ef.n_mm_type > 0 || conf.ef.n_mm > 0) {
                    if ((invert && mref_b != ci) || pref_b != ci) {
                        int m = o_parse_mismatches(b, conf.ef.n_mm_type, conf.ef.n_mm);
                        if (m == -1) { ret = -1; goto fail; }
                        if (m > 0) continue;
                    }
                }
                int rev = (b->flag & BAM_FREVERSE) != 0;
                if (pref_b == c) { if (rev) ref_rev++; else ref_fwd++; }
                else { if (rev) alt_rev++; else alt_fwd++; }

                /* count_one_record */
                int rpos = o_relative_position(b, p->qpos);
                if (rpos < 0) { ret = -1; goto fail; }
                counts_t* cc = invert ? &mc[i] : &pc[i];
                int refb = invert ? mref_b : pref_b;
                cc->total++;
                if (refb == ci) {
                    cc->nr++;
                    if (invert) m_ref_pos[rpos]++; else p_ref_pos[rpos]++;
                } else {
                    int hr;
                    unsigned k = okh_put(&cc->var, (unsigned char)ci, &hr);
                    if (hr == 0) cc->var.vals[k] += 1; else cc->var.vals[k] = 1;
                    cc->nv++;
                    if (invert) { m_alt_pos[rpos]++; m_has_var = 1; }
                    else { p_alt_pos[rpos]++; p_has_var = 1; }
                }
                switch (ci) {
                case 'A': cc->na++; break;
                case 'C': cc->nc++; break;
                case 'G': cc->ng++; break;
                case 'T': cc->nt++; break;
                default: cc->nn++; break;
                }
            }
        }

        /* calc_biases */
        if (p_has_var) {
            st[0] = oracle_calc_mwu_biasZ(p_ref_pos, p_alt_pos, NBASE_POS, 0, 1);
            st[1] = oracle_calc_vdb(p_alt_pos, NBASE_POS);
            st[2] = oracle_calc_sor(ref_fwd, ref_rev, alt_fwd, alt_rev);
        }
        if (m_has_var) {
            st[3] = oracle_calc_mwu_biasZ(m_ref_pos, m_alt_pos, NBASE_POS, 0, 1);
            st[4] = oracle_calc_vdb(m_alt_pos, NBASE_POS);
            st[5] = oracle_calc_sor(ref_fwd, ref_rev, alt_fwd, alt_rev);
        }

        /* store_counts */
        int write_only_v = 0, write_p = 0, write_m = 0, p_has_v = 0, m_has_v = 0, p_is_ma = 0, m_is_ma = 0;
        for (i = 0; i < a->nbam; ++i)
            if (conf.only_keep_variants[i]) write_only_v = 1;
        for (i = 0; i < a->nbam; ++i) {
            if (pc[i].total >= conf.min_depth && pc[i].nv >= conf.ef.min_var_reads) write_p = 1;
            if (mc[i].total >= conf.min_depth && mc[i].nv >= conf.ef.min_var_reads) write_m = 1;
            if (conf.min_af > 0) { prune_af(&pc[i], conf.min_af); prune_af(&mc[i], conf.min_af); }
            int vs = (int)pc[i].var.size;
            if (vs > 0) { if (conf.only_keep_variants[i]) p_has_v = 1; if (vs > 1) p_is_ma = 1; }
            vs = (int)mc[i].var.size;
            if (vs > 0) { if (conf.only_keep_variants[i]) m_has_v = 1; if (vs > 1) m_is_ma = 1; }
        }
        if (write_only_v) { write_p = p_has_v && write_p; write_m = m_has_v && write_m; }
        if (!conf.report_multiallelics) { write_p = !p_is_ma && write_p; write_m = !m_is_ma && write_m; }
        if (write_p) {
            for (i = 0; i < a->nbam; ++i) prow[i] = &pc[i];
            sink_add(&sink, prow, tid, pos, pref_b, '+', st);
        }
        if (write_m) {
            for (i = 0; i < a->nbam; ++i) prow[i] = &mc[i];
            sink_add(&sink, prow, tid, pos, mref_b, '-', st + 3);
        }
    }
    if (ret > 0) ret = 0;

fail:
    sink_compact(&sink);
    if (iter) mplp_destroy(iter);
    for (i = 0; i < a->nbam; ++i) {
        if (data[i]) { if (data[i]->fp) bamf_close(data[i]->fp); free(data[i]); }
        counts_clear(&pc[i]); counts_clear(&mc[i]);
        okh_free(&pc[i].var); okh_free(&mc[i].var);
        free(pc[i].umi); free(mc[i].umi);
    }
    free(data); free(pc); free(mc); free(prow); free(n_plp); free(plp);
    free(ref[0]); free(ref[1]);
    oreg_free(ridx);
    return ret;
}

void oracle_plp_result_free(oracle_plp_result* r) {
    if (!r) return;
    free(r->tid); free(r->pos); free(r->strand); free(r->ref);
    free(r->rpbz); free(r->vdb); free(r->sor); free(r->alt); free(r->counts);
    if (r->target_names) {
        for (int i = 0; i < r->n_targets; ++i) free(r->target_names[i]);
        free(r->target_names);
    }
    memset(r, 0, sizeof(*r));
}

int64_t oracle_count_aligned_bases(const char* bampath, int64_t* n_records) {
    bam_file_t* f = bamf_open(bampath);
    if (!f) return -1;
    rec_t r;
    memset(&r, 0, sizeof(r));
    int64_t n = 0, nr = 0;
    while (bamf_read1(f, &r) == 0) {
        ++nr;
        if (r.tid < 0 || (r.flag & BAM_FUNMAP)) continue;
        for (int k = 0; k < r.n_cigar; ++k) {
            int op = cig_op(r.cigar[k]);
            if (op == OP_M || op == OP_EQ || op == OP_X) n += cig_len(r.cigar[k]);
        }
    }
    rec_free(&r);
    bamf_close(f);
    if (n_records) *n_records = nr;
    return n;
}
