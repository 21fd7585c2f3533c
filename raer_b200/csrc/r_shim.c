/*
 * r_shim.c -- the R-facing side of the drop-in raer.so: SEXP <-> C-ABI marshalling only.
 *
 * Where R is installed:  R CMD SHLIB -o raer.so r_shim.c -L. -lraer_b200   (see INTEGRATION.md). The build image has no
 * R: the tests compile this file against tests/r_stub/ (a small stand-in for the R API that implements every entry
 * point used here) and drive the registered routines through it, so the marshalling is compiled and exercised.
 *
 * It registers the same four .Call routines with the same arities as the reference's
 * src/init.c:12-24, validates arguments with the reference's messages (src/plp.c:998-1055,
 * src/sc-plp.c:874-940) and returns the same objects:
 *   .pileup   -> VECSXP[n+1]: [[1]] list(rpbz, vdb, sor); [[1+i]] the 13 named vectors of
 *                src/plp_data.c:285-289
 *   .scpileup -> ScalarInteger(status); results are the three text files
 * All device work happens inside libraer_b200.so; no R API is touched from worker threads and
 * Rf_error is raised only after every native resource has been released.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdlib.h>
#include <string.h>

#include "raer_b200.h"

static const char** chr_vec(SEXP x) {
    int n = LENGTH(x);
    const char** v = (const char**)R_alloc(n ? n : 1, sizeof(char*));
    for (int i = 0; i < n; ++i) v[i] = translateChar(STRING_ELT(x, i));
    return v;
}

/* Interrupt hook: R_CheckUserInterrupt longjmps, so it runs under R_ToplevelExec exactly like the reference's
 * checkInterrupt (src/plp_utils.c:7-12); the library calls this from the calling thread between read batches and
 * unwinds with RAER_ERR_INTERRUPT when it returns nonzero. */
static void chk_int_fn(void* dummy) { (void)dummy; R_CheckUserInterrupt(); }
static int poll_r(void* ctx) { (void)ctx; return R_ToplevelExec(chk_int_fn, NULL) == FALSE; }

/* The native result lives outside R's heap: an allocVector() that fails longjmps past us. An external pointer with a C
 * finalizer owns it from the moment raer_pileup() returns; the normal path frees it and clears the pointer. */
static void result_finalizer(SEXP xp) {
    raer_pileup_result* r = (raer_pileup_result*)R_ExternalPtrAddr(xp);
    if (r) {
        raer_pileup_result_free(r);
        free(r);
        R_ClearExternalPtr(xp);
    }
}

/* argument checks of .pileup (src/plp.c:998-1055) + unpacking; shared by the three pileup routines */
static void unpack_pileup_args(raer_pileup_args* a, SEXP bampaths, SEXP indexes, SEXP n, SEXP fapath, SEXP region, SEXP lst,
                               SEXP int_args, SEXP dbl_args, SEXP lgl_args, SEXP libtype, SEXP only_keep_variants, SEXP min_mapQ,
                               SEXP umi) {
    if (!isInteger(n) || LENGTH(n) != 1) Rf_error("'n' must be integer(1)");
    int nf = INTEGER(n)[0];
    if (!isString(bampaths) || LENGTH(bampaths) != nf)
        Rf_error("'bampaths' must be character vector equal in length to number of bam files");
    if (!isString(indexes) || LENGTH(indexes) != nf)
        Rf_error("'indexes' must be character vector equal in length to number of bam files");
    if (!isString(fapath) || LENGTH(fapath) != 1) Rf_error("'fapath' must be character(1)");
    if (!isString(region) || LENGTH(region) > 1) Rf_error("'region' must be character of length 0 or 1");
    if (!isInteger(int_args) || LENGTH(int_args) != 14) Rf_error("'int_args' must be integer of length 14");
    if (!isReal(dbl_args) || LENGTH(dbl_args) != 5) Rf_error("'dbl_args' must be numeric of length 5");
    if (!isLogical(lgl_args) || LENGTH(lgl_args) != 2) Rf_error("'lgl_args' must be logical of length 2");
    if (!isInteger(libtype) || LENGTH(libtype) != nf) Rf_error("'lib_type' must be integer of same length as bamfiles");
    if (!isLogical(only_keep_variants) || LENGTH(only_keep_variants) != nf)
        Rf_error("'only_keep_variants' must be logical of same length as bamfiles");
    if (!isInteger(min_mapQ) || LENGTH(min_mapQ) != nf) Rf_error("'min_mapQ' must be integer of same length as bamfiles");
    if (!isString(umi) || LENGTH(umi) > 1) Rf_error("'umi' must be character of length 0 or 1");

    memset(a, 0, sizeof(*a));
    a->nbam = nf;
    a->bampaths = chr_vec(bampaths);
    a->indexes = chr_vec(indexes);
    a->fapath = translateChar(STRING_ELT(fapath, 0));
    a->region = LENGTH(region) ? translateChar(STRING_ELT(region, 0)) : NULL;
    if (!isNull(lst) && length(VECTOR_ELT(lst, 0)) > 0) {
        if (length(lst) != 3) Rf_error("'lst' must contain seqnames, start, and end");
        SEXP sn = VECTOR_ELT(lst, 0), st = VECTOR_ELT(lst, 1), en = VECTOR_ELT(lst, 2);
        a->n_regions = LENGTH(sn);
        if (!isString(sn) || !isInteger(st) || !isInteger(en) || LENGTH(st) != a->n_regions || LENGTH(en) != a->n_regions)
            Rf_error("'seqnames', 'start' and 'end' must have equal lengths");
        a->reg_seqnames = chr_vec(sn);
        a->reg_start = INTEGER(st);
        a->reg_end = INTEGER(en);
    }
    memcpy(a->int_args, INTEGER(int_args), sizeof(a->int_args));
    memcpy(a->dbl_args, REAL(dbl_args), sizeof(a->dbl_args));
    a->lgl_args[0] = LOGICAL(lgl_args)[0];
    a->lgl_args[1] = LOGICAL(lgl_args)[1];
    a->libtype = INTEGER(libtype);
    a->only_keep_variants = LOGICAL(only_keep_variants);
    a->min_mapq = INTEGER(min_mapQ);
    a->umi_tag = LENGTH(umi) ? translateChar(STRING_ELT(umi, 0)) : NULL;
    a->poll = poll_r;
    a->poll_ctx = NULL;
}

/* runs the engine; returns the PROTECTed external pointer that owns the native result (caller UNPROTECTs 1) */
static SEXP run_pileup(const raer_pileup_args* a, raer_pileup_result** out) {
    raer_pileup_result* r = (raer_pileup_result*)calloc(1, sizeof(*r));
    if (!r) Rf_error("[raer internal] out of memory");
    int rc = raer_pileup(a, r);
    if (rc != RAER_OK) {
        free(r);
        if (rc == RAER_ERR_INTERRUPT) Rf_onintr(); /* the interrupt the poll saw, raised where R expects it; does not return */
        Rf_error("[raer internal] error detected during pileup: %s", raer_last_error());
    }
    SEXP xp = PROTECT(R_MakeExternalPtr(r, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(xp, result_finalizer, TRUE);
    *out = r;
    return xp;
}

/* ALT strings: a call has a handful of distinct values ("-", "G", "A,G", ...) over up to 1e8 rows. One CHARSXP per
 * distinct string, found through a small open-addressing table keyed by the 12 bytes of the column entry, instead of
 * one mkChar() -- a lookup in R's global string hash -- per row and file. */
#define ALT_SLOTS 512
typedef struct {
    char key[ALT_SLOTS][12];
    SEXP val[ALT_SLOTS];
    SEXP keeper; /* STRSXP that keeps the cached CHARSXPs reachable */
    int used;
} alt_cache;
static SEXP alt_chr(alt_cache* c, const char* s12) {
    unsigned h = 2166136261u;
    for (int i = 0; i < 12 && s12[i]; ++i) h = (h ^ (unsigned char)s12[i]) * 16777619u;
    for (unsigned i = h & (ALT_SLOTS - 1), probes = 0; probes < ALT_SLOTS; i = (i + 1) & (ALT_SLOTS - 1), ++probes) {
        if (!c->val[i]) {
            if (c->used >= ALT_SLOTS / 2) break; /* table full enough: fall back to R's own cache */
            strncpy(c->key[i], s12, 12);
            c->val[i] = mkChar(s12);
            SET_STRING_ELT(c->keeper, c->used++, c->val[i]);
            return c->val[i];
        }
        if (!strncmp(c->key[i], s12, 12)) return c->val[i];
    }
    return mkChar(s12);
}

static SEXP names_vec(const char* const* nm, int n) {
    SEXP v = PROTECT(allocVector(STRSXP, n));
    for (int i = 0; i < n; ++i) SET_STRING_ELT(v, i, mkChar(nm[i]));
    UNPROTECT(1);
    return v;
}

SEXP pileup(SEXP bampaths, SEXP indexes, SEXP n, SEXP fapath, SEXP region, SEXP lst, SEXP int_args, SEXP dbl_args,
            SEXP lgl_args, SEXP libtype, SEXP only_keep_variants, SEXP min_mapQ, SEXP umi) {
    raer_pileup_args a;
    unpack_pileup_args(&a, bampaths, indexes, n, fapath, region, lst, int_args, dbl_args, lgl_args, libtype, only_keep_variants,
                       min_mapQ, umi);
    const int nf = a.nbam;
    raer_pileup_result* rp;
    SEXP owner = run_pileup(&a, &rp); /* PROTECTed */
    const raer_pileup_result r = *rp;

    static const char* NM[] = {"seqname", "pos", "strand", "REF", "ALT", "nRef", "nAlt", "nA", "nT", "nC", "nG", "nN", "nX"};
    static const char* SNM[] = {"rpbz", "vdb", "sor"};
    R_xlen_t nr = (R_xlen_t)r.nrows;
    SEXP result = PROTECT(allocVector(VECSXP, nf + 1));
    {
        SEXP sd = PROTECT(allocVector(VECSXP, 3));
        SEXP rpz = PROTECT(allocVector(REALSXP, nr)), vd = PROTECT(allocVector(REALSXP, nr)), so = PROTECT(allocVector(REALSXP, nr));
        if (nr) {
            memcpy(REAL(rpz), r.rpbz, nr * sizeof(double));
            memcpy(REAL(vd), r.vdb, nr * sizeof(double));
            memcpy(REAL(so), r.sor, nr * sizeof(double));
        }
        SET_VECTOR_ELT(sd, 0, rpz); SET_VECTOR_ELT(sd, 1, vd); SET_VECTOR_ELT(sd, 2, so);
        setAttrib(sd, R_NamesSymbol, names_vec(SNM, 3));
        SET_VECTOR_ELT(result, 0, sd);
        UNPROTECT(4);
    }
    /* few distinct strings: build each CHARSXP once and reuse the pointer (hard part 3) */
    SEXP seq_chr = PROTECT(allocVector(STRSXP, r.n_targets ? r.n_targets : 1));
    for (int t = 0; t < r.n_targets; ++t) SET_STRING_ELT(seq_chr, t, mkChar(r.target_names[t]));
    SEXP plus = PROTECT(mkChar("+")), minus = PROTECT(mkChar("-"));
    SEXP base_chr[256];
    memset(base_chr, 0, sizeof(base_chr));
    SEXP keep = PROTECT(allocVector(STRSXP, 256));
    alt_cache* ac = (alt_cache*)R_alloc(1, sizeof(alt_cache));
    memset(ac, 0, sizeof(*ac));
    ac->keeper = PROTECT(allocVector(STRSXP, ALT_SLOTS));
    for (int f = 0; f < nf; ++f) {
        SEXP l = PROTECT(allocVector(VECSXP, 13));
        setAttrib(l, R_NamesSymbol, names_vec(NM, 13));
        SEXP seqname = PROTECT(allocVector(STRSXP, nr)), pos = PROTECT(allocVector(INTSXP, nr));
        SEXP strand = PROTECT(allocVector(STRSXP, nr)), ref = PROTECT(allocVector(STRSXP, nr)), alt = PROTECT(allocVector(STRSXP, nr));
        SEXP cnt[8];
        for (int k = 0; k < 8; ++k) cnt[k] = PROTECT(allocVector(INTSXP, nr));
        const char* altp = r.alt + (size_t)f * nr * 12;
        const int32_t* cp = r.counts + (size_t)f * nr * 8;
        for (R_xlen_t i = 0; i < nr; ++i) {
            SET_STRING_ELT(seqname, i, STRING_ELT(seq_chr, r.tid[i]));
            INTEGER(pos)[i] = r.pos[i];
            SET_STRING_ELT(strand, i, r.strand[i] == '+' ? plus : minus);
            unsigned char rc8 = (unsigned char)r.ref[i];
            if (!base_chr[rc8]) {
                char b[2] = {(char)rc8, 0};
                base_chr[rc8] = mkChar(b);
                SET_STRING_ELT(keep, rc8, base_chr[rc8]);
            }
            SET_STRING_ELT(ref, i, base_chr[rc8]);
            SET_STRING_ELT(alt, i, alt_chr(ac, altp + i * 12));
            for (int k = 0; k < 8; ++k) INTEGER(cnt[k])[i] = cp[i * 8 + k];
        }
        SET_VECTOR_ELT(l, 0, seqname); SET_VECTOR_ELT(l, 1, pos); SET_VECTOR_ELT(l, 2, strand);
        SET_VECTOR_ELT(l, 3, ref); SET_VECTOR_ELT(l, 4, alt);
        for (int k = 0; k < 8; ++k) SET_VECTOR_ELT(l, 5 + k, cnt[k]);
        SET_VECTOR_ELT(result, 1 + f, l);
        UNPROTECT(14);
    }
    result_finalizer(owner); /* normal path: free now, the finalizer finds a cleared pointer */
    UNPROTECT(7);
    return result;
}

/* .pileup_matrix -- SURVEY.md 8(f-1). Every file carries the same rows at this boundary (src/plp.c:445-459), so the
 * per-file GRanges + unique / sort / findOverlaps(type = "equal") + NA fill of lists_to_grs / merge_pileups /
 * fill_matrices (R/pileup.R:654-795) collapse to stacking the columns: this routine hands R the rowRanges columns once
 * and one rows x files matrix per assay. Same 13 arguments as .pileup. Result: named list
 * seqname, pos, strand, REF, rpbz, vdb, sor (row vectors) + ALT (character matrix) + nRef nAlt nA nT nC nG nN nX
 * (integer matrices), column-major like every R matrix. */
SEXP pileup_matrix(SEXP bampaths, SEXP indexes, SEXP n, SEXP fapath, SEXP region, SEXP lst, SEXP int_args, SEXP dbl_args,
                   SEXP lgl_args, SEXP libtype, SEXP only_keep_variants, SEXP min_mapQ, SEXP umi) {
    raer_pileup_args a;
    unpack_pileup_args(&a, bampaths, indexes, n, fapath, region, lst, int_args, dbl_args, lgl_args, libtype, only_keep_variants,
                       min_mapQ, umi);
    const int nf = a.nbam;
    raer_pileup_result* rp;
    SEXP owner = run_pileup(&a, &rp); /* PROTECTed */
    const raer_pileup_result r = *rp;
    const R_xlen_t nr = (R_xlen_t)r.nrows;
    static const char* NM[] = {"seqname", "pos", "strand", "REF", "rpbz", "vdb", "sor", "ALT",
                               "nRef",    "nAlt", "nA",    "nT",  "nC",   "nG",  "nN",  "nX"};
    SEXP out = PROTECT(allocVector(VECSXP, 16));
    setAttrib(out, R_NamesSymbol, names_vec(NM, 16));
    SEXP seq_chr = PROTECT(allocVector(STRSXP, r.n_targets ? r.n_targets : 1));
    for (int t = 0; t < r.n_targets; ++t) SET_STRING_ELT(seq_chr, t, mkChar(r.target_names[t]));
    SEXP plus = PROTECT(mkChar("+")), minus = PROTECT(mkChar("-"));
    SEXP keep = PROTECT(allocVector(STRSXP, 256));
    SEXP base_chr[256];
    memset(base_chr, 0, sizeof(base_chr));
    alt_cache* ac = (alt_cache*)R_alloc(1, sizeof(alt_cache));
    memset(ac, 0, sizeof(*ac));
    ac->keeper = PROTECT(allocVector(STRSXP, ALT_SLOTS));
    SEXP seqname = PROTECT(allocVector(STRSXP, nr)), pos = PROTECT(allocVector(INTSXP, nr));
    SEXP strand = PROTECT(allocVector(STRSXP, nr)), ref = PROTECT(allocVector(STRSXP, nr));
    for (R_xlen_t i = 0; i < nr; ++i) {
        SET_STRING_ELT(seqname, i, STRING_ELT(seq_chr, r.tid[i]));
        INTEGER(pos)[i] = r.pos[i];
        SET_STRING_ELT(strand, i, r.strand[i] == '+' ? plus : minus);
        unsigned char rc8 = (unsigned char)r.ref[i];
        if (!base_chr[rc8]) {
            char b[2] = {(char)rc8, 0};
            base_chr[rc8] = mkChar(b);
            SET_STRING_ELT(keep, rc8, base_chr[rc8]);
        }
        SET_STRING_ELT(ref, i, base_chr[rc8]);
    }
    SET_VECTOR_ELT(out, 0, seqname); SET_VECTOR_ELT(out, 1, pos); SET_VECTOR_ELT(out, 2, strand); SET_VECTOR_ELT(out, 3, ref);
    UNPROTECT(4);
    const double* st[3] = {r.rpbz, r.vdb, r.sor};
    for (int k = 0; k < 3; ++k) {
        SEXP v = PROTECT(allocVector(REALSXP, nr));
        if (nr) memcpy(REAL(v), st[k], nr * sizeof(double));
        SET_VECTOR_ELT(out, 4 + k, v);
        UNPROTECT(1);
    }
    SEXP dim = PROTECT(allocVector(INTSXP, 2));
    INTEGER(dim)[0] = (int)nr;
    INTEGER(dim)[1] = nf;
    {
        SEXP alt = PROTECT(allocVector(STRSXP, nr * nf));
        for (int f = 0; f < nf; ++f) {
            const char* altp = r.alt + (size_t)f * nr * 12;
            for (R_xlen_t i = 0; i < nr; ++i) SET_STRING_ELT(alt, (R_xlen_t)f * nr + i, alt_chr(ac, altp + i * 12));
        }
        setAttrib(alt, R_DimSymbol, dim);
        SET_VECTOR_ELT(out, 7, alt);
        UNPROTECT(1);
    }
    for (int k = 0; k < 8; ++k) {
        SEXP m = PROTECT(allocVector(INTSXP, nr * nf));
        int* mp = INTEGER(m);
        for (int f = 0; f < nf; ++f) {
            const int32_t* cp = r.counts + (size_t)f * nr * 8;
            for (R_xlen_t i = 0; i < nr; ++i) mp[(R_xlen_t)f * nr + i] = cp[i * 8 + k];
        }
        setAttrib(m, R_DimSymbol, dim);
        SET_VECTOR_ELT(out, 8 + k, m);
        UNPROTECT(1);
    }
    result_finalizer(owner);
    UNPROTECT(8);
    return out;
}

/* .pileup_aei -- SURVEY.md 8(f-2). calc_AEI only needs, per BAM file, the counts of the written rows summed by
 * (REF of the row, counted base) (R/calc_AEI.R:274-323): the engine reduces them on the device and no row crosses the
 * boundary. Same 13 arguments as .pileup (the site list = the Alu ranges minus the SNPs). Result: integer-valued
 * numeric array [4, 4, n]: [REF A C G T, nA nC nG nT, file]. */
SEXP pileup_aei(SEXP bampaths, SEXP indexes, SEXP n, SEXP fapath, SEXP region, SEXP lst, SEXP int_args, SEXP dbl_args,
                SEXP lgl_args, SEXP libtype, SEXP only_keep_variants, SEXP min_mapQ, SEXP umi) {
    raer_pileup_args a;
    unpack_pileup_args(&a, bampaths, indexes, n, fapath, region, lst, int_args, dbl_args, lgl_args, libtype, only_keep_variants,
                       min_mapQ, umi);
    a.aei_mode = 1;
    raer_pileup_result* rp;
    SEXP owner = run_pileup(&a, &rp); /* PROTECTed */
    SEXP out = PROTECT(allocVector(REALSXP, 16 * (R_xlen_t)a.nbam));
    for (int f = 0; f < a.nbam; ++f)
        for (int rr = 0; rr < 4; ++rr)
            for (int b = 0; b < 4; ++b) REAL(out)[f * 16 + b * 4 + rr] = rp->aei ? (double)rp->aei[f * 16 + rr * 4 + b] : 0.0;
    SEXP dim = PROTECT(allocVector(INTSXP, 3));
    INTEGER(dim)[0] = 4; INTEGER(dim)[1] = 4; INTEGER(dim)[2] = a.nbam;
    setAttrib(out, R_DimSymbol, dim);
    result_finalizer(owner);
    UNPROTECT(3);
    return out;
}

SEXP scpileup(SEXP bampaths, SEXP indexes, SEXP query_region, SEXP lst, SEXP barcodes, SEXP cbtag, SEXP int_args,
              SEXP dbl_args, SEXP libtype, SEXP outfns, SEXP umi, SEXP remove_overlaps, SEXP min_mapq, SEXP min_counts) {
    if (!isString(bampaths) || LENGTH(bampaths) < 1) Rf_error("'bampaths' must be character");
    if (!isString(indexes) || LENGTH(indexes) < 1) Rf_error("'indexes' must be character");
    if (LENGTH(bampaths) != LENGTH(indexes)) Rf_error("'bampaths' and 'indexes' must be same length");
    if (!isString(query_region) || LENGTH(query_region) > 1) Rf_error("'qregion' must be character of length 0 or 1");
    if (isNull(lst)) Rf_error("'lst' must be non-null");
    if (!isString(barcodes) || LENGTH(barcodes) < 1) Rf_error("'barcodes' must be character");
    if (!isString(cbtag) || LENGTH(cbtag) > 1) Rf_error("'cbtag' must be character of length 0 or 1");
    if (!isInteger(int_args) || LENGTH(int_args) != 14) Rf_error("'int_args' must be integer of length 14");
    if (!isReal(dbl_args) || LENGTH(dbl_args) != 5) Rf_error("'dbl_args' must be numeric of length 5");
    if (!isInteger(libtype) || LENGTH(libtype) != 1) Rf_error("'lib_type' must be integer(1)");
    if (!isString(outfns) || LENGTH(outfns) != 3) Rf_error("'outfns' must be character(3)");
    if (!isString(umi) || LENGTH(umi) > 1) Rf_error("'umi' must be character of length 0 or 1");
    if (!isLogical(remove_overlaps) || LENGTH(remove_overlaps) != 1) Rf_error("'remove_overlaps' must be logical(1)");
    if (!isInteger(min_mapq) || LENGTH(min_mapq) != 1) Rf_error("'min_mapq' must be integer(1)");
    if (!isInteger(min_counts) || LENGTH(min_counts) != 1) Rf_error("'min_counts' must be integer(1)");
    if (length(lst) != 6) Rf_error("'lst' must contain seqnames, pos, strand, ref, alt, and rowidx");

    raer_scpileup_args a;
    memset(&a, 0, sizeof(a));
    a.nbam = LENGTH(bampaths);
    a.bampaths = chr_vec(bampaths);
    a.indexes = chr_vec(indexes);
    a.region = LENGTH(query_region) ? translateChar(STRING_ELT(query_region, 0)) : NULL;
    a.n_sites = LENGTH(VECTOR_ELT(lst, 0));
    a.site_seqnames = chr_vec(VECTOR_ELT(lst, 0));
    a.site_pos = INTEGER(VECTOR_ELT(lst, 1));
    a.site_strand = INTEGER(VECTOR_ELT(lst, 2));
    a.site_ref = chr_vec(VECTOR_ELT(lst, 3));
    a.site_alt = chr_vec(VECTOR_ELT(lst, 4));
    a.site_idx = INTEGER(VECTOR_ELT(lst, 5));
    a.n_barcodes = LENGTH(barcodes);
    a.barcodes = chr_vec(barcodes);
    a.cb_tag = LENGTH(cbtag) ? translateChar(STRING_ELT(cbtag, 0)) : NULL;
    memcpy(a.int_args, INTEGER(int_args), sizeof(a.int_args));
    memcpy(a.dbl_args, REAL(dbl_args), sizeof(a.dbl_args));
    a.libtype = INTEGER(libtype)[0];
    for (int i = 0; i < 3; ++i) a.outfns[i] = R_ExpandFileName(translateChar(STRING_ELT(outfns, i)));
    a.umi_tag = LENGTH(umi) ? translateChar(STRING_ELT(umi, 0)) : NULL;
    a.remove_overlaps = LOGICAL(remove_overlaps)[0];
    a.min_mapq = INTEGER(min_mapq)[0];
    a.min_counts = INTEGER(min_counts)[0];
    a.poll = poll_r;
    a.poll_ctx = NULL;
    int res = raer_scpileup(&a);
    if (res == RAER_ERR_INTERRUPT) Rf_onintr();
    if (res < 0) REprintf("[raer internal] error detected during pileup, %d: %s\n", res, raer_last_error());
    return ScalarInteger(res);
}

SEXP get_region(SEXP region) {
    if (!isString(region) || length(region) != 1) Rf_error("'region' must be character");
    const char* cregion = translateChar(STRING_ELT(region, 0));
    char chrom[1024];
    int32_t beg, end;
    if (raer_get_region(cregion, chrom, sizeof(chrom), &beg, &end) != RAER_OK) Rf_error("could not parse region:%s", cregion);
    const char* names[] = {"chrom", "start", "end", ""};
    SEXP res = PROTECT(mkNamed(VECSXP, names));
    SET_VECTOR_ELT(res, 0, mkString(chrom));
    SET_VECTOR_ELT(res, 1, ScalarInteger(beg));
    SET_VECTOR_ELT(res, 2, ScalarInteger(end));
    UNPROTECT(1);
    return res;
}

SEXP fisher_exact(SEXP mat) {
    if (!isMatrix(mat) || nrows(mat) != 4) Rf_error("'mat' must be matrix with 4 rows");
    if (!isInteger(mat)) Rf_error("'mat' must be an integer matrix");
    int nc = ncols(mat);
    SEXP res = PROTECT(allocVector(REALSXP, nc));
    raer_fisher_exact(INTEGER(mat), nc, REAL(res));
    UNPROTECT(1);
    return res;
}

static const R_CallMethodDef CallEntries[] = {
    {".get_region", (DL_FUNC)&get_region, 1},
    {".fisher_exact", (DL_FUNC)&fisher_exact, 1},
    {".pileup", (DL_FUNC)&pileup, 13},
    {".scpileup", (DL_FUNC)&scpileup, 14},
    /* additions: the reference's R code does not call them; raer_b200's R patch does */
    {".pileup_matrix", (DL_FUNC)&pileup_matrix, 13},
    {".pileup_aei", (DL_FUNC)&pileup_aei, 13},
    {NULL, NULL, 0}};

void R_init_raer(DllInfo* dll) {
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}

/* library.dynam.unload(): give the device memory pool and the page-locked staging buffers back */
void R_unload_raer(DllInfo* dll) {
    (void)dll;
    if (raer_device_count() > 0) raer_trim_device_cache(0);
}
