/*
 * r_shim.c -- the R-facing side of the drop-in raer.so: SEXP <-> C-ABI marshalling only.
 *
 * NOT compiled in this repository's build (no R headers in the build image); it compiles where
 * Rinternals.h is available:  R CMD SHLIB -o raer.so r_shim.c -L. -lraer_b200   (see INTEGRATION.md).
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

SEXP pileup(SEXP bampaths, SEXP indexes, SEXP n, SEXP fapath, SEXP region, SEXP lst, SEXP int_args, SEXP dbl_args,
            SEXP lgl_args, SEXP libtype, SEXP only_keep_variants, SEXP min_mapQ, SEXP umi) {
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

    raer_pileup_args a;
    memset(&a, 0, sizeof(a));
    a.nbam = nf;
    a.bampaths = chr_vec(bampaths);
    a.indexes = chr_vec(indexes);
    a.fapath = translateChar(STRING_ELT(fapath, 0));
    a.region = LENGTH(region) ? translateChar(STRING_ELT(region, 0)) : NULL;
    if (!isNull(lst) && length(VECTOR_ELT(lst, 0)) > 0) {
        if (length(lst) != 3) Rf_error("'lst' must contain seqnames, start, and end");
        SEXP sn = VECTOR_ELT(lst, 0), st = VECTOR_ELT(lst, 1), en = VECTOR_ELT(lst, 2);
        a.n_regions = LENGTH(sn);
        if (!isString(sn) || !isInteger(st) || !isInteger(en) || LENGTH(st) != a.n_regions || LENGTH(en) != a.n_regions)
            Rf_error("'seqnames', 'start' and 'end' must have equal lengths");
        a.reg_seqnames = chr_vec(sn);
        a.reg_start = INTEGER(st);
        a.reg_end = INTEGER(en);
    }
    memcpy(a.int_args, INTEGER(int_args), sizeof(a.int_args));
    memcpy(a.dbl_args, REAL(dbl_args), sizeof(a.dbl_args));
    a.lgl_args[0] = LOGICAL(lgl_args)[0];
    a.lgl_args[1] = LOGICAL(lgl_args)[1];
    a.libtype = INTEGER(libtype);
    a.only_keep_variants = LOGICAL(only_keep_variants);
    a.min_mapq = INTEGER(min_mapQ);
    a.umi_tag = LENGTH(umi) ? translateChar(STRING_ELT(umi, 0)) : NULL;

    raer_pileup_result r;
    int rc = raer_pileup(&a, &r);
    if (rc != RAER_OK) Rf_error("[raer internal] error detected during pileup: %s", raer_last_error());

    static const char* NM[] = {"seqname", "pos", "strand", "REF", "ALT", "nRef", "nAlt", "nA", "nT", "nC", "nG", "nN", "nX"};
    R_xlen_t nr = (R_xlen_t)r.nrows;
    SEXP result = PROTECT(allocVector(VECSXP, nf + 1));
    {
        SEXP sd = PROTECT(allocVector(VECSXP, 3)), nms = PROTECT(allocVector(STRSXP, 3));
        SEXP rp = PROTECT(allocVector(REALSXP, nr)), vd = PROTECT(allocVector(REALSXP, nr)), so = PROTECT(allocVector(REALSXP, nr));
        if (nr) {
            memcpy(REAL(rp), r.rpbz, nr * sizeof(double));
            memcpy(REAL(vd), r.vdb, nr * sizeof(double));
            memcpy(REAL(so), r.sor, nr * sizeof(double));
        }
        SET_VECTOR_ELT(sd, 0, rp); SET_VECTOR_ELT(sd, 1, vd); SET_VECTOR_ELT(sd, 2, so);
        SET_STRING_ELT(nms, 0, mkChar("rpbz")); SET_STRING_ELT(nms, 1, mkChar("vdb")); SET_STRING_ELT(nms, 2, mkChar("sor"));
        setAttrib(sd, R_NamesSymbol, nms);
        SET_VECTOR_ELT(result, 0, sd);
        UNPROTECT(5);
    }
    /* few distinct strings: build each CHARSXP once and reuse the pointer (hard part 3) */
    SEXP seq_chr = PROTECT(allocVector(STRSXP, r.n_targets ? r.n_targets : 1));
    for (int t = 0; t < r.n_targets; ++t) SET_STRING_ELT(seq_chr, t, mkChar(r.target_names[t]));
    SEXP plus = PROTECT(mkChar("+")), minus = PROTECT(mkChar("-"));
    SEXP base_chr[256];
    memset(base_chr, 0, sizeof(base_chr));
    SEXP keep = PROTECT(allocVector(STRSXP, 256));
    for (int f = 0; f < nf; ++f) {
        SEXP l = PROTECT(allocVector(VECSXP, 13)), nms = PROTECT(allocVector(STRSXP, 13));
        for (int k = 0; k < 13; ++k) SET_STRING_ELT(nms, k, mkChar(NM[k]));
        setAttrib(l, R_NamesSymbol, nms);
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
            SET_STRING_ELT(alt, i, mkChar(altp + i * 12)); /* R's global CHARSXP cache dedups "-" / "G" ... */
            for (int k = 0; k < 8; ++k) INTEGER(cnt[k])[i] = cp[i * 8 + k];
        }
        SET_VECTOR_ELT(l, 0, seqname); SET_VECTOR_ELT(l, 1, pos); SET_VECTOR_ELT(l, 2, strand);
        SET_VECTOR_ELT(l, 3, ref); SET_VECTOR_ELT(l, 4, alt);
        for (int k = 0; k < 8; ++k) SET_VECTOR_ELT(l, 5 + k, cnt[k]);
        SET_VECTOR_ELT(result, 1 + f, l);
        UNPROTECT(15);
    }
    raer_pileup_result_free(&r);
    UNPROTECT(5);
    return result;
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
    int res = raer_scpileup(&a);
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
    {NULL, NULL, 0}};

void R_init_raer(DllInfo* dll) {
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
