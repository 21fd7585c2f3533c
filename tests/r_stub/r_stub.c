/* tests/r_stub/r_stub.c -- TEST INFRASTRUCTURE: the object model behind tests/r_stub/Rinternals.h plus a harness that
 * a Python test drives through ctypes: build argument vectors, .Call a registered routine by name (errors and
 * interrupts come back as status codes, the way R's condition system would unwind them), read the result back.
 * Objects are never freed (a test process); external-pointer finalizers run on request so that a test can see that a
 * result abandoned by a failing allocation is released. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "R_ext/Rdynload.h"
#include "Rinternals.h"

struct SEXPREC {
    int type;
    R_xlen_t len;
    void* data;      /* int / double / SEXP array, or char* for CHARSXP */
    SEXP names, dim; /* the two attributes the shim sets */
    void* ext;
    R_CFinalizer_t fin;
};
static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL};
static struct SEXPREC names_sym = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL}, dim_sym = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_rec, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;

static jmp_buf* g_top = NULL;      /* where Rf_error / Rf_onintr unwind to */
static jmp_buf* g_toplevel = NULL; /* innermost R_ToplevelExec */
static char g_msg[1024];
static long g_alloc_count = 0, g_fail_alloc_at = -1;
static long g_checks = 0, g_interrupt_at = -1;
static SEXP g_extptrs[4096];
static int g_n_ext = 0;
static long g_finalized = 0;

static SEXP new_obj(int type, R_xlen_t n, size_t esz) {
    SEXP x = (SEXP)calloc(1, sizeof(struct SEXPREC));
    x->type = type;
    x->len = n;
    x->data = calloc(n > 0 ? (size_t)n : 1, esz);
    x->names = x->dim = R_NilValue;
    return x;
}
SEXP Rf_allocVector(unsigned type, R_xlen_t n) {
    ++g_alloc_count;
    if (g_fail_alloc_at >= 0 && g_alloc_count >= g_fail_alloc_at) {
        g_fail_alloc_at = -1;
        Rf_error("cannot allocate vector of size %ld", (long)n);
    }
    switch (type) {
    case LGLSXP: case INTSXP: return new_obj((int)type, n, sizeof(int));
    case REALSXP: return new_obj(REALSXP, n, sizeof(double));
    case STRSXP: {
        SEXP x = new_obj(STRSXP, n, sizeof(SEXP));
        for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)x->data)[i] = Rf_mkChar("");
        return x;
    }
    case VECSXP: {
        SEXP x = new_obj(VECSXP, n, sizeof(SEXP));
        for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)x->data)[i] = R_NilValue;
        return x;
    }
    default: Rf_error("r_stub: unsupported type %u", type);
    }
}
int* INTEGER(SEXP x) { return (int*)x->data; }
int* LOGICAL(SEXP x) { return (int*)x->data; }
double* REAL(SEXP x) { return (double*)x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP*)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP*)x->data)[i] = v; return v; }
int LENGTH(SEXP x) { return (int)x->len; }
R_xlen_t Rf_length(SEXP x) { return x->len; }
SEXP Rf_mkChar(const char* s) {
    SEXP x = (SEXP)calloc(1, sizeof(struct SEXPREC));
    x->type = CHARSXP;
    x->len = (R_xlen_t)strlen(s);
    x->data = strdup(s);
    x->names = x->dim = R_NilValue;
    return x;
}
SEXP Rf_mkString(const char* s) { SEXP x = Rf_allocVector(STRSXP, 1); SET_STRING_ELT(x, 0, Rf_mkChar(s)); return x; }
SEXP Rf_mkNamed(unsigned type, const char** names) {
    int n = 0;
    while (names[n][0]) ++n;
    SEXP x = Rf_allocVector(type, n), nm = Rf_allocVector(STRSXP, n);
    for (int i = 0; i < n; ++i) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
    x->names = nm;
    return x;
}
SEXP Rf_ScalarInteger(int v) { SEXP x = Rf_allocVector(INTSXP, 1); INTEGER(x)[0] = v; return x; }
const char* Rf_translateChar(SEXP x) { return (const char*)x->data; }
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val) {
    if (name == R_NamesSymbol) x->names = val;
    else if (name == R_DimSymbol) x->dim = val;
    return val;
}
Rboolean Rf_isInteger(SEXP x) { return x->type == INTSXP; }
Rboolean Rf_isReal(SEXP x) { return x->type == REALSXP; }
Rboolean Rf_isLogical(SEXP x) { return x->type == LGLSXP; }
Rboolean Rf_isString(SEXP x) { return x->type == STRSXP; }
Rboolean Rf_isNull(SEXP x) { return x->type == NILSXP; }
Rboolean Rf_isMatrix(SEXP x) { return x->dim != R_NilValue && x->dim->len == 2; }
int Rf_nrows(SEXP x) { return Rf_isMatrix(x) ? INTEGER(x->dim)[0] : (int)x->len; }
int Rf_ncols(SEXP x) { return Rf_isMatrix(x) ? INTEGER(x->dim)[1] : 1; }
SEXP Rf_protect(SEXP x) { return x; }
void Rf_unprotect(int n) { (void)n; }
void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_msg, sizeof(g_msg), fmt, ap);
    va_end(ap);
    if (!g_top) { fprintf(stderr, "r_stub: Rf_error outside a call: %s\n", g_msg); abort(); }
    longjmp(*g_top, 1);
}
void Rf_onintr(void) {
    snprintf(g_msg, sizeof(g_msg), "interrupt");
    if (!g_top) abort();
    longjmp(*g_top, 2);
}
void REprintf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_msg, sizeof(g_msg), fmt, ap);
    va_end(ap);
}
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)size); }
const char* R_ExpandFileName(const char* s) { return s; }
void R_CheckUserInterrupt(void) {
    ++g_checks;
    if (g_interrupt_at >= 0 && g_checks >= g_interrupt_at) {
        g_interrupt_at = -1;
        if (g_toplevel) longjmp(*g_toplevel, 1);
        Rf_onintr();
    }
}
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) {
    jmp_buf jb, *saved = g_toplevel;
    g_toplevel = &jb;
    Rboolean ok = TRUE;
    if (setjmp(jb) == 0) fun(data);
    else ok = FALSE;
    g_toplevel = saved;
    return ok;
}
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
    (void)tag; (void)prot;
    SEXP x = new_obj(EXTPTRSXP, 0, 1);
    x->ext = p;
    if (g_n_ext < 4096) g_extptrs[g_n_ext++] = x;
    return x;
}
void* R_ExternalPtrAddr(SEXP s) { return s->ext; }
void R_ClearExternalPtr(SEXP s) { s->ext = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) { (void)onexit; s->fin = fun; }

static const R_CallMethodDef* g_calls = NULL;
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
    (void)info; (void)c; (void)f; (void)e;
    g_calls = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value) { (void)info; (void)value; return TRUE; }

/* ---------------------------------------------------------------- harness (ctypes) */
void R_init_raer(DllInfo* dll);
int rstub_init(void) { if (!g_calls) R_init_raer(NULL); return g_calls != NULL; }
int rstub_registered(int i, const char** name, int* arity) {
    if (!g_calls || !g_calls[i].name) return 0;
    *name = g_calls[i].name;
    *arity = g_calls[i].numArgs;
    return 1;
}
SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_int(int n, const int* v) { SEXP x = Rf_allocVector(INTSXP, n); if (n) memcpy(x->data, v, (size_t)n * sizeof(int)); return x; }
SEXP rstub_lgl(int n, const int* v) { SEXP x = Rf_allocVector(LGLSXP, n); if (n) memcpy(x->data, v, (size_t)n * sizeof(int)); return x; }
SEXP rstub_real(int n, const double* v) { SEXP x = Rf_allocVector(REALSXP, n); if (n) memcpy(x->data, v, (size_t)n * sizeof(double)); return x; }
SEXP rstub_str(int n, const char** v) {
    SEXP x = Rf_allocVector(STRSXP, n);
    for (int i = 0; i < n; ++i) SET_STRING_ELT(x, i, Rf_mkChar(v[i]));
    return x;
}
SEXP rstub_list(int n) { return Rf_allocVector(VECSXP, n); }
void rstub_list_set(SEXP l, int i, SEXP v) { SET_VECTOR_ELT(l, i, v); }
void rstub_set_dim(SEXP x, int nr, int nc) { SEXP d = Rf_allocVector(INTSXP, 2); INTEGER(d)[0] = nr; INTEGER(d)[1] = nc; x->dim = d; }
int rstub_type(SEXP x) { return x->type; }
long rstub_len(SEXP x) { return (long)x->len; }
void* rstub_data(SEXP x) { return x->data; }
const char* rstub_chr(SEXP strsxp, long i) { return (const char*)STRING_ELT(strsxp, i)->data; }
SEXP rstub_elt(SEXP l, long i) { return VECTOR_ELT(l, i); }
SEXP rstub_names(SEXP x) { return x->names; }
SEXP rstub_dim(SEXP x) { return x->dim; }
const char* rstub_message(void) { return g_msg; }
void rstub_interrupt_at_check(long k) { g_checks = 0; g_interrupt_at = k; }
long rstub_checks(void) { return g_checks; }
void rstub_fail_alloc_at(long k) { g_alloc_count = 0; g_fail_alloc_at = k; }
/* what a garbage collection does for external pointers nobody cleared */
long rstub_run_finalizers(void) {
    long ran = 0;
    for (int i = 0; i < g_n_ext; ++i)
        if (g_extptrs[i]->ext && g_extptrs[i]->fin) { g_extptrs[i]->fin(g_extptrs[i]); ++ran; ++g_finalized; }
    g_n_ext = 0;
    return ran;
}
typedef SEXP (*F1)(SEXP);
typedef SEXP (*F13)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F14)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
/* .Call(name, args...): 0 = value in *out, 1 = R error (rstub_message), 2 = interrupt, -1 = no such routine / arity */
int rstub_call(const char* name, int nargs, SEXP* a, SEXP* out) {
    if (!rstub_init()) return -1;
    const R_CallMethodDef* d = g_calls;
    for (; d->name; ++d)
        if (!strcmp(d->name, name)) break;
    if (!d->name || d->numArgs != nargs) return -1;
    jmp_buf jb;
    g_top = &jb;
    g_msg[0] = 0;
    int how = setjmp(jb);
    if (how == 0) {
        if (nargs == 1) *out = ((F1)d->fun)(a[0]);
        else if (nargs == 13) *out = ((F13)d->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]);
        else if (nargs == 14)
            *out = ((F14)d->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13]);
        else how = -1;
    }
    g_top = NULL;
    return how;
}
