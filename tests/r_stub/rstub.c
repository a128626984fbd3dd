/* Implementation of the stand-in R API (see Rinternals.h in this directory) -- test infrastructure only. */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

static SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, NULL, NULL}, dim_rec = {SYMSXP, 0, NULL, NULL, NULL, NULL}, names_rec = {SYMSXP, 0, NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_rec, R_DimSymbol = &dim_rec, R_NamesSymbol = &names_rec;
const R_CallMethodDef* rstub_registered = NULL;
static SEXP ext_ptrs[64];
static int n_ext = 0;

static size_t elem_size(int type) { return type == REALSXP ? sizeof(double) : type == INTSXP ? sizeof(int) : type == RAWSXP ? 1 : sizeof(SEXP); }

SEXP Rf_allocVector(int type, R_xlen_t n) {
    SEXP s = (SEXP)calloc(1, sizeof(SEXPREC));
    s->type = type; s->length = n; s->dim = R_NilValue; s->names = R_NilValue;
    s->data = calloc((size_t)(n > 0 ? n : 1), elem_size(type));
    return s;
}
SEXP Rf_allocMatrix(int type, int nrow, int ncol) {
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    rstub_set_dim(s, nrow, ncol);
    return s;
}
void rstub_set_dim(SEXP x, int nrow, int ncol) {
    SEXP d = Rf_allocVector(INTSXP, 2);
    INTEGER(d)[0] = nrow; INTEGER(d)[1] = ncol;
    x->dim = d;
}
SEXP Rf_mkChar(const char* s) {
    SEXP c = (SEXP)calloc(1, sizeof(SEXPREC));
    c->type = CHARSXP; c->length = (R_xlen_t)strlen(s); c->data = strdup(s); c->dim = R_NilValue; c->names = R_NilValue;
    return c;
}
void rstub_set_names(SEXP x, const char* const* names) {
    SEXP n = Rf_allocVector(STRSXP, x->length);
    for (R_xlen_t i = 0; i < x->length; ++i) SET_STRING_ELT(n, i, Rf_mkChar(names[i]));
    x->names = n;
}
SEXP Rf_getAttrib(SEXP x, SEXP name) { return name == R_DimSymbol ? x->dim : name == R_NamesSymbol ? x->names : R_NilValue; }
double Rf_asReal(SEXP x) { return x->type == REALSXP ? REAL(x)[0] : x->type == INTSXP ? (double)INTEGER(x)[0] : 0.0; }
int Rf_asInteger(SEXP x) { return (int)Rf_asReal(x); }
void REprintf(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); }
void Rprintf(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stdout, fmt, ap); va_end(ap); }
void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "Error: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    exit(2);
}
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)size); }
void R_CheckUserInterrupt(void) {}
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
    (void)tag; (void)prot;
    SEXP s = (SEXP)calloc(1, sizeof(SEXPREC));
    s->type = EXTPTRSXP; s->data = p; s->dim = R_NilValue; s->names = R_NilValue;
    return s;
}
void* R_ExternalPtrAddr(SEXP s) { return s->data; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, void (*fun)(SEXP), Rboolean onexit) {
    (void)onexit;
    s->finalizer = fun;
    if (n_ext < 64) ext_ptrs[n_ext++] = s;
}
void rstub_run_finalizers(void) {
    for (int i = 0; i < n_ext; ++i)
        if (ext_ptrs[i]->finalizer) ext_ptrs[i]->finalizer(ext_ptrs[i]);
    n_ext = 0;
}
int R_registerRoutines(DllInfo* info, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
    (void)info; (void)c; (void)f; (void)e;
    rstub_registered = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* info, Rboolean value) { (void)info; return value; }
