/* Minimal stand-in for R's C API -- TEST INFRASTRUCTURE ONLY (R and its headers are absent from the build image).
 * Just enough of Rinternals.h / R.h / R_ext/Rdynload.h for r-package/src/rshim.c to compile, link against
 * libsgmcmc_b200.so and be driven by tests/r_stub/shim_driver.c: vectors of doubles / integers / strings, generic
 * lists, the dim and names attributes, external pointers with C finalizers, R_alloc, Rf_error. */
#ifndef SGM_R_STUB_RINTERNALS_H
#define SGM_R_STUB_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>

typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE = 1 } Rboolean;
enum { NILSXP = 0, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22, RAWSXP = 24, CHARSXP = 9, SYMSXP = 1 };

typedef struct SEXPREC {
    int type;
    R_xlen_t length;
    void* data;                 /* double* / int* / SEXP* / char* / external address */
    struct SEXPREC* dim;        /* attributes the shim reads */
    struct SEXPREC* names;
    void (*finalizer)(struct SEXPREC*);
} SEXPREC, *SEXP;

extern SEXP R_NilValue, R_DimSymbol, R_NamesSymbol;

SEXP Rf_allocVector(int type, R_xlen_t n);
SEXP Rf_allocMatrix(int type, int nrow, int ncol);
SEXP Rf_getAttrib(SEXP x, SEXP name);
SEXP Rf_mkChar(const char* s);
double Rf_asReal(SEXP x);
int Rf_asInteger(SEXP x);
void Rf_error(const char* fmt, ...) __attribute__((noreturn));
void REprintf(const char* fmt, ...);
void Rprintf(const char* fmt, ...);
char* R_alloc(size_t n, int size);
void R_CheckUserInterrupt(void);
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot);
void* R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, void (*fun)(SEXP), Rboolean onexit);

#define REAL(x) ((double*)(x)->data)
#define RAW(x) ((unsigned char*)(x)->data)
#define INTEGER(x) ((int*)(x)->data)
#define XLENGTH(x) ((x)->length)
#define LENGTH(x) ((int)(x)->length)
#define VECTOR_ELT(x, i) (((SEXP*)(x)->data)[i])
#define SET_VECTOR_ELT(x, i, v) (((SEXP*)(x)->data)[i] = (v))
#define STRING_ELT(x, i) (((SEXP*)(x)->data)[i])
#define SET_STRING_ELT(x, i, v) (((SEXP*)(x)->data)[i] = (v))
#define CHAR(x) ((const char*)(x)->data)
#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))

/* stub-only helpers for the driver */
void rstub_set_dim(SEXP x, int nrow, int ncol);
void rstub_set_names(SEXP x, const char* const* names);
void rstub_run_finalizers(void);
#endif
