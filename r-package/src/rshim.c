/*
 * .Call shim between the R files of this package and the C ABI of libsgmcmc_b200.so
 * (include/sgmcmc_b200.h).  Marshalling only: every numeric operation happens in the library.
 * UNVERIFIED in the build environment (no R headers there).
 *
 * Replaces the reference's `tf$...` -> reticulate -> Python boundary (R/update.R:44,127-129).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>
#include "sgmcmc_b200.h"

static void check(int rc) {
    if (rc != SGM_OK) Rf_error("%s", sgm_last_error());   /* same message text as the reference's stop() */
}

static void finalize_handle(SEXP ptr) {
    sgm_handle* h = (sgm_handle*)R_ExternalPtrAddr(ptr);
    if (h) { sgm_destroy(h); R_ClearExternalPtr(ptr); }
}

static sgm_handle* handle_of(SEXP ptr) {
    sgm_handle* h = (sgm_handle*)R_ExternalPtrAddr(ptr);
    if (!h) Rf_error("sgmcmc session was closed");
    return h;
}

static void fill_dims(SEXP x, int32_t* ndim, int64_t dims[4]) {
    SEXP d = Rf_getAttrib(x, R_DimSymbol);
    if (d == R_NilValue) {                       /* R vector: a length-1 vector is a scalar parameter */
        *ndim = (XLENGTH(x) > 1) ? 1 : 0;
        dims[0] = XLENGTH(x);
    } else {
        *ndim = (int32_t)XLENGTH(d);
        for (int i = 0; i < *ndim && i < 4; ++i) dims[i] = INTEGER(d)[i];
    }
}

/* spec: list(model, algo, cv, minibatchSize, L, optStepsize, nItersOpt, seed (or NULL), nChains, device, flags,
 *            usePrior, priorSd, precision (or NULL)); dataset / params: named lists of double arrays;
 * stepsize / alpha / a: double vectors aligned with params. */
SEXP C_sgm_create(SEXP spec, SEXP dataset, SEXP params, SEXP stepsize, SEXP alpha, SEXP a) {
    sgm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = SGM_ABI_VERSION;
#define NUM(i) Rf_asReal(VECTOR_ELT(spec, (i)))
    cfg.model = (int32_t)NUM(0); cfg.algo = (int32_t)NUM(1); cfg.cv = (int32_t)NUM(2);
    cfg.minibatch_size = NUM(3); cfg.L = (int32_t)NUM(4); cfg.opt_stepsize = NUM(5); cfg.n_iters_opt = (int64_t)NUM(6);
    SEXP seed = VECTOR_ELT(spec, 7);
    cfg.seeded = seed != R_NilValue; cfg.seed = cfg.seeded ? (uint64_t)Rf_asReal(seed) : 0;
    cfg.n_chains = (int32_t)NUM(8); cfg.device = (int32_t)NUM(9); cfg.flags = (uint32_t)NUM(10);
    cfg.use_prior = (int32_t)NUM(11); cfg.prior_sd = NUM(12);
    SEXP prec = VECTOR_ELT(spec, 13);
    cfg.precision = (prec == R_NilValue) ? NULL : REAL(prec);
#undef NUM
    int nd = (int)XLENGTH(dataset), np = (int)XLENGTH(params);
    sgm_array* arr = (sgm_array*)R_alloc(nd, sizeof(sgm_array));
    sgm_param* par = (sgm_param*)R_alloc(np, sizeof(sgm_param));
    SEXP dn = Rf_getAttrib(dataset, R_NamesSymbol), pn = Rf_getAttrib(params, R_NamesSymbol);
    for (int i = 0; i < nd; ++i) {
        SEXP x = VECTOR_ELT(dataset, i);          /* the R file coerced every entry with as.double */
        memset(&arr[i], 0, sizeof arr[i]);
        arr[i].name = CHAR(STRING_ELT(dn, i)); arr[i].data = REAL(x);
        fill_dims(x, &arr[i].ndim, arr[i].dims);
        if (arr[i].ndim == 0) arr[i].ndim = 1;
    }
    for (int i = 0; i < np; ++i) {
        SEXP x = VECTOR_ELT(params, i);
        memset(&par[i], 0, sizeof par[i]);
        par[i].name = CHAR(STRING_ELT(pn, i)); par[i].init = REAL(x);
        fill_dims(x, &par[i].ndim, par[i].dims);
        par[i].stepsize = REAL(stepsize)[i]; par[i].alpha = REAL(alpha)[i]; par[i].a = REAL(a)[i];
    }
    cfg.n_data = nd; cfg.data = arr; cfg.n_params = np; cfg.params = par;
    sgm_handle* h = NULL;
    check(sgm_create(&cfg, &h));
    SEXP ptr = PROTECT(R_MakeExternalPtr(h, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, finalize_handle, TRUE);
    UNPROTECT(1);
    return ptr;
}

/* checkOptDivergence / checkDivergence progress lines (reference R/controlVariates.R:121-123, R/update.R:204-206): the library
 * computes the log posterior estimate on the device every 100 iterations and calls back on this (the R main) thread */
static void progress_line(void* user, int32_t phase, int64_t iter, const double* logpost, int32_t n) {
    (void)phase;
    if (!user || !*(const int*)user) return;   /* verbose = FALSE */
    double v = 0;
    for (int i = 0; i < n; ++i) v += logpost[i];
    REprintf("Iteration: %lld\t\tLog posterior estimate: %g\n", (long long)iter, n > 0 ? v / n : v);   /* message() writes to stderr */
}

SEXP C_sgm_init(SEXP ptr, SEXP verbose) {
    sgm_handle* h = handle_of(ptr);
    int v = Rf_asInteger(verbose);
    check(sgm_set_progress(h, progress_line, &v));
    const int rc = sgm_init(h);
    sgm_set_progress(h, NULL, NULL);
    check(rc);
    return R_NilValue;
}

/* runSGMCMC's whole loop in one call (reference R/update.R:145-163): n_iters steps, every state stored, the chain streamed to the host
 * while later iterations run, divergence check + progress line every 100 iterations.  Returns [iters x P x nChains] doubles. */
SEXP C_sgm_sample(SEXP ptr, SEXP n_iters, SEXP n_chains, SEXP verbose) {
    sgm_handle* h = handle_of(ptr);
    const R_xlen_t iters = (R_xlen_t)Rf_asReal(n_iters);
    const int P = sgm_num_scalars(h), C = Rf_asInteger(n_chains);
    int v = Rf_asInteger(verbose);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, iters * P * C));
    check(sgm_set_progress(h, progress_line, &v));
    const int rc = sgm_sample(h, (int64_t)iters, 100, REAL(out));
    sgm_set_progress(h, NULL, NULL);
    check(rc);
    R_CheckUserInterrupt();
    UNPROTECT(1);
    return out;
}

/* logLik(current params, testset) per chain on the device (replaces the user-built TensorFlow tensors of the reference's vignettes,
 * vignettes/nn.Rmd:113-123, logisticRegression.Rmd:89-115).  testset: named list of double arrays like the dataset. */
SEXP C_sgm_heldout_loglik(SEXP ptr, SEXP testset, SEXP n_chains) {
    sgm_handle* h = handle_of(ptr);
    const int nd = (int)XLENGTH(testset);
    sgm_array* arr = (sgm_array*)R_alloc(nd, sizeof(sgm_array));
    SEXP dn = Rf_getAttrib(testset, R_NamesSymbol);
    for (int i = 0; i < nd; ++i) {
        SEXP x = VECTOR_ELT(testset, i);
        memset(&arr[i], 0, sizeof arr[i]);
        arr[i].name = CHAR(STRING_ELT(dn, i)); arr[i].data = REAL(x);
        fill_dims(x, &arr[i].ndim, arr[i].dims);
        if (arr[i].ndim == 0) arr[i].ndim = 1;
    }
    SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_chains)));
    check(sgm_heldout_loglik(h, arr, nd, REAL(out)));
    UNPROTECT(1);
    return out;
}
SEXP C_sgm_step(SEXP ptr) { check(sgm_step(handle_of(ptr))); return R_NilValue; }

/* runSGMCMC's inner loop: n_iters steps on the device, then back to R (progress line, R_CheckUserInterrupt) */
SEXP C_sgm_run(SEXP ptr, SEXP n_iters, SEXP store) {
    check(sgm_run(handle_of(ptr), (int64_t)Rf_asReal(n_iters), Rf_asInteger(store)));
    R_CheckUserInterrupt();
    return R_NilValue;
}

SEXP C_sgm_reserve_chain(SEXP ptr, SEXP n_iters) { check(sgm_reserve_chain(handle_of(ptr), (int64_t)Rf_asReal(n_iters))); return R_NilValue; }
SEXP C_sgm_reset_chain(SEXP ptr) { check(sgm_reset_chain(handle_of(ptr))); return R_NilValue; }

/* flat [P x nChains] doubles; the R side splits them into the named list getParams returns */
SEXP C_sgm_get_params(SEXP ptr, SEXP n_chains) {
    sgm_handle* h = handle_of(ptr);
    int P = sgm_num_scalars(h), C = Rf_asInteger(n_chains);
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, P, C));
    check(sgm_get_params(h, REAL(out)));
    UNPROTECT(1);
    return out;
}

/* [iters x P x nChains] doubles: per chain, every parameter already in R's (nIters, dims...) column-major layout */
SEXP C_sgm_copy_chain(SEXP ptr, SEXP n_chains) {
    sgm_handle* h = handle_of(ptr);
    R_xlen_t iters = (R_xlen_t)sgm_chain_length(h);
    int P = sgm_num_scalars(h), C = Rf_asInteger(n_chains);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, iters * P * C));
    check(sgm_copy_chain(h, REAL(out)));
    UNPROTECT(1);
    return out;
}

SEXP C_sgm_est_logpost(SEXP ptr, SEXP n_chains) {
    SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(n_chains)));
    check(sgm_est_logpost(handle_of(ptr), REAL(out)));
    UNPROTECT(1);
    return out;
}

/* The whole device state as a raw vector / restored from one (sgm_save_state, sgm_load_state): what saveRDS can persist between R
 * sessions.  The reference leaves this to the user of the step-by-step API, who can save getParams() only (NEWS.md:19). */
SEXP C_sgm_save_state(SEXP ptr) {
    sgm_handle* h = handle_of(ptr);
    size_t need = 0;
    check(sgm_save_state(h, NULL, 0, &need));
    SEXP out = PROTECT(Rf_allocVector(RAWSXP, (R_xlen_t)need));
    check(sgm_save_state(h, RAW(out), need, &need));
    UNPROTECT(1);
    return out;
}
SEXP C_sgm_load_state(SEXP ptr, SEXP blob) {
    check(sgm_load_state(handle_of(ptr), RAW(blob), (size_t)XLENGTH(blob)));
    return R_NilValue;
}

SEXP C_sgm_info(SEXP ptr) {
    sgm_handle* h = handle_of(ptr);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 3));
    REAL(out)[0] = (double)sgm_num_rows(h); REAL(out)[1] = (double)sgm_minibatch(h); REAL(out)[2] = (double)sgm_num_scalars(h);
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"C_sgm_create", (DL_FUNC)&C_sgm_create, 6}, {"C_sgm_init", (DL_FUNC)&C_sgm_init, 2},
    {"C_sgm_sample", (DL_FUNC)&C_sgm_sample, 4}, {"C_sgm_heldout_loglik", (DL_FUNC)&C_sgm_heldout_loglik, 3},
    {"C_sgm_step", (DL_FUNC)&C_sgm_step, 1}, {"C_sgm_run", (DL_FUNC)&C_sgm_run, 3},
    {"C_sgm_reserve_chain", (DL_FUNC)&C_sgm_reserve_chain, 2}, {"C_sgm_reset_chain", (DL_FUNC)&C_sgm_reset_chain, 1},
    {"C_sgm_get_params", (DL_FUNC)&C_sgm_get_params, 2}, {"C_sgm_copy_chain", (DL_FUNC)&C_sgm_copy_chain, 2},
    {"C_sgm_est_logpost", (DL_FUNC)&C_sgm_est_logpost, 2}, {"C_sgm_info", (DL_FUNC)&C_sgm_info, 1},
    {"C_sgm_save_state", (DL_FUNC)&C_sgm_save_state, 1}, {"C_sgm_load_state", (DL_FUNC)&C_sgm_load_state, 2},
    {NULL, NULL, 0}};

void R_init_sgmcmc(DllInfo* dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
