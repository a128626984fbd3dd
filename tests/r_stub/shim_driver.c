/* Drives r-package/src/rshim.c the way the package's R files do (.Call C_sgm_create / init / step / run / get_params /
 * copy_chain / est_logpost), on top of the stand-in R API in this directory.  Test infrastructure only.
 *
 *   shim_driver check          registration table only (no device)
 *   shim_driver run            logistic SGLD-CV, N = 2000, d = 6, n = 100, seed 11, 3 chains: prints the flat parameters after
 *                              init + 1 step + 20 stored iterations and the chain length, for comparison with the same calls
 *                              made through the Python mirror (tests/test_r_shim.py)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"

SEXP C_sgm_create(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP C_sgm_init(SEXP, SEXP);
SEXP C_sgm_sample(SEXP, SEXP, SEXP, SEXP);
SEXP C_sgm_heldout_loglik(SEXP, SEXP, SEXP);
SEXP C_sgm_step(SEXP);
SEXP C_sgm_run(SEXP, SEXP, SEXP);
SEXP C_sgm_get_params(SEXP, SEXP);
SEXP C_sgm_copy_chain(SEXP, SEXP);
SEXP C_sgm_est_logpost(SEXP, SEXP);
SEXP C_sgm_info(SEXP);
SEXP C_sgm_save_state(SEXP);
SEXP C_sgm_load_state(SEXP, SEXP);
void R_init_sgmcmc(DllInfo*);

static SEXP num(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }

/* the same little generator as tests/test_r_shim.py: x = sin-hash in [-1, 1), y from a planted direction */
static double hash01(unsigned i, unsigned j) {
    unsigned x = i * 2654435761u ^ (j + 1u) * 2246822519u;
    x ^= x >> 15; x *= 2654435761u; x ^= x >> 13;
    return (double)(x & 0xFFFFFFu) / 16777216.0;
}

int main(int argc, char** argv) {
    R_init_sgmcmc(NULL);
    int n_reg = 0;
    for (const R_CallMethodDef* m = rstub_registered; m && m->name; ++m) ++n_reg;
    printf("registered %d .Call routines\n", n_reg);
    if (argc < 2 || strcmp(argv[1], "run") != 0) return n_reg == 14 ? 0 : 1;

    const int N = 2000, d = 6, chains = 3;
    SEXP X = Rf_allocMatrix(REALSXP, N, d), y = Rf_allocVector(REALSXP, N);
    for (int i = 0; i < N; ++i) {
        double z = 0.1;
        for (int j = 0; j < d; ++j) { const double v = 2.0 * hash01((unsigned)i, (unsigned)j) - 1.0; REAL(X)[i + (size_t)N * j] = v; z += v * (0.3 * (j + 1) - 1.0); }
        REAL(y)[i] = hash01((unsigned)i, 77u) < 1.0 / (1.0 + exp(-z)) ? 1.0 : 0.0;
    }
    SEXP dataset = Rf_allocVector(VECSXP, 2);
    SET_VECTOR_ELT(dataset, 0, X); SET_VECTOR_ELT(dataset, 1, y);
    const char* dnames[] = {"X", "y"};
    rstub_set_names(dataset, dnames);
    SEXP bias = num(0.0), beta = Rf_allocMatrix(REALSXP, d, 1);
    SEXP params = Rf_allocVector(VECSXP, 2);
    SET_VECTOR_ELT(params, 0, bias); SET_VECTOR_ELT(params, 1, beta);
    const char* pnames[] = {"bias", "beta"};
    rstub_set_names(params, pnames);
    SEXP step = Rf_allocVector(REALSXP, 2), alpha = Rf_allocVector(REALSXP, 2), a = Rf_allocVector(REALSXP, 2);
    for (int i = 0; i < 2; ++i) { REAL(step)[i] = 1e-4; REAL(alpha)[i] = 0.01; REAL(a)[i] = 0.01; }
    /* spec: model, algo, cv, minibatchSize, L, optStepsize, nItersOpt, seed, nChains, device, flags, usePrior, priorSd, precision */
    SEXP spec = Rf_allocVector(VECSXP, 14);
    const double vals[13] = {3 /* logistic */, 1 /* sgld */, 1 /* cv */, 100, 1, 1e-4, 50, 11, chains, 0, 0, 1, 0.0};
    for (int i = 0; i < 13; ++i) SET_VECTOR_ELT(spec, i, num(vals[i]));
    SET_VECTOR_ELT(spec, 13, R_NilValue);

    SEXP h = C_sgm_create(spec, dataset, params, step, alpha, a);
    SEXP info = C_sgm_info(h);
    printf("info N=%.0f n=%.0f P=%.0f\n", REAL(info)[0], REAL(info)[1], REAL(info)[2]);
    C_sgm_init(h, num(0));
    C_sgm_step(h);
    C_sgm_run(h, num(20), num(1));
    SEXP th = C_sgm_get_params(h, num(chains));
    SEXP ch = C_sgm_copy_chain(h, num(chains));
    SEXP lp = C_sgm_est_logpost(h, num(chains));
    printf("chain_doubles %ld\n", (long)XLENGTH(ch));
    printf("params");
    for (R_xlen_t i = 0; i < XLENGTH(th); ++i) printf(" %.17g", REAL(th)[i]);
    printf("\nlast_stored");
    {   /* per chain: [iters x P] column-major; print the last stored row of chain 0 */
        const int P = (int)REAL(info)[2];
        const long iters = (long)XLENGTH(ch) / (P * chains);
        for (int p = 0; p < P; ++p) printf(" %.17g", REAL(ch)[(iters - 1) + iters * p]);
    }
    printf("\nlogpost_finite %d\n", isfinite(REAL(lp)[0]) ? 1 : 0);
    {   /* runSGMCMC as the R file makes it: a second sampler, initSess (verbose: progress lines on stderr), ONE C_sgm_sample call */
        SEXP h2 = C_sgm_create(spec, dataset, params, step, alpha, a);
        C_sgm_init(h2, num(1));
        SEXP ch2 = C_sgm_sample(h2, num(230), num(chains), num(1));
        const int P = (int)REAL(info)[2];
        printf("sampled_doubles %ld\nsampled_last", (long)XLENGTH(ch2));
        for (int p = 0; p < P; ++p) printf(" %.17g", REAL(ch2)[229 + 230 * p]);
        SEXP th2 = C_sgm_get_params(h2, num(chains));
        printf("\nsampled_state");
        for (int p = 0; p < P; ++p) printf(" %.17g", REAL(th2)[p]);
        /* heldoutLogLik on the first 300 rows */
        SEXP Xt = Rf_allocMatrix(REALSXP, 300, d), yt = Rf_allocVector(REALSXP, 300);
        for (int i = 0; i < 300; ++i) { for (int j = 0; j < d; ++j) REAL(Xt)[i + 300 * (size_t)j] = REAL(X)[i + (size_t)N * j]; REAL(yt)[i] = REAL(y)[i]; }
        SEXP testset = Rf_allocVector(VECSXP, 2);
        SET_VECTOR_ELT(testset, 0, Xt); SET_VECTOR_ELT(testset, 1, yt);
        rstub_set_names(testset, dnames);
        SEXP ho = C_sgm_heldout_loglik(h2, testset, num(chains));
        printf("\nheldout");
        for (int c = 0; c < chains; ++c) printf(" %.17g", REAL(ho)[c]);
        printf("\n");
        /* saveState / loadState: a third sampler restored from h2's blob continues h2's run exactly */
        SEXP blob = C_sgm_save_state(h2);
        for (int k = 0; k < 3; ++k) C_sgm_step(h2);
        SEXP h3 = C_sgm_create(spec, dataset, params, step, alpha, a);
        C_sgm_load_state(h3, blob);
        for (int k = 0; k < 3; ++k) C_sgm_step(h3);
        SEXP ta = C_sgm_get_params(h2, num(chains)), tb = C_sgm_get_params(h3, num(chains));
        int same = XLENGTH(ta) == XLENGTH(tb);
        for (R_xlen_t i = 0; same && i < XLENGTH(ta); ++i) same = REAL(ta)[i] == REAL(tb)[i];
        printf("state_blob_bytes %ld\nresume_equal %d\n", (long)XLENGTH(blob), same);
    }
    rstub_run_finalizers();   /* what R's garbage collector does with the external pointer */
    return 0;
}
