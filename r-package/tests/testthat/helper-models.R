# Shared set-up for the ports of the reference's testthat files (tests/testthat/test1dGauss.R, test3dGauss.R,
# testStepByStep.R, testOptional.R, testSeeds.R).  UNVERIFIED here: no R in the build image; the same bounds are enforced
# through the Python mirror in tests/test_statistics_gpu.py.  Tuning constants follow test1dGauss.R:29-83.
skipWithoutDevice = function() {
    ok = tryCatch( { sgldSetup( gaussMeanLogLik(), list( X = rnorm( 10 ) ), list( theta = 0 ), 1e-4, minibatchSize = 5 ); TRUE },
                   error = function( e ) FALSE )
    if ( !ok ) skip( "no CUDA device / libsgmcmc_b200.so" )
}

samplers = list(
    sgld    = function( ll, d, p, lp, ... ) sgld( ll, d, p, 1e-4, logPrior = lp, minibatchSize = 100, nIters = 1100, verbose = FALSE, seed = 13, ... ),
    sgldcv  = function( ll, d, p, lp, ... ) sgldcv( ll, d, p, 1e-4, 1e-5, logPrior = lp, minibatchSize = 100, nIters = 1100, nItersOpt = 1000, verbose = FALSE, seed = 13, ... ),
    sghmc   = function( ll, d, p, lp, ... ) sghmc( ll, d, p, 1e-5, logPrior = lp, minibatchSize = 100, alpha = 0.1, L = 3, nIters = 1100, verbose = FALSE, seed = 13, ... ),
    sghmccv = function( ll, d, p, lp, ... ) sghmccv( ll, d, p, 5e-5, 1e-5, logPrior = lp, minibatchSize = 100, alpha = 0.1, L = 3, nIters = 1100, nItersOpt = 1000, verbose = FALSE, seed = 13, ... ),
    sgnht   = function( ll, d, p, lp, ... ) sgnht( ll, d, p, 5e-6, logPrior = lp, minibatchSize = 100, a = 1e-2, nIters = 1100, verbose = FALSE, seed = 13, ... ),
    sgnhtcv = function( ll, d, p, lp, ... ) sgnhtcv( ll, d, p, 1e-4, 1e-5, logPrior = lp, minibatchSize = 100, a = 1e-2, nIters = 1100, nItersOpt = 1000, verbose = FALSE, seed = 13, ... ) )

# the reference's acceptance region for a Gaussian-mean posterior centred at 0 (test1dGauss.R:92-95): after 100 burn-in draws the
# chain stays within +-0.5 and its mean is within sqrt(0.1) of 0
expectNearZero = function( chain ) {
    kept = chain[ -( 1:100 ), , drop = FALSE ]
    for ( j in seq_len( ncol( kept ) ) ) {
        expect_gte( min( kept[ , j ] ), -0.5 )
        expect_lte( max( kept[ , j ] ), 0.5 )
        expect_lte( mean( kept[ , j ] )^2, 0.1 )
    }
}
