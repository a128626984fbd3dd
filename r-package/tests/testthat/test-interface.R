context( "step-by-step interface, optional arguments, seeds, rejected closures (testStepByStep.R, testOptional.R, testSeeds.R, testFloat64.R)" )

test_that( "Setup / initSess / sgmcmcStep / getParams reproduce the one-shot bounds", {
    skipWithoutDevice()
    set.seed( 13 )
    dataset = list( X = matrix( rnorm( 3 * 10^4 ), ncol = 3 ) )
    sg = sgldSetup( gaussMeanLogLik(), dataset, list( theta = rnorm( 3, 0, 0.1 ) ), 1e-4, logPrior = normalPrior( 5 ), minibatchSize = 100, seed = 13 )
    expect_equal( class( sg ), c( "sgld", "sgmcmc" ) )
    sess = initSess( sg, FALSE )
    draws = matrix( 0, 1100, 3 )
    for ( i in 1:1100 ) { sgmcmcStep( sg, sess ); draws[ i, ] = getParams( sg, sess )$theta }
    expectNearZero( draws )
} )

test_that( "defaults: no logPrior, minibatchSize = 0.01, scalar stepsize", {
    skipWithoutDevice()
    dataset = list( X = matrix( rnorm( 2 * 10^3 ), ncol = 2 ) )
    out = sgld( gaussMeanLogLik(), dataset, list( theta = c( 0.1, 0.1 ) ), 1e-4, nIters = 10, verbose = FALSE )
    expect_equal( dim( out$theta ), c( 10, 2 ) )
    expect_error( sgldSetup( gaussMeanLogLik(), list( X = rnorm( 10 ) ), list( theta = 0 ), 1e-4, minibatchSize = 0.001 ),
                  "too small for dataset size" )                        # convertProp, setup.R:128-140
} )

test_that( "the same seed gives the same chain", {
    skipWithoutDevice()
    dataset = list( X = rnorm( 1000 ) )
    run = function() sgnht( gaussMeanLogLik(), dataset, list( theta = 0 ), 5e-6, logPrior = normalPrior( 5 ), minibatchSize = 5, nIters = 10, verbose = FALSE, seed = 1 )
    expect_identical( run()$theta, run()$theta )
} )

test_that( "an ordinary closure is rejected, never differentiated on the CPU", {
    expect_error( sgldSetup( function( params, dataset ) 0, list( X = rnorm( 10 ) ), list( theta = 0 ), 1e-4 ), "cannot be differentiated on the device" )
} )

test_that( "several chains come back with a leading chain dimension; one chain keeps the reference's shapes", {
    skipWithoutDevice()
    data = list( "X" = matrix( rnorm( 2000 ), ncol = 2 ) )
    params = list( "theta" = c( 0, 0 ) )
    one = sgld( gaussMeanLogLik(), data, params, 1e-4, logPrior = normalPrior( 10 ), minibatchSize = 50, nIters = 120, verbose = FALSE, seed = 1 )
    expect_equal( dim( one$theta ), c( 120, 2 ) )
    sg = sgldSetup( gaussMeanLogLik(), data, params, 1e-4, logPrior = normalPrior( 10 ), minibatchSize = 50, seed = 1, nChains = 4L )
    many = sgmcmc:::runSGMCMC( sg, params, list( "nIters" = 120, "verbose" = FALSE ) )
    expect_equal( dim( many$theta ), c( 4, 120, 2 ) )
    expect_equal( many$theta[ 1, , ], one$theta )           # chain 1 of a multi-chain run is the single-chain run (same stream)
} )

test_that( "heldoutLogLik replaces the user-built TensorFlow tensors of the vignettes", {
    skipWithoutDevice()
    X = matrix( rnorm( 4000 ), ncol = 4 ); y = as.numeric( runif( 1000 ) < 0.5 )
    sg = sgldSetup( logisticLogLik(), list( "X" = X, "y" = y ), list( "bias" = 0, "beta" = matrix( 0, 4, 1 ) ), 1e-4,
            logPrior = laplacePrior(), minibatchSize = 100, seed = 2 )
    sess = initSess( sg, verbose = FALSE )
    ll = heldoutLogLik( sg, sess, list( "X" = X[ 1:200, ], "y" = y[ 1:200 ] ) )
    expect_equal( ll, 200 * log( 0.5 ), tolerance = 1e-6 )    # bias = 0, beta = 0: every label has probability 1/2
} )
