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
