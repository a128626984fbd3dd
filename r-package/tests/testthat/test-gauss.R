context( "Gaussian-mean posteriors, all six samplers (ports of test1dGauss.R and test3dGauss.R)" )

test_that( "1-d Gaussian, Normal(0, 5) prior", {
    skipWithoutDevice()
    set.seed( 13 )
    dataset = list( X = rnorm( 10^4 ) )
    for ( name in names( samplers ) ) {
        out = samplers[[ name ]]( gaussMeanLogLik(), dataset, list( theta = rnorm( 1, 0, 0.1 ) ), normalPrior( 5 ) )
        expect_equal( dim( out$theta ), c( 1100, 1 ), info = name )      # scalar parameter -> (nIters, 1), storage.R:42-51
        expectNearZero( out$theta )
    }
} )

test_that( "3-d Gaussian with full covariance", {
    skipWithoutDevice()
    set.seed( 13 )
    Sigma = matrix( c( 1, 0.3, 0.5, 0.3, 1, 0.2, 0.5, 0.2, 1 ), 3 )
    dataset = list( X = matrix( rnorm( 3 * 10^4 ), ncol = 3 ) %*% chol( Sigma ) )
    for ( name in names( samplers ) ) {
        out = samplers[[ name ]]( gaussMeanLogLik( Sigma ), dataset, list( theta = rnorm( 3, 0, 0.1 ) ), normalPrior( 5 ) )
        expect_equal( dim( out$theta ), c( 1100, 3 ), info = name )
        expectNearZero( out$theta )
    }
} )
