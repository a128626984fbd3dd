# Same signatures and defaults as the reference (R/sgld.R:50-58,114-123; R/sghmc.R; R/sgnht.R; R/sgldcv.R; ...).
sgld = function( logLik, dataset, params, stepsize, logPrior = NULL, minibatchSize = 0.01,
            nIters = 10^4L, verbose = TRUE, seed = NULL ) {
    sgld = sgldSetup( logLik, dataset, params, stepsize, logPrior, minibatchSize, seed )
    runSGMCMC( sgld, params, list( "nIters" = nIters, "verbose" = verbose ) )
}
sgldSetup = function( logLik, dataset, params, stepsize, logPrior = NULL, minibatchSize = 0.01, seed = NULL, ... ) {
    sgld = createSGMCMC( logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, algo = 1L, ... )
    class( sgld ) = c( "sgld", "sgmcmc" )
    return( sgld )
}
sghmc = function( logLik, dataset, params, stepsize, logPrior = NULL, minibatchSize = 0.01,
            alpha = 0.01, L = 5L, nIters = 10^4L, verbose = TRUE, seed = NULL ) {
    sghmc = sghmcSetup( logLik, dataset, params, stepsize, logPrior, minibatchSize, alpha, L, seed )
    runSGMCMC( sghmc, params, list( "nIters" = nIters, "verbose" = verbose ) )
}
sghmcSetup = function( logLik, dataset, params, stepsize, logPrior = NULL, minibatchSize = 0.01,
            alpha = 0.01, L = 5L, seed = NULL, ... ) {
    sghmc = createSGMCMC( logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, algo = 2L, alpha = alpha, L = L, ... )
    sghmc$alpha = convertList( alpha, params ); sghmc$L = L
    class( sghmc ) = c( "sghmc", "sgmcmc" )
    return( sghmc )
}
sgnht = function( logLik, dataset, params, stepsize, logPrior = NULL, minibatchSize = 0.01,
            a = 0.01, nIters = 10^4L, verbose = TRUE, seed = NULL ) {
    sgnht = sgnhtSetup( logLik, dataset, params, stepsize, logPrior, minibatchSize, a, seed )
    runSGMCMC( sgnht, params, list( "nIters" = nIters, "verbose" = verbose ) )
}
sgnhtSetup = function( logLik, dataset, params, stepsize, logPrior = NULL, minibatchSize = 0.01, a = 0.01, seed = NULL, ... ) {
    sgnht = createSGMCMC( logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, algo = 3L, a = a, ... )
    sgnht$a = convertList( a, params )
    class( sgnht ) = c( "sgnht", "sgmcmc" )
    return( sgnht )
}
sgldcv = function( logLik, dataset, params, stepsize, optStepsize, logPrior = NULL, minibatchSize = 0.01,
            nIters = 10^4L, nItersOpt = 10^4L, verbose = TRUE, seed = NULL ) {
    sgldcv = sgldcvSetup( logLik, dataset, params, stepsize, optStepsize, logPrior, minibatchSize, nItersOpt, verbose, seed )
    runSGMCMC( sgldcv, params, list( "nIters" = nIters, "nItersOpt" = nItersOpt, "verbose" = verbose ) )
}
sgldcvSetup = function( logLik, dataset, params, stepsize, optStepsize, logPrior = NULL, minibatchSize = 0.01,
            nItersOpt = 10^4L, verbose = TRUE, seed = NULL, ... ) {
    sgldcv = createSGMCMC( logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, algo = 1L, cv = TRUE,
            optStepsize = optStepsize, nItersOpt = nItersOpt, ... )
    sgldcv$nItersOpt = nItersOpt
    class( sgldcv ) = c( "sgld", "sgmcmccv", "sgmcmc" )
    return( sgldcv )
}
sghmccv = function( logLik, dataset, params, stepsize, optStepsize, logPrior = NULL, minibatchSize = 0.01,
            alpha = 0.01, L = 5L, nIters = 10^4L, nItersOpt = 10^4L, verbose = TRUE, seed = NULL ) {
    sghmccv = sghmccvSetup( logLik, dataset, params, stepsize, optStepsize, logPrior, minibatchSize, alpha, L, nItersOpt, verbose, seed )
    runSGMCMC( sghmccv, params, list( "nIters" = nIters, "nItersOpt" = nItersOpt, "verbose" = verbose ) )
}
sghmccvSetup = function( logLik, dataset, params, stepsize, optStepsize, logPrior = NULL, minibatchSize = 0.01,
            alpha = 0.01, L = 5L, nItersOpt = 10^4L, verbose = TRUE, seed = NULL, ... ) {
    sghmccv = createSGMCMC( logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, algo = 2L, cv = TRUE,
            optStepsize = optStepsize, nItersOpt = nItersOpt, alpha = alpha, L = L, ... )
    sghmccv$alpha = convertList( alpha, params ); sghmccv$L = L; sghmccv$nItersOpt = nItersOpt
    class( sghmccv ) = c( "sghmc", "sgmcmccv", "sgmcmc" )
    return( sghmccv )
}
sgnhtcv = function( logLik, dataset, params, stepsize, optStepsize, logPrior = NULL, minibatchSize = 0.01,
            a = 0.01, nIters = 10^4L, nItersOpt = 10^4L, verbose = TRUE, seed = NULL ) {
    sgnhtcv = sgnhtcvSetup( logLik, dataset, params, stepsize, optStepsize, logPrior, minibatchSize, a, nItersOpt, verbose, seed )
    runSGMCMC( sgnhtcv, params, list( "nIters" = nIters, "nItersOpt" = nItersOpt, "verbose" = verbose ) )
}
sgnhtcvSetup = function( logLik, dataset, params, stepsize, optStepsize, logPrior = NULL, minibatchSize = 0.01,
            a = 0.01, nItersOpt = 10^4L, verbose = TRUE, seed = NULL, ... ) {
    sgnhtcv = createSGMCMC( logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, algo = 3L, cv = TRUE,
            optStepsize = optStepsize, nItersOpt = nItersOpt, a = a, ... )
    sgnhtcv$a = convertList( a, params ); sgnhtcv$nItersOpt = nItersOpt
    class( sgnhtcv ) = c( "sgnht", "sgmcmccv", "sgmcmc" )
    return( sgnhtcv )
}
