# Model-spec objects that stand where the reference takes logLik / logPrior closures.
# R closures cannot be differentiated on the device: the four vignette families are tagged
# constructors routed to hand-written gradient kernels; anything else stops with a clear message.
gaussMeanLogLik = function( Sigma = NULL ) structure( list( model = 1L, Sigma = Sigma, dataNames = "X" ), class = "sgmcmc_loglik" )
mixture2LogLik = function() structure( list( model = 2L, dataNames = "X" ), class = "sgmcmc_loglik" )
logisticLogLik = function() structure( list( model = 3L, dataNames = c( "X", "y" ) ), class = "sgmcmc_loglik" )
softmaxNetLogLik = function() structure( list( model = 4L, dataNames = c( "X", "y" ) ), class = "sgmcmc_loglik" )
normalPrior = function( sd ) structure( list( kind = "normal", sd = sd ), class = "sgmcmc_logprior" )
laplacePrior = function() structure( list( kind = "laplace" ), class = "sgmcmc_logprior" )
gammaPrecisionPrior = function() structure( list( kind = "gammaPrecision" ), class = "sgmcmc_logprior" )

notRouted = paste( "logLik/logPrior must be model-spec objects (gaussMeanLogLik, mixture2LogLik, logisticLogLik,",
        "softmaxNetLogLik and their priors): arbitrary closures cannot be differentiated on the device",
        "and there is no CPU fallback" )

resolveModel = function( logLik, logPrior ) {
    if ( !inherits( logLik, "sgmcmc_loglik" ) ) stop( notRouted )
    if ( !is.null( logPrior ) && !inherits( logPrior, "sgmcmc_logprior" ) ) stop( notRouted )
    want = c( "normal", "normal", "laplace", "gammaPrecision" )[ logLik$model ]
    if ( !is.null( logPrior ) && logPrior$kind != want ) stop( paste0( "this logLik is routed with the ", want, " prior (or logPrior = NULL)" ) )
    precision = if ( !is.null( logLik$Sigma ) ) solve( as.matrix( logLik$Sigma ) ) else NULL
    list( model = logLik$model, usePrior = as.integer( !is.null( logPrior ) ),
            priorSd = if ( !is.null( logPrior ) && logPrior$kind == "normal" ) logPrior$sd else 0, precision = precision,
            dataNames = logLik$dataNames )
}
