# createSGMCMC (reference R/setup.R:2-28): same fields, but the "graph" is a handle to device state.
createSGMCMC = function( logLik, logPrior, dataset, params, stepsize, minibatchSize, seed, algo, cv = FALSE,
        optStepsize = 0, nItersOpt = 0, alpha = 0.01, L = 5L, a = 0.01, nChains = 1L, device = 0L, sharedMode = FALSE ) {
    m = resolveModel( logLik, logPrior )
    stepsize = convertList( stepsize, params )
    alpha = convertList( alpha, params )
    a = convertList( a, params )
    data = lapply( dataset[ m$dataNames ], function( x ) { storage.mode( x ) = "double"; x } )
    if ( any( sapply( data, is.null ) ) ) stop( paste0( "dataset entries required: ", paste( m$dataNames, collapse = ", " ) ) )
    pars = lapply( params, function( x ) { storage.mode( x ) = "double"; x } )
    spec = list( m$model, algo, as.integer( cv ), minibatchSize, as.integer( L ), optStepsize, nItersOpt, seed,
            as.integer( nChains ), as.integer( device ), as.integer( sharedMode ), m$usePrior, m$priorSd, m$precision )
    if ( !is.null( seed ) ) set.seed( seed )
    handle = .Call( C_sgm_create, spec, data, pars, as.double( unlist( stepsize[ names( params ) ] ) ),
            as.double( unlist( alpha[ names( params ) ] ) ), as.double( unlist( a[ names( params ) ] ) ) )
    info = .Call( C_sgm_info, handle )
    list( "N" = info[1], "data" = dataset, "n" = info[2], "stepsize" = stepsize, "params" = params,
            "handle" = handle, "nChains" = as.integer( nChains ), "P" = info[3], "logLik" = logLik )
}

# If stepsizes or a or alpha tuning constants are not in a list format then convert (reference R/setup.R:160-170)
convertList = function( tuningConst, params ) {
    if ( typeof( tuningConst ) == "list" ) return( tuningConst )
    out = list()
    for ( pname in names( params ) ) out[[pname]] = tuningConst
    return( out )
}

getShape = function( input ) {
    shapeInput = dim( input )
    if ( is.null( shapeInput ) ) shapeInput = c( length( input ) )
    return( shapeInput )
}
