# sgmcmcStep / initSess / runSGMCMC (reference R/update.R:37-207): same generics and S3 methods; `sess` is the
# external pointer to the device state (replaces tf$Session()).
sgmcmcStep = function( sgmcmc, sess ) UseMethod( "sgmcmcStep" )
sgmcmcStep.sgld = function( sgmcmc, sess ) invisible( .Call( C_sgm_step, sess ) )
sgmcmcStep.sghmc = function( sgmcmc, sess ) invisible( .Call( C_sgm_step, sess ) )
sgmcmcStep.sgnht = function( sgmcmc, sess ) invisible( .Call( C_sgm_step, sess ) )

initSess = function( sgmcmc, verbose = TRUE ) UseMethod( "initSess" )
initSess.sgmcmc = function( sgmcmc, verbose = TRUE ) {
    .Call( C_sgm_init, sgmcmc$handle, as.integer( verbose ) )
    return( sgmcmc$handle )
}
initSess.sgmcmccv = function( sgmcmc, verbose = TRUE ) {
    if ( verbose ) message( "\nFinding initial MAP estimates..." )
    # mode finder (progress line + NaN check every 100 iterations, printed by the shim), full-data gradient, theta := theta_hat: on the device
    .Call( C_sgm_init, sgmcmc$handle, as.integer( verbose ) )
    return( sgmcmc$handle )
}

# reference R/update.R:145-163.  The loop, storeState and checkDivergence (every 100 iterations) run inside ONE library call; the chain
# streams to the host while later iterations are sampled.  "Chain diverged, try decreasing stepsize" arrives as an R error.
runSGMCMC = function( sgmcmc, paramsRaw, options ) {
    sess = initSess( sgmcmc, options$verbose )
    if ( options$verbose ) message( "\nSimulating from SGMCMC algorithm..." )
    buf = .Call( C_sgm_sample, sess, options$nIters, sgmcmc$nChains, as.integer( options$verbose ) )
    return( splitChains( sgmcmc, buf, paramsRaw, options$nIters ) )
}

# sess$run( sgmcmc$estLogPost, fresh minibatch ) of the reference (R/update.R:198-199), one value per chain
estLogPost = function( sgmcmc, sess ) .Call( C_sgm_est_logpost, sess, sgmcmc$nChains )

# The whole sampler state as a raw vector (parameters, momenta, thermostats, the control-variate mode and its gradient, seed and
# position in the random streams): saveRDS( saveState( sgmcmc, sess ), file ) ... loadState( sgmcmc, readRDS( file ) ) on an object
# made by the same *Setup call takes the place of initSess -- no second mode search -- and the run continues bit for bit.
# (The reference leaves persistence to the user of the step-by-step API, who can keep getParams() output only: NEWS.md:19.)
saveState = function( sgmcmc, sess ) {
    .Call( C_sgm_save_state, sess )
}
loadState = function( sgmcmc, blob ) {
    .Call( C_sgm_load_state, sgmcmc$handle, blob )
    sgmcmc$handle
}

# logLik( current params, testset ) per chain, on the device: replaces the TensorFlow tensors a user of the reference builds from
# sgld$params to monitor a test-set log loss (vignettes/nn.Rmd:113-123, logisticRegression.Rmd:89-115)
heldoutLogLik = function( sgmcmc, sess, testset ) {
    m = resolveModel( sgmcmc$logLik, NULL )
    data = lapply( testset[ m$dataNames ], function( x ) { storage.mode( x ) = "double"; x } )
    if ( any( sapply( data, is.null ) ) ) stop( paste0( "testset entries required: ", paste( m$dataNames, collapse = ", " ) ) )
    .Call( C_sgm_heldout_loglik, sess, data, sgmcmc$nChains )
}
