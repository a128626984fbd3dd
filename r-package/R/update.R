# sgmcmcStep / initSess / runSGMCMC (reference R/update.R:37-207): same generics and S3 methods; `sess` is the
# external pointer to the device state (replaces tf$Session()).
sgmcmcStep = function( sgmcmc, sess ) UseMethod( "sgmcmcStep" )
sgmcmcStep.sgld = function( sgmcmc, sess ) invisible( .Call( C_sgm_step, sess ) )
sgmcmcStep.sghmc = function( sgmcmc, sess ) invisible( .Call( C_sgm_step, sess ) )
sgmcmcStep.sgnht = function( sgmcmc, sess ) invisible( .Call( C_sgm_step, sess ) )

initSess = function( sgmcmc, verbose = TRUE ) UseMethod( "initSess" )
initSess.sgmcmc = function( sgmcmc, verbose = TRUE ) {
    .Call( C_sgm_init, sgmcmc$handle )
    return( sgmcmc$handle )
}
initSess.sgmcmccv = function( sgmcmc, verbose = TRUE ) {
    if ( verbose ) message( "\nFinding initial MAP estimates..." )
    .Call( C_sgm_init, sgmcmc$handle )    # mode finder + full-data gradient + theta := theta_hat, all on the device
    return( sgmcmc$handle )
}

runSGMCMC = function( sgmcmc, paramsRaw, options ) {
    sess = initSess( sgmcmc, options$verbose )
    .Call( C_sgm_reset_chain, sess )
    .Call( C_sgm_reserve_chain, sess, options$nIters )
    if ( options$verbose ) message( "\nSimulating from SGMCMC algorithm..." )
    done = 0
    while ( done < options$nIters ) {
        chunk = min( 100 - done %% 100, options$nIters - done )      # the reference's check / progress cadence
        .Call( C_sgm_run, sess, chunk, 1L )
        done = done + chunk
        if ( done %% 100 == 0 ) checkDivergence( sess, sgmcmc, done, options$verbose )
    }
    return( getChains( sgmcmc, sess, paramsRaw, options$nIters ) )
}

checkDivergence = function( sess, sgmcmc, iter, verbose ) {
    currentEstimate = .Call( C_sgm_est_logpost, sess, sgmcmc$nChains )
    if ( any( is.nan( currentEstimate ) ) ) stop( "Chain diverged, try decreasing stepsize" )
    if ( verbose ) message( paste0( "Iteration: ", iter, "\t\tLog posterior estimate: ", currentEstimate[1] ) )
}
