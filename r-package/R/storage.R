# getParams / chain copy-out (reference R/storage.R:37-80).  The chain lives on the device and is copied out
# once, already in the reference's (nIters, dims...) column-major layout per parameter.
# nChains = 1 (the reference's behaviour) returns exactly the reference's shapes; nChains > 1 (extension) adds a LEADING chain dimension.
getParams = function( sgmcmc, sess ) {
    flat = .Call( C_sgm_get_params, sess, sgmcmc$nChains )       # P x nChains
    C = sgmcmc$nChains
    out = list(); o = 0
    for ( pname in names( sgmcmc$params ) ) {
        shp = dim( sgmcmc$params[[pname]] ); sz = length( sgmcmc$params[[pname]] )
        vals = flat[ ( o + 1 ):( o + sz ), , drop = FALSE ]       # sz x nChains
        if ( C == 1 ) out[[pname]] = if ( is.null( shp ) ) vals[ , 1 ] else array( vals[ , 1 ], dim = shp )
        else out[[pname]] = aperm( array( vals, dim = c( if ( is.null( shp ) ) sz else shp, C ) ), c( length( shp ) + 1 + is.null( shp ), seq_len( max( 1, length( shp ) ) ) ) )
        o = o + sz
    }
    return( out )
}

# buf: per chain, for each parameter an (nIters, dims...) column-major block (the layout of initStorage, reference R/storage.R:42-51)
splitChains = function( sgmcmc, buf, paramsRaw, nIters ) {
    C = sgmcmc$nChains; P = sgmcmc$P
    out = list(); o = 0
    for ( pname in names( paramsRaw ) ) {
        shp = c( nIters, getShape( paramsRaw[[pname]] ) ); sz = prod( shp )
        if ( C == 1 ) {
            out[[pname]] = array( buf[ ( o + 1 ):( o + sz ) ], dim = shp )
        } else {      # nChains x nIters x dims...
            per = array( 0, dim = c( C, shp ) )
            for ( c in seq_len( C ) ) {
                base = ( c - 1 ) * nIters * P + o
                slab = array( buf[ ( base + 1 ):( base + sz ) ], dim = shp )
                per[ matrix( c( rep( c, sz ), as.matrix( expand.grid( lapply( shp, seq_len ) ) ) ), nrow = sz ) ] = slab
            }
            out[[pname]] = per
        }
        o = o + sz
    }
    return( out )
}

# the device-resident chain of the step-by-step interface (sgm_run with store = 1): one copy-out
getChains = function( sgmcmc, sess, paramsRaw, nIters ) {
    buf = .Call( C_sgm_copy_chain, sess, sgmcmc$nChains )
    return( splitChains( sgmcmc, buf, paramsRaw, nIters ) )
}
