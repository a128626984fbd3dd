# getParams / chain copy-out (reference R/storage.R:37-80).  The chain lives on the device and is copied out
# once, already in the reference's (nIters, dims...) column-major layout per parameter.
getParams = function( sgmcmc, sess ) {
    flat = .Call( C_sgm_get_params, sess, sgmcmc$nChains )       # P x nChains
    out = list(); o = 0
    for ( pname in names( sgmcmc$params ) ) {
        shp = dim( sgmcmc$params[[pname]] ); sz = length( sgmcmc$params[[pname]] )
        vals = flat[ ( o + 1 ):( o + sz ), 1 ]
        out[[pname]] = if ( is.null( shp ) ) vals else array( vals, dim = shp )
        o = o + sz
    }
    return( out )
}

getChains = function( sgmcmc, sess, paramsRaw, nIters ) {
    buf = .Call( C_sgm_copy_chain, sess, sgmcmc$nChains )        # per chain: for each parameter an (nIters, dims...) block
    out = list(); o = 0
    for ( pname in names( paramsRaw ) ) {
        shp = c( nIters, getShape( paramsRaw[[pname]] ) ); sz = prod( shp )
        out[[pname]] = array( buf[ ( o + 1 ):( o + sz ) ], dim = shp )   # chain 1 (nChains = 1 reproduces the reference)
        o = o + sz
    }
    return( out )
}
