# Generates tests/golden/reference_v1.json FROM THE REFERENCE ITSELF (STOR-i/sgmcmc + TensorFlow), on a machine that has both.
# Neither R nor TensorFlow exists in the build image, so this script has never been run there; the oracle is "parity unpinned"
# until its output is committed.  With the file present, tests/test_reference_fixtures.py checks the oracle (and through it the CUDA
# path) against the reference's own numbers.
#
#   Rscript tests/golden/make_reference_fixtures.R          # needs library(sgmcmc), library(jsonlite)
#
# What is pinned: everything on the hot path that is deterministic in the reference --
#   * estLogPost on a given minibatch                                   (R/setup.R:116-125)
#   * d estLogPost / d theta_p per parameter tensor                     (getGradients.sgmcmc, R/setup.R:36-42)
#   * the full-data gradient at theta_hat and theta := theta_hat        (calcFullGrads, R/controlVariates.R:105-111)
#   * the control-variate estimate at theta != theta_hat                (getGradients.sgmcmccv, R/setup.R:45-62)
# The injected noise cannot be pinned (tf$random_normal inside the graph, R/dynamics.R:135-142).
# Inputs are closed-form (no RNG), and are stored in the file next to the outputs.
suppressMessages({ library(sgmcmc); library(tensorflow); library(jsonlite) })
tfd <- tryCatch(tf$distributions, error = function(e) import("tensorflow_probability")$distributions)

det_matrix <- function(n, d, a = 0.37, b = 1.31) { m <- outer(1:n, 1:d, function(i, j) sin(a * i + b * j) + 0.25 * cos(0.11 * i * j)); m }

families <- list()

# ---- Gaussian mean (vignettes/mvGauss.Rmd:37-54)
families$gauss2d <- list(
  dataset = list(X = det_matrix(64, 2)),
  params = list(theta = c(0.3, -0.2)),
  logLik = function(params, dataset) {
    tf$reduce_sum(tfd$MultivariateNormalDiag(params$theta, c(1, 1))$log_prob(dataset$X))
  },
  logPrior = function(params) { tf$reduce_sum(tfd$Normal(0, 10)$log_prob(params$theta)) },
  theta2 = list(theta = c(0.25, -0.1)))

# ---- two-component mixture (vignettes/gaussMixture.Rmd:48-76)
families$mixture <- list(
  dataset = list(X = det_matrix(64, 2, 0.53, 0.77)),
  params = list(theta1 = c(0, 0), theta2 = c(0.1, 0.1)),
  logLik = function(params, dataset) {
    cat <- tfd$Categorical(c(0.5, 0.5))
    comps <- list(tfd$MultivariateNormalDiag(params$theta1, c(1, 1)), tfd$MultivariateNormalDiag(params$theta2, c(1, 1)))
    tf$reduce_sum(tfd$Mixture(cat, comps)$log_prob(dataset$X))
  },
  logPrior = function(params) {
    p <- tfd$MultivariateNormalDiag(c(0, 0), c(10, 10))
    p$log_prob(params$theta1) + p$log_prob(params$theta2)
  },
  theta2 = list(theta1 = c(0.05, -0.02), theta2 = c(0.12, 0.07)))

# ---- logistic regression, Laplace prior (vignettes/logisticRegression.Rmd:45-57)
Xl <- det_matrix(64, 5, 0.29, 0.91)
families$logistic5 <- list(
  dataset = list(X = Xl, y = matrix(as.numeric(sin(0.7 * (1:64)) > 0), ncol = 1)),
  params = list(bias = 0.05, beta = matrix(c(0.02, -0.03, 0.01, 0.04, -0.02), nrow = 5)),
  logLik = function(params, dataset) {
    yEst <- 1 / (1 + tf$exp(-tf$squeeze(params$bias + tf$matmul(dataset$X, params$beta))))
    tf$reduce_sum(dataset$y[, 1] * tf$log(yEst) + (1 - dataset$y[, 1]) * tf$log(1 - yEst))
  },
  logPrior = function(params) { -(tf$reduce_sum(tf$abs(params$beta)) + tf$reduce_sum(tf$abs(params$bias))) },
  theta2 = list(bias = 0.07, beta = matrix(c(0.03, -0.01, 0.02, 0.03, -0.04), nrow = 5)))

idx <- c(3, 7, 12, 20, 21, 33, 47, 60)   # 1-based rows of the minibatch (n = 8)

feed_for <- function(obj, dataset, sel) {
  fd <- dict()
  for (nm in names(obj$placeholders)) {
    v <- dataset[[nm]]
    fd[[obj$placeholders[[nm]]]] <- if (is.null(dim(v))) v[sel] else v[sel, , drop = FALSE]
  }
  fd
}

out <- list(meta = list(sgmcmc = as.character(packageVersion("sgmcmc")), tensorflow = tf$version$VERSION, idx0 = idx - 1))
for (fam in names(families)) {
  f <- families[[fam]]
  rec <- list(dataset = f$dataset, params = f$params, theta2 = f$theta2, n = length(idx))
  # ---- plain estimator
  sg <- sgldSetup(f$logLik, f$dataset, f$params, 1e-4, logPrior = f$logPrior, minibatchSize = length(idx), seed = 1)
  sess <- initSess(sg, verbose = FALSE)
  fd <- feed_for(sg, f$dataset, idx)
  rec$estLogPost <- sess$run(sg$estLogPost, feed_dict = fd)
  rec$grad <- list()
  for (p in names(sg$params)) rec$grad[[p]] <- sess$run(tf$gradients(sg$estLogPost, sg$params[[p]])[[1]], feed_dict = fd)
  # ---- control variates: full gradient at theta_hat = params, then the estimate at theta2
  cv <- sgldcvSetup(f$logLik, f$dataset, f$params, 1e-4, 1e-5, logPrior = f$logPrior, minibatchSize = length(idx), nItersOpt = 100, seed = 1)
  sess2 <- tf$Session(); sess2$run(tf$global_variables_initializer())
  sgmcmc:::calcFullGrads(cv, sess2)
  rec$fullGrad <- list(); rec$cvGrad <- list()
  for (p in names(cv$params)) rec$fullGrad[[p]] <- sess2$run(cv$logPostOptGrad[[p]])
  for (p in names(cv$params)) sess2$run(cv$params[[p]]$assign(f$theta2[[p]]))
  class(cv) <- c("sgld", "sgmcmccv", "sgmcmc")
  g <- sgmcmc:::getGradients(cv)
  fd2 <- feed_for(cv, f$dataset, idx)
  for (p in names(cv$params)) rec$cvGrad[[p]] <- sess2$run(g[[p]], feed_dict = fd2)
  out[[fam]] <- rec
}
# run from the repository root
write_json(out, file.path("tests", "golden", "reference_v1.json"), digits = NA, auto_unbox = TRUE, matrix = "rowmajor")
