library(testthat)
library(sgmcmc)
test_check("sgmcmc")
