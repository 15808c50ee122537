// RcppArmadillo.h (shim) -- TEST INFRASTRUCTURE: see armadillo_min.hpp
#pragma once
#include "armadillo_min.hpp"
#include "Rcpp.h"
