// Host mirror of the reference's Lynch-estimator interface (lynch.hpp:32-46): same names and meaning; the
// optimiser, the objective and the per-profile likelihoods run on the GPU (sidgpu_lynch_fit, sidgpu_lynch_objective,
// sidgpu_profile_loglik).  compoundLikelihood takes (pi, eps) directly instead of a gsl_vector (GSL is not a
// dependency here).
#pragma once
#include <array>
#include <vector>

#include "pileup.hpp"

typedef struct GenotypeLikelihood {                    // lynch.hpp:32-35
    long double L_homozygous;
    long double L_heterozygous;
} GenotypeLikelihood;

typedef struct {                                       // lynch.hpp:37-41
    double heterozygosity;
    double error_rate;
    std::vector<GenotypeLikelihood> profile_likelihoods;
} ProfileGenotypeLikelihoods;

// lynch.cpp:17-35: Nelder-Mead from (1e-3, 1e-3) on compoundLikelihood, then L_hom / L_het of every profile at the
// fitted error rate.  The likelihoods are computed in log space on the device and exponentiated here, so values the
// reference's long double still represents below 1e-308 come out as their long double exponentials as well.
ProfileGenotypeLikelihoods estimateProfileGenotypeLikelihoods(const std::vector<UniqueProfile>&, const std::array<double, 4>);

// lynch.cpp:37-61: -sum count * log((1 - pi) L_hom + pi L_het); DBL_MAX outside [0,1]^2.
double compoundLikelihood(double pi, double epsilon, const std::vector<UniqueProfile>&, const std::array<double, 4>);
