// Host mirror of the reference's statistics interface used on the calling path (stats.hpp:8,11); both run on the
// GPU (sidgpu_lr_test, sidgpu_bh_adjust).
#pragma once
#include <vector>

// stats.cpp:29-37: p-value of the likelihood ratio test with one degree of freedom.
double likelihoodRatioTest(long double l_H0, long double l_H1);
// stats.cpp:58-80
std::vector<double> adjustBenjaminiHochberg(const std::vector<double>&);
