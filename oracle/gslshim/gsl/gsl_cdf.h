/* TEST INFRASTRUCTURE ONLY -- stand-in for <gsl/gsl_cdf.h> (reference: stats.cpp:3,33,35). */
#ifndef SIDB200_GSLSHIM_CDF_H
#define SIDB200_GSLSHIM_CDF_H
#ifdef __cplusplus
extern "C" {
#endif
double gsl_cdf_chisq_Q(double x, double nu);
#ifdef __cplusplus
}
#endif
#endif
