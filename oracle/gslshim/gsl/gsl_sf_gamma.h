/* TEST INFRASTRUCTURE ONLY -- stand-in for <gsl/gsl_sf_gamma.h> (reference: lynch.hpp:9,18,26). */
#ifndef SIDB200_GSLSHIM_SF_GAMMA_H
#define SIDB200_GSLSHIM_SF_GAMMA_H
#ifdef __cplusplus
extern "C" {
#endif
double gsl_sf_lngamma(double x);
#ifdef __cplusplus
}
#endif
#endif
