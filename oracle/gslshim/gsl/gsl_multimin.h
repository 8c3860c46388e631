/* TEST INFRASTRUCTURE ONLY -- stand-in for the subset of GNU GSL's <gsl/gsl_multimin.h>
 * that the reference imports (optimization.hpp:6,19-31,37-46,51-95; lynch.cpp:37-39).
 * GSL is not installed in this image and is not vendored by the reference (configure.ac:14-16
 * only AC_CHECK_LIBs it, version unpinned), so this header + gslshim.cpp restate the 13 symbols
 * the reference objects import.  "Parity unpinned" against real GSL -- see DESIGN.md.
 * Written from the published algorithm description (Nelder-Mead "nmsimplex2"), not from GSL sources.
 */
#ifndef SIDB200_GSLSHIM_MULTIMIN_H
#define SIDB200_GSLSHIM_MULTIMIN_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_EBADFUNC = 9 };

typedef struct { size_t size; double* data; } gsl_vector;
gsl_vector* gsl_vector_alloc(size_t n);
void gsl_vector_free(gsl_vector* v);
double gsl_vector_get(const gsl_vector* v, size_t i);
void gsl_vector_set(gsl_vector* v, size_t i, double x);

typedef struct {
    double (*f)(const gsl_vector* x, void* params);
    size_t n;
    void* params;
} gsl_multimin_function;

typedef struct { const char* name; } gsl_multimin_fminimizer_type;
extern const gsl_multimin_fminimizer_type* gsl_multimin_fminimizer_nmsimplex2;

typedef struct {
    const gsl_multimin_fminimizer_type* type;
    gsl_multimin_function* f;
    double fval;
    gsl_vector* x;
    double size;
    void* state;
} gsl_multimin_fminimizer;

gsl_multimin_fminimizer* gsl_multimin_fminimizer_alloc(const gsl_multimin_fminimizer_type* T, size_t n);
void gsl_multimin_fminimizer_free(gsl_multimin_fminimizer* s);
int gsl_multimin_fminimizer_set(gsl_multimin_fminimizer* s, gsl_multimin_function* f,
                                const gsl_vector* x, const gsl_vector* step_size);
int gsl_multimin_fminimizer_iterate(gsl_multimin_fminimizer* s);
double gsl_multimin_fminimizer_size(const gsl_multimin_fminimizer* s);
int gsl_multimin_test_size(double size, double epsabs);

#ifdef __cplusplus
}
#endif
#endif
