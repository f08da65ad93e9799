/* TEST INFRASTRUCTURE ONLY (oracle/): gsl_sf_lngamma stand-in (ELBO path only,
 * src/gpbase.hh:373-383,729-737,964-966). */
#ifndef HPF_SHIM_GSL_SF_H
#define HPF_SHIM_GSL_SF_H
#include <math.h>
#include "gsl_sf_psi.h"
#ifdef __cplusplus
extern "C" {
#endif
static inline double gsl_sf_lngamma(double x) { return lgamma(x); }
#ifdef __cplusplus
}
#endif
#endif
