/* TEST INFRASTRUCTURE ONLY (oracle/): gsl_sf_psi stand-in, see gsl_rng.h for
 * why this exists.  Call sites: src/gpbase.hh:260,337,355,593,694,712,923.
 *
 * Digamma for x > 0 in double precision: upward recurrence
 * psi(x) = psi(x+1) - 1/x until x >= 10, then the Stirling (Bernoulli)
 * asymptotic series through x^-14.  Checked against scipy.special.digamma to
 * <= 1e-13 relative in tests/test_oracle.py (SURVEY.md section 8c mitigation).
 * The reference never evaluates psi at x <= 0 (make_nonzero floors at 1e-30,
 * src/gpbase.hh:27-44). */
#ifndef HPF_SHIM_GSL_SF_PSI_H
#define HPF_SHIM_GSL_SF_PSI_H
#include <math.h>
#ifdef __cplusplus
extern "C" {
#endif

static inline double gsl_sf_psi(double x)
{
  double r = 0.0;
  while (x < 10.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  double f = 1.0 / (x * x);
  double t = f * (-1.0 / 12 + f * (1.0 / 120 + f * (-1.0 / 252 + f * (1.0 / 240 +
             f * (-1.0 / 132 + f * (691.0 / 32760 + f * (-1.0 / 12)))))));
  return r + log(x) - 0.5 / x + t;
}

#ifdef __cplusplus
}
#endif
#endif
