/* TEST INFRASTRUCTURE ONLY -- part of oracle/: a minimal stand-in for the GNU
 * Scientific Library symbols that premgopalan/hgaprec uses, so that the
 * UNMODIFIED reference sources under /root/reference/src can be compiled into
 * oracle/_ref/ without GSL being installed (GSL is an un-vendored, un-pinned
 * system dependency of the reference: configure.ac:17-19).
 *
 * Call sites this serves: src/hgaprec.cc:34-38 (env_setup/alloc/set),
 * src/gpbase.hh:299-352,659-710,933-945 (gsl_rng_uniform),
 * src/hgaprec.cc:1718 (gsl_rng_uniform_int).
 *
 * Semantics restated from the published GSL behaviour: the default generator
 * is mt19937 (Matsumoto & Nishimura 2002 initialisation), seed 0 maps to 4357,
 * uniform() = get()/2^32 in [0,1), uniform_int() rejects with
 * scale = 0xffffffff / n.  RNG exactness only fixes the start state, which the
 * oracle and the engine share, so any deterministic stream
 * is sufficient; "parity unpinned" applies to GSL itself, not to the CAVI path.
 */
#ifndef HPF_SHIM_GSL_RNG_H
#define HPF_SHIM_GSL_RNG_H
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { const char *name; } gsl_rng_type;
typedef struct {
  unsigned long mt[624];
  int mti;
} gsl_rng;

static const gsl_rng_type hpf_shim_mt19937 = { "mt19937" };
static const gsl_rng_type *gsl_rng_default = &hpf_shim_mt19937;
static unsigned long gsl_rng_default_seed = 0;

static inline void gsl_rng_set(gsl_rng *r, unsigned long s)
{
  if (s == 0) s = 4357;
  r->mt[0] = s & 0xffffffffUL;
  for (int i = 1; i < 624; ++i)
    r->mt[i] = (1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
  r->mti = 624;
}

static inline const gsl_rng_type *gsl_rng_env_setup(void)
{
  const char *e = getenv("GSL_RNG_SEED");
  if (e) gsl_rng_default_seed = strtoul(e, 0, 0);
  return gsl_rng_default;
}

static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *t)
{
  (void)t;
  gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
  gsl_rng_set(r, gsl_rng_default_seed);
  return r;
}

static inline void gsl_rng_free(gsl_rng *r) { free(r); }

static inline unsigned long gsl_rng_get(gsl_rng *r)
{
  unsigned long *mt = r->mt;
  if (r->mti >= 624) {
    int k;
    for (k = 0; k < 624 - 397; ++k) {
      unsigned long y = (mt[k] & 0x80000000UL) | (mt[k + 1] & 0x7fffffffUL);
      mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    for (; k < 623; ++k) {
      unsigned long y = (mt[k] & 0x80000000UL) | (mt[k + 1] & 0x7fffffffUL);
      mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    {
      unsigned long y = (mt[623] & 0x80000000UL) | (mt[0] & 0x7fffffffUL);
      mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    r->mti = 0;
  }
  unsigned long v = mt[r->mti++];
  v ^= (v >> 11);
  v ^= (v << 7) & 0x9d2c5680UL;
  v ^= (v << 15) & 0xefc60000UL;
  v ^= (v >> 18);
  return v & 0xffffffffUL;
}

static inline double gsl_rng_uniform(gsl_rng *r)
{
  return (double)gsl_rng_get(r) / 4294967296.0;
}

static inline unsigned long gsl_rng_uniform_int(gsl_rng *r, unsigned long n)
{
  unsigned long scale = 0xffffffffUL / n, k;
  do {
    k = gsl_rng_get(r) / scale;
  } while (k >= n);
  return k;
}

#ifdef __cplusplus
}
#endif
#endif
