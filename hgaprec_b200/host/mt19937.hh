// mt19937.hh -- the generator behind HGAPRec's start state.
//
// The reference draws every initial shape / rate from GSL's default generator
// (gsl_rng_default == mt19937, src/hgaprec.cc:34-38) through gsl_rng_uniform /
// gsl_rng_uniform_int (src/gpbase.hh:299-352, 659-710, 933-945;
// src/hgaprec.cc:1718).  This is a from-scratch MT19937 (Matsumoto & Nishimura,
// 2002 initialisation) with GSL's documented conventions: seed 0 means 4357,
// uniform() = next()/2^32 in [0,1), uniform_int(n) rejects above scale*n.
#ifndef HPF_HOST_MT19937_HH
#define HPF_HOST_MT19937_HH
#include <stdint.h>

namespace hpfhost {

class Mt19937 {
public:
  explicit Mt19937(unsigned long seed = 0) { set(seed); }

  void set(unsigned long seed)
  {
    if (seed == 0) seed = 4357;
    s_[0] = (uint32_t)(seed & 0xffffffffUL);
    for (int i = 1; i < N; ++i) s_[i] = 1812433253u * (s_[i - 1] ^ (s_[i - 1] >> 30)) + (uint32_t)i;
    pos_ = N;
  }

  uint32_t next()
  {
    if (pos_ >= N) refill();
    uint32_t y = s_[pos_++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }

  double uniform() { return next() / 4294967296.0; }

  uint32_t uniform_int(uint32_t n)
  {
    const uint32_t scale = 0xffffffffu / n;
    uint32_t k;
    do {
      k = next() / scale;
    } while (k >= n);
    return k;
  }

private:
  enum { N = 624, M = 397 };
  void refill()
  {
    for (int i = 0; i < N; ++i) {
      const uint32_t y = (s_[i] & 0x80000000u) | (s_[(i + 1) % N] & 0x7fffffffu);
      s_[i] = s_[(i + M) % N] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    pos_ = 0;
  }
  uint32_t s_[N];
  int pos_;
};

// digamma for x > 0, fp64 (replaces gsl_sf_psi in the host-side start state,
// src/gpbase.hh:337,355): recurrence up to x >= 10, then the asymptotic series.
inline double digamma(double x)
{
  double r = 0.0;
  while (x < 10.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12 + f * (1.0 / 120 + f * (-1.0 / 252 + f * (1.0 / 240 + f * (-1.0 / 132 +
                   f * (691.0 / 32760 + f * (-1.0 / 12)))))));
  return r + __builtin_log(x) - 0.5 / x + t;
}

} // namespace hpfhost
#endif
