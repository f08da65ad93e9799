// modelstate.cc -- the start state (host only; no engine dependency)
#include "hgaprec.hh"

namespace hpfhost {

// draw order of HGAPRec::initialize (153-204)
void ModelState::initialize(Mt19937 &rng, uint32_t n, uint32_t m, uint32_t k, bool hier, bool bias)
{
  if (!hier) {
    beta.initialize(rng);
    theta.initialize(rng);
    beta.initialize_exp(rng);
    theta.initialize_exp(rng);
  } else {
    thetarate.initialize2(rng, k);
    thetarate.compute_expectations();
    betarate.initialize2(rng, k);
    betarate.compute_expectations();
    beta.initialize(rng);
    beta.initialize_exp(rng);
    theta.initialize(rng);
    theta.initialize_exp(rng);
  }
  if (bias) {
    thetabias.initialize2(rng, m);
    thetabias.compute_expectations();
    betabias.initialize2(rng, n);
    betabias.compute_expectations();
  }
}

} // namespace hpfhost
