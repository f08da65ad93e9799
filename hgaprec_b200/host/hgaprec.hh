// hgaprec.hh -- host driver of the B200 engine, mirroring the reference's HGAPRec
// (src/hgaprec.hh:10-48): same public entry points (vb, vb_bias, vb_hier,
// gen_ranking_for_users), same report files, same stopping rule; the loop bodies
// are replaced by calls through the C ABI of include/hpf_cuda.h.
#ifndef HPF_HOST_HGAPREC_HH
#define HPF_HOST_HGAPREC_HH
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <time.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/hpf_cuda.h"
#include "gpstate.hh"
#include "mt19937.hh"
#include "ratings.hh"

namespace hpfhost {

// the subset of the reference's Env (src/env.hh) that the path reads
struct Options {
  std::string dir, label;
  uint32_t n, m, k;
  uint32_t rfreq, max_iterations, rating_threshold;
  double seed;
  double a, b, c, d;    // parsed and used in the directory name only, as in the reference
  bool binary_data, bias, hier, vb, logl, gen_ranking;
  bool csr_cache; // -csr-cache (not in the reference): keep / reuse <dir>/train.tsv.hpfcsr, see ratings.hh
  int device;
  int gpus; // > 1: shard the users over GPUs 0 .. gpus-1 (one ctx, hpf_config.n_devices)
  std::string prefix;   // output directory, Env's naming (src/env.hh:283-369)
  volatile sig_atomic_t *save_state_now;
  Options()
      : n(0), m(0), k(0), rfreq(10), max_iterations(1000), rating_threshold(1), seed(0), a(0.3), b(0.3), c(0.3), d(0.3),
        binary_data(false), bias(false), hier(false), vb(true), logl(false), gen_ranking(false), csr_cache(false), device(0), gpus(1), save_state_now(0) {}
  std::string make_prefix() const;
};

// every parameter set HGAPRec owns on this path (src/hgaprec.hh:105-115) and the
// start state HGAPRec::initialize() gives them (src/hgaprec.cc:153-204)
struct ModelState {
  GammaMatrix theta, beta;          // _theta/_beta (GPMatrixGR) or _htheta/_hbeta (GPMatrix)
  GammaMatrix thetabias, betabias;  // n x 1, m x 1
  GammaArray thetarate, betarate;   // xi, eta
  ModelState(uint32_t n, uint32_t m, uint32_t k, bool hier)
      : theta(hier ? "htheta" : "theta", 0.3, 0.3, n, k, hier), beta(hier ? "hbeta" : "beta", 0.3, 0.3, m, k, hier),
        thetabias("thetabias", 0.3, 0.3, n, 1, true), betabias("betabias", 0.3, 0.3, m, 1, true),
        thetarate("thetarate", 0.3, 0.3, n), betarate("betarate", 0.3, 0.3, m) {}
  void initialize(Mt19937 &rng, uint32_t n, uint32_t m, uint32_t k, bool hier, bool bias);
};

class HGAPRec {
public:
  HGAPRec(Options &opt, Ratings &ratings);
  ~HGAPRec();

  void vb();       // src/hgaprec.cc:919-980
  void vb_bias();  // src/hgaprec.cc:1219-1319
  void vb_hier();  // src/hgaprec.cc:1321-1436
  void gen_ranking_for_users(bool load); // src/hgaprec.cc:2087-2112

private:
  void run(bool honour_max_iterations);    // the shared loop skeleton
  void report();                           // the rfreq block, src/hgaprec.cc:1418-1428
  void compute_likelihood(bool validation);// src/hgaprec.cc:1439-1501
  void compute_precision(bool save_ranking_file); // src/hgaprec.cc:1703-1848
  void compute_itemrank(bool final);       // src/hgaprec.cc:1607-1701
  void logl();                             // src/hgaprec.cc:2160-2255
  void save_model();                       // src/hgaprec.cc:2137-2158
  bool load_beta_and_theta();              // src/hgaprec.cc:2114-2135
  void do_on_stop();                       // src/hgaprec.cc:1572-1577
  void upload_state();
  void download_state();
  void exclusions_of(const std::vector<uint32_t> &users, std::vector<uint64_t> *ptr, std::vector<uint32_t> *idx) const;
  void die(const char *what);              // lerr + exit(-1), the reference's error behaviour
  int duration() const { return (int)(time(0) - start_time_); }
  std::string out(const std::string &f) const { return opt_.prefix + f; }

  Options &opt_;
  Ratings &ratings_;
  uint32_t n_, m_, k_;
  uint32_t iter_;
  time_t start_time_;
  ModelState st_;
  Mt19937 rng_;
  HeldoutMap validation_map_, test_map_;
  std::map<uint32_t, bool> sampled_users_;
  double prev_h_;
  int nh_;
  uint32_t topn_by_user_;
  FILE *vf_, *tf_, *pf_, *af_, *logf_;
  hpf_ctx *ctx_;
};

} // namespace hpfhost
#endif
