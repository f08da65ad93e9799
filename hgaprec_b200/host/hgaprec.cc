// hgaprec.cc -- see hgaprec.hh.  Reference line numbers are into src/hgaprec.cc.
#include "hgaprec.hh"

#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include <algorithm>
#include <sstream>

namespace hpfhost {

// Env::Env's directory name (src/env.hh:283-369): n/m as given on the command line
std::string Options::make_prefix() const
{
  std::ostringstream sa;
  sa << "n" << n << "-m" << m << "-k" << k;
  if (label != "") sa << "-" << label;
  else if (dir.length() > 3) {
    const std::string q = dir.substr(0, 2);
    if (isalpha(q[0])) sa << "-" << q;
  }
  if (a != 0.3) sa << "-a" << a;
  if (b != 0.3) sa << "-b" << b;
  if (c != 0.3) sa << "-c" << c;
  if (d != 0.3) sa << "-d" << d;
  sa << "-batch";
  if (binary_data) sa << "-bin";
  if (bias) sa << "-bias";
  if (hier) sa << "-hier";
  if (vb) sa << "-vb";
  if (seed) sa << "-seed" << seed;
  return sa.str();
}

void HGAPRec::die(const char *what)
{
  const char *msg = hpf_last_error(ctx_);
  fprintf(stderr, "hgaprec: %s: %s\n", what, msg);
  if (logf_) {
    fprintf(logf_, "%s: %s\n", what, msg);
    fflush(logf_);
  }
  exit(-1);
}

HGAPRec::HGAPRec(Options &opt, Ratings &ratings)
    : opt_(opt), ratings_(ratings), n_(ratings.n()), m_(ratings.m()), k_(opt.k), iter_(0), start_time_(time(0)),
      st_(n_, m_, k_, opt.hier), rng_(0), prev_h_(0.0), nh_(0),
      topn_by_user_(100), vf_(0), tf_(0), pf_(0), af_(0), logf_(0), ctx_(0)
{
  // gsl_rng_alloc + gsl_rng_set(seed) only when the seed is non-zero (8-38); the
  // seed is a double truncated to unsigned long
  if (opt_.seed) rng_.set((unsigned long)opt_.seed);
  logf_ = fopen(out("/infer.log").c_str(), "a");
  // the reference opens (and leaves empty) heldout/ndcg/rmse as well (40-74)
  const char *empties[] = { "/heldout.txt", "/ndcg.txt", "/rmse.txt" };
  for (size_t i = 0; i < 3; ++i) {
    FILE *f = fopen(out(empties[i]).c_str(), "w");
    if (f) fclose(f);
  }
  af_ = fopen(out("/logl.txt").c_str(), "w"); // _af (56-60): one "%.5f" line per report window under -logl
  if (!af_) {
    printf("cannot open logl file:%s\n", strerror(errno));
    exit(-1);
  }
  vf_ = fopen(out("/validation.txt").c_str(), "w");
  tf_ = fopen(out("/test.txt").c_str(), "w");
  pf_ = fopen(out("/precision.txt").c_str(), "w");
  if (!vf_ || !tf_ || !pf_) {
    printf("cannot open heldout file:%s\n", strerror(errno));
    exit(-1);
  }
  // load_validation_and_test_sets (110-151)
  std::string err;
  if (!ratings_.read_heldout(opt_.dir + "/validation.tsv", &validation_map_, &err) ||
      !ratings_.read_heldout(opt_.dir + "/test.tsv", &test_map_, &err)) {
    fprintf(stderr, "hgaprec: %s\n", err.c_str());
    exit(-1);
  }
  printf("+ loaded validation and test sets from %s\n", opt_.dir.c_str());
  fflush(stdout);

  hpf_config cfg;
  hpf_config_default(&cfg);
  cfg.n_users = n_;
  cfg.n_items = m_;
  cfg.k = k_;
  cfg.flags = (opt_.hier ? HPF_HIER : 0u) | (opt_.bias ? HPF_BIAS : 0u) | (opt_.binary_data ? HPF_BINARY : 0u) |
              (!opt_.vb ? HPF_JACOBI : 0u) | (opt_.logl ? HPF_LOGL : 0u);
  cfg.device = opt_.device;
  if (opt_.gpus > 1) {
    if (opt_.gpus > HPF_MAX_DEVICES) opt_.gpus = HPF_MAX_DEVICES;
    cfg.n_devices = (uint32_t)opt_.gpus;
    for (int g = 0; g < opt_.gpus; ++g) cfg.devices[g] = g;
  }
  if (hpf_create(&cfg, &ctx_) != 0) die("hpf_create");
}

HGAPRec::~HGAPRec()
{
  if (vf_) fclose(vf_);
  if (tf_) fclose(tf_);
  if (pf_) fclose(pf_);
  if (af_) fclose(af_);
  if (logf_) fclose(logf_);
  hpf_destroy(ctx_);
}

void HGAPRec::upload_state()
{
  if (hpf_set_state(ctx_, HPF_THETA, st_.theta.shape_curr.data(), st_.theta.rate_curr.data(), st_.theta.expected_v.data(),
                    st_.theta.expected_logv.data()) != 0 ||
      hpf_set_state(ctx_, HPF_BETA, st_.beta.shape_curr.data(), st_.beta.rate_curr.data(), st_.beta.expected_v.data(),
                    st_.beta.expected_logv.data()) != 0)
    die("hpf_set_state");
  if (opt_.hier &&
      (hpf_set_state(ctx_, HPF_THETARATE, st_.thetarate.shape_curr.data(), st_.thetarate.rate_curr.data(), st_.thetarate.expected_v.data(),
                     st_.thetarate.expected_logv.data()) != 0 ||
       hpf_set_state(ctx_, HPF_BETARATE, st_.betarate.shape_curr.data(), st_.betarate.rate_curr.data(), st_.betarate.expected_v.data(),
                     st_.betarate.expected_logv.data()) != 0))
    die("hpf_set_state");
  if (opt_.bias &&
      (hpf_set_state(ctx_, HPF_THETABIAS, st_.thetabias.shape_curr.data(), st_.thetabias.rate_curr.data(), st_.thetabias.expected_v.data(),
                     st_.thetabias.expected_logv.data()) != 0 ||
       hpf_set_state(ctx_, HPF_BETABIAS, st_.betabias.shape_curr.data(), st_.betabias.rate_curr.data(), st_.betabias.expected_v.data(),
                     st_.betabias.expected_logv.data()) != 0))
    die("hpf_set_state");
}

// device -> the GPMatrix fields every host consumer reads (save_model)
void HGAPRec::download_state()
{
  if (hpf_get_state(ctx_, HPF_THETA, st_.theta.shape_curr.data(), st_.theta.rate_curr.data(), st_.theta.expected_v.data(),
                    st_.theta.expected_logv.data()) != 0 ||
      hpf_get_state(ctx_, HPF_BETA, st_.beta.shape_curr.data(), st_.beta.rate_curr.data(), st_.beta.expected_v.data(),
                    st_.beta.expected_logv.data()) != 0)
    die("hpf_get_state");
  if (opt_.hier &&
      (hpf_get_state(ctx_, HPF_THETARATE, st_.thetarate.shape_curr.data(), st_.thetarate.rate_curr.data(), st_.thetarate.expected_v.data(),
                     st_.thetarate.expected_logv.data()) != 0 ||
       hpf_get_state(ctx_, HPF_BETARATE, st_.betarate.shape_curr.data(), st_.betarate.rate_curr.data(), st_.betarate.expected_v.data(),
                     st_.betarate.expected_logv.data()) != 0))
    die("hpf_get_state");
  if (opt_.bias &&
      (hpf_get_state(ctx_, HPF_THETABIAS, st_.thetabias.shape_curr.data(), st_.thetabias.rate_curr.data(), st_.thetabias.expected_v.data(),
                     st_.thetabias.expected_logv.data()) != 0 ||
       hpf_get_state(ctx_, HPF_BETABIAS, st_.betabias.shape_curr.data(), st_.betabias.rate_curr.data(), st_.betabias.expected_v.data(),
                     st_.betabias.expected_logv.data()) != 0))
    die("hpf_get_state");
}

void HGAPRec::vb() { run(false); }
void HGAPRec::vb_bias() { run(false); }
void HGAPRec::vb_hier() { run(true); }

// The three reference loops share one skeleton: sweep + updates (now hpf_iterate),
// then every rfreq iterations the report block, then the SIGTERM poll.  Only
// vb_hier checks max_iterations (1337-1339) and it runs iterations
// 0..max_iterations inclusive; vb / vb_bias end through the stopping rule.
// Between two reports nothing on the host reads the state, so the device runs
// the whole window in one call.
void HGAPRec::run(bool honour_max_iterations)
{
  st_.initialize(rng_, n_, m_, k_, opt_.hier, opt_.bias);
  std::vector<uint64_t> row_ptr;
  std::vector<uint32_t> col_idx;
  std::vector<uint8_t> y;
  ratings_.to_csr(&row_ptr, &col_idx, &y);
  if (hpf_set_ratings_csr(ctx_, row_ptr.data(), col_idx.data(), opt_.binary_data ? NULL : y.data()) != 0) die("hpf_set_ratings_csr");
  upload_state();
  const uint32_t rfreq = opt_.rfreq ? opt_.rfreq : 1;
  for (;;) {
    if (honour_max_iterations && iter_ > opt_.max_iterations) exit(0);
    // iterations iter_ .. last, where `last` is the next report point
    uint32_t last = (iter_ % rfreq == 0) ? iter_ : (iter_ / rfreq + 1) * rfreq;
    if (honour_max_iterations && last > opt_.max_iterations) last = opt_.max_iterations;
    if (opt_.save_state_now && *opt_.save_state_now) last = iter_; // after SIGTERM the reference re-saves every iteration
    if (hpf_iterate(ctx_, last - iter_ + 1) != 0) die("hpf_iterate");
    iter_ = last;
    printf("\r iteration %d", iter_);
    fflush(stdout);
    if (iter_ % rfreq == 0) report();
    if (opt_.save_state_now && *opt_.save_state_now) {
      if (logf_) fprintf(logf_, "Saving state at iteration %d duration %d secs\n", iter_, duration());
      do_on_stop();
    }
    iter_++;
  }
}

void HGAPRec::report()
{
  compute_likelihood(true);
  compute_likelihood(false);
  save_model();
  compute_precision(false);
  if (opt_.hier || !opt_.bias) compute_itemrank(false); // vb_bias's report block has no compute_itemrank (1301-1310)
  if (opt_.logl) logl();
}

// HGAPRec::logl (2160-2255): the sweep over the training nonzeros and the Gamma terms run on the device
void HGAPRec::logl()
{
  double s = 0.0;
  if (hpf_elbo(ctx_, &s) != 0) die("hpf_elbo");
  fprintf(af_, "%.5f\n", s);
  fflush(af_);
}

void HGAPRec::compute_likelihood(bool validation)
{
  const HeldoutMap &mp = validation ? validation_map_ : test_map_;
  FILE *ff = validation ? vf_ : tf_;
  std::vector<uint32_t> u, i;
  std::vector<uint8_t> yy;
  u.reserve(mp.size()); i.reserve(mp.size()); yy.reserve(mp.size());
  for (HeldoutMap::const_iterator it = mp.begin(); it != mp.end(); ++it) {
    u.push_back(it->first.first);
    i.push_back(it->first.second);
    yy.push_back(it->second);
  }
  double s = 0.0;
  if (hpf_heldout_loglik(ctx_, u.data(), i.data(), yy.data(), u.size(), &s) != 0) die("hpf_heldout_loglik");
  const uint32_t k = (uint32_t)u.size();
  fprintf(ff, "%d\t%d\t%.9f\t%d\n", iter_, duration(), s / k, k);
  fflush(ff);
  const double a = s / k;
  if (!validation) return;
  // the stopping rule (1476-1491)
  bool stop = false;
  int why = -1;
  if (iter_ > 30) {
    if (a > prev_h_ && prev_h_ != 0 && fabs((a - prev_h_) / prev_h_) < 0.000001) {
      stop = true;
      why = 0;
    } else if (a < prev_h_)
      nh_++;
    else if (a > prev_h_)
      nh_ = 0;
    if (nh_ > 2) {
      why = 1;
      stop = true;
    }
  }
  prev_h_ = a;
  FILE *f = fopen(out("/max.txt").c_str(), "w");
  if (f) {
    fprintf(f, "%d\t%d\t%.5f\t%d\n", iter_, duration(), a, why);
    fclose(f);
  }
  if (stop) {
    do_on_stop();
    exit(0);
  }
}

void HGAPRec::do_on_stop()
{
  save_model();
  gen_ranking_for_users(false);
}

void HGAPRec::save_model()
{
  download_state();
  bool ok = true;
  if (opt_.hier) {
    ok &= st_.beta.save_state(opt_.prefix, ratings_.seq2item());
    ok &= st_.betarate.save_state(opt_.prefix, ratings_.seq2item());
    ok &= st_.theta.save_state(opt_.prefix, ratings_.seq2user());
    ok &= st_.thetarate.save_state(opt_.prefix, ratings_.seq2user());
  } else {
    ok &= st_.beta.save_state(opt_.prefix, ratings_.seq2item());
    ok &= st_.theta.save_state(opt_.prefix, ratings_.seq2user());
  }
  if (opt_.bias) {
    ok &= st_.betabias.save_state(opt_.prefix, ratings_.seq2item());
    ok &= st_.thetabias.save_state(opt_.prefix, ratings_.seq2user());
  }
  if (!ok && logf_) fprintf(logf_, "cannot write model files under %s\n", opt_.prefix.c_str());
}

// per listed user: the items compute_precision skips -- training and validation
// (1729: _ratings.r(n,m) > 0 || is_validation(r))
void HGAPRec::exclusions_of(const std::vector<uint32_t> &users, std::vector<uint64_t> *ptr, std::vector<uint32_t> *idx) const
{
  ptr->assign(users.size() + 1, 0);
  idx->clear();
  for (size_t a = 0; a < users.size(); ++a) {
    const uint32_t u = users[a];
    const std::vector<uint32_t> &tr = ratings_.items_of(u);
    for (size_t j = 0; j < tr.size(); ++j)
      if (ratings_.value_at(u, j) > 0) idx->push_back(tr[j]);
    for (HeldoutMap::const_iterator it = validation_map_.lower_bound(Pair(u, 0)); it != validation_map_.end() && it->first.first == u; ++it)
      idx->push_back(it->first.second);
    (*ptr)[a + 1] = idx->size();
  }
}

void HGAPRec::compute_precision(bool save_ranking_file)
{
  if (iter_ % 100 == 0 && iter_ > 0) save_ranking_file = true;
  if (!save_ranking_file) {
    // 1000 distinct users (or n/2) drawn from the SAME generator as the start state (1715-1721)
    sampled_users_.clear();
    do {
      sampled_users_[rng_.uniform_int(n_)] = true;
    } while (sampled_users_.size() < 1000 && sampled_users_.size() < n_ / 2);
  }
  std::vector<uint32_t> users;
  for (std::map<uint32_t, bool>::const_iterator it = sampled_users_.begin(); it != sampled_users_.end(); ++it) users.push_back(it->first);
  FILE *f = save_ranking_file ? fopen(out("/ranking.tsv").c_str(), "w") : NULL;
  const uint32_t topn = topn_by_user_;
  std::vector<uint64_t> eptr;
  std::vector<uint32_t> eidx, items((size_t)users.size() * topn);
  std::vector<float> scores((size_t)users.size() * topn);
  exclusions_of(users, &eptr, &eidx);
  if (!users.empty() && hpf_topn(ctx_, users.data(), (uint32_t)users.size(), eptr.data(), eidx.data(), topn, items.data(), scores.data()) != 0)
    die("hpf_topn");
  double mhits10 = 0, mhits100 = 0;
  uint32_t total_users = 0;
  for (size_t a = 0; a < users.size(); ++a) {
    const uint32_t u = users[a];
    uint32_t hits10 = 0, hits100 = 0;
    for (uint32_t j = 0; j < topn && j < m_; ++j) {
      const uint32_t it = items[a * topn + j];
      if (it == 0xffffffffu) break;
      const double pred = scores[a * topn + j];
      int v = 0;
      HeldoutMap::const_iterator t = test_map_.find(Pair(u, it));
      if (t != test_map_.end()) {
        v = ratings_.test_hit(t->second) ? 1 : 0;
        if (j < 10) {
          if (v > 0) { hits10++; hits100++; }
        } else if (j < 100) {
          if (v > 0) hits100++;
        }
      }
      if (f && ratings_.r(u, it) == 0) fprintf(f, "%d\t%d\t%.5f\t%d\n", ratings_.user_id(u), ratings_.item_id(it), pred, v);
    }
    mhits10 += (double)hits10 / 10;
    mhits100 += (double)hits100 / 100;
    total_users++;
  }
  if (f) fclose(f);
  fprintf(pf_, "%d\t%.5f\t%.5f\n", total_users, (double)mhits10 / total_users, (double)mhits100 / total_users);
  fflush(pf_);
}

// compute_itemrank (1607-1701): position of every test hit in the user's full descending list
void HGAPRec::compute_itemrank(bool final)
{
  if (iter_ % 100 == 0 && iter_ > 0) final = true;
  if (!final) return;
  FILE *f = fopen(out("/itemrank.tsv").c_str(), "w");
  FILE *itemf = fopen(out("/meanrank.txt").c_str(), "w");
  if (!f || !itemf) {
    printf("cannot open logl file:%s\n", strerror(errno));
    exit(-1);
  }
  std::vector<uint32_t> users;
  for (std::map<uint32_t, bool>::const_iterator it = sampled_users_.begin(); it != sampled_users_.end(); ++it) users.push_back(it->first);
  std::vector<uint64_t> eptr, qptr(users.size() + 1, 0);
  std::vector<uint32_t> eidx, qidx;
  exclusions_of(users, &eptr, &eidx);
  for (size_t a = 0; a < users.size(); ++a) { // the user's test items that count as hits (test_hit)
    const uint32_t u = users[a];
    for (HeldoutMap::const_iterator t = test_map_.lower_bound(Pair(u, 0)); t != test_map_.end() && t->first.first == u; ++t)
      if (ratings_.test_hit(t->second)) qidx.push_back(t->first.second);
    qptr[a + 1] = qidx.size();
  }
  std::vector<uint32_t> rank(qidx.size());
  std::vector<float> pred(qidx.size());
  if (!users.empty() &&
      hpf_item_ranks(ctx_, users.data(), (uint32_t)users.size(), eptr.data(), eidx.data(), qptr.data(), qidx.data(), rank.data(), pred.data()) != 0)
    die("hpf_item_ranks");
  double sum_rank = 0, sum_reciprocal_rank = 0;
  uint32_t total_users = 0;
  for (size_t a = 0; a < users.size(); ++a) {
    const uint32_t u = users[a];
    // items the sorted list counts as ranked: those with no training rating (1662-1663)
    std::vector<uint32_t> tr;
    const std::vector<uint32_t> &it = ratings_.items_of(u);
    for (size_t j = 0; j < it.size(); ++j)
      if (ratings_.value_at(u, j) > 0) tr.push_back(it[j]);
    std::sort(tr.begin(), tr.end());
    const uint32_t nranked = m_ - (uint32_t)(std::unique(tr.begin(), tr.end()) - tr.begin());
    // the reference walks the list top-down: emit the hits by ascending position
    std::vector<std::pair<uint32_t, uint64_t> > order;
    for (uint64_t q = qptr[a]; q < qptr[a + 1]; ++q) order.push_back(std::make_pair(rank[q], q));
    std::sort(order.begin(), order.end());
    double rank_ui = 0, reciprocal_rank_ui = 0;
    uint32_t ntestitems = 0;
    for (size_t o = 0; o < order.size(); ++o) {
      const uint64_t q = order[o].second;
      const uint32_t j = order[o].first;
      ntestitems++;
      fprintf(f, "%d\t%d\t%.5f\t%d\t%d\n", u, qidx[q], pred[q], j, (int)ratings_.item_degree(qidx[q]));
      rank_ui += (j + 1);
      reciprocal_rank_ui += 1 / (j + 1); // integer division, as in the reference (1683)
    }
    if (ntestitems > 0 && nranked > 0) {
      sum_rank += (rank_ui / nranked) / ntestitems;
      sum_reciprocal_rank += reciprocal_rank_ui / ntestitems;
      total_users++;
    }
  }
  fclose(f);
  fprintf(itemf, "%d\t%.5f\t%.5f\n", total_users, (double)sum_rank / total_users, (double)sum_reciprocal_rank / total_users);
  fclose(itemf);
}

bool HGAPRec::load_beta_and_theta()
{
  bool ok = true;
  if (!opt_.hier) {
    ok &= st_.beta.load();
    ok &= st_.theta.load();
  } else {
    ok &= st_.thetarate.load();
    ok &= st_.betarate.load();
    ok &= st_.beta.load();
    ok &= st_.theta.load();
  }
  if (opt_.bias) {
    // GPMatrix::load reads only E[v] (src/gpbase.hh:410-415)
    ok &= load_tsv("betabias.tsv", st_.betabias.expected_v.data(), m_, 1);
    ok &= load_tsv("thetabias.tsv", st_.thetabias.expected_v.data(), n_, 1);
  }
  return ok;
}

void HGAPRec::gen_ranking_for_users(bool load)
{
  if (load) {
    if (!load_beta_and_theta()) {
      fprintf(stderr, "hgaprec: cannot load the model files from the current directory\n");
      exit(-1);
    }
    upload_state();
  }
  sampled_users_.clear();
  if (!ratings_.read_test_users(opt_.dir + "/test_users.tsv", &sampled_users_)) {
    fprintf(stderr, "cannot read %s/test_users.tsv (missing or malformed)\n", opt_.dir.c_str());
    return;
  }
  compute_precision(true);
  compute_itemrank(true);
  if (logf_) fprintf(logf_, "DONE writing ranking.tsv in output directory\n");
}

} // namespace hpfhost
