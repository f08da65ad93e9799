// gpstate.hh -- host mirrors of the reference's Gamma variational-parameter
// containers, kept only for what stays on the host: the start state, the
// device <-> host transfer buffers and the on-disk TSV model format.
//
//   GammaMatrix  <->  GPMatrix   (rate rows x k, -hier)   src/gpbase.hh:54-147
//                     GPMatrixGR (rate is a k-vector)     src/gpbase.hh:441-519
//   GammaArray   <->  GPArray                             src/gpbase.hh:789-855
//
// Same field names as the reference (shape_curr / rate_curr / expected_v /
// expected_logv); storage is contiguous row-major fp64, which is what
// hpf_set_state / hpf_get_state exchange.  The update arithmetic
// (update_shape_next*, update_rate_next*, swap, compute_expectations, sum_rows,
// sum_cols) is NOT here: it runs on the GPU behind hpf_iterate.
#ifndef HPF_HOST_GPSTATE_HH
#define HPF_HOST_GPSTATE_HH
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "mt19937.hh"

namespace hpfhost {

// "seq \t id \t v0 \t v1 ... \n" with %.8f values: D2Array<double>::save /
// D1Array<double>::save (src/matrix.hh:1140-1166, 725-744).  ids maps a row
// number to the external id; rows past its end print the row number.
inline bool save_tsv(const std::string &path, const double *v, size_t rows, size_t cols, const std::vector<uint32_t> &ids)
{
  FILE *f = fopen(path.c_str(), "w");
  if (!f) return false;
  std::vector<char> buf(1 << 20);
  setvbuf(f, buf.data(), _IOFBF, buf.size());
  for (size_t i = 0; i < rows; ++i) {
    fprintf(f, "%d\t%d\t", (int)i, (int)(i < ids.size() ? ids[i] : (uint32_t)i));
    for (size_t k = 0; k < cols; ++k) fprintf(f, k + 1 == cols ? "%.8f\n" : "%.8f\t", v[i * cols + k]);
  }
  fclose(f);
  return true;
}

// reads the same format back, skipping the two leading columns
// (D2Array<double>::load / D1Array<double>::load, src/matrix.hh:1198-1266, 767-803)
inline bool load_tsv(const std::string &path, double *v, size_t rows, size_t cols)
{
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return false;
  char *line = NULL;
  size_t cap = 0;
  size_t i = 0;
  while (i < rows && getline(&line, &cap, f) > 0) {
    char *p = line;
    size_t col = 0;
    for (;;) {
      char *q = NULL;
      const double d = strtod(p, &q);
      if (q == p) break;
      p = q;
      if (col >= 2 && col - 2 < cols) v[i * cols + (col - 2)] = d;
      col++;
    }
    i++;
  }
  free(line);
  fclose(f);
  return true;
}

struct GammaMatrix {
  std::string name;
  uint32_t rows, k;
  bool per_row_rate; // GPMatrix (true) or GPMatrixGR (false)
  double sprior, rprior;
  std::vector<double> shape_curr, rate_curr, expected_v, expected_logv;

  GammaMatrix(const std::string &nm, double a, double b, uint32_t r, uint32_t kk, bool hier)
      : name(nm), rows(r), k(kk), per_row_rate(hier), sprior(a), rprior(b), shape_curr((size_t)r * kk, a),
        rate_curr(hier ? (size_t)r * kk : kk, b), expected_v((size_t)r * kk, 0.0), expected_logv((size_t)r * kk, 0.0) {}

  // GPMatrix::initialize / GPMatrixGR::initialize (src/gpbase.hh:292-308, 655-663):
  // shapes prior + 0.01 U row-major, then k rate draws prior + 0.1 U
  void initialize(Mt19937 &r)
  {
    for (size_t e = 0; e < shape_curr.size(); ++e) shape_curr[e] = sprior + 0.01 * r.uniform();
    std::vector<double> b(k);
    for (uint32_t q = 0; q < k; ++q) b[q] = rprior + 0.1 * r.uniform();
    if (per_row_rate)
      for (uint32_t i = 0; i < rows; ++i) memcpy(&rate_curr[(size_t)i * k], b.data(), sizeof(double) * k);
    else
      rate_curr = b;
  }
  // initialize2(v) (src/gpbase.hh:310-322, 665-676): shapes random, rate = prior + v
  void initialize2(Mt19937 &r, double v)
  {
    for (size_t e = 0; e < shape_curr.size(); ++e) shape_curr[e] = sprior + 0.01 * r.uniform();
    for (size_t e = 0; e < rate_curr.size(); ++e) rate_curr[e] = rprior + v;
  }
  // initialize_exp (src/gpbase.hh:324-340, 694-708): expectations from a FRESH random
  // rate per element -- not from rate_curr
  void initialize_exp(Mt19937 &r)
  {
    for (size_t e = 0; e < shape_curr.size(); ++e) {
      const double b = rprior + 0.1 * r.uniform();
      expected_v[e] = shape_curr[e] / b;
      expected_logv[e] = digamma(shape_curr[e]) - log(b);
    }
  }
  // compute_expectations (src/gpbase.hh:248-262, 581-600) -- host copy used only for
  // the bias start state and for models reloaded from shape + rate files
  void compute_expectations()
  {
    for (uint32_t i = 0; i < rows; ++i)
      for (uint32_t q = 0; q < k; ++q) {
        const size_t e = (size_t)i * k + q;
        double a = shape_curr[e], b = per_row_rate ? rate_curr[e] : rate_curr[q];
        if (!(a > 0)) a = 1e-30;
        if (!(b > 0)) b = 1e-30;
        expected_v[e] = a / b;
        expected_logv[e] = digamma(a) - log(b);
      }
  }
  // save_state (src/gpbase.hh:389-398, 743-752); the GR rate file has k rows whose
  // id column runs through the entity id map, as in the reference
  bool save_state(const std::string &dir, const std::vector<uint32_t> &ids) const
  {
    bool ok = save_tsv(dir + "/" + name + "_shape.tsv", shape_curr.data(), rows, k, ids);
    if (per_row_rate) ok &= save_tsv(dir + "/" + name + "_rate.tsv", rate_curr.data(), rows, k, ids);
    else ok &= save_tsv(dir + "/" + name + "_rate.tsv", rate_curr.data(), k, 1, ids);
    return ok & save_tsv(dir + "/" + name + ".tsv", expected_v.data(), rows, k, ids);
  }
  // load (src/gpbase.hh:410-415 reads only E[v]; 754-764 reads shape + rate and
  // recomputes the expectations); files are looked up in the current directory
  bool load()
  {
    if (per_row_rate) return load_tsv(name + ".tsv", expected_v.data(), rows, k);
    if (!load_tsv(name + "_shape.tsv", shape_curr.data(), rows, k) || !load_tsv(name + "_rate.tsv", rate_curr.data(), k, 1))
      return false;
    compute_expectations();
    return true;
  }
};

struct GammaArray {
  std::string name;
  uint32_t n;
  double sprior, rprior;
  std::vector<double> shape_curr, rate_curr, expected_v, expected_logv;

  GammaArray(const std::string &nm, double a, double b, uint32_t nn)
      : name(nm), n(nn), sprior(a), rprior(b), shape_curr(nn, a), rate_curr(nn, b), expected_v(nn, 0.0), expected_logv(nn, 0.0) {}

  // GPArray::initialize2 (src/gpbase.hh:939-949) + compute_expectations (912-925)
  void initialize2(Mt19937 &r, double v)
  {
    for (uint32_t i = 0; i < n; ++i) {
      shape_curr[i] = sprior + 0.01 * r.uniform();
      rate_curr[i] = rprior + v;
    }
  }
  void compute_expectations()
  {
    for (uint32_t i = 0; i < n; ++i) {
      double a = shape_curr[i], b = rate_curr[i];
      if (!(a > 0)) a = 1e-30;
      if (!(b > 0)) b = 1e-30;
      expected_v[i] = a / b;
      expected_logv[i] = digamma(a) - log(b);
    }
  }
  bool save_state(const std::string &dir, const std::vector<uint32_t> &ids) const
  {
    return save_tsv(dir + "/" + name + "_shape.tsv", shape_curr.data(), n, 1, ids) &
           save_tsv(dir + "/" + name + "_rate.tsv", rate_curr.data(), n, 1, ids) &
           save_tsv(dir + "/" + name + ".tsv", expected_v.data(), n, 1, ids);
  }
  bool load()
  {
    if (!load_tsv(name + "_shape.tsv", shape_curr.data(), n, 1) || !load_tsv(name + "_rate.tsv", rate_curr.data(), n, 1)) return false;
    compute_expectations();
    return true;
  }
};

} // namespace hpfhost
#endif
