/* TEST INFRASTRUCTURE ONLY -- see hpf_oracle.h.  Plain-C fp64 restatement of
 * the reference's CAVI hot path.  Every function cites the reference lines it
 * restates (paths relative to /root/reference/).  Nothing here is shipped or
 * measured as the product. */
#include "hpf_oracle.h"
#include "gsl_shim/gsl/gsl_rng.h"    /* own mt19937 (same stream the _ref build uses) */
#include "gsl_shim/gsl/gsl_sf_psi.h" /* own digamma  (same one the _ref build uses)  */
#include "gsl_shim/gsl/gsl_sf.h"     /* gsl_sf_lngamma stand-in (same one the _ref build uses) */

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PRIOR_SHAPE 0.3 /* src/hgaprec.cc:13-20: every (shape, rate) prior is (0.3, 0.3); */
#define PRIOR_RATE  0.3 /* the -a/-b/-c/-d flags never reach HGAPRec         */

double hpf_oracle_digamma(double x) { return gsl_sf_psi(x); }

/* GPBase::make_nonzero, src/gpbase.hh:27-44 */
static inline double floor30(double v) { return v > 0.0 ? v : 1e-30; }

/* GPMatrix/GPMatrixGR/GPArray::compute_expectations for one element,
 * src/gpbase.hh:248-262, 581-600, 912-925 */
static inline void expect1(double shape, double rate, double *ev, double *elogv)
{
  double a = floor30(shape), b = floor30(rate);
  *ev = a / b;
  *elogv = gsl_sf_psi(a) - log(b);
}

/* ------------------------------------------------------------------ init */

/* GPArray::initialize2(v) + compute_expectations, src/gpbase.hh:939-949 */
static void init_gparray2(double *g[4], uint32_t n, double v, gsl_rng *r)
{
  for (uint32_t i = 0; i < n; ++i) {
    g[HPF_O_SHAPE][i] = PRIOR_SHAPE + 0.01 * gsl_rng_uniform(r);
    g[HPF_O_RATE][i] = PRIOR_RATE + v;
  }
  for (uint32_t i = 0; i < n; ++i)
    expect1(g[HPF_O_SHAPE][i], g[HPF_O_RATE][i], &g[HPF_O_EV][i], &g[HPF_O_ELOGV][i]);
}

/* GPMatrix::initialize (src/gpbase.hh:292-308) / GPMatrixGR::initialize
 * (678-688): shapes row-major, then k rate draws.  per_row_rate != 0 copies the
 * k-vector to every row (GPMatrix), otherwise the rate stays a k-vector. */
static void init_shapes(double *g[4], uint32_t rows, uint32_t k, int per_row_rate, gsl_rng *r)
{
  for (size_t e = 0; e < (size_t)rows * k; ++e)
    g[HPF_O_SHAPE][e] = PRIOR_SHAPE + 0.01 * gsl_rng_uniform(r);
  for (uint32_t j = 0; j < k; ++j)
    g[HPF_O_RATE][j] = PRIOR_RATE + 0.1 * gsl_rng_uniform(r);
  if (per_row_rate)
    for (uint32_t i = 1; i < rows; ++i)
      memcpy(g[HPF_O_RATE] + (size_t)i * k, g[HPF_O_RATE], sizeof(double) * k);
}

/* GPMatrix::initialize_exp (src/gpbase.hh:324-340) / GPMatrixGR (706-715):
 * expectations from a FRESH random rate per element; the stored rate is not used. */
static void init_exp(double *g[4], uint32_t rows, uint32_t k, gsl_rng *r)
{
  for (size_t e = 0; e < (size_t)rows * k; ++e) {
    double b = PRIOR_RATE + 0.1 * gsl_rng_uniform(r);
    double a = g[HPF_O_SHAPE][e];
    g[HPF_O_EV][e] = a / b;
    g[HPF_O_ELOGV][e] = gsl_sf_psi(a) - log(b);
  }
}

/* HGAPRec::initialize, src/hgaprec.cc:153-204 */
void hpf_oracle_init(hpf_oracle_state *s, unsigned long seed)
{
  gsl_rng *r = gsl_rng_alloc(gsl_rng_default);
  if (seed) gsl_rng_set(r, seed); /* hgaprec.cc:37-38 */
  const uint32_t n = s->n, m = s->m, k = s->k;
  if (!(s->flags & HPF_O_HIER)) {
    init_shapes(s->beta, m, k, 0, r);  /* _beta.initialize()      */
    init_shapes(s->theta, n, k, 0, r); /* _theta.initialize()     */
    init_exp(s->beta, m, k, r);        /* _beta.initialize_exp()  */
    init_exp(s->theta, n, k, r);       /* _theta.initialize_exp() */
  } else {
    init_gparray2(s->thetarate, n, (double)k, r);
    init_gparray2(s->betarate, m, (double)k, r);
    init_shapes(s->beta, m, k, 1, r);
    init_exp(s->beta, m, k, r);
    init_shapes(s->theta, n, k, 1, r);
    init_exp(s->theta, n, k, r);
  }
  if (s->flags & HPF_O_BIAS) {
    /* GPMatrix n x 1: initialize2(v) (gpbase.hh:310-322) has the same draw
     * pattern as GPArray::initialize2 for a single column */
    init_gparray2(s->thetabias, n, (double)m, r); /* hgaprec.cc:198 */
    init_gparray2(s->betabias, m, (double)n, r);  /* hgaprec.cc:201 */
  }
  gsl_rng_free(r);
}

/* ------------------------------------------------------------- iteration */

/* D1Array<double>::logsum, src/matrix.hh:367-381: streaming log-add-exp */
static inline double logsum_stream(const double *v, uint32_t len)
{
  double r = v[0];
  for (uint32_t i = 1; i < len; ++i) {
    if (v[i] < r)
      r = r + log(1 + exp(v[i] - r));
    else
      r = v[i] + log(1 + exp(r - v[i]));
  }
  return r;
}

/* one user row of the nnz loop: get_phi (hgaprec.cc:206-239), lognormalize
 * (matrix.hh:383-389), scale (399-406), update_shape_next1/3 (gpbase.hh:175-193) */
static void sweep_rows(const hpf_oracle_state *s, const uint64_t *row_ptr,
                       const uint32_t *col_idx, const uint8_t *y,
                       uint32_t u0, uint32_t u1, double *phi,
                       double *t_snext, double *b_snext, double *tb_snext, double *bb_snext)
{
  const uint32_t k = s->k;
  const int bias = (s->flags & HPF_O_BIAS) != 0;
  const uint32_t width = bias ? k + 2 : k;
  const double *tl = s->theta[HPF_O_ELOGV], *bl = s->beta[HPF_O_ELOGV];
  for (uint32_t u = u0; u < u1; ++u) {
    for (uint64_t j = row_ptr[u]; j < row_ptr[u + 1]; ++j) {
      const uint32_t i = col_idx[j];
      const double yv = y ? (double)y[j] : 1.0;
      const double *tu = tl + (size_t)u * k, *bi = bl + (size_t)i * k;
      for (uint32_t q = 0; q < k; ++q) phi[q] = tu[q] + bi[q];
      if (bias) {
        phi[k] = s->thetabias[HPF_O_ELOGV][u];
        phi[k + 1] = s->betabias[HPF_O_ELOGV][i];
      }
      const double lz = logsum_stream(phi, width);
      for (uint32_t q = 0; q < width; ++q) phi[q] = exp(phi[q] - lz);
      if (yv > 1)
        for (uint32_t q = 0; q < width; ++q) phi[q] *= yv;
      double *ts = t_snext + (size_t)u * k, *bs = b_snext + (size_t)i * k;
      for (uint32_t q = 0; q < k; ++q) ts[q] += phi[q];
      for (uint32_t q = 0; q < k; ++q) bs[q] += phi[q];
      if (bias) {
        tb_snext[u] += phi[k];
        bb_snext[i] += phi[k + 1];
      }
    }
  }
}

static void fill(double *p, size_t cnt, double v)
{
  for (size_t e = 0; e < cnt; ++e) p[e] = v;
}

/* GPMatrix::sum_rows (gpbase.hh:264-271): v[k] = sum over rows of Ev */
static void col_totals(const double *ev, uint32_t rows, uint32_t k, double *v)
{
  fill(v, k, 0.0);
  for (uint32_t i = 0; i < rows; ++i)
    for (uint32_t q = 0; q < k; ++q) v[q] += ev[(size_t)i * k + q];
}

/* swap + compute_expectations for a hier GPMatrix whose next-rate is
 * prior_rate[row] + add[k]  (set_prior_rate 163-173, update_rate_next 218-223,
 * swap 240-246, compute_expectations 248-262) */
static void finish_hier(double *g[4], const double *snext, uint32_t rows, uint32_t k,
                        const double *row_prior, const double *add)
{
  for (uint32_t i = 0; i < rows; ++i)
    for (uint32_t q = 0; q < k; ++q) {
      size_t e = (size_t)i * k + q;
      g[HPF_O_SHAPE][e] = snext[e];
      g[HPF_O_RATE][e] = row_prior[i] + add[q];
    }
}

/* same for GPMatrixGR: rate k-vector = prior + add (gpbase.hh:560-564, 573-579) */
static void finish_gr(double *g[4], const double *snext, uint32_t rows, uint32_t k, const double *add)
{
  memcpy(g[HPF_O_SHAPE], snext, sizeof(double) * (size_t)rows * k);
  for (uint32_t q = 0; q < k; ++q) g[HPF_O_RATE][q] = PRIOR_RATE + add[q];
}

static void expectations(double *g[4], uint32_t rows, uint32_t k, int rate_is_vector)
{
  for (uint32_t i = 0; i < rows; ++i)
    for (uint32_t q = 0; q < k; ++q) {
      size_t e = (size_t)i * k + q;
      double rate = rate_is_vector ? g[HPF_O_RATE][q] : g[HPF_O_RATE][e];
      expect1(g[HPF_O_SHAPE][e], rate, &g[HPF_O_EV][e], &g[HPF_O_ELOGV][e]);
    }
}

/* bias GPMatrix (rows x 1): update_rate_next_all(0, v) (gpbase.hh:225-231), swap */
static void finish_bias(double *g[4], const double *snext, uint32_t rows, double v)
{
  for (uint32_t i = 0; i < rows; ++i) {
    g[HPF_O_SHAPE][i] = snext[i];
    g[HPF_O_RATE][i] = PRIOR_RATE + v;
  }
}

/* GPArray thetarate/betarate step, hgaprec.cc:1398-1414; gpbase.hh:877-889, 897-925 */
static void rate_prior_step(double *g[4], const double *ev, uint32_t rows, uint32_t k)
{
  for (uint32_t i = 0; i < rows; ++i) {
    double tot = 0.0; /* sum_cols, gpbase.hh:273-280 */
    for (uint32_t q = 0; q < k; ++q) tot += ev[(size_t)i * k + q];
    g[HPF_O_SHAPE][i] = PRIOR_SHAPE + k * PRIOR_SHAPE; /* hgaprec.cc:1400: k * sprior */
    g[HPF_O_RATE][i] = PRIOR_RATE + tot;
    expect1(g[HPF_O_SHAPE][i], g[HPF_O_RATE][i], &g[HPF_O_EV][i], &g[HPF_O_ELOGV][i]);
  }
}

void hpf_oracle_iterate(hpf_oracle_state *s, const uint64_t *row_ptr,
                        const uint32_t *col_idx, const uint8_t *y,
                        uint32_t niters, int nthreads)
{
  const uint32_t n = s->n, m = s->m, k = s->k;
  const int hier = (s->flags & HPF_O_HIER) != 0;
  const int bias = (s->flags & HPF_O_BIAS) != 0;
  const int jacobi = (s->flags & HPF_O_JACOBI) != 0 && !hier; /* only vb_bias has -novb */
  if (nthreads < 1) nthreads = 1;
#ifndef _OPENMP
  nthreads = 1;
#endif
  const size_t nk = (size_t)n * k, mk = (size_t)m * k;
  double *t_snext = malloc(sizeof(double) * nk);
  double *b_snext = malloc(sizeof(double) * mk);
  double *tb_snext = malloc(sizeof(double) * (n ? n : 1));
  double *bb_snext = malloc(sizeof(double) * (m ? m : 1));
  double *ksum = malloc(sizeof(double) * k), *ksum2 = malloc(sizeof(double) * k);
  /* user ranges of equal nnz for the threaded variant */
  uint32_t *cut = malloc(sizeof(uint32_t) * (nthreads + 1));
  cut[0] = 0;
  for (int t = 1; t <= nthreads; ++t) {
    uint64_t target = row_ptr[n] * (uint64_t)t / nthreads;
    uint32_t u = cut[t - 1];
    while (u < n && row_ptr[u] < target) ++u;
    cut[t] = (t == nthreads) ? n : u;
  }
  double **priv_b = calloc(nthreads, sizeof(double *));
  double **priv_bb = calloc(nthreads, sizeof(double *));
  for (int t = 1; t < nthreads; ++t) {
    priv_b[t] = malloc(sizeof(double) * mk);
    priv_bb[t] = malloc(sizeof(double) * (m ? m : 1));
  }

  for (uint32_t it = 0; it < niters; ++it) {
    /* "next" buffers hold the prior after swap()/set_to_prior() */
    fill(t_snext, nk, PRIOR_SHAPE);
    fill(b_snext, mk, PRIOR_SHAPE);
    fill(tb_snext, n, PRIOR_SHAPE);
    fill(bb_snext, m, PRIOR_SHAPE);

    /* nnz loop: hgaprec.cc:1340-1366 (hier), 928-942 (vb), 1227-1248 (vb_bias) */
    if (nthreads == 1) {
      double *phi = malloc(sizeof(double) * (k + 2));
      sweep_rows(s, row_ptr, col_idx, y, 0, n, phi, t_snext, b_snext, tb_snext, bb_snext);
      free(phi);
    } else {
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
      {
        int t = omp_get_thread_num();
        double *phi = malloc(sizeof(double) * (k + 2));
        double *bs = t == 0 ? b_snext : priv_b[t];
        double *bbs = t == 0 ? bb_snext : priv_bb[t];
        if (t > 0) { fill(bs, mk, 0.0); fill(bbs, m, 0.0); }
        sweep_rows(s, row_ptr, col_idx, y, cut[t], cut[t + 1], phi, t_snext, bs, tb_snext, bbs);
        free(phi);
      }
      for (int t = 1; t < nthreads; ++t) {
        for (size_t e = 0; e < mk; ++e) b_snext[e] += priv_b[t][e];
        for (uint32_t i = 0; i < m; ++i) bb_snext[i] += priv_bb[t][i];
      }
#endif
    }

    if (hier) {
      /* hgaprec.cc:1370-1378: theta rate = E[xi_u] + sum_i E[beta_ik] (old beta) */
      col_totals(s->beta[HPF_O_EV], m, k, ksum);
      finish_hier(s->theta, t_snext, n, k, s->thetarate[HPF_O_EV], ksum);
      expectations(s->theta, n, k, 0);
      /* hgaprec.cc:1380-1386: beta rate = E[eta_i] + sum_u E[theta_uk] (NEW theta) */
      col_totals(s->theta[HPF_O_EV], n, k, ksum);
      finish_hier(s->beta, b_snext, m, k, s->betarate[HPF_O_EV], ksum);
      expectations(s->beta, m, k, 0);
      if (bias) { /* hgaprec.cc:1388-1396 */
        finish_bias(s->thetabias, tb_snext, n, (double)m);
        expectations(s->thetabias, n, 1, 0);
        finish_bias(s->betabias, bb_snext, m, (double)n);
        expectations(s->betabias, m, 1, 0);
      }
      rate_prior_step(s->thetarate, s->theta[HPF_O_EV], n, k); /* 1398-1405 */
      rate_prior_step(s->betarate, s->beta[HPF_O_EV], m, k);   /* 1407-1414 */
    } else if (!jacobi) {
      /* vb(): hgaprec.cc:944-956; vb_bias() Gauss-Seidel branch: 1250-1271 */
      col_totals(s->beta[HPF_O_EV], m, k, ksum);
      finish_gr(s->theta, t_snext, n, k, ksum);
      expectations(s->theta, n, k, 1);
      col_totals(s->theta[HPF_O_EV], n, k, ksum);
      finish_gr(s->beta, b_snext, m, k, ksum);
      expectations(s->beta, m, k, 1);
      if (bias) {
        finish_bias(s->thetabias, tb_snext, n, (double)m);
        expectations(s->thetabias, n, 1, 0);
        finish_bias(s->betabias, bb_snext, m, (double)n);
        expectations(s->betabias, m, 1, 0);
      }
    } else {
      /* vb_bias() -novb branch: hgaprec.cc:1276-1297: both sums from OLD expectations */
      col_totals(s->beta[HPF_O_EV], m, k, ksum);
      col_totals(s->theta[HPF_O_EV], n, k, ksum2);
      finish_gr(s->theta, t_snext, n, k, ksum);
      finish_gr(s->beta, b_snext, m, k, ksum2);
      if (bias) {
        finish_bias(s->thetabias, tb_snext, n, (double)m);
        finish_bias(s->betabias, bb_snext, m, (double)n);
      }
      expectations(s->theta, n, k, 1);
      expectations(s->beta, m, k, 1);
      if (bias) {
        expectations(s->thetabias, n, 1, 0);
        expectations(s->betabias, m, 1, 0);
      }
    }
  }
  for (int t = 1; t < nthreads; ++t) { free(priv_b[t]); free(priv_bb[t]); }
  free(priv_b); free(priv_bb); free(cut);
  free(t_snext); free(b_snext); free(tb_snext); free(bb_snext); free(ksum); free(ksum2);
}

/* ------------------------------------------------------------ evaluation */

/* HGAPRec::log_factorial, hgaprec.cc:1563-1570 */
static double log_factorial(uint32_t v)
{
  double r = log(1);
  for (uint32_t i = 2; i <= v; ++i) r += log(i);
  return r;
}

/* prediction_score[_hier] / the rate inside rating_likelihood[_hier]:
 * hgaprec.cc:1850-1880, 1969-1991, 1503-1560 */
static inline double pair_rate(const hpf_oracle_state *s, uint32_t u, uint32_t i)
{
  const uint32_t k = s->k;
  const double *tu = s->theta[HPF_O_EV] + (size_t)u * k;
  const double *bi = s->beta[HPF_O_EV] + (size_t)i * k;
  double r = 0.0;
  for (uint32_t q = 0; q < k; ++q) r += tu[q] * bi[q];
  if (s->flags & HPF_O_BIAS) r += s->thetabias[HPF_O_EV][u] + s->betabias[HPF_O_EV][i];
  return r;
}

double hpf_oracle_heldout(const hpf_oracle_state *s, const uint32_t *u,
                          const uint32_t *i, const uint8_t *y, uint64_t npairs)
{
  double tot = 0.0;
  for (uint64_t p = 0; p < npairs; ++p) {
    double r = pair_rate(s, u[p], i[p]);
    if (r < 1e-30) r = 1e-30;
    if (s->flags & HPF_O_BINARY)
      tot += y[p] == 0 ? -r : log(1 - exp(-r));
    else
      tot += y[p] * log(r) - r - log_factorial(y[p]);
  }
  return tot;
}

typedef struct { uint32_t item; double score; } scored;

static int by_score_desc(const void *a, const void *b)
{
  const scored *x = a, *z = b;
  if (x->score != z->score) return x->score < z->score ? 1 : -1;
  return x->item < z->item ? -1 : (x->item > z->item);
}

void hpf_oracle_topn(const hpf_oracle_state *s, const uint32_t *users, uint32_t nu,
                     const uint64_t *excl_ptr, const uint32_t *excl_idx,
                     uint32_t topn, uint32_t *items_out, double *scores_out)
{
  const uint32_t m = s->m;
  scored *lst = malloc(sizeof(scored) * m);
  for (uint32_t a = 0; a < nu; ++a) {
    const uint32_t u = users[a];
    for (uint32_t i = 0; i < m; ++i) {
      lst[i].item = i;
      lst[i].score = pair_rate(s, u, i);
    }
    /* training and validation items keep their slot with score 0 (hgaprec.cc:1729-1735) */
    for (uint64_t e = excl_ptr[a]; e < excl_ptr[a + 1]; ++e) lst[excl_idx[e]].score = 0.0;
    qsort(lst, m, sizeof(scored), by_score_desc);
    for (uint32_t j = 0; j < topn; ++j) {
      items_out[(size_t)a * topn + j] = j < m ? lst[j].item : 0xffffffffu;
      scores_out[(size_t)a * topn + j] = j < m ? lst[j].score : 0.0;
    }
  }
  free(lst);
}

/* ------------------------------------------------------------------ ELBO */

/* GPMatrix::compute_elbo_term_helper (src/gpbase.hh:360-387) for a rows x k set whose rate is a
 * matrix (rate_is_vector == 0: GPMatrix) or a k-vector (GPMatrixGR::compute_elbo_term_helper,
 * 717-741).  row_prior / row_log_prior: _hier_rprior / _hier_log_rprior (only a GPMatrix after
 * set_prior_rate, i.e. htheta / hbeta); NULL selects the constant (_sprior, _rprior) branch. */
static double elbo_matrix(double *const g[4], uint32_t rows, uint32_t k, int rate_is_vector,
                          const double *row_prior, const double *row_log_prior)
{
  double s = 0.0;
  for (uint32_t n = 0; n < rows; ++n) {
    const double *ev = g[HPF_O_EV] + (size_t)n * k, *el = g[HPF_O_ELOGV] + (size_t)n * k;
    for (uint32_t q = 0; q < k; ++q) {
      if (row_prior) {
        s += PRIOR_SHAPE * row_log_prior[n] + (PRIOR_SHAPE - 1) * el[q];
        s -= row_prior[n] * ev[q] + gsl_sf_lngamma(PRIOR_SHAPE);
      } else {
        s += PRIOR_SHAPE * log(PRIOR_RATE) + (PRIOR_SHAPE - 1) * el[q];
        s -= PRIOR_RATE * ev[q] + gsl_sf_lngamma(PRIOR_SHAPE);
      }
    }
    for (uint32_t q = 0; q < k; ++q) {
      const double a = floor30(g[HPF_O_SHAPE][(size_t)n * k + q]);
      const double b = floor30(rate_is_vector ? g[HPF_O_RATE][q] : g[HPF_O_RATE][(size_t)n * k + q]);
      s -= a * log(b) + (a - 1) * el[q];
      s += b * ev[q] + gsl_sf_lngamma(a);
    }
  }
  return s;
}

/* GPArray::compute_elbo_term_helper, src/gpbase.hh:951-969 */
static double elbo_array(double *const g[4], uint32_t rows)
{
  double s = 0.0;
  for (uint32_t n = 0; n < rows; ++n) {
    const double a = floor30(g[HPF_O_SHAPE][n]), b = floor30(g[HPF_O_RATE][n]);
    const double ev = g[HPF_O_EV][n], el = g[HPF_O_ELOGV][n];
    s += PRIOR_SHAPE * log(PRIOR_RATE) + (PRIOR_SHAPE - 1) * el;
    s -= PRIOR_RATE * ev + gsl_sf_lngamma(PRIOR_SHAPE);
    s -= a * log(b) + (a - 1) * el;
    s += b * ev + gsl_sf_lngamma(a);
  }
  return s;
}

/* HGAPRec::logl, src/hgaprec.cc:2160-2255, as coded: phi is scaled by y BEFORE the entropy-like
 * term (2214-2220), so a rating y > 1 enters as y * (y phi_k) * (Elog - log(y phi_k)). */
double hpf_oracle_elbo(const hpf_oracle_state *s, const uint64_t *row_ptr, const uint32_t *col_idx,
                       const uint8_t *y, const double *xi_ev, const double *xi_elog,
                       const double *eta_ev, const double *eta_elog)
{
  const uint32_t n = s->n, m = s->m, k = s->k;
  const int hier = (s->flags & HPF_O_HIER) != 0;
  const int bias = (s->flags & HPF_O_BIAS) != 0;
  const uint32_t width = bias ? k + 2 : k;
  double *phi = malloc(sizeof(double) * (k + 2));
  double tot = 0.0;
  for (uint32_t u = 0; u < n; ++u) {
    for (uint64_t j = row_ptr[u]; j < row_ptr[u + 1]; ++j) {
      const uint32_t i = col_idx[j];
      const double yv = y ? (double)y[j] : 1.0;
      const double *tl = s->theta[HPF_O_ELOGV] + (size_t)u * k, *bl = s->beta[HPF_O_ELOGV] + (size_t)i * k;
      const double *te = s->theta[HPF_O_EV] + (size_t)u * k, *be = s->beta[HPF_O_EV] + (size_t)i * k;
      for (uint32_t q = 0; q < k; ++q) phi[q] = tl[q] + bl[q];
      if (bias) {
        phi[k] = s->thetabias[HPF_O_ELOGV][u];
        phi[k + 1] = s->betabias[HPF_O_ELOGV][i];
      }
      const double lz = logsum_stream(phi, width);
      for (uint32_t q = 0; q < width; ++q) phi[q] = exp(phi[q] - lz);
      if (yv > 1)
        for (uint32_t q = 0; q < width; ++q) phi[q] *= yv;
      double v = 0.0;
      for (uint32_t q = 0; q < k; ++q) v += yv * phi[q] * (tl[q] + bl[q] - log(phi[q]));
      tot += v;
      if (bias) {
        tot += yv * phi[k] * (s->thetabias[HPF_O_ELOGV][u] - log(phi[k]));
        tot += yv * phi[k + 1] * (s->betabias[HPF_O_ELOGV][i] - log(phi[k + 1]));
      }
      for (uint32_t q = 0; q < k; ++q) tot -= te[q] * be[q];
      if (bias) {
        tot -= s->thetabias[HPF_O_EV][u];
        tot -= s->betabias[HPF_O_EV][i];
      }
    }
  }
  free(phi);
  if (!hier) {
    tot += elbo_matrix(s->theta, n, k, 1, NULL, NULL);
    tot += elbo_matrix(s->beta, m, k, 1, NULL, NULL);
  } else {
    tot += elbo_matrix(s->theta, n, k, 0, xi_ev, xi_elog);
    tot += elbo_matrix(s->beta, m, k, 0, eta_ev, eta_elog);
    tot += elbo_array(s->thetarate, n);
    tot += elbo_array(s->betarate, m);
  }
  if (bias) { /* n x 1 / m x 1 GPMatrix that never saw set_prior_rate: constant-prior branch */
    tot += elbo_matrix(s->thetabias, n, 1, 0, NULL, NULL);
    tot += elbo_matrix(s->betabias, m, 1, 0, NULL, NULL);
  }
  return tot;
}
