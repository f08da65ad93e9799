// hostcheck.cc -- dumps what the host side hands to the engine (the CSR, the id
// maps, the held-out maps and the start state) without touching a GPU, so the
// CPU test-suite can compare the host logic with the reference's own dumps.
//   hgaprec_hostcheck -dir D -n N -m M -k K [-hier] [-bias] [-binary-data]
//                     [-rating-threshold V] [-seed S] [-csr-cache] -out FILE
// FILE is a flat tagged container ("HPFDUMP1": name, dtype, dims, data per record)
// that the test-suite reads back.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "hgaprec.hh"

using namespace hpfhost;

static void put(FILE *f, const char *name, uint32_t dtype, const void *data, size_t itemsize, const std::vector<uint64_t> &dims)
{
  const uint32_t nl = (uint32_t)strlen(name), nd = (uint32_t)dims.size();
  fwrite(&nl, 4, 1, f);
  fwrite(name, 1, nl, f);
  fwrite(&dtype, 4, 1, f);
  fwrite(&nd, 4, 1, f);
  size_t cnt = 1;
  for (size_t i = 0; i < dims.size(); ++i) { fwrite(&dims[i], 8, 1, f); cnt *= dims[i]; }
  if (cnt) fwrite(data, itemsize, cnt, f);
}
static void put_f64(FILE *f, const std::string &name, const std::vector<double> &v, uint64_t rows, uint64_t cols)
{
  std::vector<uint64_t> d;
  d.push_back(rows);
  if (cols) d.push_back(cols);
  put(f, name.c_str(), 0, v.data(), 8, d);
}
template <class G> static void put_group(FILE *f, const std::string &nm, const G &g, uint64_t rows, uint64_t cols, uint64_t rate_rows)
{
  put_f64(f, nm + ".shape", g.shape_curr, rows, cols);
  put_f64(f, nm + ".rate", g.rate_curr, rate_rows, rate_rows == rows ? cols : 0);
  put_f64(f, nm + ".Ev", g.expected_v, rows, cols);
  put_f64(f, nm + ".Elogv", g.expected_logv, rows, cols);
}
static void put_map(FILE *f, const char *nm, const HeldoutMap &mp)
{
  std::vector<uint32_t> u, i;
  std::vector<uint8_t> y;
  for (HeldoutMap::const_iterator it = mp.begin(); it != mp.end(); ++it) {
    u.push_back(it->first.first); i.push_back(it->first.second); y.push_back(it->second);
  }
  std::vector<uint64_t> d(1, u.size());
  put(f, (std::string(nm) + ".u").c_str(), 1, u.data(), 4, d);
  put(f, (std::string(nm) + ".i").c_str(), 1, i.data(), 4, d);
  put(f, (std::string(nm) + ".y").c_str(), 2, y.data(), 1, d);
}

int main(int argc, char **argv)
{
  Options o;
  std::string outp;
  for (int i = 1; i < argc; ++i) {
    const char *a = argv[i];
    #define NEXT (i + 1 < argc ? argv[++i] : "")
    if (!strcmp(a, "-dir")) o.dir = NEXT;
    else if (!strcmp(a, "-n")) o.n = atoi(NEXT);
    else if (!strcmp(a, "-m")) o.m = atoi(NEXT);
    else if (!strcmp(a, "-k")) o.k = atoi(NEXT);
    else if (!strcmp(a, "-seed")) o.seed = atof(NEXT);
    else if (!strcmp(a, "-rating-threshold")) o.rating_threshold = atoi(NEXT);
    else if (!strcmp(a, "-hier")) o.hier = true;
    else if (!strcmp(a, "-bias")) o.bias = true;
    else if (!strcmp(a, "-binary-data")) o.binary_data = true;
    else if (!strcmp(a, "-novb")) o.vb = false;
    else if (!strcmp(a, "-csr-cache")) o.csr_cache = true;
    else if (!strcmp(a, "-out")) outp = NEXT;
    else { fprintf(stderr, "unknown option %s\n", a); return 2; }
    #undef NEXT
  }
  Ratings ratings(o.n, o.m, o.binary_data, o.rating_threshold);
  std::string err;
  if (!ratings.read_train(o.dir, &err, o.csr_cache)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  if (o.csr_cache) printf("csr cache: %s\n", ratings.cache_note().c_str());
  HeldoutMap val, tst;
  if (!ratings.read_heldout(o.dir + "/validation.tsv", &val, &err) || !ratings.read_heldout(o.dir + "/test.tsv", &tst, &err)) {
    fprintf(stderr, "%s\n", err.c_str());
    return 1;
  }
  const uint32_t n = ratings.n(), m = ratings.m(), k = o.k;
  Mt19937 rng(0);
  if (o.seed) rng.set((unsigned long)o.seed);
  ModelState st(n, m, k, o.hier);
  st.initialize(rng, n, m, k, o.hier, o.bias);
  std::vector<uint64_t> rp;
  std::vector<uint32_t> ci;
  std::vector<uint8_t> y;
  ratings.to_csr(&rp, &ci, &y);

  FILE *f = fopen(outp.c_str(), "wb");
  if (!f) { fprintf(stderr, "cannot write %s\n", outp.c_str()); return 1; }
  fwrite("HPFDUMP1", 1, 8, f);
  std::vector<uint64_t> d1(1);
  d1[0] = rp.size(); put(f, "csr.row_ptr", 3, rp.data(), 8, d1);
  d1[0] = ci.size(); put(f, "csr.col_idx", 1, ci.data(), 4, d1);
  d1[0] = y.size(); put(f, "csr.y", 2, y.data(), 1, d1);
  d1[0] = n; put(f, "seq2user", 1, ratings.seq2user().data(), 4, d1);
  d1[0] = m; put(f, "seq2movie", 1, ratings.seq2item().data(), 4, d1);
  {
    std::vector<uint32_t> deg(m);
    std::vector<uint64_t> tot(m);
    for (uint32_t i = 0; i < m; ++i) { deg[i] = ratings.item_degree(i); tot[i] = ratings.item_total(i); }
    d1[0] = m; put(f, "item_degree", 1, deg.data(), 4, d1);
    put(f, "item_total", 3, tot.data(), 8, d1);
  }
  put_map(f, "validation", val);
  put_map(f, "test", tst);
  const std::string tn = o.hier ? "htheta" : "theta", bn = o.hier ? "hbeta" : "beta";
  put_group(f, tn, st.theta, n, k, o.hier ? n : k);
  put_group(f, bn, st.beta, m, k, o.hier ? m : k);
  if (o.hier) {
    put_group(f, "thetarate", st.thetarate, n, 0, n);
    put_group(f, "betarate", st.betarate, m, 0, m);
  }
  if (o.bias) {
    put_group(f, "thetabias", st.thetabias, n, 0, n);
    put_group(f, "betabias", st.betabias, m, 0, m);
  }
  // next draw of the generator: pins the number of draws initialize() consumed
  const double nxt = rng.uniform();
  std::vector<uint64_t> one(1, 1);
  put(f, "rng.next_uniform", 0, &nxt, 8, one);
  fclose(f);
  return 0;
}
