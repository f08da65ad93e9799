#include "ratings.hh"

#include <errno.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

namespace hpfhost {

// ---- binary cache of the parsed training matrix --------------------------------
namespace {

struct CacheHeader {
  char magic[8];                 // "HPFCSR01"
  uint32_t max_users, max_items; // the -n / -m caps the parse ran under
  uint32_t binary, threshold;    // -binary-data, -rating-threshold
  uint32_t n, m;                 // users / items that received a sequence number
  uint64_t nratings;             // lines kept == CSR entries
  uint64_t tsv_size;             // train.tsv the cache was made from
  int64_t tsv_mtime_s, tsv_mtime_ns;
  uint64_t checksum;             // over the payload that follows the header
};

// 64-bit FNV-1a over 8-byte words (tail bytes one by one)
uint64_t mix(uint64_t h, const void *data, size_t bytes)
{
  const unsigned char *p = (const unsigned char *)data;
  size_t i = 0;
  for (; i + 8 <= bytes; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    h = (h ^ w) * 1099511628211ull;
  }
  for (; i < bytes; ++i) h = (h ^ p[i]) * 1099511628211ull;
  return h;
}
const uint64_t kMixSeed = 14695981039346656037ull;

bool stat_tsv(const std::string &tsv, CacheHeader *h)
{
  struct stat sb;
  if (stat(tsv.c_str(), &sb) != 0) return false;
  h->tsv_size = (uint64_t)sb.st_size;
  h->tsv_mtime_s = (int64_t)sb.st_mtim.tv_sec;
  h->tsv_mtime_ns = (int64_t)sb.st_mtim.tv_nsec;
  return true;
}

template <class T> bool read_vec(FILE *f, std::vector<T> *v, size_t count, uint64_t *h)
{
  v->resize(count);
  if (count && fread(v->data(), sizeof(T), count, f) != count) return false;
  *h = mix(*h, v->data(), count * sizeof(T));
  return true;
}
template <class T> bool write_vec(FILE *f, const std::vector<T> &v, uint64_t *h)
{
  *h = mix(*h, v.data(), v.size() * sizeof(T));
  return v.empty() || fwrite(v.data(), sizeof(T), v.size(), f) == v.size();
}

} // namespace

bool Ratings::load_cache(const std::string &tsv, const std::string &cache)
{
  FILE *f = fopen(cache.c_str(), "rb");
  if (!f) return false;
  CacheHeader h, now;
  bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, "HPFCSR01", 8) == 0 && stat_tsv(tsv, &now) &&
            h.tsv_size == now.tsv_size && h.tsv_mtime_s == now.tsv_mtime_s && h.tsv_mtime_ns == now.tsv_mtime_ns &&
            h.max_users == max_users_ && h.max_items == max_items_ && h.binary == (binary_ ? 1u : 0u) &&
            h.threshold == threshold_ && h.n <= max_users_ && h.m <= max_items_;
  if (ok) { // the header must account for every byte of the file before anything is sized from it
    struct stat sb;
    const uint64_t want = sizeof h + 4ull * h.n + 4ull * h.m + 8ull * ((uint64_t)h.n + 1) + 5ull * h.nratings;
    ok = fstat(fileno(f), &sb) == 0 && (uint64_t)sb.st_size == want;
  }
  std::vector<uint32_t> s2u, s2i, col;
  std::vector<uint64_t> rp;
  std::vector<uint8_t> val;
  uint64_t sum = kMixSeed;
  ok = ok && read_vec(f, &s2u, h.n, &sum) && read_vec(f, &s2i, h.m, &sum) && read_vec(f, &rp, (size_t)h.n + 1, &sum) &&
       rp[0] == 0 && rp[h.n] == h.nratings && read_vec(f, &col, h.nratings, &sum) && read_vec(f, &val, h.nratings, &sum) &&
       sum == h.checksum && fgetc(f) == EOF;
  fclose(f);
  if (!ok) return false;
  for (uint32_t u = 0; u < h.n; ++u)
    if (rp[u + 1] < rp[u]) return false;
  for (uint64_t j = 0; j < h.nratings; ++j)
    if (col[j] >= h.m) return false;
  seq2user_.swap(s2u);
  seq2item_.swap(s2i);
  user2seq_.clear(); item2seq_.clear();
  user2seq_.reserve(h.n); item2seq_.reserve(h.m);
  for (uint32_t u = 0; u < h.n; ++u) user2seq_[seq2user_[u]] = u;
  for (uint32_t i = 0; i < h.m; ++i) item2seq_[seq2item_[i]] = i;
  items_.assign(h.n, std::vector<uint32_t>());
  vals_.assign(h.n, std::vector<uint8_t>());
  for (uint32_t u = 0; u < h.n; ++u) {
    items_[u].assign(col.begin() + rp[u], col.begin() + rp[u + 1]);
    vals_[u].assign(val.begin() + rp[u], val.begin() + rp[u + 1]);
  }
  nratings_ = h.nratings;
  finalize(); // values are already fixed up; this rebuilds the per-item degree and total
  return true;
}

void Ratings::save_cache(const std::string &tsv, const std::string &cache)
{
  CacheHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, "HPFCSR01", 8);
  if (!stat_tsv(tsv, &h)) { cache_note_ = "not written: cannot stat " + tsv; return; }
  h.max_users = max_users_; h.max_items = max_items_;
  h.binary = binary_ ? 1u : 0u; h.threshold = threshold_;
  h.n = n(); h.m = m(); h.nratings = nratings_;
  std::vector<uint64_t> rp(n() + 1, 0);
  std::vector<uint32_t> col;
  std::vector<uint8_t> val;
  col.reserve(nratings_); val.reserve(nratings_);
  for (uint32_t u = 0; u < n(); ++u) {
    col.insert(col.end(), items_[u].begin(), items_[u].end());
    val.insert(val.end(), vals_[u].begin(), vals_[u].end());
    rp[u + 1] = col.size();
  }
  char tmp[64];
  snprintf(tmp, sizeof tmp, ".tmp.%ld", (long)getpid());
  const std::string part = cache + tmp; // written aside, renamed when complete: a reader never sees half a cache
  FILE *f = fopen(part.c_str(), "wb");
  if (!f) { cache_note_ = "not written: cannot create " + part + ": " + strerror(errno); return; }
  uint64_t sum = kMixSeed;
  bool ok = fwrite(&h, sizeof h, 1, f) == 1 && write_vec(f, seq2user_, &sum) && write_vec(f, seq2item_, &sum) &&
            write_vec(f, rp, &sum) && write_vec(f, col, &sum) && write_vec(f, val, &sum);
  h.checksum = sum;
  ok = ok && fseek(f, 0, SEEK_SET) == 0 && fwrite(&h, sizeof h, 1, f) == 1;
  ok = (fclose(f) == 0) && ok;
  if (ok && rename(part.c_str(), cache.c_str()) == 0) cache_note_ = "written " + cache;
  else {
    cache_note_ = "not written: " + std::string(strerror(errno));
    remove(part.c_str());
  }
}

// one unsigned decimal (strtoul-like: optional sign, wraps modulo 2^32); false at end of input
bool TripleReader::number(uint32_t *out)
{
  int ch = get();
  while (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') ch = get();
  if (ch < 0) return false;
  bool neg = false;
  if (ch == '+' || ch == '-') {
    neg = ch == '-';
    ch = get();
  }
  if (ch < '0' || ch > '9') { // not a number: the caller must not mistake this for the end of the input
    bad_ = true;
    bad_at_ = base_ + pos_ - (ch < 0 ? 0 : 1);
    return false;
  }
  uint32_t v = 0;
  while (ch >= '0' && ch <= '9') {
    v = v * 10u + (uint32_t)(ch - '0');
    ch = get();
  }
  if (!(ch < 0 || ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r')) { // "2.5", "12abc": not a number either
    bad_ = true;
    bad_at_ = base_ + pos_ - 1;
    return false;
  }
  *out = neg ? 0u - v : v;
  return true;
}

bool Ratings::read_train(const std::string &dir, std::string *err, bool use_cache)
{
  const std::string path = dir + "/train.tsv";
  const std::string cache = path + ".hpfcsr";
  cache_note_.clear();
  if (use_cache && load_cache(path, cache)) {
    cache_note_ = "loaded " + cache;
    return true;
  }
  FILE *f = fopen(path.c_str(), "r");
  if (!f) {
    *err = "cannot open file " + path + ": " + strerror(errno);
    return false;
  }
  TripleReader rd(f);
  uint32_t uid, iid, rating;
  uint32_t last_uid = 0, last_u = 0;
  bool have_last = false; // lines of one user usually come together: skip the map lookup
  while (rd.next(&uid, &iid, &rating)) {
    uint32_t u = 0, i = 0;
    bool new_user = false, new_item = false;
    if (have_last && uid == last_uid) u = last_u;
    else {
      std::unordered_map<uint32_t, uint32_t>::iterator ut = user2seq_.find(uid);
      if (ut == user2seq_.end()) new_user = true; else u = ut->second;
    }
    std::unordered_map<uint32_t, uint32_t>::iterator it = item2seq_.find(iid);
    if (it == item2seq_.end()) new_item = true; else i = it->second;
    if ((new_user && seq2user_.size() >= max_users_) || (new_item && seq2item_.size() >= max_items_)) continue;
    if (rating_class(rating) == 0) continue;
    if (new_user) {
      u = (uint32_t)seq2user_.size();
      user2seq_[uid] = u;
      seq2user_.push_back(uid);
      items_.push_back(std::vector<uint32_t>());
      vals_.push_back(std::vector<uint8_t>());
    }
    if (new_item) {
      i = (uint32_t)seq2item_.size();
      item2seq_[iid] = i;
      seq2item_.push_back(iid);
    }
    last_uid = uid; last_u = u; have_last = true;
    nratings_++;
    items_[u].push_back(i);
    vals_[u].push_back(binary_ ? (uint8_t)1 : (uint8_t)rating);
  }
  fclose(f);
  if (rd.bad()) { // never train on (or cache) a silently truncated matrix
    *err = rd.complaint(path);
    return false;
  }
  finalize();
  if (use_cache) save_cache(path, cache);
  return true;
}

void Ratings::finalize()
{
  const uint32_t M = m();
  item_degree_.assign(M, 0);
  item_total_.assign(M, 0);
  std::vector<uint32_t> last_pos(M, 0xffffffffu); // scratch: last position of an item inside the current user
  for (uint32_t u = 0; u < n(); ++u) {
    std::vector<uint32_t> &v = items_[u];
    std::vector<uint8_t> &w = vals_[u];
    bool dup = false;
    for (size_t j = 0; j < v.size(); ++j) {
      if (last_pos[v[j]] != 0xffffffffu) dup = true;
      last_pos[v[j]] = (uint32_t)j;
    }
    if (dup) // every occurrence reads the value of the last line (the reference's map was overwritten)
      for (size_t j = 0; j < v.size(); ++j) w[j] = w[last_pos[v[j]]];
    for (size_t j = 0; j < v.size(); ++j) {
      last_pos[v[j]] = 0xffffffffu;
      item_degree_[v[j]]++;
      item_total_[v[j]] += w[j];
    }
  }
}

bool Ratings::read_heldout(const std::string &path, HeldoutMap *out, std::string *err) const
{
  FILE *f = fopen(path.c_str(), "r");
  if (!f) {
    *err = "cannot open file " + path + ": " + strerror(errno);
    return false;
  }
  TripleReader rd(f);
  uint32_t uid, iid, rating;
  while (rd.next(&uid, &iid, &rating)) {
    std::unordered_map<uint32_t, uint32_t>::const_iterator ut = user2seq_.find(uid), it = item2seq_.find(iid);
    if (ut == user2seq_.end() || it == item2seq_.end()) continue; // capacity is exhausted after training
    if (rating_class(rating) == 0) continue;
    (*out)[Pair(ut->second, it->second)] = binary_ ? (uint8_t)1 : (uint8_t)rating;
  }
  fclose(f);
  if (rd.bad()) {
    *err = rd.complaint(path);
    return false;
  }
  return true;
}

bool Ratings::read_test_users(const std::string &path, std::map<uint32_t, bool> *out) const
{
  FILE *f = fopen(path.c_str(), "r");
  if (!f) return false;
  TripleReader rd(f);
  uint32_t uid;
  while (rd.number(&uid)) {
    std::unordered_map<uint32_t, uint32_t>::const_iterator ut = user2seq_.find(uid);
    if (ut != user2seq_.end()) (*out)[ut->second] = true;
  }
  fclose(f);
  return !rd.bad();
}

void Ratings::to_csr(std::vector<uint64_t> *row_ptr, std::vector<uint32_t> *col_idx, std::vector<uint8_t> *y) const
{
  row_ptr->assign(n() + 1, 0);
  col_idx->clear();
  y->clear();
  col_idx->reserve(nratings_);
  y->reserve(nratings_);
  for (uint32_t u = 0; u < n(); ++u) {
    const std::vector<uint32_t> &v = items_[u];
    const std::vector<uint8_t> &w = vals_[u];
    col_idx->insert(col_idx->end(), v.begin(), v.end());
    for (size_t j = 0; j < v.size(); ++j) y->push_back((uint8_t)(w[j] == 0 ? 1 : w[j]));
    (*row_ptr)[u + 1] = col_idx->size();
  }
}

void Ratings::write_marginals(const std::string &outdir) const
{
  FILE *f = fopen((outdir + "/byusers.tsv").c_str(), "w");
  if (f) {
    for (uint32_t u = 0; u < n(); ++u) {
      const std::vector<uint8_t> &w = vals_[u];
      if (w.empty()) continue;
      uint32_t t = 0;
      for (size_t j = 0; j < w.size(); ++j) t += w[j];
      fprintf(f, "%d\t%d\t%d\t%d\n", u, seq2user_[u], (int)w.size(), t);
    }
    fclose(f);
  }
  f = fopen((outdir + "/byitems.tsv").c_str(), "w");
  if (f) {
    for (uint32_t i = 0; i < m(); ++i) {
      if (item_degree_[i] == 0) continue;
      fprintf(f, "%d\t%d\t%d\t%d\n", i, seq2item_[i], (int)item_degree_[i], (uint32_t)item_total_[i]);
    }
    fclose(f);
  }
}

} // namespace hpfhost
