#include "ratings.hh"

#include <errno.h>
#include <string.h>

namespace hpfhost {

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
  if (ch < '0' || ch > '9') return false;
  uint32_t v = 0;
  while (ch >= '0' && ch <= '9') {
    v = v * 10u + (uint32_t)(ch - '0');
    ch = get();
  }
  *out = neg ? 0u - v : v;
  return true;
}

bool Ratings::read_train(const std::string &dir, std::string *err)
{
  const std::string path = dir + "/train.tsv";
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
  finalize();
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
  return true;
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
