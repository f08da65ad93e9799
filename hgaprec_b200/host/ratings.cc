#include "ratings.hh"

#include <errno.h>
#include <string.h>

namespace hpfhost {

namespace {
// one "%u\t%u\t%u" record per call; false at end of input
bool next_triple(FILE *f, uint32_t *a, uint32_t *b, uint32_t *c)
{
  return fscanf(f, "%u %u %u", a, b, c) == 3;
}
} // namespace

bool Ratings::read_train(const std::string &dir, std::string *err)
{
  const std::string path = dir + "/train.tsv";
  FILE *f = fopen(path.c_str(), "r");
  if (!f) {
    *err = "cannot open file " + path + ": " + strerror(errno);
    return false;
  }
  uint32_t uid, iid, rating;
  while (next_triple(f, &uid, &iid, &rating)) {
    std::unordered_map<uint32_t, uint32_t>::iterator ut = user2seq_.find(uid), it = item2seq_.find(iid);
    if ((ut == user2seq_.end() && seq2user_.size() >= max_users_) || (it == item2seq_.end() && seq2item_.size() >= max_items_))
      continue;
    if (rating_class(rating) == 0) continue;
    uint32_t u, i;
    if (ut == user2seq_.end()) {
      u = (uint32_t)seq2user_.size();
      user2seq_[uid] = u;
      seq2user_.push_back(uid);
      user_items_.push_back(std::vector<uint32_t>());
    } else
      u = ut->second;
    if (it == item2seq_.end()) {
      i = (uint32_t)seq2item_.size();
      item2seq_[iid] = i;
      seq2item_.push_back(iid);
      item_users_.push_back(std::vector<uint32_t>());
    } else
      i = it->second;
    nratings_++;
    value_[((uint64_t)u << 32) | i] = binary_ ? (uint8_t)1 : (uint8_t)rating;
    user_items_[u].push_back(i);
    item_users_[i].push_back(u);
  }
  fclose(f);
  return true;
}

bool Ratings::read_heldout(const std::string &path, HeldoutMap *out, std::string *err) const
{
  FILE *f = fopen(path.c_str(), "r");
  if (!f) {
    *err = "cannot open file " + path + ": " + strerror(errno);
    return false;
  }
  uint32_t uid, iid, rating;
  while (next_triple(f, &uid, &iid, &rating)) {
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
  uint32_t uid;
  while (fscanf(f, "%u", &uid) == 1) {
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
    const std::vector<uint32_t> &v = user_items_[u];
    for (size_t j = 0; j < v.size(); ++j) {
      col_idx->push_back(v[j]);
      const uint32_t val = r(u, v[j]);
      y->push_back((uint8_t)(val == 0 ? 1 : val));
    }
    (*row_ptr)[u + 1] = col_idx->size();
  }
}

void Ratings::write_marginals(const std::string &outdir) const
{
  FILE *f = fopen((outdir + "/byusers.tsv").c_str(), "w");
  if (f) {
    for (uint32_t u = 0; u < n(); ++u) {
      const std::vector<uint32_t> &v = user_items_[u];
      if (v.empty()) continue;
      uint32_t t = 0;
      for (size_t j = 0; j < v.size(); ++j) t += r(u, v[j]);
      fprintf(f, "%d\t%d\t%d\t%d\n", u, seq2user_[u], (int)v.size(), t);
    }
    fclose(f);
  }
  f = fopen((outdir + "/byitems.tsv").c_str(), "w");
  if (f) {
    for (uint32_t i = 0; i < m(); ++i) {
      const std::vector<uint32_t> &v = item_users_[i];
      if (v.empty()) continue;
      uint32_t t = 0;
      for (size_t j = 0; j < v.size(); ++j) t += r(v[j], i);
      fprintf(f, "%d\t%d\t%d\t%d\n", i, seq2item_[i], (int)v.size(), t);
    }
    fclose(f);
  }
}

} // namespace hpfhost
