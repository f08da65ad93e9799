// ratings.hh -- host-side data step in front of the hot path: reads the
// reference's TSV data set and lays the training matrix out as the CSR that
// hpf_set_ratings_csr takes.
//
// Mirrors Ratings::read_generic (src/ratings.cc:63-119) and the id <-> seq maps
// (src/ratings.hh:117-151): sequence numbers are handed out in first-appearance
// order of train.tsv, rows keep file order, a repeated (user, item) line keeps
// its slot in the walk but carries the LAST rating (the reference stores values
// in a std::map that the later line overwrites, and walks the vector), ratings
// are truncated to uint8 (yval_t, src/env.hh:20), ratings of class 0 are
// dropped (src/ratings.hh:191-197), and held-out lines whose user or item was
// never seen in training are dropped (src/ratings.cc:79-81).
#ifndef HPF_HOST_RATINGS_HH
#define HPF_HOST_RATINGS_HH
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace hpfhost {

typedef std::pair<uint32_t, uint32_t> Pair;          // (user seq, item seq): the reference's Rating
typedef std::map<Pair, uint8_t> HeldoutMap;          // CountMap, iterated in (user, item) order

struct Options;

class Ratings {
public:
  Ratings(uint32_t max_users, uint32_t max_items, bool binary, uint32_t rating_threshold)
      : max_users_(max_users), max_items_(max_items), binary_(binary), threshold_(rating_threshold), nratings_(0) {}

  // train.tsv -> adjacency in file order.  Returns false if the file cannot be read.
  bool read_train(const std::string &dir, std::string *err);
  // validation.tsv / test.tsv -> map; unseen users / items are skipped
  bool read_heldout(const std::string &path, HeldoutMap *out, std::string *err) const;
  // test_users.tsv -> set of user seqs (Ratings::read_test_users, src/ratings.cc:273-292)
  bool read_test_users(const std::string &path, std::map<uint32_t, bool> *out) const;

  uint32_t n() const { return (uint32_t)user_items_.size(); }
  uint32_t m() const { return (uint32_t)item_users_.size(); }
  uint64_t nratings() const { return nratings_; }
  uint32_t user_id(uint32_t seq) const { return seq2user_[seq]; }
  uint32_t item_id(uint32_t seq) const { return seq2item_[seq]; }
  const std::vector<uint32_t> &seq2user() const { return seq2user_; }
  const std::vector<uint32_t> &seq2item() const { return seq2item_; }
  const std::vector<uint32_t> &items_of(uint32_t u) const { return user_items_[u]; }
  const std::vector<uint32_t> &users_of(uint32_t i) const { return item_users_[i]; }
  // Ratings::r(n, m): 0 when absent (src/ratings.hh:153-165)
  uint32_t r(uint32_t u, uint32_t i) const
  {
    std::unordered_map<uint64_t, uint8_t>::const_iterator it = value_.find(((uint64_t)u << 32) | i);
    return it == value_.end() ? 0u : it->second;
  }
  bool test_hit(uint32_t v) const { return binary_ ? v >= 1 : v >= threshold_; }

  // CSR in the reference's walk order; y is what vb*() would read through r(n, m),
  // with a wrapped-to-zero rating standing in as 1 (the loop only scales when y > 1)
  void to_csr(std::vector<uint64_t> *row_ptr, std::vector<uint32_t> *col_idx, std::vector<uint8_t> *y) const;

  // byusers.tsv / byitems.tsv (Ratings::write_marginal_distributions, src/ratings.cc:217-271)
  void write_marginals(const std::string &outdir) const;

private:
  uint32_t rating_class(uint32_t v) const { return binary_ ? (v >= threshold_ ? 1u : 0u) : v; }
  uint32_t max_users_, max_items_;
  bool binary_;
  uint32_t threshold_;
  uint64_t nratings_;
  std::unordered_map<uint32_t, uint32_t> user2seq_, item2seq_;
  std::vector<uint32_t> seq2user_, seq2item_;
  std::vector<std::vector<uint32_t> > user_items_, item_users_;
  std::unordered_map<uint64_t, uint8_t> value_;
};

} // namespace hpfhost
#endif
