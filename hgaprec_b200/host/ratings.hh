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
//
// Unlike the reference (fscanf + a std::map node per rating, minutes and tens of
// GB at Netflix scale) the reader parses the file from a large buffer and keeps
// 5 bytes per rating: per user the item list and a parallel value list.
//
// Binary cache (the reference has no counterpart): with use_cache the parsed
// training matrix -- id maps, CSR, post-fix-up values -- is kept next to the data as
// <dir>/train.tsv.hpfcsr and reloaded by the next run instead of parsing the text again.  A cache is
// only used when it was written for the same train.tsv (size and modification time), the same -n / -m
// caps and the same -binary-data / -rating-threshold, and its checksum holds; otherwise the text is
// parsed and the cache rewritten.  Whatever is loaded is bit-identical to what the parser builds.
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

// whitespace-separated unsigned integers from a file, three per record ("%u\t%u\t%u\n")
class TripleReader {
public:
  explicit TripleReader(FILE *f) : f_(f), buf_(1 << 22), pos_(0), end_(0), base_(0), bad_(false), bad_at_(0) {}
  // one line of three numbers; false at the end of the input AND at a malformed token (bad() tells which): a line
  // cut short by the end of the file counts as malformed
  bool next(uint32_t *a, uint32_t *b, uint32_t *c)
  {
    if (!number(a)) return false;
    if (number(b) && number(c)) return true;
    if (!bad_) { bad_ = true; bad_at_ = base_ + pos_; }
    return false;
  }
  bool number(uint32_t *out);
  // a token that is not an unsigned decimal stopped the reader (a header line, a float, a stray character)
  bool bad() const { return bad_; }
  uint64_t bad_offset() const { return bad_at_; }
  std::string complaint(const std::string &path) const
  {
    return "malformed input in " + path + " at byte " + std::to_string(bad_at_) + ": expected unsigned decimal numbers";
  }

private:
  int get()
  {
    if (pos_ == end_) {
      base_ += end_;
      end_ = fread(buf_.data(), 1, buf_.size(), f_);
      pos_ = 0;
      if (end_ == 0) return -1;
    }
    return (unsigned char)buf_[pos_++];
  }
  FILE *f_;
  std::vector<char> buf_;
  size_t pos_, end_;
  uint64_t base_; // file offset of buf_[0]
  bool bad_;
  uint64_t bad_at_;
};

class Ratings {
public:
  Ratings(uint32_t max_users, uint32_t max_items, bool binary, uint32_t rating_threshold)
      : max_users_(max_users), max_items_(max_items), binary_(binary), threshold_(rating_threshold), nratings_(0) {}

  // train.tsv -> adjacency in file order.  Returns false if the file cannot be read.
  bool read_train(const std::string &dir, std::string *err, bool use_cache = false);
  // what the cache did for the last read_train: "", "loaded <path>", "written <path>" or "not written: <why>"
  const std::string &cache_note() const { return cache_note_; }
  // validation.tsv / test.tsv -> map; unseen users / items are skipped
  bool read_heldout(const std::string &path, HeldoutMap *out, std::string *err) const;
  // test_users.tsv -> set of user seqs (Ratings::read_test_users, src/ratings.cc:273-292)
  bool read_test_users(const std::string &path, std::map<uint32_t, bool> *out) const;

  uint32_t n() const { return (uint32_t)items_.size(); }
  uint32_t m() const { return (uint32_t)seq2item_.size(); }
  uint64_t nratings() const { return nratings_; }
  uint32_t user_id(uint32_t seq) const { return seq2user_[seq]; }
  uint32_t item_id(uint32_t seq) const { return seq2item_[seq]; }
  const std::vector<uint32_t> &seq2user() const { return seq2user_; }
  const std::vector<uint32_t> &seq2item() const { return seq2item_; }
  // Ratings::get_movies(n): the user's items in file order (a repeated line appears twice)
  const std::vector<uint32_t> &items_of(uint32_t u) const { return items_[u]; }
  // value the reference's loops read for the j-th entry of user u: Ratings::r(u, items_of(u)[j])
  uint32_t value_at(uint32_t u, size_t j) const { return vals_[u][j]; }
  // Ratings::get_users(m)->size(): lines that named the item
  uint32_t item_degree(uint32_t i) const { return item_degree_[i]; }
  uint64_t item_total(uint32_t i) const { return item_total_[i]; } // sum of the item's ratings (byitems.tsv)
  // Ratings::r(n, m): 0 when absent (src/ratings.hh:153-165)
  uint32_t r(uint32_t u, uint32_t i) const
  {
    const std::vector<uint32_t> &v = items_[u];
    for (size_t j = 0; j < v.size(); ++j)
      if (v[j] == i) return vals_[u][j];
    return 0u;
  }
  bool test_hit(uint32_t v) const { return binary_ ? v >= 1 : v >= threshold_; }

  // CSR in the reference's walk order; y is what vb*() would read through r(n, m),
  // with a wrapped-to-zero rating standing in as 1 (the loop only scales when y > 1)
  void to_csr(std::vector<uint64_t> *row_ptr, std::vector<uint32_t> *col_idx, std::vector<uint8_t> *y) const;

  // byusers.tsv / byitems.tsv (Ratings::write_marginal_distributions, src/ratings.cc:217-271)
  void write_marginals(const std::string &outdir) const;

private:
  uint32_t rating_class(uint32_t v) const { return binary_ ? (v >= threshold_ ? 1u : 0u) : v; }
  void finalize(); // repeated (user, item) lines take the last value; per-item degree and rating total
  bool load_cache(const std::string &tsv, const std::string &cache);
  void save_cache(const std::string &tsv, const std::string &cache);
  std::string cache_note_;
  uint32_t max_users_, max_items_;
  bool binary_;
  uint32_t threshold_;
  uint64_t nratings_;
  std::unordered_map<uint32_t, uint32_t> user2seq_, item2seq_;
  std::vector<uint32_t> seq2user_, seq2item_;
  std::vector<std::vector<uint32_t> > items_;
  std::vector<std::vector<uint8_t> > vals_;
  std::vector<uint32_t> item_degree_;
  std::vector<uint64_t> item_total_;
};

} // namespace hpfhost
#endif
