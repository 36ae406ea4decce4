// rpe/Utility.hpp — host-side samplers with the reference's names and call signatures.
//
// Mirrors /root/reference/pose/Utility.hpp:
//   sortIndexes            :107-118   indices sorted by descending value
//   RandomElements<T>      :125-156   run(m, &v): m distinct indices of [0,n), partial Fisher-Yates from the tail
//   ProsacSampler<T>       :161-250   PROSAC growth function of Chum & Matas
// The draws feed the GPU as an H x 4 int32 sample table (rpe_c_api.h), so given the same random
// source the GPU pipeline evaluates exactly the hypotheses the CPU loop would.
//
// Random source. The reference calls ::rand() and never seeds it. Both samplers here take an
// optional rpe::RandSource*: by default (nullptr) they call ::rand() like the reference, so a
// program that swaps headers sees the same stream; rpe::GlibcRandom restates glibc's TYPE_3
// generator (the algorithm behind ::rand() on Linux) so that a table can be produced
// re-entrantly, without touching libc's hidden global state, from an explicit seed.
//
// Differences kept deliberately small: RandomElements no longer re-initialises its O(n) index
// array on every call (Utility.hpp:141-143) — it undoes its m swaps instead, which yields the same
// draws — and its destructor uses delete[] (the reference mismatches new[]/delete, :132-135).
#ifndef RPE_UTILITY_HPP_
#define RPE_UTILITY_HPP_

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace rpe {

// Snapshot of a generator. The GPU path draws sample rows ahead of the iterations the reference's early-stopping loop
// would actually run; afterwards the generator is put back and advanced by exactly the draws of the iterations the
// reference executes, so that whatever the program draws next (the next estimator, the Simulator) is unchanged.
struct RandState {
  uint32_t w[68];
  int words;
  void* where;
  RandState() : words(0), where(nullptr) {}
};

struct RandSource {
  virtual ~RandSource() {}
  virtual int next() = 0;  // same contract as ::rand(): uniform in [0, 2^31)
  virtual bool save(RandState*) { return false; }
  virtual bool load(const RandState&) { return false; }
};

// libc's hidden rand()/random() state. glibc keeps it in a user-replaceable array (initstate/setstate, POSIX): switching
// to a scratch array makes libc write the current rear index into the header word of the array it leaves, after which
// header + ring can be copied; copying them back and switching to that array again rewinds the generator.
struct LibcRandom : RandSource {
  int next() override { return ::rand(); }
  static char* scratch() {
    static char buf[128];
    return buf;
  }
  bool save(RandState* st) override {
#if defined(__GLIBC__)
    char* prev = ::initstate(1u, scratch(), 128);
    if (!prev) return false;
    uint32_t header;
    memcpy(&header, prev, sizeof(header));
    static const int kDegree[5] = {0, 7, 15, 31, 63};
    const int type = (int)(header % 5u);
    st->words = kDegree[type] + 1;
    st->where = prev;
    memcpy(st->w, prev, (size_t)st->words * 4);
    ::setstate(prev);
    return true;
#else
    (void)st;
    return false;
#endif
  }
  bool load(const RandState& st) override {
#if defined(__GLIBC__)
    if (!st.where || st.words <= 0) return false;
    char* cur = ::initstate(1u, scratch(), 128);  // leave the live array (libc rewrites its header, restored below)
    if (cur != st.where) {                        // the program installed another state array meanwhile
      if (cur) ::setstate(cur);
      return false;
    }
    memcpy(st.where, st.w, (size_t)st.words * 4);
    ::setstate((char*)st.where);
    return true;
#else
    (void)st;
    return false;
#endif
  }
};
inline bool rand_save(RandSource* src, RandState* st) {
  LibcRandom libc;
  return src ? src->save(st) : libc.save(st);
}
inline bool rand_load(RandSource* src, const RandState& st) {
  LibcRandom libc;
  return src ? src->load(st) : libc.load(st);
}

// glibc random_r TYPE_3: 31-word additive feedback, taps 3 and 31, 310 outputs discarded after seeding.
class GlibcRandom : public RandSource {
 public:
  explicit GlibcRandom(uint32_t seed = 1) { seed_with(seed); }
  void seed_with(uint32_t seed) {
    if (seed == 0) seed = 1;
    int32_t word = (int32_t)seed;
    ring_[0] = (uint32_t)word;
    for (int i = 1; i < kDeg; ++i) {
      // word = 16807 * word mod (2^31 - 1) without overflow (Schrage)
      const int32_t hi = word / 127773, lo = word % 127773;
      word = 16807 * lo - 2836 * hi;
      if (word < 0) word += 2147483647;
      ring_[i] = (uint32_t)word;
    }
    front_ = kSep;
    rear_ = 0;
    for (int i = 0; i < 10 * kDeg; ++i) (void)next();
  }
  int next() override {
    ring_[front_] += ring_[rear_];
    const int out = (int)(ring_[front_] >> 1);
    if (++front_ == kDeg) front_ = 0;
    if (++rear_ == kDeg) rear_ = 0;
    return out;
  }
  bool save(RandState* st) override {
    for (int i = 0; i < kDeg; ++i) st->w[i] = ring_[i];
    st->w[kDeg] = (uint32_t)front_;
    st->w[kDeg + 1] = (uint32_t)rear_;
    st->words = kDeg + 2;
    st->where = this;
    return true;
  }
  bool load(const RandState& st) override {
    if (st.where != this || st.words != kDeg + 2) return false;
    for (int i = 0; i < kDeg; ++i) ring_[i] = st.w[i];
    front_ = (int)st.w[kDeg];
    rear_ = (int)st.w[kDeg + 1];
    return true;
  }

 private:
  static const int kDeg = 31, kSep = 3;
  uint32_t ring_[kDeg];
  int front_, rear_;
};

}  // namespace rpe

template <typename T>
std::vector<int> sortIndexes(const std::vector<T>& v) {
  std::vector<int> order(v.size());
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&v](int a, int b) { return v[a] > v[b]; });
  return order;
}

template <class T>
class RandomElements {
 public:
  explicit RandomElements(int n, rpe::RandSource* src = nullptr) : n_(n), src_(src), perm_(new T[n > 0 ? n : 1]) {
    for (int i = 0; i < n_; ++i) perm_[i] = (T)i;
  }
  ~RandomElements() { delete[] perm_; }
  RandomElements(const RandomElements&) = delete;
  RandomElements& operator=(const RandomElements&) = delete;

  void run(int m, std::vector<T>* picked) {
    picked->clear();
    if (m > n_) m = n_;
    undo_.clear();
    for (int j = n_ - 1; j > n_ - m - 1; --j) {
      const int r = draw() % (j + 1);
      std::swap(perm_[r], perm_[j]);
      undo_.push_back(r);
      picked->push_back(perm_[j]);
    }
    // restore the identity permutation: revert the swaps last-to-first
    for (int k = (int)undo_.size() - 1, j = n_ - (int)undo_.size(); k >= 0; --k, ++j) std::swap(perm_[undo_[k]], perm_[j]);
  }

 private:
  int draw() { return src_ ? src_->next() : ::rand(); }
  int n_;
  rpe::RandSource* src_;
  T* perm_;
  std::vector<int> undo_;
};

template <class T>
class ProsacSampler {
 public:
  ProsacSampler(const int min_num_samples, const int num_datapoints, rpe::RandSource* src = nullptr)
      : m_(min_num_samples), N_(num_datapoints), src_(src) {
    restart(1);
  }
  // Jump to the k-th PROSAC sample (Eq. 6 of the paper); the growth state is rebuilt from scratch.
  void setSampleNumber(int k) { restart(k); }

  bool sample(std::vector<int>* subset_indices) {
    // advance the growth function to sample number t_ (incremental form of the reference's
    // `for (t = 1; t <= _t; t++)` recomputation: the state after t-1 steps is carried over)
    while (grown_to_ < t_) {
      ++grown_to_;
      if (grown_to_ > t_n_prime_ && n_ < N_) {
        const T t_n_plus1 = (t_n_ * (n_ + 1.0)) / (n_ + 1.0 - m_);
        t_n_prime_ += std::ceil(t_n_plus1 - t_n_);
        t_n_ = t_n_plus1;
        n_++;
      }
    }
    subset_indices->reserve(m_);
    std::vector<int> used;
    if (t_n_prime_ < t_) {
      for (int i = 0; i < m_; i++) subset_indices->push_back(fresh(n_, &used));
    } else {
      for (int i = 0; i < m_ - 1; i++) subset_indices->push_back(fresh(n_ - 1, &used));
      subset_indices->push_back(n_);  // the reference pushes index n (Utility.hpp:238), which can equal N
    }
    t_++;
    return true;
  }

 private:
  void restart(int k) {
    t_ = k;
    grown_to_ = 0;
    n_ = m_;
    t_n_ = (T)20000;  // _T_N
    for (int i = 0; i < m_; i++) t_n_ *= static_cast<T>(n_ - i) / (N_ - i);
    t_n_prime_ = 1.0;
  }
  int fresh(int bound, std::vector<int>* used) {
    int r;
    do {
      r = (src_ ? src_->next() : ::rand()) % bound;
    } while (std::find(used->begin(), used->end(), r) != used->end());
    used->push_back(r);
    return r;
  }
  int m_, N_;
  rpe::RandSource* src_;
  int t_, grown_to_, n_;
  T t_n_, t_n_prime_;
};

#endif  // RPE_UTILITY_HPP_
