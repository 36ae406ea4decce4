// oracle/sampler_model.hpp — TEST INFRASTRUCTURE (see oracle/README.md).
//
// Restatement of the reference's samplers (pose/Utility.hpp:125-250) and of the libc
// generator they draw from. The reference calls ::rand() and never calls srand(), i.e. it
// runs glibc's TYPE_3 additive-feedback generator from seed 1. glibc source is not in the
// container; the published algorithm (r[i] = r[i-3] + r[i-31], output >> 1, 310 discarded
// warm-up values) is restated here and PINNED bit-for-bit against this container's
// ::rand() in tests/test_oracle_rand.py.
#ifndef ORACLE_SAMPLER_MODEL_HPP_
#define ORACLE_SAMPLER_MODEL_HPP_

#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

namespace orc {

class GlibcRand {
 public:
  explicit GlibcRand(unsigned seed = 1) { reseed(seed); }
  void reseed(unsigned seed) {
    if (seed == 0) seed = 1;
    int32_t r[34];
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; ++i) {
      // 16807 * r[i-1] mod 2147483647 via Schrage, as glibc __srandom_r does
      const int64_t hi = r[i - 1] / 127773;
      const int64_t lo = r[i - 1] % 127773;
      int64_t word = 16807 * lo - 2836 * hi;
      if (word < 0) word += 2147483647;
      r[i] = (int32_t)word;
    }
    for (int i = 0; i < 31; ++i) state_[i] = (uint32_t)r[i];
    f_ = 3;
    b_ = 0;
    for (int i = 0; i < 310; ++i) (void)next();
  }
  // returns what ::rand() returns: 31 bits
  int next() {
    state_[f_] += state_[b_];
    const uint32_t result = state_[f_] >> 1;
    f_ = (f_ + 1) % 31;
    b_ = (b_ + 1) % 31;
    return (int)result;
  }

 private:
  uint32_t state_[31];
  int f_, b_;
};

// Utility.hpp:125-156. Semantics of run(m): identity permutation of [0,n), then for
// j = n-1 ... n-m: r = rand() % (j+1); swap idx[r], idx[j]; emit idx[j].
// The O(n) re-initialisation per call (Utility.hpp:141-143) is kept literally here.
class RandomElementsModel {
 public:
  RandomElementsModel(int n, GlibcRand* rng) : idx_(n), n_(n), rng_(rng) {}
  void run(int m, std::vector<int>* out) {
    out->clear();
    for (int i = 0; i < n_; ++i) idx_[i] = i;
    for (int j = n_ - 1; j > n_ - m - 1; --j) {
      const int ridx = rng_->next() % (j + 1);
      const int temp = idx_[ridx];
      idx_[ridx] = idx_[j];
      idx_[j] = temp;
      out->push_back(temp);
    }
  }

 private:
  std::vector<int> idx_;
  int n_;
  GlibcRand* rng_;
};

// Utility.hpp:161-250 (Chum & Matas PROSAC growth function). T is the reference's Tp.
template <class T>
class ProsacSamplerModel {
 public:
  ProsacSamplerModel(int min_num_samples, int num_datapoints, GlibcRand* rng)
      : N_(num_datapoints), T_N_(20000), t_(1), m_(min_num_samples), rng_(rng) {}
  void sample(std::vector<int>* subset) {
    T t_n = (T)T_N_;
    int n = m_;
    for (int i = 0; i < m_; i++) t_n *= static_cast<T>(n - i) / (N_ - i);
    T t_n_prime = 1.0;
    for (int t = 1; t <= t_; t++) {
      if (t > t_n_prime && n < N_) {
        T t_n_plus1 = (t_n * (n + 1.0)) / (n + 1.0 - m_);
        t_n_prime += std::ceil(t_n_plus1 - t_n);
        t_n = t_n_plus1;
        n++;
      }
    }
    if (t_n_prime < t_) {
      std::vector<int> used;
      for (int i = 0; i < m_; i++) {
        int r;
        while (std::find(used.begin(), used.end(), (r = rng_->next() % n)) != used.end()) {
        }
        used.push_back(r);
        subset->push_back(r);
      }
    } else {
      std::vector<int> used;
      for (int i = 0; i < m_ - 1; i++) {
        int r;
        while (std::find(used.begin(), used.end(), (r = rng_->next() % (n - 1))) != used.end()) {
        }
        used.push_back(r);
        subset->push_back(r);
      }
      subset->push_back(n);  // Utility.hpp:238 — index n (can equal N: reference quirk, kept)
    }
    t_++;
  }

 private:
  int N_, T_N_, t_, m_;
  GlibcRand* rng_;
};

}  // namespace orc

#endif  // ORACLE_SAMPLER_MODEL_HPP_
