// Snapshot / rewind of the random sources behind the samplers (include/rpe/Utility.hpp): libc's hidden rand() state
// and the re-entrant GlibcRandom. Prints "ok" lines; any "FAIL" makes the Python test fail.
#include <cstdio>
#include <vector>

#include "rpe/Utility.hpp"

static int check(const char* what, bool ok) {
  printf("%s %s\n", ok ? "ok" : "FAIL", what);
  return ok ? 0 : 1;
}

int main() {
  int bad = 0;
  // libc rand(), unseeded like the reference, after some use
  for (int i = 0; i < 1000; ++i) (void)rand();
  rpe::RandState st;
  const bool saved = rpe::rand_save(nullptr, &st);
  bad += check("libc save", saved);
  std::vector<int> a(500);
  for (int& v : a) v = rand();
  bad += check("libc load", rpe::rand_load(nullptr, st));
  bool same = true;
  for (int v : a) same = same && rand() == v;
  bad += check("libc stream repeats after rewind", same);
  // partial re-consumption: rewind, draw 123, the next value is a[123]
  bad += check("libc load 2", rpe::rand_load(nullptr, st));
  for (int i = 0; i < 123; ++i) (void)rand();
  bad += check("libc continues from the rewound position", rand() == a[123]);
  // srand afterwards still behaves (the state array is libc's own again)
  srand(7);
  const int s7 = rand();
  srand(7);
  bad += check("srand after rewind", rand() == s7);
  // RandomElements through the default source: rows repeat after a rewind
  {
    RandomElements<int> re(1000);
    rpe::RandState s2;
    rpe::rand_save(nullptr, &s2);
    std::vector<int> r1, r2;
    re.run(3, &r1);
    rpe::rand_load(nullptr, s2);
    re.run(3, &r2);
    bad += check("RandomElements repeats", r1 == r2);
  }
  // GlibcRandom (explicit source) equals libc for the same seed, and rewinds
  {
    rpe::GlibcRandom g(42);
    srand(42);
    bool eq = true;
    for (int i = 0; i < 100; ++i) eq = eq && g.next() == rand();
    bad += check("GlibcRandom == libc", eq);
    rpe::RandState s3;
    bad += check("GlibcRandom save", g.save(&s3));
    const int x = g.next();
    (void)g.next();
    bad += check("GlibcRandom load", g.load(s3));
    bad += check("GlibcRandom repeats", g.next() == x);
  }
  return bad;
}
