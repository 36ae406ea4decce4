// CPU check of the product's host+device solver templates (include/rpe/solvers*.h) through the header API:
// prints the bits of shinji / kneip_main / kneip / nl_2p / o4_roots results for inputs read from stdin, so that
// tests/test_host_solvers.py can compare them bit for bit with the oracle in DET mode. No GPU involved.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "rpe/AbsoluteOrientationNormal.hpp"
#include "rpe/MinimalSolvers.hpp"

#ifndef REAL
#define REAL float  // -DREAL=double: the binary64 instantiation (what the device's f64 path runs)
#endif
typedef REAL real;

static unsigned long long bits(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
static unsigned long long bits(double f) {
  unsigned long long u;
  memcpy(&u, &f, 8);
  return u;
}
static void print_pose(const char* tag, const rpe::SE3<real>& s) {
  const rpe::Quaternion<real> q = s.so3().unit_quaternion();
  const rpe::Vec3<real> t = s.translation();
  printf("%s %llu %llu %llu %llu %llu %llu %llu\n", tag, bits(q.x()), bits(q.y()), bits(q.z()), bits(q.w()), bits(t[0]), bits(t[1]),
         bits(t[2]));
}

int main() {
  int cases;
  if (scanf("%d", &cases) != 1) return 1;
  for (int c = 0; c < cases; ++c) {
    rpe::MatrixX<real> Xw(3, 4), Xc(3, 4), bv(3, 4), Nw(3, 4), Nc(3, 4);
    rpe::MatrixX<real>* arr[5] = {&Xw, &Xc, &bv, &Nw, &Nc};
    for (int a = 0; a < 5; ++a)
      for (int i = 0; i < 12; ++i) {
        unsigned long long u;
        if (scanf("%llu", &u) != 1) return 1;
        real f;
        if (sizeof(real) == 4) {
          const unsigned u32 = (unsigned)u;
          memcpy(&f, &u32, 4);
        } else {
          memcpy(&f, &u, sizeof(real));
        }
        (*arr[a])(i) = f;
      }
    print_pose("shinji", shinji<real>(Xw, Xc, 3));
    std::vector<rpe::SE3<real> > sols;
    kneip_main<real>(Xw, bv, &sols);
    printf("kneip_main_count %d\n", (int)sols.size());
    for (size_t i = 0; i < sols.size(); ++i) print_pose("kneip_main", sols[i]);
    rpe::SE3<real> s4;
    const bool ok = kneip<real>(Xw, bv, &s4);
    printf("kneip4_ok %d\n", ok ? 1 : 0);
    if (ok) print_pose("kneip4", s4);
    rpe::SE3<real> snl;
    nl_2p<real>(Xc.col(0), Nc.col(0), Xc.col(1), Xw.col(0), Nw.col(0), Xw.col(1), &snl);
    print_pose("nl_2p", snl);
  }
  return 0;
}
