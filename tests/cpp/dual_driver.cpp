// One driver, two builds: a program written ONLY in the reference's vocabulary — Eigen matrices, Sophus::SO3, the four
// adapters, the estimators and refits with the reference's signatures, samples drawn from ::rand() — compiled
//   (A) against the reference's own headers:   -I oracle/ref_shim -I /root/reference/pose        (CPU, Eigen stand-in)
//   (B) against the drop-in headers:           -I oracle/ref_shim -I include/rpe -lrpe_b200      (B200)
// without changing a line. tests/test_dual_driver.py builds both, runs (A) here and checks it against the oracle; the
// output of (A) is committed as tests/golden/dual_driver_golden.json for the GPU box, where (B) is run
// (tests/test_gpu_dual_driver.py). Eigen is not installed in this image: `Eigen/Dense` is the stand-in of oracle/ref_shim in
// both builds (the drop-in headers only need data() / rows() / cols() of it).
//
// usage: dual_driver <input.bin>     input: int32 n, then bv, xc, nc, xw, nw (3 x n float32 each), weights (n x 3 float32)
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include <Eigen/Dense>
#include "AbsoluteOrientationNormal.hpp"

using namespace Eigen;
typedef float data_type;
typedef Matrix<data_type, Dynamic, Dynamic> MX;

static unsigned fnv(unsigned h, int v) { return (h ^ (unsigned)v) * 16777619u; }

template <class Adapter>
static void report(const char* name, Adapter& adapter, int iter, unsigned mask_hash, int n0, int n1, int n2) {
  const Sophus::SO3<data_type> R = adapter.getRcw();
  printf("{\"case\": \"%s\", \"max_votes\": %d, \"iter\": %d, \"mask_hash\": %u, \"n_idx\": [%d, %d, %d], "
         "\"q\": [\"%a\", \"%a\", \"%a\", \"%a\"], \"t\": [\"%a\", \"%a\", \"%a\"]}\n",
         name, adapter.getMaxVotes(), iter, mask_hash, n0, n1, n2, (double)R.unit_quaternion().x(),
         (double)R.unit_quaternion().y(), (double)R.unit_quaternion().z(), (double)R.unit_quaternion().w(),
         (double)adapter.gettw()[0], (double)adapter.gettw()[1], (double)adapter.gettw()[2]);
}

static bool load(FILE* f, MX* m, int rows, int cols) {
  m->resize(rows, cols);
  return fread(m->data(), sizeof(data_type), (size_t)rows * cols, f) == (size_t)rows * cols;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* f = fopen(argv[1], "rb");
  int n = 0;
  if (!f || fread(&n, sizeof(int), 1, f) != 1 || n <= 0) return 2;
  MX U, P, N, Q, M, all_weights;
  if (!load(f, &U, 3, n) || !load(f, &P, 3, n) || !load(f, &N, 3, n) || !load(f, &Q, 3, n) || !load(f, &M, 3, n) ||
      !load(f, &all_weights, n, 3))
    return 2;
  fclose(f);
  const data_type focal = 585.f, thre_3d = 0.2f, thre_2d = 8.f, thre_nl = 0.1f, confidence = 0.99f;
  const int iteration = 1000;

  {  // SimpleMain.cpp:66-83
    AOOnlyPoseAdapter<data_type> adapter(P, Q);
    adapter.setFocal(focal, focal);
    adapter.setWeights(all_weights);
    int updated_iter = iteration;
    srand(101);
    shinji_prosac<data_type>(adapter, thre_3d, updated_iter, confidence);
    unsigned h = 2166136261u;
    for (int i = 0; i < n; i++) h = fnv(h, adapter.isInlier33(i));
    report("shinji_prosac", adapter, updated_iter, h, -1, (int)adapter.getInlierIdx().size(), -1);
    shinji_ls1<data_type>(adapter);
    report("shinji_prosac+shinji_ls1", adapter, updated_iter, h, -1, (int)adapter.getInlierIdx().size(), -1);
    updated_iter = iteration;
    srand(102);
    shinji_ransac2<data_type>(adapter, thre_3d, updated_iter, confidence);
    h = 2166136261u;
    for (int i = 0; i < n; i++) h = fnv(h, adapter.isInlier33(i));
    report("shinji_ransac2", adapter, updated_iter, h, -1, (int)adapter.getInlierIdx().size(), -1);
    shinji_ls1<data_type>(adapter);
    report("shinji_ransac2+shinji_ls1", adapter, updated_iter, h, -1, (int)adapter.getInlierIdx().size(), -1);
  }
  {  // SimpleMain.cpp:140-151
    PnPPoseAdapter<data_type> adapter(U, Q);
    adapter.setFocal(focal, focal);
    adapter.setWeights(all_weights);
    int updated_iter = iteration;
    srand(103);
    kneip_prosac<data_type>(adapter, thre_2d, updated_iter, confidence);
    unsigned h = 2166136261u;
    for (int i = 0; i < n; i++) h = fnv(h, adapter.isInlier23(i));
    report("kneip_prosac", adapter, updated_iter, h, (int)adapter.getInlierIdx().size(), -1, -1);
    updated_iter = iteration;
    srand(104);
    kneip_ransac<data_type>(adapter, thre_2d, updated_iter, confidence);
    h = 2166136261u;
    for (int i = 0; i < n; i++) h = fnv(h, adapter.isInlier23(i));
    report("kneip_ransac", adapter, updated_iter, h, (int)adapter.getInlierIdx().size(), -1, -1);
  }
  {  // SimpleMain.cpp:213-231
    AOPoseAdapter<data_type> adapter(U, P, Q);
    adapter.setFocal(focal, focal);
    adapter.setWeights(all_weights);
    int updated_iter = iteration;
    srand(105);
    shinji_kneip_ransac<data_type>(adapter, thre_3d, thre_2d, updated_iter, confidence);
    unsigned h = 2166136261u;
    for (int i = 0; i < n; i++) h = fnv(fnv(h, adapter.isInlier23(i)), adapter.isInlier33(i));
    PnPPoseAdapter<data_type>* p2 = &adapter;
    report("shinji_kneip_ransac", adapter, updated_iter, h, (int)p2->getInlierIdx().size(), (int)adapter.getInlierIdx().size(), -1);
    shinji_ls<data_type>(adapter);
    report("shinji_kneip_ransac+shinji_ls", adapter, updated_iter, h, (int)p2->getInlierIdx().size(),
           (int)adapter.getInlierIdx().size(), -1);
  }
  {  // TestMain.cpp:163-221
    NormalAOPoseAdapter<data_type> adapter(U, P, N, Q, M);
    adapter.setFocal(focal, focal);
    PnPPoseAdapter<data_type>* p2 = &adapter;
    AOPoseAdapter<data_type>* p3 = &adapter;
    const char* names[3] = {"nl_kneip_ransac", "nl_shinji_ransac", "nl_shinji_kneip_ransac"};
    for (int which = 0; which < 3; which++) {
      int updated_iter = iteration;
      srand(106 + which);
      if (which == 0) nl_kneip_ransac<data_type>(adapter, thre_2d, thre_nl, updated_iter, confidence);
      if (which == 1) nl_shinji_ransac<data_type>(adapter, thre_3d, thre_nl, updated_iter, confidence);
      if (which == 2) nl_shinji_kneip_ransac<data_type>(adapter, thre_3d, thre_2d, thre_nl, updated_iter, confidence);
      unsigned h = 2166136261u;
      for (int i = 0; i < n; i++) h = fnv(fnv(fnv(h, adapter.isInlier23(i)), adapter.isInlier33(i)), adapter.isInlierNN(i));
      report(names[which], adapter, updated_iter, h, (int)p2->getInlierIdx().size(), (int)p3->getInlierIdx().size(),
             (int)adapter.getInlierIdx().size());
    }
    nl_shinji_kneip_ls<data_type>(adapter);
    report("nl_shinji_kneip_ransac+nl_shinji_kneip_ls", adapter, 0, 0u, (int)p2->getInlierIdx().size(),
           (int)p3->getInlierIdx().size(), (int)adapter.getInlierIdx().size());
    adapter.setWeights(all_weights);
    nl_shinji_kneip_ls<data_type>(adapter);
    report("nl_shinji_kneip_ls(dynamic weights)", adapter, 0, 0u, (int)p2->getInlierIdx().size(), (int)p3->getInlierIdx().size(),
           (int)adapter.getInlierIdx().size());
  }
  return 0;
}
