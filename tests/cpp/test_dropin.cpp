// Drop-in test: the reference's own call pattern (SimpleMain.cpp:66-82, :140-151, :213-231; TestMain.cpp:163-221)
// written against include/rpe/*.hpp. Prints one JSON object per case; tests/test_gpu_cpp_api.py compares them with
// the CPU oracle on the same inputs and the same ::rand() sample stream.
#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "rpe/AbsoluteOrientationNormal.hpp"
#include "rpe/MinimalSolvers.hpp"
#include "rpe/Simulator.hpp"

typedef float data_type;

template <class Adapter>
static void report(const char* name, Adapter& adapter, int iter, int n_inl) {
  const rpe::Quaternion<data_type> q = adapter.getRcw().unit_quaternion();
  const rpe::Vec3<data_type> t = adapter.gettw();
  printf("{\"case\": \"%s\", \"max_votes\": %d, \"iter\": %d, \"n_inliers\": %d, \"q\": [%.9g, %.9g, %.9g, %.9g], \"t\": [%.9g, %.9g, %.9g]}\n",
         name, adapter.getMaxVotes(), iter, n_inl, q.x(), q.y(), q.z(), q.w(), t[0], t[1], t[2]);
}

int main(int argc, char** argv) {
  const int total = argc > 1 ? atoi(argv[1]) : 1000;
  const int iteration = argc > 2 ? atoi(argv[2]) : 100000;
  const data_type f = 585., min_depth = 0.4f, max_depth = 8.f;
  // inputs through the C-ABI generators so that the Python side can rebuild exactly the same arrays
  float qg[4], tg[3];
  rpe_sim_pose(11, (float)(M_PI / 2), 5.0f, qg, tg);
  rpe::MatrixX<data_type> Q(3, total), P(3, total), U(3, total), M(3, total), N(3, total), Pgt(3, total);
  rpe::MatrixX<data_type> all_weights(total, 3);

  // ---- test_3d_3d (SimpleMain.cpp:21-93): AOOnlyPoseAdapter + shinji_ransac2 + shinji_ls1
  rpe_sim_3d_3d(12, qg, tg, total, 0.1f, 0.5f, min_depth, max_depth, f, 1, Q.data(), P.data(), all_weights.data());
  {
    AOOnlyPoseAdapter<data_type> adapter(P, Q);
    adapter.setFocal(f, f);
    adapter.setWeights(all_weights);
    int updated_iter = iteration;
    shinji_ransac2<data_type>(adapter, 0.25f, updated_iter, 0.9999f);
    report("shinji_ransac2", adapter, updated_iter, (int)adapter.getInlierIdx().size());
    shinji_ls1<data_type>(adapter);
    report("shinji_ls1", adapter, updated_iter, (int)adapter.getInlierIdx().size());
    shinji_ls2<data_type>(adapter);
    report("shinji_ls2", adapter, updated_iter, (int)adapter.getInlierIdx().size());
  }
  // ---- test_3d_2d (SimpleMain.cpp:95-166): PnPPoseAdapter + kneip_ransac
  rpe_sim_2d_3d(13, qg, tg, total, 1.0f, 0.3f, min_depth, max_depth, f, 1, Q.data(), U.data(), Pgt.data(), all_weights.data());
  {
    PnPPoseAdapter<data_type> adapter(U, Q);
    adapter.setFocal(f, f);
    adapter.setWeights(all_weights);
    int updated_iter = 2000;
    kneip_ransac<data_type>(adapter, 8.f, updated_iter, 0.99f);
    report("kneip_ransac", adapter, updated_iter, (int)adapter.getInlierIdx().size());
    refine_lm<data_type>(adapter, nullptr, 8);
    report("kneip_ransac+lm", adapter, updated_iter, (int)adapter.getInlierIdx().size());
  }
  // ---- TestMain.cpp:163-221: NormalAOPoseAdapter, hybrid and normal-aware estimators, multi-modal refinement
  rpe_sim_2d_3d_nl(14, qg, tg, total, 1.0f, 0.3f, 0.05f, 0.3f, (float)(2.0 * M_PI / 180.), 0.3f, min_depth, max_depth, f, 1,
                   Q.data(), M.data(), P.data(), N.data(), U.data(), all_weights.data());
  {
    NormalAOPoseAdapter<data_type> adapter(U, P, N, Q, M);
    adapter.setFocal(f, f);
    int updated_iter = 300;
    shinji_kneip_ransac<data_type>(adapter, 0.2f, 8.f, updated_iter, 0.99f);
    report("shinji_kneip_ransac", adapter, updated_iter, (int)((AOPoseAdapter<data_type>&)adapter).getInlierIdx().size());
    updated_iter = 300;
    nl_kneip_ransac<data_type>(adapter, 8.f, 0.1f, updated_iter, 0.99f);
    report("nl_kneip_ransac", adapter, updated_iter, (int)adapter.getInlierIdx().size());
    updated_iter = 300;
    nl_shinji_ransac<data_type>(adapter, 0.2f, 0.1f, updated_iter, 0.99f);
    report("nl_shinji_ransac", adapter, updated_iter, (int)adapter.getInlierIdx().size());
    updated_iter = 300;
    nl_shinji_kneip_ransac<data_type>(adapter, 0.2f, 8.f, 0.1f, updated_iter, 0.99f);
    report("nl_shinji_kneip_ransac", adapter, updated_iter, (int)adapter.getInlierIdx().size());
    nl_shinji_kneip_ls<data_type>(adapter);
    report("nl_shinji_kneip_ls", adapter, updated_iter, (int)adapter.getInlierIdx().size());
    adapter.setWeights(all_weights);
    nl_shinji_kneip_ls<data_type>(adapter);
    report("nl_shinji_kneip_ls_dw", adapter, updated_iter, (int)adapter.getInlierIdx().size());
  }
  // ---- Tp = double (TestMain.cpp instantiates its estimators as <double>): binary64 arrays read from a file written
  // by the Python test (five 3 x n blocks: bv, xc, nc, xw, nw), decided in binary64 on the device
  if (argc > 3) {
    FILE* fp = fopen(argv[3], "rb");
    if (!fp) return 2;
    rpe::MatrixX<double> Ud(3, total), Pd(3, total), Nd(3, total), Qd(3, total), Md(3, total);
    rpe::MatrixX<double>* arrs[5] = {&Ud, &Pd, &Nd, &Qd, &Md};
    for (int a = 0; a < 5; ++a)
      if (fread(arrs[a]->data(), sizeof(double), (size_t)3 * total, fp) != (size_t)3 * total) return 3;
    fclose(fp);
    NormalAOPoseAdapter<double> adapter(Ud, Pd, Nd, Qd, Md);
    adapter.setFocal(585., 585.);
    int updated_iter = 300;
    nl_shinji_kneip_ransac<double>(adapter, 0.2, 8., 0.1, updated_iter, 0.99);
    const rpe::Quaternion<double> q = adapter.getRcw().unit_quaternion();
    const rpe::Vec3<double> t = adapter.gettw();
    printf("{\"case\": \"nl_shinji_kneip_ransac_f64\", \"max_votes\": %d, \"iter\": %d, \"n_inliers\": %d, \"q\": [%.17g, %.17g, %.17g, %.17g], \"t\": [%.17g, %.17g, %.17g]}\n",
           adapter.getMaxVotes(), updated_iter, (int)adapter.getInlierIdx().size(), q.x(), q.y(), q.z(), q.w(), t[0], t[1], t[2]);
    nl_shinji_kneip_ls<double>(adapter);  // refits of a double adapter: float copies on the device, binary64 accumulation
    {
      const rpe::Quaternion<double> ql = adapter.getRcw().unit_quaternion();
      const rpe::Vec3<double> tl = adapter.gettw();
      printf("{\"case\": \"nl_shinji_kneip_ls_f64\", \"max_votes\": %d, \"iter\": %d, \"n_inliers\": 0, \"q\": [%.17g, %.17g, %.17g, %.17g], \"t\": [%.17g, %.17g, %.17g]}\n",
             adapter.getMaxVotes(), updated_iter, ql.x(), ql.y(), ql.z(), ql.w(), tl[0], tl[1], tl[2]);
    }
    PnPPoseAdapter<double> pnp(Ud, Qd);
    pnp.setFocal(585., 585.);
    updated_iter = 500;
    kneip_ransac<double>(pnp, 8., updated_iter, 0.99);
    const rpe::Quaternion<double> q2 = pnp.getRcw().unit_quaternion();
    const rpe::Vec3<double> t2 = pnp.gettw();
    printf("{\"case\": \"kneip_ransac_f64\", \"max_votes\": %d, \"iter\": %d, \"n_inliers\": %d, \"q\": [%.17g, %.17g, %.17g, %.17g], \"t\": [%.17g, %.17g, %.17g]}\n",
           pnp.getMaxVotes(), updated_iter, (int)pnp.getInlierIdx().size(), q2.x(), q2.y(), q2.z(), q2.w(), t2[0], t2[1], t2[2]);
  }
  return 0;
}
