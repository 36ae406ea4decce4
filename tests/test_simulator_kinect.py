"""CPU: simulate_kinect_2d_3d_nl_correspondences (Simulator.hpp:368-436) — the Kinect lateral / axial noise model of
Nguyen, Izadi & Lovell restated in include/rpe/sim_core.hpp and exported as rpe_sim_kinect_2d_3d_nl."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _R(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_kinect_noise_model_statistics(rpe):
    n, f = 200000, 585.0
    q, t = rpe.sim_pose(21)
    d = rpe.sim_kinect_2d_3d_nl(22, q, t, n, n2d=1.0, or2d=0.0, or3d=0.2, nnl=float(np.deg2rad(2.0)), ornl=0.0)
    R, tt = _R(q), np.asarray(t, np.float64)
    P = d["xc"].astype(np.float64)
    Pgt = d["xw"].astype(np.float64) @ R.T + tt          # world points are exact: camera ground truth
    e = P - Pgt
    w = d["weights"][1].astype(np.float64)
    inl = np.abs(e).max(axis=1) < 0.5                     # 3-D outliers are raw frustum points
    assert abs((~inl).mean() - 0.2) < 0.01
    z = Pgt[:, 2]
    assert (w > 0).all() and w.max() <= 1.0 + 1e-6       # sigma_axial(0, min_depth) / sigma_axial
    # axial sigma: .0012 + .0019 (z - .4)^2 for theta <= 60 deg — recover it from the weights and from the z residuals
    sig_a = 0.0012 / w
    base = 0.0012 + 0.0019 * (z - 0.4) ** 2
    assert (sig_a >= base * (1 - 1e-4)).all()             # the grazing-angle term only adds
    frontal = inl & (np.abs(sig_a / base - 1) < 1e-4)
    assert frontal.mean() > 0.3
    def robust_sigma(v):  # a few 3-D outliers land within 0.5 m of the truth: use the median absolute deviation
        return 1.4826 * np.median(np.abs(v))

    for lo, hi, tol in [(0.5, 1.5, 0.15), (3.0, 4.0, 0.05), (6.5, 7.5, 0.03)]:
        sel = frontal & (z > lo) & (z < hi)
        ratio = robust_sigma(e[sel, 2] / base[sel])
        assert abs(ratio - 1) < tol, (lo, hi, ratio)
    # lateral sigma = (.8 + .035 theta / (pi/2 - theta)) z / f >= .8 z / f, and x / y share it
    sel = inl & (z > 3.0) & (z < 4.0)
    sx, sy = robust_sigma(e[sel, 0]), robust_sigma(e[sel, 1])
    assert sx > 0.8 * 3.0 / f and abs(sx / sy - 1) < 0.05
    # bearing vectors: unit, consistent with the camera ground truth to the pixel noise
    bv = d["bv"].astype(np.float64)
    assert np.abs(np.linalg.norm(bv, axis=1) - 1).max() < 1e-5
    px = f * (bv[:, :2] / bv[:, 2:3] - Pgt[:, :2] / Pgt[:, 2:3])
    assert abs(px.std() - 1.0) < 0.05


def test_kinect_header_api_compiles_and_runs(tmp_path):
    src = tmp_path / "k.cpp"
    src.write_text('''
#include <cstdio>
#include "rpe/Simulator.hpp"
int main() {
  rpe::SO3<double> R = generate_random_rotation<double>(M_PI / 2, false);
  rpe::Vec3<double> t = generate_random_translation_uniform<double>(5.0);
  rpe::MatrixX<double> Q, M, P, N, U, W;
  simulate_kinect_2d_3d_nl_correspondences<double>(R, t, 500, 1.0, 0.1, 0.1, 0.03, 0.1, 0.4, 8.0, 585.0, &Q, &M, &P, &N, &U, &W);
  rpe::MatrixX<double> uv = project_point_cloud<double>(P, 585.0);
  rpe::Vec3<double> c = generate_a_random_point<double>(0.4, 8.0, 320. / 585., 240. / 585.);
  if (uv.rows() != 2 || uv.cols() != 500 || !(c[2] >= 0.4 && c[2] <= 8.0)) return 2;
  const double sa = axial_noise_kinect<double>(0.0, 0.4), sl = lateral_noise_kinect<double>(0.0, 2.0, 585.0);
  printf("%d %d %.6f %.6f\\n", (int)P.cols(), (int)W.rows(), sa, sl);
  return (P.cols() == 500 && W.rows() == 500 && W.cols() == 3) ? 0 : 1;
}
''')
    exe = str(tmp_path / "k")
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)],
                   check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "500" and abs(float(out[2]) - 0.0012) < 1e-6 and abs(float(out[3]) - 0.8 * 2.0 / 585.0) < 1e-6
