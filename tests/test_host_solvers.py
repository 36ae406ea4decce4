"""CPU: the product's host+device solver templates (include/rpe/solvers.h, solvers_p3p.h) through the header-only
API, against the oracle in DET mode — bit for bit. The same templates are what the CUDA generators instantiate."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, real="float"):
    exe = os.path.join(tmp_path, "test_host_solvers_" + real)
    subprocess.run(["g++", "-std=c++11", "-O2", "-ffp-contract=off", "-DREAL=" + real, "-I", os.path.join(ROOT, "include"), "-o",
                    exe, os.path.join(ROOT, "tests", "cpp", "test_host_solvers.cpp")], check=True)
    return exe


def test_header_api_compiles_cxx11(tmp_path):
    """The drop-in program written against include/rpe/*.hpp compiles as plain C++11 (no nvcc, no Eigen)."""
    obj = os.path.join(tmp_path, "dropin.o")
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", "-o", obj,
                    os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp")], check=True)


import pytest


@pytest.mark.parametrize("real", ["float", "double"])
def test_host_solvers_bit_exact_vs_oracle(orc, rpe, tmp_path, real):
    """float: what the device's binary32 generators run; double: the instantiation of the binary64 path."""
    exe = _build(str(tmp_path), real)
    npdt, npbits = (np.float32, np.uint32) if real == "float" else (np.float64, np.uint64)
    orc.set_math_mode(orc.DET)
    cases = 300
    rng = np.random.default_rng(3)
    q, t = rpe.sim_pose(5)
    d = rpe.sim_2d_3d_nl(6, q, t, 400, or2d=0.1, or3d=0.1, ornl=0.1)
    if real == "double":  # directions that are unit vectors to double precision, as a double Simulator produces them
        d = {k: v.astype(np.float64) for k, v in d.items()}
        for k in ("bv", "nw", "nc"):
            d[k] = d[k] / np.linalg.norm(d[k], axis=1, keepdims=True)
    lines = [str(cases)]
    picks = []
    for _ in range(cases):
        idx = rng.choice(400, 4, replace=False)
        picks.append(idx)
        for name in ("xw", "xc", "bv", "nw", "nc"):
            lines.append(" ".join(str(int(v)) for v in np.ascontiguousarray(d[name][idx]).view(npbits).ravel()))
    out = subprocess.run([exe], input="\n".join(lines), capture_output=True, text=True, check=True).stdout.split("\n")
    it = iter([l for l in out if l])

    def pose_bits(q_, t_):
        return list(np.asarray(q_, npdt).view(npbits)) + list(np.asarray(t_, npdt).view(npbits))

    n_kneip = 0
    for idx in picks:
        Xw, Xc, bv, Nw, Nc = (d[k][idx] for k in ("xw", "xc", "bv", "nw", "nc"))
        tag, *vals = next(it).split()
        qs, ts, ok = orc.shinji(Xw[:3], Xc[:3], K=3, cols=4, dt=npdt)
        assert tag == "shinji" and [int(v) for v in vals] == pose_bits(qs, ts)
        tag, cnt = next(it).split()
        qk, tk = orc.kneip_main(Xw, bv, dt=npdt)
        assert tag == "kneip_main_count" and int(cnt) == len(qk)
        for i in range(int(cnt)):
            tag, *vals = next(it).split()
            assert [int(v) for v in vals] == pose_bits(qk[i], tk[i])
        tag, okv = next(it).split()
        q4, t4, ok4 = orc.kneip4(Xw, bv, dt=npdt)
        assert tag == "kneip4_ok" and int(okv) == int(ok4)
        if ok4:
            tag, *vals = next(it).split()
            assert [int(v) for v in vals] == pose_bits(q4, t4)
            n_kneip += 1
        tag, *vals = next(it).split()
        qn, tn = orc.nl_2p(Xc[0], Nc[0], Xc[1], Xw[0], Nw[0], Xw[1], dt=npdt)
        assert tag == "nl_2p" and [int(v) for v in vals] == pose_bits(qn, tn)
    orc.set_math_mode(orc.LIBM)
    assert n_kneip > cases // 2


def test_random_source_snapshot_and_rewind(tmp_path):
    """The drop-in headers draw whole passes of sample rows ahead and then put ::rand() back to where the reference's
    early-stopping loop would have left it (include/rpe/Utility.hpp RandState)."""
    exe = str(tmp_path / "test_rand_state")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "test_rand_state.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "FAIL" not in out.stdout, out.stdout
    assert out.stdout.count("ok ") >= 11


def test_every_public_header_is_self_contained(tmp_path):
    """Each header of include/ compiles on its own as C++11 with -Wall -Werror (a caller may include just one)."""
    import glob
    heads = sorted(glob.glob(os.path.join(ROOT, "include", "rpe", "*.h*"))) + [os.path.join(ROOT, "include", "rpe_c_api.h")]
    assert len(heads) >= 20
    for h in heads:
        src = tmp_path / "one.cpp"
        src.write_text(f'#include "{h}"\nint main() {{ return 0; }}\n')
        subprocess.run(["g++", "-std=c++11", "-O0", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", "-o",
                        str(tmp_path / "one.o"), str(src)], check=True)


def test_host_stopping_rule_equals_oracle_for_float_and_double(orc, tmp_path):
    """RANSACUpdateNumIters<float> / <double> of the header API (what a caller can evaluate on the host) are the rules the
    device replays; both equal the DET-mode oracle over a sweep of vote counts."""
    src = tmp_path / "rule.cpp"
    src.write_text('#include <cstdio>\n#include "rpe/P3P.hpp"\nint main() { for (int v = 1; v < 1000; ++v) { double ep = (1000.0 - v) / 1000.0; '
                   'printf("%d %d\\n", RANSACUpdateNumIters<double>(0.99, ep, 3, 100000), '
                   'RANSACUpdateNumIters<float>(0.99f, (float)ep, 4, 100000)); } return 0; }\n')
    exe = str(tmp_path / "rule")
    libdir = os.path.join(ROOT, "rgbd_pose_estimation_b200")
    subprocess.run(["g++", "-std=c++11", "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src), "-L", libdir,
                    "-lrpe_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    orc.set_math_mode(orc.DET)
    k = 0
    for v in range(1, 1000):
        ep = (1000.0 - v) / 1000.0
        assert int(out[k]) == orc.update_num_iters(0.99, ep, 3, 100000, dt=np.float64), v
        assert int(out[k + 1]) == orc.update_num_iters(np.float32(0.99), np.float32(ep), 4, 100000, dt=np.float32), v
        k += 2
    orc.set_math_mode(orc.LIBM)
