"""tests/cpp/dual_driver.cpp — ONE driver written in the reference's vocabulary only — builds against the reference's
own headers and against the drop-in headers without changing a line (CPU part; the GPU part is
tests/test_gpu_dual_driver.py). Here: both builds succeed, the drop-in build refuses to run without a GPU (no CPU
fallback), and the reference build's output is what the oracle computes on the same inputs and ::rand() seeds."""
import os
import subprocess

import numpy as np
import pytest

from tests import dual_driver_common as dd

F = 585.0
needs_reference = pytest.mark.skipif(not os.path.isdir(dd.REFERENCE), reason="/root/reference not present")


def test_dropin_build_links_and_fails_loudly_without_a_gpu(tmp_path, rpe):
    exe = str(tmp_path / "driver_b200")
    dd.build_dropin(exe)
    dd.write_input(rpe, str(tmp_path / "in.bin"))
    import ctypes
    lib = ctypes.CDLL(rpe.lib_path)
    if lib.rpe_device_count() > 0:
        pytest.skip("a GPU is present: the run itself is checked by tests/test_gpu_dual_driver.py")
    p = subprocess.run([exe, str(tmp_path / "in.bin")], capture_output=True, text=True)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr


@needs_reference
def test_reference_build_matches_the_oracle(tmp_path, rpe, orc):
    from tests import refshim
    assert refshim.available()
    exe = str(tmp_path / "driver_ref")
    dd.build_reference(exe)
    q, t, d = dd.write_input(rpe, str(tmp_path / "in.bin"))
    got = dd.parse(subprocess.run([exe, str(tmp_path / "in.bin")], capture_output=True, text=True, check=True).stdout)
    assert len(got) == 13
    arrs = {k: d[k] for k in ("bv", "xc", "nc", "xw", "nw")}
    n, w = dd.N, np.ascontiguousarray(d["weights"], np.float32)
    ct, cn = refshim.cos_thr(8.0, F), refshim.cos_nl(0.1)
    plan = [("shinji_prosac", 0, 101, w[1]), ("shinji_ransac2", 0, 102, None), ("kneip_prosac", 6, 103, w[0]),
            ("kneip_ransac", 1, 104, None), ("shinji_kneip_ransac", 2, 105, None), ("nl_kneip_ransac", 3, 106, None),
            ("nl_shinji_ransac", 4, 107, None), ("nl_shinji_kneip_ransac", 5, 108, None)]
    res = {}
    for name, method, seed, weights in plan:
        m = 3 if method == 0 else 4
        S = orc.prosac_table(seed, n, m, 1000, weights) if weights is not None else orc.sample_table(seed, n, m, 1000)
        r = orc.ransac(method, S, thr3d=0.2, cos_thr=ct, cos_nl=cn, confidence=0.99, full=False, **arrs)
        g = got[name]
        assert (g["max_votes"], g["iter"]) == (r["max_votes"], r["iter_final"]), name
        assert np.array_equal(g["q"], r["q"]) and np.array_equal(g["t"], r["t"]), name
        sums = [int(v) for v in r["mask"].sum(axis=1)]
        rows = {0: [None, 1], 1: [0], 6: [0], 2: [0, 1]}.get(method, [0, 1, 2])
        for lst in dd.FRESH_LISTS[name]:
            assert g["n_idx"][lst] == sums[rows.index(lst)], (name, lst)
        res[name] = r
    # refits
    for name, base in (("shinji_prosac+shinji_ls1", "shinji_prosac"), ("shinji_ransac2+shinji_ls1", "shinji_ransac2"),
                       ("shinji_kneip_ransac+shinji_ls", "shinji_kneip_ransac")):
        qq, tt, ok = orc.shinji_ls(arrs["xc"], arrs["xw"], res[base]["mask"][1])
        assert ok and np.array_equal(got[name]["q"], qq) and np.array_equal(got[name]["t"], tt), name
    r = res["nl_shinji_kneip_ransac"]
    q1, t1 = orc.nl_shinji_kneip_ls(r["q"], r["t"], r["mask"], r["max_votes"], **arrs)
    g = got["nl_shinji_kneip_ransac+nl_shinji_kneip_ls"]
    assert np.array_equal(g["q"], q1) and np.array_equal(g["t"], t1)
    q2, t2 = orc.nl_shinji_kneip_ls(q1, t1, r["mask"], r["max_votes"], weights3=w, **arrs)
    g = got["nl_shinji_kneip_ls(dynamic weights)"]
    assert np.array_equal(g["q"], q2) and np.array_equal(g["t"], t2)
    # and the committed golden file (what the GPU box compares the drop-in build with) is this very output
    import json
    gold = json.load(open(dd.GOLDEN))
    assert sorted(gold["cases"]) == sorted(got)
    for name, g in got.items():
        c = gold["cases"][name]
        assert (c["max_votes"], c["iter"], c["mask_hash"], c["n_idx"]) == (g["max_votes"], g["iter"], g["mask_hash"], g["n_idx"])
        assert c["q_hex"] == [float(v).hex() for v in g["q"]] and c["t_hex"] == [float(v).hex() for v in g["t"]]
