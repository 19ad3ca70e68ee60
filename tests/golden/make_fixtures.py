#!/usr/bin/env python
"""Extract the DATA (not code) the reference's own tests pin this path with.

Run in the build container, where /root/reference is mounted:
    python tests/golden/make_fixtures.py
Writes small fixtures next to this script; they are committed so the GPU box (which has no
/root/reference) and the CPU test suite can use them.

Sources (all under /root/reference):
  tst/curve_fitting.cpp:9-79          67 (t, y) samples "From Ceres"; expected optimum :116-117
  tst/camera_calibration.cpp:29-30    pinhole intrinsics; :66-76 five LiDAR/pixel pairs;
                                      :91-98 matlab / ceres solutions
  tst/simple_model.cpp:24-25          Michaelis-Menten 7 samples (also covariance.cpp:9-10,
                                      loss_function.cpp:41-42); differentiation.cpp:48-49 9 samples
  tst/point2point.cpp:93-101          ground-truth transform of the fachada test
  tst/data/fachada.txt                29 310-point cloud (x y z r g b), md5 747d84d2...
"""
import hashlib
import json
import os
import re

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def _floats(text):
    return [float(t) for t in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", text)]


def curve():
    src = _read("tst/curve_fitting.cpp")
    body = re.search(r"const double data\[\] = \{(.*?)\};", src, re.S).group(1)
    vals = _floats(body)
    assert len(vals) == 134, len(vals)
    return {"t": vals[0::2], "y": vals[1::2], "expected": [0.291861, 0.131439],
            "tol_ic1": 5e-5, "tol_ic2": 1e-4, "x0_ic2": [1.2, 2.0], "iters_ic2": 50}


def camera():
    src = _read("tst/camera_calibration.cpp")
    K = _floats(re.search(r"camera_model <<(.*?);", src, re.S).group(1))
    assert len(K) == 12
    pts = [_floats(m)[:3] for m in re.findall(r"Eigen::Vector4d\(([^)]*)\)\);", src)]
    pix = [_floats(m) for m in re.findall(r"Eigen::Vector2i\(([^)]*)\)\);", src)]
    assert len(pts) == 5 and len(pix) == 5
    ceres = _floats(re.search(r"ceres_solution\[MODEL_PARAMETERS\] = \{(.*?)\};", src, re.S).group(1))
    matlab = _floats(re.search(r"matlab_solution\[MODEL_PARAMETERS\] = \{(.*?)\};", src, re.S).group(1))
    return {"K": K, "points": pts, "pixels": pix, "ceres_solution": ceres,
            "matlab_solution": matlab, "tol": 5e-5, "bad_x0": [0.5, 0.5, 0.5, 0.2, 0.5, 0.5],
            "bad_iters": 50,
            "frame_conversion": "Rx(pi/2) * Rz(pi/2)  (tst/camera_calibration.cpp:24-27)"}


def michaelis_menten():
    s = _read("tst/simple_model.cpp")
    x = _floats(re.search(r"x_data\[7\] = \{(.*?)\};", s).group(1))
    y = _floats(re.search(r"y_data\[7\] = \{(.*?)\};", s).group(1))
    d = _read("tst/differentiation.cpp")
    x9 = _floats(re.search(r"x_data\[\] = \{(.*?)\};", d).group(1))
    y9 = _floats(re.search(r"y_data\[\] = \{(.*?)\};", d).group(1))
    assert len(x) == 7 and len(y) == 7 and len(x9) == 9 and len(y9) == 9
    return {"t7": x, "y7": y, "t9": x9, "y9": y9, "expected": [0.362, 0.556], "tol": 0.01,
            "x0_a": [0.9, 0.2], "x0_b": [1.9, 1.5], "gm_threshold": 100.0}


def fachada():
    path = os.path.join(REF, "tst/data/fachada.txt")
    raw = open(path, "rb").read()
    md5 = hashlib.md5(raw).hexdigest()
    assert md5 == "747d84d2a7f92db4f727414699b69e07", md5
    arr = np.loadtxt(path, dtype=np.float64)
    assert arr.shape == (29310, 6)
    # raw little-endian float64, n x 3 row-major: readable by numpy.fromfile and by fread in tests/cpp
    np.ascontiguousarray(arr[:, :3], dtype="<f8").tofile(os.path.join(OUT, "fachada_xyz.f64"))
    return {"n": 29310, "md5_txt": md5,
            "gt_euler_xyz": [0.3, 0.4, 0.5], "gt_translation": [10.5, 10.2, 0.1],
            "note": "tgt = T * src with R = Rx(.3) Ry(.4) Rz(.5) (tst/point2point.cpp:93-103)"}


def main():
    fx = {"curve": curve(), "camera": camera(), "michaelis_menten": michaelis_menten(),
          "fachada": fachada(),
          "powell": {"x0": [3, -1, 0, 4], "iters": 25, "tol": 5e-5, "cov_scale": 0.01}}
    with open(os.path.join(OUT, "reference_fixtures.json"), "w") as f:
        json.dump(fx, f, indent=1)
    # the same numbers as "name count v0 v1 ..." lines for the C++ tests (tests/cpp/fixtures.h)
    flat = {
        "curve_t": fx["curve"]["t"], "curve_y": fx["curve"]["y"], "curve_expected": fx["curve"]["expected"],
        "camera_K": fx["camera"]["K"], "camera_points": sum(fx["camera"]["points"], []),
        "camera_pixels": sum(fx["camera"]["pixels"], []), "camera_ceres": fx["camera"]["ceres_solution"],
        "camera_bad_x0": fx["camera"]["bad_x0"],
        "mm_t7": fx["michaelis_menten"]["t7"], "mm_y7": fx["michaelis_menten"]["y7"],
        "mm_t9": fx["michaelis_menten"]["t9"], "mm_y9": fx["michaelis_menten"]["y9"],
        "mm_expected": fx["michaelis_menten"]["expected"],
        "fachada_gt_euler": fx["fachada"]["gt_euler_xyz"], "fachada_gt_t": fx["fachada"]["gt_translation"],
    }
    with open(os.path.join(OUT, "reference_fixtures.txt"), "w") as f:
        for k, v in flat.items():
            f.write(k + " " + str(len(v)) + " " + " ".join(repr(float(x)) for x in v) + "\n")
    print("wrote reference_fixtures.json/.txt and fachada_xyz.f64 in", OUT)


if __name__ == "__main__":
    main()
