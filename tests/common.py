"""Shared fixtures/helpers for the test-suite: golden data of the reference's own tests
(tests/golden/, extracted by tests/golden/make_fixtures.py) and oracle cost builders."""
import json
import os

import numpy as np

from oracle import oracle_py as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLDEN, "reference_fixtures.json")) as _f:
    FX = json.load(_f)


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def fachada():
    """src cloud + tgt = T_gt * src, as tst/point2point.cpp:87-103 builds them (fp64)."""
    src = np.fromfile(os.path.join(GOLDEN, "fachada_xyz.f64"), dtype="<f8").reshape(-1, 3)
    ex, ey, ez = FX["fachada"]["gt_euler_xyz"]
    R = rot_x(ex) @ rot_y(ey) @ rot_z(ez)
    t = np.array(FX["fachada"]["gt_translation"])
    tgt = src @ R.T + t
    return np.ascontiguousarray(src), np.ascontiguousarray(tgt), R, t


def camera_consts():
    """K (3x4 row-major) ++ C (4x4 row-major), tst/camera_calibration.cpp:22-30."""
    K = np.array(FX["camera"]["K"], dtype=np.float64)
    Cm = np.eye(4)
    Cm[:3, :3] = rot_x(np.pi / 2) @ rot_z(np.pi / 2)
    return np.concatenate([K, Cm.reshape(-1)])


def curve_cost(lo=0, hi=67, **kw):
    t = np.array(FX["curve"]["t"][lo:hi], dtype=np.float64)
    y = np.array(FX["curve"]["y"][lo:hi], dtype=np.float64)
    return orc.Cost(orc.EXP_CURVE, 2, 1, len(t), a=t, b=y, **kw)


def camera_cost(**kw):
    pts = np.array(FX["camera"]["points"], dtype=np.float64)
    pix = np.array(FX["camera"]["pixels"], dtype=np.float64)
    return orc.Cost(orc.PINHOLE, 6, 2, 5, a=pts, b=pix, consts=camera_consts(), **kw)


def mm_cost(n=7, dtype=np.float32, **kw):
    key = "7" if n == 7 else "9"
    t = np.array(FX["michaelis_menten"]["t" + key], dtype=dtype)
    y = np.array(FX["michaelis_menten"]["y" + key], dtype=dtype)
    return orc.Cost(orc.MICHAELIS_MENTEN, 2, 1, n, a=t, b=y, **kw)


def powell_cost(**kw):
    return orc.Cost(orc.POWELL, 4, 4, 1, **kw)


def rel_err(a, b):
    """max|a-b| / max|b| — the relative measure used for H and b throughout the suite."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))
