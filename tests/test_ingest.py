"""Point-cloud ingest, CPU-only: the multi-threaded text loader must read exactly what the
reference's `while (file >> x >> y >> z >> r >> g >> b)` loop (tst/point2point.cpp:125-138) reads — here
checked against numpy's correctly-rounded parser and the committed fachada fixture — including its behaviour on
ragged / malformed input."""
import os

import numpy as np
import pytest

from tests.common import GOLDEN


@pytest.fixture(scope="module")
def capi():
    from moptimizer_0_b200 import build, capi as c
    build.build()
    return c


def write_cloud(path, xyz, rgb=None, fmt="%.8f"):
    rgb = np.zeros((xyz.shape[0], 3), dtype=int) if rgb is None else rgb
    with open(path, "w") as f:
        for p, c in zip(xyz, rgb):
            f.write(" ".join(fmt % v for v in p) + " %d %d %d\n" % tuple(c))


def test_fachada_text_roundtrip(capi, tmp_path):
    src = np.fromfile(os.path.join(GOLDEN, "fachada_xyz.f64"), dtype="<f8").reshape(-1, 3)
    path = str(tmp_path / "fachada.txt")
    rng = np.random.default_rng(0)
    write_cloud(path, src, rng.integers(0, 256, src.shape))  # same "x y z r g b" layout / 8 decimals as the fixture
    got = capi.cloud_read_text(path)
    assert got.shape == (29310, 3)
    assert np.array_equal(got, src)                         # decimal -> double is correctly rounded, bit for bit
    assert np.array_equal(got, np.loadtxt(path)[:, :3])
    got32 = capi.cloud_read_text(path, dtype=np.float32)
    assert got32.dtype == np.float32 and np.array_equal(got32, src.astype(np.float32))
    six = capi.cloud_read_text(path, keep=6)
    assert six.shape == (29310, 6) and np.array_equal(six[:, :3], src)


def test_large_file_parses_in_parallel_chunks(capi, tmp_path):
    rng = np.random.default_rng(1)
    xyz = rng.normal(scale=50.0, size=(200_000, 3))
    path = str(tmp_path / "big.txt")
    write_cloud(path, xyz, fmt="%.17g")
    got = capi.cloud_read_text(path)
    assert np.array_equal(got, xyz)


def test_ragged_and_malformed_input_matches_stream_semantics(capi, tmp_path):
    p = str(tmp_path / "ragged.txt")
    with open(p, "w") as f:
        f.write("1 2 3 0 0 0\n  4.5\t-5e-1 +6 1 1 1   \n\n7 8 9 2 2 2\n10 11\n")   # last record is short: dropped
    got = capi.cloud_read_text(p)
    assert np.array_equal(got, [[1, 2, 3], [4.5, -0.5, 6], [7, 8, 9]])
    with open(p, "w") as f:
        f.write("1 2 3 0 0 0\n4 5 x 0 0 0\n7 8 9 0 0 0\n")                          # extraction fails at 'x': stop
    assert np.array_equal(capi.cloud_read_text(p), [[1, 2, 3]])
    with open(p, "w") as f:
        f.write("")
    assert capi.cloud_read_text(p).shape == (0, 3)
    with pytest.raises(capi.MoptError, match="not a file"):
        capi.cloud_read_text(str(tmp_path / "missing.txt"))
    with pytest.raises(capi.MoptError, match="not a file"):  # a directory opens with "rb": it must not be sized and read
        capi.cloud_read_text(str(tmp_path))


def test_binary_cache_roundtrip(capi, tmp_path):
    rng = np.random.default_rng(2)
    for dtype in (np.float32, np.float64):
        a = rng.normal(size=(1234, 3)).astype(dtype)
        p = str(tmp_path / "c.bin")
        capi.cloud_write_binary(p, a)
        b = capi.cloud_read_binary(p)
        assert b.dtype == dtype and np.array_equal(a, b)
    with open(str(tmp_path / "bad.bin"), "wb") as f:
        f.write(b"not a cloud")
    with pytest.raises(capi.MoptError, match="MOPTCLD1"):
        capi.cloud_read_binary(str(tmp_path / "bad.bin"))
