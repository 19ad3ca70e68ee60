"""The C++ mirror of the reference API (include/moptimizer/*.h over the C ABI): builds the C++ test binary
that re-states the reference's own tests (tests/cpp/reference_tests.cpp) and, on a GPU, runs it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "run_tests")


def build_cpp_tests():
    from moptimizer_0_b200 import build
    return build.build_cpp_tests()


def test_cpp_mirror_compiles_and_links():
    exe = build_cpp_tests()
    assert os.path.exists(exe)


def test_cpp_binary_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = build_cpp_tests()
    r = subprocess.run([exe], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 77 and "no usable CUDA device" in r.stdout


@pytest.mark.gpu
def test_cpp_reference_tests_pass_on_gpu():
    exe = EXE if os.path.exists(EXE) else build_cpp_tests()
    r = subprocess.run([exe], capture_output=True, text=True, cwd=ROOT, timeout=600)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " 0 failed" in r.stdout


def test_so3_mirror_host_only(tmp_path):
    """include/moptimizer/so3.h (mirror of src/so3.cpp:7-155) against closed forms and Exp/Log round trips; pure host."""
    exe = str(tmp_path / "so3_host_test")
    src = os.path.join(ROOT, "tests", "cpp", "so3_host_test.cpp")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe],
                   check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout


def test_eigen_typed_overloads_host_only(tmp_path):
    """The reference's Eigen-typed signatures of so3::* and isDeltaSmall (include/moptimizer/so3.h:8-41, delta.h:11-16)
    compile in the reference's own call forms (MOPTIMIZER_USE_EIGEN) and forward to the raw-array functions.  The
    image has no Eigen: tests/cpp/eigen_stub/ provides the minimal <Eigen/Dense> surface they touch."""
    exe = str(tmp_path / "eigen_overloads_test")
    src = os.path.join(ROOT, "tests", "cpp", "eigen_overloads_test.cpp")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-DMOPTIMIZER_USE_EIGEN", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "tests", "cpp", "eigen_stub"), src, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "eigen overloads ok" in r.stdout, r.stdout


def test_affine_fd_model_hooks_host_only(tmp_path):
    """The camera models' AFFINE_FD hooks (csrc/mopt_models.cuh: affine / finish_diff / tail_partials) against the
    literal difference quotient of linearization.h:97-111 taken in long double; nvcc-compiled, runs on the CPU."""
    exe = str(tmp_path / "affine_fd_host_test")
    src = os.path.join(ROOT, "tests", "cpp", "affine_fd_host_test.cu")
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "moptimizer_0_b200", "csrc"),
                    src, "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout


def _build_example(tmp_path):
    exe = str(tmp_path / "p2p_example")
    src = os.path.join(ROOT, "examples", "point2point_registration.cpp")
    pkg = os.path.join(ROOT, "moptimizer_0_b200")
    build_cpp_tests()  # makes sure libmopt_b200.so exists
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-L", pkg,
                    "-lmopt_b200", "-Wl,-rpath," + pkg, "-o", exe], check=True)
    return exe


def test_public_headers_are_self_contained(tmp_path):
    """Every header under include/ compiles on its own (-Wall -Werror): a user of the reference includes them one by one."""
    inc = os.path.join(ROOT, "include")
    for d, _, files in os.walk(inc):
        for f in files:
            if f.endswith(".h"):
                rel = os.path.relpath(os.path.join(d, f), inc)
                tu = tmp_path / "hdr.cpp"
                tu.write_text("#include <%s>\nint main() { return 0; }\n" % rel)
                r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(tu)],
                                   capture_output=True, text=True)
                assert r.returncode == 0, rel + "\n" + r.stderr[:2000]


def test_integration_example_compiles_and_fails_loudly_without_gpu(tmp_path):
    """examples/point2point_registration.cpp is the INTEGRATION.md snippet as a program."""
    import torch
    exe = _build_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([exe, "1000"], capture_output=True, text=True)
    assert r.returncode == 77 and "moptimizer::Exception" in r.stdout


@pytest.mark.gpu
def test_integration_example_runs_on_gpu(tmp_path):
    exe = _build_example(tmp_path)
    r = subprocess.run([exe, "200000"], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
