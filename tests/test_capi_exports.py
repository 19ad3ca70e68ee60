"""CPU-only: the C-ABI library builds, loads and exports every symbol include/mopt_capi.h declares.
No compute calls are made here (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from moptimizer_0_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mopt_capi.h")).read()
    return sorted(set(re.findall(r"MOPT_API\s+[\w\s\*]+?\b(mopt_\w+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "mopt_linearize" in syms and "mopt_lm_minimize" in syms and len(syms) >= 25


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"missing exports: {missing}"


def test_python_binding_lists_the_same_symbols():
    from moptimizer_0_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_struct_sizes_match_header(libpath):
    # sizes the C compiler gives the ABI structs vs the ctypes mirrors
    import subprocess, tempfile
    from moptimizer_0_b200 import capi
    src = r'''
#include <stdio.h>
#include "mopt_capi.h"
int main(void) { printf("%zu %zu %zu %zu %zu\n", sizeof(mopt_problem), sizeof(mopt_lm_options),
  sizeof(mopt_lm_trial), sizeof(mopt_lm_report), sizeof(mopt_synth)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(v) for v in out]
    assert sizes == [ctypes.sizeof(capi.Problem), ctypes.sizeof(capi.LmOptions), ctypes.sizeof(capi.LmTrial),
                     ctypes.sizeof(capi.LmReport), ctypes.sizeof(capi.Synth)]


def test_header_is_plain_c99_and_enums_match_the_python_binding(tmp_path):
    """The boundary is a C ABI: the header must compile as strict C99, and the enum values the ctypes binding
    hard-codes must be the header's."""
    import subprocess
    from moptimizer_0_b200 import capi
    names = ["MOPT_F32", "MOPT_F64", "MOPT_MODEL_POINT2POINT", "MOPT_MODEL_PINHOLE_DISTORT", "MOPT_JAC_CENTRAL",
             "MOPT_P2P_LEFT", "MOPT_MANIFOLD_SO3_LEFT", "MOPT_LOSS_HUBER", "MOPT_FLAG_GENERIC_KERNEL",
             "MOPT_FLAG_STABLE_FD", "MOPT_MAX_PARAMETERS", "MOPT_MAX_OUTPUTS", "MOPT_MAX_COSTS", "MOPT_MAX_TRACE",
             "MOPT_NCCL_ID_BYTES"]
    c = tmp_path / "e.c"
    c.write_text('#include <stdio.h>\n#include "mopt_capi.h"\nint main(void) { printf("' + " ".join(["%d"] * len(names)) +
                 '\\n", ' + ", ".join("(int)" + n for n in names) + "); return 0; }\n")
    exe = str(tmp_path / "e")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(c), "-o", exe], check=True)
    vals = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert vals == [capi.F32, capi.F64, capi.MODEL_POINT2POINT, capi.MODEL_PINHOLE_DISTORT, capi.JAC_CENTRAL,
                    capi.P2P_LEFT, capi.MANIFOLD_SO3_LEFT, capi.LOSS_HUBER, capi.FLAG_GENERIC_KERNEL,
                    capi.FLAG_STABLE_FD, capi.MAX_PARAMETERS, capi.MAX_OUTPUTS, capi.MAX_COSTS, capi.MAX_TRACE,
                    capi.NCCL_ID_BYTES]


def test_no_cuda_device_is_a_loud_error(libpath):
    from moptimizer_0_b200 import capi
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.MoptError):
        capi.Context(0)
