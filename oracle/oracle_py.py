"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes view of ``oracle/liboracle.so`` (the plain-C++ restatement of the reference's
linearization + LM path, see ``moptimizer_oracle.hpp``).  Imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

# model kinds / enums (mirror oracle_capi.cpp)
P2P, EXP_CURVE, MICHAELIS_MENTEN, PINHOLE, POWELL, POINT_DIST, PINHOLE_DISTORT, P2P_ICP = range(8)
LOSS_NONE, LOSS_GM, LOSS_HUBER = range(3)
JAC_ANALYTICAL, JAC_FORWARD, JAC_CENTRAL = range(3)
P2P_EXACT, P2P_REFTEST, P2P_REFTEST_COLMAJOR, P2P_LEFT = range(4)
MANIFOLD_ADDITIVE, MANIFOLD_SO3_LEFT = 0, 1
F32, F64 = 0, 1
STATUS = ["CONVERGED", "MAXIMUM_ITERATIONS_REACHED", "SMALL_DELTA", "NUMERIC_ERROR", "FATAL_ERROR"]


class _OrcCost(C.Structure):
    _fields_ = [
        ("model", C.c_int), ("variant", C.c_int),
        ("P", C.c_int), ("O", C.c_int), ("n", C.c_int),
        ("jac_mode", C.c_int), ("loss", C.c_int), ("loss_param", C.c_double),
        ("cov", C.POINTER(C.c_double)),
        ("a", C.c_void_p), ("b", C.c_void_p), ("data_f32", C.c_int),
        ("consts", C.POINTER(C.c_double)),
        ("cost_threads", C.c_int), ("float_carry", C.c_int), ("manifold", C.c_int),
        ("target", C.c_void_p), ("target_m", C.c_int), ("max_dist", C.c_double),
        ("update_x", C.POINTER(C.c_double)),
        ("lin_threads", C.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ only)."""
    src_newer = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("oracle_capi.cpp", "moptimizer_oracle.hpp", "Makefile"))
    if force or src_newer:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.orc_linearize.argtypes = [C.POINTER(_OrcCost), C.c_int, dp, dp, dp, dp, C.c_int]
        L.orc_compute_cost.argtypes = [C.POINTER(_OrcCost), C.c_int, dp, dp, C.c_int]
        L.orc_lm_minimize.argtypes = [C.POINTER(_OrcCost), C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_int, dp, ip, ip, dp, C.c_int, ip, C.c_int]
        L.orc_ldlt_solve.argtypes = [C.c_int, dp, dp, dp]
        L.orc_so3_convert6dof.argtypes = [dp, dp]
        L.orc_so3_left_jacobian_full.argtypes = [dp, dp]
        L.orc_time_linearize.argtypes = [C.POINTER(_OrcCost), C.c_int, dp, C.c_int, C.c_int, dp,
                                         dp, dp]
        L.orc_time_linearize.restype = C.c_double
        L.orc_hardware_concurrency.restype = C.c_int
        L.orc_generate_p2p.argtypes = [C.c_uint64, C.c_int64, C.c_int64, dp, dp, dp, C.c_double, C.c_double,
                                       C.c_double, dp, dp, C.c_int]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@dataclass
class Cost:
    """One cost term = model + data + loss + covariance (cost_function.h:15-59)."""
    model: int
    P: int
    O: int
    n: int
    a: Optional[np.ndarray] = None
    b: Optional[np.ndarray] = None
    jac_mode: int = JAC_FORWARD
    variant: int = P2P_EXACT
    loss: int = LOSS_NONE
    loss_param: float = 0.0
    cov: Optional[np.ndarray] = None      # O x O (symmetric; column-major == row-major)
    consts: Optional[np.ndarray] = None   # pinhole: K(3x4 row-major) ++ C(4x4 row-major)
    cost_threads: int = 1
    float_carry: bool = False
    manifold: int = 0
    target: Optional[np.ndarray] = None   # P2P_ICP: fixed target cloud (m, 3)
    max_dist: float = 0.0                 # P2P_ICP: maximum correspondence distance
    update_x: Optional[Sequence[float]] = None  # P2P_ICP: where cost->update(x) ran (default: evaluation point)
    lin_threads: int = 1                  # lm_minimize: threads of the linearization (1 = the reference's serial loop)
    _keep: list = field(default_factory=list, repr=False)

    def c_struct(self) -> _OrcCost:
        s = _OrcCost()
        s.model, s.variant, s.P, s.O, s.n = self.model, self.variant, self.P, self.O, self.n
        s.jac_mode, s.loss, s.loss_param = self.jac_mode, self.loss, float(self.loss_param)
        self._keep.clear()
        f32 = None
        for name in ("a", "b"):
            arr = getattr(self, name)
            if arr is None:
                setattr(s, name, None)
                continue
            if arr.dtype not in (np.float32, np.float64):
                arr = arr.astype(np.float64)
            arr = np.ascontiguousarray(arr)
            is32 = arr.dtype == np.float32
            if f32 is None:
                f32 = is32
            elif f32 != is32:
                arr = arr.astype(np.float32 if f32 else np.float64)
            self._keep.append(arr)
            setattr(s, name, arr.ctypes.data)
        s.data_f32 = 1 if f32 else 0
        if self.cov is not None:
            cv = np.asfortranarray(np.asarray(self.cov, dtype=np.float64).reshape(self.O, self.O))
            flat = np.ascontiguousarray(cv.T.reshape(-1))  # column-major flattening
            self._keep.append(flat)
            s.cov = _dp(flat)
        else:
            s.cov = None
        if self.consts is not None:
            k = np.ascontiguousarray(np.asarray(self.consts, dtype=np.float64).reshape(-1))
            self._keep.append(k)
            s.consts = _dp(k)
        else:
            s.consts = None
        s.cost_threads = int(self.cost_threads)
        s.float_carry = 1 if self.float_carry else 0
        s.manifold = int(self.manifold)
        s.lin_threads = int(self.lin_threads)
        s.target, s.target_m, s.max_dist, s.update_x = None, 0, float(self.max_dist), None
        if self.target is not None:
            t = np.ascontiguousarray(np.asarray(self.target, dtype=np.float32 if f32 else np.float64))
            self._keep.append(t)
            s.target, s.target_m = t.ctypes.data, t.shape[0]
        if self.update_x is not None:
            u = np.ascontiguousarray(np.asarray(self.update_x, dtype=np.float64))
            self._keep.append(u)
            s.update_x = _dp(u)
        return s


def linearize(cost: Cost, x: Sequence[float], scalar: int = F64, nthreads: int = 1):
    """-> (H[P,P], b[P], sum).  nthreads>1 selects the threaded (non-reference) variant."""
    P = cost.P
    xs = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    H = np.zeros((P, P), dtype=np.float64)
    b = np.zeros(P, dtype=np.float64)
    s = C.c_double(0)
    cs = cost.c_struct()
    rc = lib().orc_linearize(C.byref(cs), scalar, _dp(xs), _dp(H), _dp(b), C.byref(s), nthreads)
    if rc:
        raise RuntimeError("oracle: unsupported cost description")
    return H.T.copy(), b, s.value  # H is symmetric for symmetric C; .T undoes column-major


def compute_cost(cost: Cost, x: Sequence[float], scalar: int = F64, parallel: bool = True):
    xs = np.ascontiguousarray(np.asarray(x if len(x) else [0.0], dtype=np.float64))
    s = C.c_double(0)
    cs = cost.c_struct()
    rc = lib().orc_compute_cost(C.byref(cs), scalar, _dp(xs), C.byref(s), 1 if parallel else 0)
    if rc:
        raise RuntimeError("oracle: unsupported cost description")
    return s.value


@dataclass
class LmResult:
    x: np.ndarray
    status: str
    executed_iterations: int
    trace: np.ndarray  # rows {outer_it, k, y0, yi, rho, lambda, nu, accepted}

    @property
    def sequence(self) -> str:
        return "".join("A" if r[7] else "R" for r in self.trace)


def lm_minimize(costs: Sequence[Cost], x0: Sequence[float], max_iterations: int = 15,
                lm_iterations: int = 3, scalar: int = F64, stagnation_stop: bool = False) -> LmResult:
    """The reference's LM loop; `stagnation_stop` restates the product's MOPT_LM_STAGNATION_STOP extension."""
    P = costs[0].P
    arr = (_OrcCost * len(costs))(*[c.c_struct() for c in costs])
    x = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).copy())
    status, executed, ntrace = C.c_int(0), C.c_int(0), C.c_int(0)
    max_trace = max(1, max_iterations * max(1, lm_iterations))
    trace = np.zeros((max_trace, 8), dtype=np.float64)
    rc = lib().orc_lm_minimize(arr, len(costs), scalar, P, max_iterations, lm_iterations, _dp(x),
                               C.byref(status), C.byref(executed), _dp(trace), max_trace,
                               C.byref(ntrace), 1 if stagnation_stop else 0)
    if rc:
        raise RuntimeError("oracle: unsupported cost description")
    return LmResult(x, STATUS[status.value], executed.value, trace[: ntrace.value].copy())


def ldlt_solve(A: np.ndarray, rhs: np.ndarray) -> np.ndarray:
    n = A.shape[0]
    Af = np.ascontiguousarray(np.asarray(A, dtype=np.float64).T.reshape(-1))  # column-major
    r = np.ascontiguousarray(np.asarray(rhs, dtype=np.float64))
    out = np.zeros(n, dtype=np.float64)
    lib().orc_ldlt_solve(n, _dp(Af), _dp(r), _dp(out))
    return out


def so3_convert6dof(x) -> np.ndarray:
    xs = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    T = np.zeros(16, dtype=np.float64)
    lib().orc_so3_convert6dof(_dp(xs), _dp(T))
    return T.reshape(4, 4)


def so3_left_jacobian_full(w) -> np.ndarray:
    ws = np.ascontiguousarray(np.asarray(w, dtype=np.float64))
    J = np.zeros(9, dtype=np.float64)
    lib().orc_so3_left_jacobian_full(_dp(ws), _dp(J))
    return J.reshape(3, 3)


def time_linearize(cost: Cost, x, nthreads: int = 1, reps: int = 1):
    """-> (seconds, H, b, sum) for `reps` fp64 linearizations (model build excluded)."""
    P = cost.P
    xs = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    H = np.zeros((P, P), dtype=np.float64)
    b = np.zeros(P, dtype=np.float64)
    s = C.c_double(0)
    cs = cost.c_struct()
    t = lib().orc_time_linearize(C.byref(cs), F64, _dp(xs), nthreads, reps, _dp(H), _dp(b),
                                 C.byref(s))
    if t < 0:
        raise RuntimeError("oracle: unsupported cost description")
    return t, H.T.copy(), b, s.value


def hardware_concurrency() -> int:
    return int(lib().orc_hardware_concurrency())


def generate_p2p(seed: int, n: int, x_gt, lo=(0, 0, 0), hi=(10, 10, 10), first_index: int = 0, noise_sigma: float = 0.0,
                 outlier_fraction: float = 0.0, outlier_range: float = 0.0, nthreads: int = 0):
    """Host restatement of the product's synthetic point2point generator (csrc/mopt_store.cu): returns (src, tgt) as
    (n, 3) float64 arrays whose values are bit for bit the floats a device store generated with the same arguments
    holds."""
    src = np.empty((n, 3), dtype=np.float64)
    tgt = np.empty((n, 3), dtype=np.float64)
    g = np.ascontiguousarray(np.asarray(x_gt, dtype=np.float64))
    l = np.ascontiguousarray(np.asarray(lo, dtype=np.float64))
    h = np.ascontiguousarray(np.asarray(hi, dtype=np.float64))
    rc = lib().orc_generate_p2p(int(seed), int(first_index), int(n), _dp(g), _dp(l), _dp(h), float(noise_sigma),
                                float(outlier_fraction), float(outlier_range), _dp(src), _dp(tgt),
                                int(nthreads) if nthreads > 0 else max(1, hardware_concurrency()))
    if rc:
        raise RuntimeError("oracle: bad generator arguments")
    return src, tgt
