"""moptimizer_0_b200 — B200-native (sm_100a) linearization + Levenberg-Marquardt hot path of
Marcus-Forte/moptimizer_0, behind a C ABI (include/mopt_capi.h).  See DESIGN.md."""
from . import capi  # noqa: F401

__all__ = ["capi"]
