"""ginkgo_b200: B200-native (sm_100a) SpMV + Krylov hot path behind Ginkgo's
Executor / LinOp / solver-factory API.  See DESIGN.md."""
from ._lib import B200Error, lib, call, check  # noqa: F401
