"""nekcem_b200: B200-native (sm_100a) Maxwell SEDG right-hand side + LSRK step of NekCEM.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of
``libnekcem_b200.so``), ``api`` (the host-side mirror of the reference's operator
interface) and ``boxcase`` (synthetic periodic-box inputs for bench/smoke)."""
from .api import MaxwellB200, NekcemB200Error, comm_unique_id, lib  # noqa: F401
