"""qmc4_b200 — B200-native (sm_100a) diffusion Monte Carlo walker propagation behind a C ABI.

The compute path is the CUDA library qmc4_b200/libqmc4b200.so (sources in qmc4_b200/csrc/,
declared in include/qmc4_b200.h).  This Python package is plumbing for tests and bench.py; it has
no CPU implementation of anything and fails loudly if the CUDA library is missing."""
from . import capi  # noqa: F401
from .system import System, create_id, plan_rebalance  # noqa: F401
