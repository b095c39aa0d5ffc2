"""ctypes binding of include/qmc4_b200.h (libqmc4b200.so).

Plumbing only: every call goes straight to the C ABI; there is no Python/CPU implementation of
any of it.  Importing this module fails loudly when the CUDA library has not been built.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QMC4_B200_LIB", os.path.join(_HERE, "libqmc4b200.so"))

# qmc_jastrow_kind / qmc_wavefunction_kind / qmc_potential_kind / qmc_estimator_kind
J_GAUSSIAN, J_GAUSSGAUSS, J_SQUAREWELL, J_HARDSPHEREGAUSS, J_BSPLINE, J_POSCHTELLER, J_HARDSPHERE, J_SQUAREWELL_BOUND = range(1, 9)
J_DELTA_PHONONS, J_DELTA_IN_TRAP, J_DELTA_BOUND_STATE_PHONONS = 9, 10, 11
J_DIPOLAR_REP = 12
WF_JASTROW1B, WF_JASTROW2B_UNDIS, WF_JASTROW2B_DIS, WF_JASTROW3B_UNDIS, WF_JASTROW3B_DIS = range(1, 6)
(POT_HARMONIC1B, POT_SQUAREWELL2B, POT_DIPOLAR2B, POT_POSCHTELLER2B, POT_TWOBODY_SQUAREWELL,
 POT_TWOBODY_BARRIER, POT_THREEBODY_UNDIS_SQUAREWELL, POT_THREEBODY_UNDIS_BARRIER, POT_THREEBODY_DIS_SQUAREWELL,
 POT_THREEBODY_DIS_BARRIER) = range(1, 11)
(EST_PAIRCORRELATION, EST_TRIMERCORRELATION_UNDIS, EST_STRUCTUREFACTOR, EST_TRIMERCORRELATION, EST_TRIMERCORRELATION_DIS,
 EST_TRIMERHYPERRADIUS, EST_CENTEROFMASSSQUARED) = range(1, 8)

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class BlockResult(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "energy_sum", "energy_count", "n_moves", "n_accepted", "population", "energy_shift",
        "steps_done", "overflow", "device_ms")]


class Trace(C.Structure):
    _fields_ = [(n, c_double_p) for n in (
        "xi", "Rprop", "Fprop", "logpsi", "lap", "V", "Eprop", "q", "u_metropolis", "u_branch")] + \
        [(n, c_int_p) for n in ("accepted", "valid", "n_desc", "parent")]


EXPORTS = [
    "qmc_last_error", "qmc_abi_version", "qmc_create", "qmc_destroy", "qmc_add_wavefunction_term",
    "qmc_add_potential_term", "qmc_add_estimator", "qmc_set_dmc", "qmc_eval", "qmc_eval_device",
    "qmc_accumulate_configs", "qmc_upload_walkers", "qmc_download_walkers", "qmc_step",
    "qmc_step_traced", "qmc_block_results", "qmc_get_estimator", "qmc_clear_estimators",
    "qmc_synchronize", "qmc_comm_unique_id", "qmc_comm_init", "qmc_rebalance", "qmc_set_migration", "qmc_plan_rebalance",
    "qmc_get_info", "qmc_stream", "qmc_add_forward_walking", "qmc_set_vmc", "qmc_tabulate_wavefunction",
    "qmc_add_superfluid_fraction", "qmc_get_vector_estimator", "qmc_enable_force_energy", "qmc_get_force_energy",
]
FW_INDEX_BASE = 65536
SFL_INDEX_BASE = 131072


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "qmc4_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    if hasattr(lib, "qmc_emulated") and os.environ.get("QMC4_B200_ALLOW_EMULATION") != "tests-only":
        # tests/emu builds the CUDA sources for a CPU thread emulator to test kernel logic without a
        # GPU; it is test infrastructure and must never be picked up as the product library
        raise ImportError("qmc4_b200: %s is the tests/emu emulation build, not the CUDA library "
                          "(there is no CPU fallback)" % LIB_PATH)
    vp = C.c_void_p
    lib.qmc_last_error.restype = C.c_char_p
    lib.qmc_last_error.argtypes = [vp]
    lib.qmc_create.argtypes = [C.POINTER(vp), C.c_int, c_double_p, C.c_int, c_int_p]
    lib.qmc_destroy.argtypes = [vp]
    lib.qmc_destroy.restype = None
    lib.qmc_add_wavefunction_term.argtypes = [vp, C.c_int, c_int_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_double]
    lib.qmc_add_potential_term.argtypes = [vp, C.c_int, c_int_p, C.c_int, c_double_p, C.c_int]
    lib.qmc_add_estimator.argtypes = [vp, C.c_int, c_int_p, C.c_int, C.c_int, C.c_double, C.c_double, c_double_p, C.c_int, c_int_p]
    lib.qmc_add_forward_walking.argtypes = [vp, C.c_int, C.c_int, c_int_p, C.c_int, c_int_p]
    lib.qmc_add_superfluid_fraction.argtypes = [vp, C.c_int, C.c_int, C.c_int, c_int_p]
    lib.qmc_get_vector_estimator.argtypes = [vp, C.c_int, c_double_p, c_double_p, C.c_int]
    lib.qmc_enable_force_energy.argtypes = [vp]
    lib.qmc_get_force_energy.argtypes = [vp, c_double_p, c_double_p]
    lib.qmc_set_dmc.argtypes = [vp, C.c_double, C.c_int64, C.c_int64, C.c_uint64, C.c_int, C.c_int64]
    lib.qmc_set_vmc.argtypes = [vp, C.c_double, C.c_int64, C.c_uint64, C.c_int64]
    lib.qmc_tabulate_wavefunction.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_int64, c_double_p]
    lib.qmc_eval.argtypes = [vp, c_double_p, C.c_int64] + [c_double_p] * 5 + [c_int_p]
    lib.qmc_eval_device.argtypes = [vp, vp, C.c_int64] + [vp] * 6
    lib.qmc_accumulate_configs.argtypes = [vp, c_double_p, C.c_int64]
    lib.qmc_upload_walkers.argtypes = [vp, c_double_p, C.c_int64]
    lib.qmc_download_walkers.argtypes = [vp, c_double_p, C.c_int64, C.POINTER(C.c_int64), c_double_p]
    lib.qmc_step.argtypes = [vp, C.c_int, C.c_int]
    lib.qmc_step_traced.argtypes = [vp, C.c_int, C.POINTER(Trace), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.qmc_block_results.argtypes = [vp, C.POINTER(BlockResult)]
    lib.qmc_get_estimator.argtypes = [vp, C.c_int, c_double_p, C.c_int, c_double_p]
    lib.qmc_clear_estimators.argtypes = [vp]
    lib.qmc_synchronize.argtypes = [vp]
    lib.qmc_comm_unique_id.argtypes = [vp]
    lib.qmc_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.qmc_rebalance.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.qmc_set_migration.argtypes = [vp, C.c_int, C.c_double]
    lib.qmc_plan_rebalance.argtypes = [c_int_p, C.c_int, c_int_p, c_int_p]
    lib.qmc_get_info.argtypes = [vp, C.c_char_p, c_double_p]
    lib.qmc_stream.argtypes = [vp]
    lib.qmc_stream.restype = vp
    return lib
