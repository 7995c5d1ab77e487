"""ctypes binding of release/libputranse.so (the C-ABI declared in include/putranse.h).

The reference loads its native core the same way: ``ctypes.cdll.LoadLibrary("../release/Base.so")``
(reference openke/data/TrainDataLoader.py:30-31, openke/config/Tester.py:20-21).  There is exactly
one library and no Python/NumPy implementation behind it: if the shared object is missing, or a
compute entry point is called without a CUDA device, this module raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.abspath(os.path.join(_HERE, "..", "release", "libputranse.so"))

PK_TRANSE, PK_TRANSH, PK_TRANSD = 0, 1, 2
PK_SGD, PK_ADAGRAD = 0, 1

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
c_f32p = ctypes.POINTER(ctypes.c_float)
vp = ctypes.c_void_p


class ModelCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("model", "dim", "p_norm", "norm_flag", "opt", "neg_ent", "bern", "filter", "work_threads", "reserved")]


class Tables(ctypes.Structure):
    _fields_ = [("ent", vp * 2), ("rel", vp * 2), ("ent_state", vp * 2), ("rel_state", vp * 2),
                ("n_ent", ctypes.c_int64), ("n_rel", ctypes.c_int64)]


class Sampler(ctypes.Structure):
    _fields_ = [("by_head", vp), ("by_tail", vp), ("left_mean", vp), ("right_mean", vp), ("lcg", vp),
                ("n_tri", ctypes.c_int64), ("n_ent", ctypes.c_int64), ("n_rel", ctypes.c_int64),
                ("head_off", vp), ("tail_off", vp)]


class UniverseDesc(ctypes.Structure):
    _fields_ = [("tri_off", ctypes.c_int64), ("ent_off", ctypes.c_int64), ("rel_off", ctypes.c_int64),
                ("loss_off", ctypes.c_int64), ("n_tri", ctypes.c_int32), ("n_ent", ctypes.c_int32),
                ("n_rel", ctypes.c_int32), ("batch_size", ctypes.c_int32), ("nbatches", ctypes.c_int32),
                ("epochs", ctypes.c_int32), ("margin", ctypes.c_float), ("lr", ctypes.c_float),
                ("lcg", ctypes.c_uint64 * 8)]


# the same 128-byte record as a numpy dtype: descriptor arrays are filled column-wise, not per universe
UNIVERSE_DESC_DTYPE = np.dtype([("tri_off", np.int64), ("ent_off", np.int64), ("rel_off", np.int64), ("loss_off", np.int64),
                                ("n_tri", np.int32), ("n_ent", np.int32), ("n_rel", np.int32), ("batch_size", np.int32),
                                ("nbatches", np.int32), ("epochs", np.int32), ("margin", np.float32), ("lr", np.float32),
                                ("lcg", np.uint64, (8,))])
assert UNIVERSE_DESC_DTYPE.itemsize == ctypes.sizeof(UniverseDesc) == 128


class EnergyItem(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("key_row", "universe", "fixed_local", "rel_local", "side", "reserved")]


ENERGY_ITEM_DTYPE = np.dtype([(n, np.int32) for n in ("key_row", "universe", "fixed_local", "rel_local", "side", "reserved")])


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """The process-wide handle (like the reference, every loader/tester shares one library state)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                "libputranse.so is not built (%s). Run `make -C openke-putranse_b200 -j` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`. There is no Python fallback." % LIB_PATH)
        L = ctypes.cdll.LoadLibrary(LIB_PATH)
        _declare(L)
        _lib = L
    return _lib


def _declare(L):
    I, F, B = ctypes.c_int64, ctypes.c_float, ctypes.c_int64  # the reference passes bools as int64 too
    sig = {
        # reference-compatible surface
        "setInPath": (None, [ctypes.c_char_p]), "setOutPath": (None, [ctypes.c_char_p]),
        "setWorkThreads": (None, [I]), "getWorkThreads": (I, []), "setBern": (None, [I]),
        "getEntityTotal": (I, []), "getRelationTotal": (I, []), "getTripleTotal": (I, []),
        "getTrainTotal": (I, []), "getTestTotal": (I, []), "getValidTotal": (I, []),
        "setRandomSeed": (None, [I]), "getRandomSeed": (I, []), "randReset": (None, []),
        "importTrainFiles": (None, []), "importTestFiles": (None, []),
        "sampling": (None, [vp, vp, vp, vp, I, I, I, I, B, B, B]),
        "getParallelUniverse": (None, [I, F]),
        "getEntityTotalUniverse": (I, []), "getRelationTotalUniverse": (I, []), "getTrainTotalUniverse": (I, []),
        "getEntityRemapping": (None, [vp]), "getRelationRemapping": (None, [vp]),
        "swapHelpers": (None, []), "resetUniverse": (None, []),
        "initTest": (None, []), "getHeadBatch": (None, [vp, vp, vp]), "getTailBatch": (None, [vp, vp, vp]),
        "validInit": (None, []), "getValidHeadBatch": (None, [vp, vp, vp]), "getValidTailBatch": (None, [vp, vp, vp]),
        "testHead": (None, [vp, I, B]), "testTail": (None, [vp, I, B]),
        "validHead": (None, [vp, I]), "validTail": (None, [vp, I]),
        "test_link_prediction": (None, [B]),
        "getTestLinkMRR": (F, [B]), "getTestLinkMR": (F, [B]), "getTestLinkHit10": (F, [B]),
        "getTestLinkHit3": (F, [B]), "getTestLinkHit1": (F, [B]), "getValidHit10": (F, []),
        "getNumOfNegatives": (I, [I, I, B]), "getNumOfPositives": (I, [I, I, B]),
        "getNegativeEntities": (None, [vp, I, I, B]), "getPositiveEntities": (None, [vp, I, I, B]),
        "getNumOfEntityRelations": (I, [I, B]), "getEntityRelations": (None, [vp, I, B]),
        "activateIncrementalSetting": (None, []), "initializeIncrementalSetting": (None, []), "setNumSnapshots": (None, [I]),
        "getNumSnapshots": (I, []), "setNumOperationsRate": (None, [I]), "readGlobalNumEntities": (None, []),
        "readGlobalNumRelations": (None, []), "initializeTrainingOperations": (None, [ctypes.c_int]), "evolveTrainList": (None, []),
        "loadSnapshotTriples": (None, [ctypes.c_int]), "loadTestData": (None, [ctypes.c_int]), "loadValidData": (None, [ctypes.c_int]),
        "getNumCurrentlyContainedEntities": (I, []), "initializeTripleOperations": (None, [ctypes.c_int]), "evolveTripleList": (None, []), "pk_incremental_reset": (ctypes.c_int, []),
        "pk_incremental_list": (I, [ctypes.c_int, vp]),
        "activateLoadOfAllTriples": (None, [B]), "getNegTest": (None, []), "getTestBatch": (None, [vp] * 6),
        # pk_* host
        "pk_last_error": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
        "pk_cuda_device_count": (ctypes.c_int, []), "pk_version": (ctypes.c_char_p, []),
        "pk_last_launch_count": (ctypes.c_int, []), "pk_import_count": (ctypes.c_int, []),
        "pk_train_index": (ctypes.c_int, [vp, vp, vp, vp]),
        "pk_get_lcg": (ctypes.c_int, [vp, ctypes.c_int]), "pk_set_lcg": (ctypes.c_int, [vp, ctypes.c_int]),
        "pk_f32_running_sum": (ctypes.c_float, [vp, I]),
        "pk_universe_triples": (ctypes.c_int, [vp]),
        "pk_eval_triples": (ctypes.c_int, [ctypes.c_int, vp]),
        "pk_filter_csr": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, vp, vp, c_i64p]),
        "pk_universes_build": (vp, [ctypes.c_int, vp, vp, vp, ctypes.c_int]),
        "pk_universes_build_lean": (vp, [ctypes.c_int, vp, vp, vp, ctypes.c_int]),
        "pk_universes_free": (None, [vp]), "pk_universes_count": (ctypes.c_int, [vp]),
        "pk_universes_sizes": (ctypes.c_int, [vp, vp, vp, vp, vp]),
        "pk_universes_export": (ctypes.c_int, [vp] * 9),
        "pk_python_hyper_draws": (ctypes.c_int, [ctypes.c_int, vp, I, I, ctypes.c_double, ctypes.c_double, I, I, I, I, ctypes.c_int,
                                                 ctypes.c_double, ctypes.c_double, ctypes.c_int, vp, vp, vp, vp, vp]),
        "pk_walk_cap": (ctypes.c_int, []), "pk_walk_device_check": (ctypes.c_int, []),
        "pk_walk_scratch_bytes": (ctypes.c_int, [ctypes.c_int, vp]),
        "pk_universes_walk_device": (ctypes.c_int, [ctypes.c_int] + [vp] * 13),
        "pk_walk_pack_remaps": (ctypes.c_int, [ctypes.c_int] + [vp] * 6),
        "pk_torch_init_tables": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int]),
        "pk_init_tables_device": (ctypes.c_int, [ctypes.c_int, vp, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int, vp]),
        # pk_* device
        "pk_sample_batch": (ctypes.c_int, [ctypes.POINTER(ModelCfg), ctypes.POINTER(Sampler), I, vp, vp, vp, vp]),
        "pk_workspace_create": (vp, [ctypes.POINTER(ModelCfg), I, I, I]),
        "pk_workspace_free": (None, [vp]), "pk_workspace_check": (ctypes.c_int, [vp, vp]),
        "pk_train_step": (ctypes.c_int, [ctypes.POINTER(ModelCfg), ctypes.POINTER(Tables), vp, I, vp, vp, vp, F, F, vp, vp]),
        "pk_train_steps": (ctypes.c_int, [ctypes.POINTER(ModelCfg), ctypes.POINTER(Tables), ctypes.POINTER(Sampler), vp, I, I, F, F, vp, vp]),
        "pk_train_universes": (ctypes.c_int, [ctypes.POINTER(ModelCfg), ctypes.POINTER(Tables), vp, vp, vp, vp, vp, ctypes.c_int, vp, vp]),
        "pk_universe_kernel_class": (ctypes.c_int, [ctypes.POINTER(ModelCfg), I, I, I]),
        "pk_debug_universe_timer": (ctypes.c_int, [ctypes.c_void_p]),
        "pk_rank_space": (ctypes.c_int, [ctypes.POINTER(ModelCfg), ctypes.POINTER(Tables), I, vp, vp, vp, vp, vp, vp, vp]),
        "pk_universe_energies": (ctypes.c_int, [ctypes.POINTER(ModelCfg), ctypes.POINTER(Tables), vp, vp, vp, vp, vp, I, vp, I, vp]),
        "pk_universe_tuple_scores": (ctypes.c_int, [ctypes.POINTER(ModelCfg), ctypes.POINTER(Tables), vp, vp, vp, I, vp, vp]),
        "pk_fill_missing_energies": (ctypes.c_int, [vp, I, I, vp, vp]),
        "pk_rank_from_energy": (ctypes.c_int, [vp, I, I, vp, vp, vp, vp, vp, vp]),
        "pk_rank_from_energy_masked": (ctypes.c_int, [vp, I, I, vp, vp, vp, vp, vp, vp, I, vp]),
        "pk_rank_candidate_row": (ctypes.c_int, [vp, I, vp, vp, vp, vp, vp]),
        "pk_score_batch": (ctypes.c_int, [ctypes.POINTER(ModelCfg), ctypes.POINTER(Tables), vp, I, vp, I, vp, I, ctypes.c_int, vp, vp, vp]),
        "pk_fill_inf": (ctypes.c_int, [vp, I, vp]),
        "pk_normalise_rows": (ctypes.c_int, [vp, vp, I, ctypes.c_int, vp]),
        "pk_selftest_arith": (ctypes.c_int, [I, ctypes.c_uint32, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here == the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    L._pk_symbols = sorted(sig)


def last_error():
    buf = ctypes.create_string_buffer(1024)
    lib().pk_last_error(buf, 1024)
    return buf.value.decode(errors="replace")


def check(rc, what=""):
    if rc != 0:
        raise NativeError("%s failed (%d): %s" % (what or "libputranse call", rc, last_error()))


def require_cuda():
    n = lib().pk_cuda_device_count()
    if n <= 0:
        raise NativeError("no CUDA device: the PuTransE hot path has no CPU implementation (%s)" % last_error())
    return n


def addr(a):
    """Raw address of a numpy array, as the reference passes them (``__array_interface__['data'][0]``)."""
    return a.__array_interface__["data"][0]
