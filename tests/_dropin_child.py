"""Child process of tests/test_dropin.py: the UNMODIFIED reference Python package
(/root/reference/openke/{data,config,module}, imported through a scratch directory of symlinks)
bound to libputranse.so in place of its own release/Base.so.

usage: _dropin_child.py <scratch> <libputranse.so> <reference Base.so> <wn18 dir> <golden dir>

Everything here is host work (construction, index queries, universe sampling, candidate batches):
it runs without a GPU.  The reference's Base.so is loaded beside it, as the checker.
"""
import ctypes
import os
import sys

import numpy as np

scratch, ours, ref_so, wn18, golden = sys.argv[1:6]
REF = "/root/reference"

pkg = os.path.join(scratch, "openke")
os.makedirs(os.path.join(pkg, "release"), exist_ok=True)
for name in ("__init__.py", "base", "config", "data", "module"):
    dst = os.path.join(pkg, name)
    if not os.path.lexists(dst):
        os.symlink(os.path.join(REF, "openke", name), dst)
so = os.path.join(pkg, "release", "Base.so")
if os.path.lexists(so):
    os.remove(so)
os.symlink(ours, so)        # <- the whole integration: the reference loads OUR library under its own name
sys.path.insert(0, scratch)

from openke.config import Tester, Validator                      # noqa: E402  (reference classes)
from openke.data import TrainDataLoader, TestDataLoader          # noqa: E402

# ---- construction binds every symbol the reference declares argtypes for (TrainDataLoader.py:30-103,
#      TestDataLoader.py:29-76, Tester.py:19-36, Validator.py:21-29)
dl = TrainDataLoader(in_path=wn18, nbatches=100, threads=8, sampling_mode="normal", bern_flag=1, filter_flag=1,
                     neg_ent=1, neg_rel=0, random_seed=4)
assert (dl.entTotal, dl.relTotal, dl.tripleTotal, dl.batch_size) == (40943, 18, 141442, 1414), \
    (dl.entTotal, dl.relTotal, dl.tripleTotal, dl.batch_size)
tl = TestDataLoader(wn18, "link")
vl = TestDataLoader(wn18, "link", mode="valid")
assert (tl.testTotal, vl.validTotal, tl.entTotal) == (5000, 5000, 40943)
tester = Tester(model=None, data_loader=tl, use_gpu=False)
validator = Validator(model=None, data_loader=vl)
assert os.path.realpath(dl.lib._name) == os.path.realpath(ours)
for fn in ("getNumOfNegatives", "getNegativeEntities", "getNumOfPositives", "getPositiveEntities",
           "getNumOfEntityRelations", "getEntityRelations", "getTestBatch", "activateLoadOfAllTriples"):
    getattr(dl.lib, fn)
print("bound: TrainDataLoader, TestDataLoader(test), TestDataLoader(valid), Tester, Validator")

# ---- the reference's own native core beside it, on the same files
R = ctypes.CDLL(ref_so)
for lib in (R,):
    lib.setInPath(ctypes.create_string_buffer(wn18.encode(), len(wn18) * 2))
    lib.setBern(1)
    lib.setWorkThreads(8)
    lib.setRandomSeed(4)
    lib.randReset()
    lib.importTrainFiles()
    lib.importTestFiles()
L = dl.lib
for lib in (L, R):
    for fn in ("getNumOfNegatives", "getNumOfPositives"):
        getattr(lib, fn).argtypes = [ctypes.c_int64] * 3
        getattr(lib, fn).restype = ctypes.c_int64
    for fn in ("getNegativeEntities", "getPositiveEntities"):
        getattr(lib, fn).argtypes = [ctypes.c_void_p] + [ctypes.c_int64] * 3
    lib.getNumOfEntityRelations.argtypes = [ctypes.c_int64] * 2
    lib.getNumOfEntityRelations.restype = ctypes.c_int64
    lib.getEntityRelations.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
    lib.getParallelUniverse.argtypes = [ctypes.c_int64, ctypes.c_float]
    for fn in ("getHeadBatch", "getTailBatch", "getValidHeadBatch", "getValidTailBatch"):
        getattr(lib, fn).argtypes = [ctypes.c_void_p] * 3


def addr(a):
    return a.__array_interface__["data"][0]


def compare_adjacency(n_ent, n_rel, rng, n):
    checked = 0
    for _ in range(n):
        e, r, side = int(rng.integers(n_ent)), int(rng.integers(n_rel)), int(rng.integers(2))
        for fn_n, fn_l in (("getNumOfNegatives", "getNegativeEntities"), ("getNumOfPositives", "getPositiveEntities")):
            a, b = getattr(L, fn_n)(e, r, side), getattr(R, fn_n)(e, r, side)
            assert a == b, (fn_n, e, r, side, a, b)
            if a:
                x, y = np.full(a, -7, np.int64), np.full(a, -7, np.int64)
                getattr(L, fn_l)(addr(x), e, r, side)
                getattr(R, fn_l)(addr(y), e, r, side)
                assert np.array_equal(x, y), (fn_l, e, r, side)
                checked += a
        a, b = L.getNumOfEntityRelations(e, side), R.getNumOfEntityRelations(e, side)
        assert a == b, ("getNumOfEntityRelations", e, side, a, b)
        if a:
            x, y = np.full(a, -7, np.int64), np.full(a, -7, np.int64)
            L.getEntityRelations(addr(x), e, side)
            R.getEntityRelations(addr(y), e, side)
            # the reference never advances its cursor (Base.cpp:441-468): its slot 0 = the LAST distinct relation
            assert y[0] == x[a - 1] and np.all(y[1:] == -7) and np.all(np.diff(x) > 0), (e, side, x, y)
    return checked


rng = np.random.default_rng(7)
got = compare_adjacency(40943, 18, rng, 3000)
print("adjacency queries on the global graph: identical (%d entities compared)" % got)

# ---- a universe through the reference loader's hooks (compile_universe_dataset/get_universe_mappings/swap_helpers)
U = np.load(os.path.join(golden, "universe.npz"))
seed, tc, bal = U["cases"][0]
for lib in (L, R):
    lib.setRandomSeed(int(seed))
    lib.randReset()
dl.compile_universe_dataset(int(tc), float(bal))
R.getParallelUniverse(int(tc), float(bal))
er, rr = dl.get_universe_mappings()
assert np.array_equal(er, U["u0_ent_remap"]) and np.array_equal(rr, U["u0_rel_remap"])
nE, nR = int(U["u0_sizes"][1]), int(U["u0_sizes"][2])
dl.swap_helpers()
R.swapHelpers()
assert (L.getEntityTotal(), L.getRelationTotal(), L.getTrainTotal()) == (R.getEntityTotal(), R.getRelationTotal(), R.getTrainTotal())
got = compare_adjacency(nE, nR, rng, 1500)
dl.reset_universe()
R.resetUniverse()
assert dl.batch_size == 1414 and L.getTrainTotal() == 141442
print("universe via the reference loader hooks: remaps identical; adjacency in local ids identical (%d)" % got)

# ---- candidate batches of the reference loaders (TestDataLoader.sampling_lp) against the reference core
L.initTest()
R.initTest()
L.validInit()
R.validInit()
bufs = [np.zeros(40943, np.int64) for _ in range(3)]
for i, [head, tail] in enumerate(tl):
    for fn, d in (("getHeadBatch", head), ("getTailBatch", tail)):
        getattr(R, fn)(*[addr(b) for b in bufs])
        hh, tt, rr_ = bufs
        if d["mode"] == "head_batch":
            assert np.array_equal(d["batch_h"], hh) and d["batch_t"][0] == tt[0] and d["batch_r"][0] == rr_[0]
        else:
            assert np.array_equal(d["batch_t"], tt) and d["batch_h"][0] == hh[0] and d["batch_r"][0] == rr_[0]
    if i == 40:
        break
for i, [head, tail] in enumerate(vl):
    R.getValidHeadBatch(*[addr(b) for b in bufs])
    assert np.array_equal(head["batch_h"], bufs[0]) and head["batch_t"][0] == bufs[1][0]
    R.getValidTailBatch(*[addr(b) for b in bufs])
    assert np.array_equal(tail["batch_t"], bufs[1]) and tail["batch_h"][0] == bufs[0][0]
    if i == 40:
        break
print("candidate batches (test + valid): identical")
print("DROPIN-OK")
