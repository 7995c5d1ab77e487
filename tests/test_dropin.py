"""The drop-in boundary: the reference's own, unmodified
Python classes construct and run against libputranse.so loaded under the name release/Base.so.

CPU part (needs /root/reference, i.e. this container): child process tests/_dropin_child.py.
GPU part (golden data only, /root/reference does not exist on the GPU box): the device-backed exports
driven with exactly the ctypes argtypes the reference declares."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import util

N = util.native()
REF_PKG = "/root/reference/openke"
REF_SO = os.path.join(util.REPO, "oracle", "_ref", "Base.so")


@pytest.mark.skipif(not (os.path.isdir(REF_PKG) and os.path.exists(REF_SO)),
                    reason="needs the reference tree and oracle/_ref/Base.so (build container only)")
def test_reference_python_runs_on_libputranse(wn18_dir, tmp_path):
    child = os.path.join(util.REPO, "tests", "_dropin_child.py")
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)      # the child must import the REFERENCE's openke, not this repo's
    out = subprocess.run([sys.executable, child, str(tmp_path / "refpkg"), N.LIB_PATH, REF_SO, wn18_dir, util.GOLDEN],
                         capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert out.returncode == 0 and "DROPIN-OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def _bind_like_the_reference(L):
    """The argtypes/restypes of openke/data/TrainDataLoader.py:33-103, TestDataLoader.py:38-76,
    config/Tester.py:22-36, config/Validator.py:26-28 — verbatim types, nothing of ours."""
    i64, vp, f32 = ctypes.c_int64, ctypes.c_void_p, ctypes.c_float
    L.sampling.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64]
    L.getParallelUniverse.argtypes = [i64, f32]
    L.getEntityRemapping.argtypes = [vp]
    L.getRelationRemapping.argtypes = [vp]
    for fn in ("getHeadBatch", "getTailBatch", "getValidHeadBatch", "getValidTailBatch"):
        getattr(L, fn).argtypes = [vp, vp, vp]
    L.getTestBatch.argtypes = [vp] * 6
    L.setRandomSeed.argtypes = [i64]
    L.activateLoadOfAllTriples.argtypes = [i64]
    L.testHead.argtypes = [vp, i64, i64]
    L.testTail.argtypes = [vp, i64, i64]
    L.test_link_prediction.argtypes = [i64]
    for fn in ("getTestLinkMRR", "getTestLinkMR", "getTestLinkHit10", "getTestLinkHit3", "getTestLinkHit1"):
        getattr(L, fn).argtypes = [i64]
        getattr(L, fn).restype = f32
    L.validHead.argtypes = [vp, i64]
    L.validTail.argtypes = [vp, i64]
    L.getValidHit10.restype = f32


def _addr(a):
    return a.__array_interface__["data"][0]


@pytest.mark.gpu
def test_reference_call_sequence_through_the_cabi(wn18_dir, golden):
    """TrainDataLoader.read -> sampling (Base.cpp:266-310) with caller-owned int64/float32 numpy buffers,
    TestDataLoader.read -> getHeadBatch/getTailBatch -> testHead/testTail -> test_link_prediction ->
    getTestLink* (Tester.run_link_prediction, Tester.py:70-93) and the Validator twins, on a raw
    ctypes handle bound like the reference binds it."""
    L = ctypes.cdll.LoadLibrary(N.LIB_PATH)
    _bind_like_the_reference(L)
    # a different dataset shape first resets the Bernoulli import-count drift, as a fresh process would
    import tempfile
    tiny = util.write_dataset(tempfile.mkdtemp(), [[0, 1, 0], [1, 2, 0]], [[0, 2, 0]], [[2, 0, 0]], 3, 1)
    L.setInPath(ctypes.create_string_buffer(tiny.encode(), len(tiny) * 2))
    L.importTrainFiles()
    g = golden["sampler"]
    for name, bern, filt, k in (("b0f0k1", 0, 0, 1), ("b1f1k1", 1, 1, 1), ("b0f1k2", 0, 1, 2), ("b1f0k3", 1, 0, 3)):
        L.setInPath(ctypes.create_string_buffer(wn18_dir.encode(), len(wn18_dir) * 2))
        L.setBern(bern)
        L.setWorkThreads(8)
        L.setRandomSeed(4)
        L.randReset()
        L.importTrainFiles()
        B = L.getTrainTotal() // 100
        n = B * (1 + k)
        bh, bt, br = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
        by = np.zeros(n, np.float32)
        for call in range(3):
            L.sampling(_addr(bh), _addr(bt), _addr(br), _addr(by), B, k, 0, 0, filt, 0, 0)
            want = g[name][call]
            assert np.array_equal(bh, want[0]) and np.array_equal(bt, want[1]) and np.array_equal(br, want[2]), (name, call)
            assert np.all(by[:B] == 1) and np.all(by[B:] == -1)

    # ---- link prediction exactly as Tester.run_link_prediction drives it, scores from the shipped TransH checkpoint
    R = golden["rank_transh_wn18"]
    L.setRandomSeed(4)
    L.randReset()
    L.importTrainFiles()
    L.importTestFiles()
    E, T = L.getEntityTotal(), L.getTestTotal()
    assert (E, T) == (40943, 5000)
    ent, rel, nv = R["ent_embeddings"], R["rel_embeddings"], R["norm_vector"]

    def transh_scores(h, t, r):   # openke/module/model/TransH.py:52-92 in numpy float32 (p_norm 1)
        def nrm(x):
            return x / np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), 1e-12)
        w = nrm(nv[r])
        hh, tt = ent[h], ent[t]
        hh = hh - (hh * w).sum(-1, keepdims=True) * w
        tt = tt - (tt * w).sum(-1, keepdims=True) * w
        return np.abs(nrm(hh) + nrm(rel[r]) - nrm(tt)).sum(-1).astype(np.float32)

    ph, pt, pr = (np.zeros(E, np.int64) for _ in range(3))
    L.initTest()
    n_check = 120
    ranks = np.zeros((n_check, 4), np.int64)
    for i in range(n_check):
        L.getHeadBatch(_addr(ph), _addr(pt), _addr(pr))
        assert (ph[0], pr[0], pt[0]) == tuple(R["test_sorted"][i]) and np.all(pt == pt[0]) and np.all(pr == pr[0])
        s = transh_scores(ph, pt[:1], pr[:1])
        L.testHead(_addr(s), i, 0)
        ranks[i, 0] = (s[1:] < s[0]).sum()
        L.getTailBatch(_addr(ph), _addr(pt), _addr(pr))
        assert np.all(ph == ph[0])
        s = transh_scores(ph[:1], pt, pr[:1])
        L.testTail(_addr(s), i, 0)
        ranks[i, 2] = (s[1:] < s[0]).sum()
    # the host scores above are this test's own numpy arithmetic, so the RAW ranks are known exactly; the golden
    # ranks (reference torch arithmetic) may differ from them in near-ties only
    gold = R["ranks_p1"][:n_check]
    assert (np.abs(ranks[:, [0, 2]] - gold[:, [0, 2]]) <= 2).mean() > 0.97
    L.test_link_prediction(0)
    # the accumulators hold the first n_check triples divided by testTotal (Test.h:450-454)
    filt_l, filt_r = gold[:, 1].astype(np.float64), gold[:, 3].astype(np.float64)
    want_mrr = ((1 / (filt_l + 1)).sum() / T + (1 / (filt_r + 1)).sum() / T) / 2
    want_h10 = ((filt_l < 10).sum() / T + (filt_r < 10).sum() / T) / 2
    assert abs(L.getTestLinkMRR(0) - want_mrr) < 2e-3 * want_mrr + 1e-6
    assert abs(L.getTestLinkHit10(0) - want_h10) <= 2.0 / T

    # ---- Validator.valid's loop (Validator.py:37-45)
    L.validInit()
    V = L.getValidTotal()
    csr = []
    for side in (0, 1):   # known-true candidates per valid query (Corrupt.h:188-199 _find, as a list)
        off, nc = np.zeros(V + 1, np.int64), ctypes.c_int64(0)
        N.check(N.lib().pk_filter_csr(1, side, N.addr(off), None, ctypes.byref(nc)))   # same dlopen handle, same global state
        cand = np.zeros(max(nc.value, 1), np.int32)
        N.check(N.lib().pk_filter_csr(1, side, N.addr(off), N.addr(cand), ctypes.byref(nc)))
        csr.append((off, cand))
    hits = [0, 0]

    def filtered_rank(s, cands, truth):   # s is in candidate order: slot 0 = truth, then the others ascending
        by_entity = np.empty(E, np.float32)
        others = np.delete(np.arange(E), truth)
        by_entity[truth], by_entity[others] = s[0], s[1:]
        return int((s[1:] < s[0]).sum()) - int((by_entity[cands] < s[0]).sum())

    for i in range(40):
        L.getValidHeadBatch(_addr(ph), _addr(pt), _addr(pr))
        s = transh_scores(ph, pt[:1], pr[:1])
        L.validHead(_addr(s), i)
        hits[0] += filtered_rank(s, csr[0][1][csr[0][0][i]:csr[0][0][i + 1]], int(ph[0])) < 10
        L.getValidTailBatch(_addr(ph), _addr(pt), _addr(pr))
        s = transh_scores(ph[:1], pt, pr[:1])
        L.validTail(_addr(s), i)
        hits[1] += filtered_rank(s, csr[1][1][csr[1][0][i]:csr[1][0][i + 1]], int(pt[0])) < 10
    want = (np.float32(hits[0]) / np.float32(V) + np.float32(hits[1]) / np.float32(V)) / np.float32(2)
    assert L.getValidHit10() == pytest.approx(float(want), abs=1e-7) and hits[0] + hits[1] > 0
