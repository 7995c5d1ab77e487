"""Universe subgraphs built on the GPU (csrc/walk_device.cu) against the host builder, which tests/test_host.py and
tests/test_oracle.py pin to the reference (goldens minted by the unmodified getParallelUniverse) and to the oracle.
Integer work: everything must be bit-exact — sizes, focus relation, the number of libc draws consumed, the collected
triples in collection order, the local (h,r,t) list, both remaps and the sampler streams randReset() leaves."""
import ctypes
import os
import sys

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
N = util.native()


def _host_universes(L, seeds, tcs, bals, W):
    n = len(seeds)
    handle = L.pk_universes_build_lean(n, N.addr(seeds), N.addr(tcs), N.addr(bals), 8)
    assert handle, N.last_error()
    try:
        nT, nE, nR, focus = (np.zeros(n, dtype=np.int64) for _ in range(4))
        N.check(L.pk_universes_sizes(handle, N.addr(nT), N.addr(nE), N.addr(nR), N.addr(focus)))
        by_head = np.zeros((int(nT.sum()), 3), dtype=np.int32)
        got = np.zeros((int(nT.sum()), 3), dtype=np.int32)
        er, rr = np.zeros(int(nE.sum()), dtype=np.int32), np.zeros(int(nR.sum()), dtype=np.int32)
        lcg = np.zeros((n, W), dtype=np.uint64)
        N.check(L.pk_universes_export(handle, N.addr(by_head), None, N.addr(got), N.addr(er), N.addr(rr), None, None, N.addr(lcg)))
    finally:
        L.pk_universes_free(handle)
    return nT, nE, nR, focus, by_head, got, er, rr, lcg


def _compare(L, walker, seeds, tcs, bals, W):
    import torch
    res = walker.submit(seeds, tcs, bals, W)
    sizes = res.wait()
    nT, nE, nR, focus, by_head, got, er, rr, lcg = _host_universes(L, seeds, tcs, bals, W)
    assert (sizes[:, 5] == 0).all(), sizes[sizes[:, 5] != 0][:5]
    assert np.array_equal(sizes[:, 0], nT) and np.array_equal(sizes[:, 1], nE) and np.array_equal(sizes[:, 2], nR)
    assert np.array_equal(sizes[:, 3], focus) and np.array_equal(res.focus, focus)
    assert np.array_equal(res.lcg, lcg)
    torch.cuda.synchronize()
    d_tri, d_got = res.bufs["tri"].cpu().numpy(), res.bufs["got"].cpu().numpy()
    to = np.concatenate([[0], np.cumsum(nT)])
    for i in range(len(seeds)):
        assert np.array_equal(d_got[i, :nT[i]], got[to[i]:to[i + 1]]), (i, "collected triples")
        assert np.array_equal(d_tri[i, :nT[i]], by_head[to[i]:to[i + 1]]), (i, "local (h,r,t) list")
    e2, r2 = res.packed_remaps()
    assert np.array_equal(e2, er) and np.array_equal(r2, rr)
    res.release()
    return sizes


def _hyper(n, seed0, tc_lo, tc_hi, bal_lo=0.25, bal_hi=0.5):
    from random import Random
    seeds = np.arange(seed0, seed0 + n, dtype=np.int64)
    tcs, bals = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.float32)
    for i in range(n):
        rnd = Random(int(seeds[i]))
        tcs[i] = rnd.randrange(tc_lo, tc_hi)
        bals[i] = round(rnd.uniform(bal_lo, bal_hi), 2)
    return seeds, tcs, bals


def test_device_walk_equals_host_builder_on_wn18(wn18_dir):
    import torch
    from openke.data import TrainDataLoader
    from openke.universe_walk import DeviceWalker
    dl = TrainDataLoader(in_path=wn18_dir, nbatches=20, threads=8, bern_flag=0, filter_flag=0, neg_ent=1, random_seed=4)
    L = dl.lib
    walker = DeviceWalker(L, torch.device("cuda", 0))
    assert walker.usable()
    seeds, tcs, bals = _hyper(300, 4, 500, 2000)
    sizes = _compare(L, walker, seeds, tcs, bals, 8)
    assert sizes[:, 0].min() >= 1 and sizes[:, 6].max() >= 2
    # small and degenerate requests: a handful of triples, balance 0 (no starting points: the walk collects nothing
    # and the host builder reports the error), balance 1, a threshold above the focus relation's entity count
    seeds, tcs, bals = _hyper(40, 1000, 1, 40, 0.0, 1.0)
    res = walker.submit(seeds, tcs, bals, 8)
    s = res.wait()
    for i in range(40):
        handle = L.pk_universes_build_lean(1, N.addr(seeds[i:i + 1]), N.addr(tcs[i:i + 1]), N.addr(bals[i:i + 1]), 1)
        if s[i, 5] == 3:                      # PK_WALK_EMPTY
            assert not handle
        else:
            assert s[i, 5] == 0 and handle
            nT = np.zeros(1, dtype=np.int64)
            N.check(L.pk_universes_sizes(handle, N.addr(nT), None, None, None))
            assert nT[0] == s[i, 0]
            L.pk_universes_free(handle)
    res.release()
    # a request beyond the kernel's capacity is reported, not truncated
    seeds, tcs, bals = _hyper(3, 7, 500, 2000)
    tcs[1] = 5000
    res = walker.submit(seeds, tcs, bals, 8)
    assert res.wait()[:, 5].tolist() == [0, 1, 0] and not res.ok
    res.release()


def test_device_walk_equals_host_builder_on_relation_rich_and_large_graphs(tmp_path):
    """FB15K shape (1 345 relations; dense hubs, many collected-before retries) and a 300 000-entity graph whose
    largest focus relation does not fit the shared-memory pick tree (global-memory tree, long bitmaps)."""
    import torch
    sys.path.insert(0, os.path.join(util.REPO, "tools"))
    import synth
    from openke.data import TrainDataLoader
    from openke.universe_walk import DeviceWalker
    tr, va, te, _, _ = synth.fb15k_shape()
    path = synth.write_dataset(str(tmp_path / "fb"), tr, va[:100], te[:100], 14951, 1345)
    dl = TrainDataLoader(in_path=path, nbatches=20, threads=8, bern_flag=0, filter_flag=0, neg_ent=1, random_seed=4)
    walker = DeviceWalker(dl.lib, torch.device("cuda", 0))
    seeds, tcs, bals = _hyper(200, 4, 500, 2000)
    _compare(dl.lib, walker, seeds, tcs, bals, 8)

    g = synth.power_law_graph(300_000, 6, 1_500_000, seed=99)
    path = synth.write_dataset(str(tmp_path / "big"), g[:-200], g[-200:-100], g[-100:], 300_000, 6)
    dl = TrainDataLoader(in_path=path, nbatches=20, threads=8, bern_flag=0, filter_flag=0, neg_ent=1, random_seed=4)
    walker = DeviceWalker(dl.lib, torch.device("cuda", 0))
    need = np.zeros(3, dtype=np.int64)
    N.check(dl.lib.pk_walk_scratch_bytes(1, N.addr(need)))
    assert need[2] > 0, "this graph was meant to need the global-memory pick tree"
    seeds, tcs, bals = _hyper(64, 11, 500, 2000)
    _compare(dl.lib, walker, seeds, tcs, bals, 8)


def test_orchestrator_trains_the_same_ensemble_with_device_and_host_universes(tmp_path):
    """Parallel_Universe_Config with universes built on the GPU (default) against the same configuration with the host
    builder: identical subgraphs (sizes, focus relations, local id maps), the same sampler streams and therefore the same
    training — losses and tables agree up to the order of the float reductions of repeated rows, the evaluation ranks
    the same.  Requests the kernel does not cover take the host path; an impossible request raises the host builder's
    error either way."""
    import torch
    import test_product as T
    path, _ = T._small_graph(tmp_path)
    dev_pu, host_pu = T._pu(path), T._pu(path)
    host_pu.device_walk = False
    for pu in (dev_pu, host_pu):
        pu.record_losses = True
        pu.async_training = True
        for n in (5, 4, 6):
            pu.train_parallel_universes(n)
        pu.synchronize()
    assert dev_pu._walker is not None and dev_pu._walker.launches >= 3 and host_pu._walker is None
    for u in range(15):
        a, b = dev_pu.universe_hyper[u], host_pu.universe_hyper[u]
        assert a == b, (u, a, b)
        assert dict(dev_pu.entity_id_mappings[u]) == dict(host_pu.entity_id_mappings[u])
        assert dict(dev_pu.relation_id_mappings[u]) == dict(host_pu.relation_id_mappings[u])
        assert np.allclose(dev_pu.universe_losses[u], host_pu.universe_losses[u], rtol=1e-4, atol=1e-6), u
        wa = dev_pu.trained_embedding_spaces[u].ent_embeddings.weight
        wb = host_pu.trained_embedding_spaces[u].ent_embeddings.weight
        assert wa.shape == wb.shape and float((wa - wb).abs().max()) < 1e-4, u
    dev_pu.run_link_prediction()
    host_pu.run_link_prediction()
    assert (dev_pu.last_ranks == host_pu.last_ranks).all(1).mean() > 0.95
    # beyond the kernel's capacity: the host builder is used from the start
    big = T._pu(path, min_triple_constraint=2500, max_triple_constraint=2600)
    big.train_parallel_universes(2)
    big.synchronize()
    assert (big._walker is None or big._walker.launches == 0) and big.universe_hyper[0]["nT"] >= 1
    # no starting points (balance 0): the walk collects nothing; both builders end in the same error
    for flag in (True, False):
        none = T._pu(path, min_balance=0.0, max_balance=0.0)
        none.device_walk = flag
        with pytest.raises(N.NativeError, match="collected no triples"):
            none.train_parallel_universes(2)
            none.synchronize()


def test_device_walk_against_reference_goldens_and_the_oracle(wn18_dir, golden):
    """The GPU builder pinned directly: (1) to the universes the UNMODIFIED reference built (tests/golden/universe.npz,
    minted by getParallelUniverse through the reference's own Python: sizes, both remaps, the collected triples in global
    ids, the local (h,r,t) list); (2) to the oracle's plain restatement of UniverseConstructor.h on 72 further universes
    of the static script's ranges."""
    import random
    import torch
    from openke.data import TrainDataLoader
    from openke.universe_walk import DeviceWalker
    from oracle import native as on
    dl = TrainDataLoader(in_path=wn18_dir, nbatches=20, threads=8, bern_flag=0, filter_flag=0, neg_ent=1, random_seed=4)
    walker = DeviceWalker(dl.lib, torch.device("cuda", 0))
    cap = walker.cap
    U = golden["universe"]
    cases = [(int(s_), int(tc), float(b)) for s_, tc, b in U["cases"]]
    keep = [i for i, c in enumerate(cases) if c[1] <= cap and int(np.float32(c[2]) * np.float32(c[1])) <= cap]
    assert len(keep) >= 4
    seeds = np.array([cases[i][0] for i in keep], np.int64)
    tcs = np.array([cases[i][1] for i in keep], np.int64)
    bals = np.array([cases[i][2] for i in keep], np.float32)
    res = walker.submit(seeds, tcs, bals, 8)
    s = res.wait()
    torch.cuda.synchronize()
    tri, got = res.bufs["tri"].cpu().numpy(), res.bufs["got"].cpu().numpy()
    er, rr = res.bufs["ent_remap"].cpu().numpy(), res.bufs["rel_remap"].cpu().numpy()
    for j, i in enumerate(keep):
        assert s[j, 5] == 0 and tuple(s[j, :3]) == tuple(U["u%d_sizes" % i]), (i, s[j])
        nT, nE, nR = (int(x) for x in s[j, :3])
        assert np.array_equal(er[j, :nE], U["u%d_ent_remap" % i]) and np.array_equal(rr[j, :nR], U["u%d_rel_remap" % i])
        assert np.array_equal(got[j, :nT], U["u%d_triples_global" % i])
        assert np.array_equal(tri[j, :nT], U["u%d_triples_local" % i])
    res.release()

    w = np.load(os.path.join(util.GOLDEN, "wn18.npz"))
    o = on.Oracle(threads=8, bern=0)
    o.import_train(w["train"], 40943, 18)
    rng = random.Random(7)
    more = [(1000 + i, rng.randrange(500, 2000), rng.uniform(0.25, 0.5)) for i in range(72)]
    more += [(5, 50, 0.02), (8, 700, 1.0), (10, 3, 0.5), (11, 2, 1.0)]
    seeds = np.array([c[0] for c in more], np.int64)
    tcs = np.array([c[1] for c in more], np.int64)
    bals = np.array([c[2] for c in more], np.float32)
    res = walker.submit(seeds, tcs, bals, 8)
    s = res.wait()
    torch.cuda.synchronize()
    got = res.bufs["got"].cpu().numpy()
    er, rr = res.bufs["ent_remap"].cpu().numpy(), res.bufs["rel_remap"].cpu().numpy()
    for j, (seed, tc, bal) in enumerate(more):
        o.seed(seed)
        otri, oer, orr = o.universe(tc, float(np.float32(bal)))
        assert s[j, 5] == 0, (j, seed, tc, bal, s[j])
        assert np.array_equal(got[j, :s[j, 0]], otri), (j, seed, tc, bal)
        assert np.array_equal(er[j, :s[j, 1]], oer) and np.array_equal(rr[j, :s[j, 2]], orr)
    res.release()
