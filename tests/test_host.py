"""CPU tests of the product's host side: the C-ABI library loads and exports what include/putranse.h
declares, and its integer work (reader, indexes, Bernoulli means, glibc rand clone, universe
construction, filter lists) is bit-identical to vectors minted by the unmodified reference."""
import ctypes
import os
import re

import numpy as np
import pytest

import util

N = util.native()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(util.REPO, "include", "putranse.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", hdr))
    names -= {"defined", "sizeof"}
    names = {n for n in names if not n.startswith("PK_")}
    L = ctypes.CDLL(N.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, "declared in include/putranse.h but not exported: %s" % missing
    assert len(names) > 60


def test_no_cuda_means_loud_failure_not_fallback():
    L = N.lib()
    if L.pk_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(N.NativeError):
        N.require_cuda()


def test_glibc_rand_clone_matches_libc(wn18_dir):
    """setRandomSeed/randReset expose the generator: the stream states are rand() outputs."""
    L = N.lib()
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 4, 5, 123, 2147483647, 0):
        L.setWorkThreads(8)
        L.setRandomSeed(seed)
        L.randReset()
        got = np.zeros(8, np.uint64)
        N.check(L.pk_get_lcg(N.addr(got), 8))
        libc.srand(ctypes.c_uint(seed))
        want = np.array([libc.rand() for _ in range(8)], np.uint64)
        assert np.array_equal(got, want), seed


def _load(L, path, threads=8, bern=0, seed=4):
    L.setInPath(path.encode())
    L.setBern(bern)
    L.setWorkThreads(threads)
    L.setRandomSeed(seed)
    L.randReset()
    L.importTrainFiles()


def test_reader_and_bernoulli_drift(wn18_dir, golden):
    """left_mean/right_mean after the 1st..4th import of one process."""
    L = N.lib()
    # a different dataset shape first resets the drift, as a fresh process would
    g = golden["sampler"]
    import tempfile
    tiny = util.write_dataset(tempfile.mkdtemp(), [[0, 1, 0], [1, 2, 0]], [[0, 2, 0]], [[2, 0, 0]], 3, 1)
    _load(L, tiny)
    for name in ("b0f0k1", "b1f1k1", "b0f1k2", "b1f0k3"):
        _load(L, wn18_dir)
        assert L.pk_import_count() == int(g[name + "_imports"])
        assert (L.getEntityTotal(), L.getRelationTotal(), L.getTrainTotal()) == (40943, 18, 141442)
        lm, rm = np.zeros(18, np.float32), np.zeros(18, np.float32)
        N.check(L.pk_train_index(None, None, N.addr(lm), N.addr(rm)))
        assert np.array_equal(lm, g[name + "_left_mean"]), name
        assert np.array_equal(rm, g[name + "_right_mean"]), name
    by_head, by_tail = np.zeros((141442, 3), np.int32), np.zeros((141442, 3), np.int32)
    N.check(L.pk_train_index(N.addr(by_head), N.addr(by_tail), None, None))
    k = by_head[:, 0].astype(np.int64) * 10 ** 8 + by_head[:, 1].astype(np.int64) * 10 ** 6 // 10 + 0
    assert np.all(np.lexsort((by_head[:, 2], by_head[:, 1], by_head[:, 0])) == np.arange(141442))
    assert np.all(np.lexsort((by_tail[:, 0], by_tail[:, 1], by_tail[:, 2])) == np.arange(141442))


def test_universes_bit_exact_with_reference(wn18_dir, golden):
    L = N.lib()
    _load(L, wn18_dir)
    U = golden["universe"]
    for i, (seed, tc, bal) in enumerate(U["cases"]):
        L.setRandomSeed(int(seed))
        L.randReset()
        L.getParallelUniverse(int(tc), float(bal))
        nT, nE, nR = L.getTrainTotalUniverse(), L.getEntityTotalUniverse(), L.getRelationTotalUniverse()
        assert (nT, nE, nR) == tuple(U["u%d_sizes" % i])
        er, rr = np.zeros(nE, np.int64), np.zeros(nR, np.int64)
        L.getEntityRemapping(N.addr(er))
        L.getRelationRemapping(N.addr(rr))
        tri = np.zeros((nT, 3), np.int32)
        N.check(L.pk_universe_triples(N.addr(tri)))
        assert np.array_equal(er, U["u%d_ent_remap" % i])
        assert np.array_equal(rr, U["u%d_rel_remap" % i])
        assert np.array_equal(tri, U["u%d_triples_global" % i])
        L.swapHelpers()
        assert (L.getTrainTotal(), L.getEntityTotal(), L.getRelationTotal()) == (nT, nE, nR)
        bh, lm, rm = np.zeros((nT, 3), np.int32), np.zeros(nR, np.float32), np.zeros(nR, np.float32)
        N.check(L.pk_train_index(N.addr(bh), None, N.addr(lm), N.addr(rm)))
        assert np.array_equal(bh, U["u%d_triples_local" % i])
        assert np.array_equal(lm, U["u%d_left_mean" % i]) and np.array_equal(rm, U["u%d_right_mean" % i])
        L.resetUniverse()
        assert L.getTrainTotal() == 141442


def test_batched_threaded_builder_equals_sequential(wn18_dir, golden):
    L = N.lib()
    _load(L, wn18_dir)
    U = golden["universe"]
    cases = U["cases"]
    n = len(cases)
    seeds = np.ascontiguousarray(cases[:, 0], np.int64)
    tcs = np.ascontiguousarray(cases[:, 1], np.int64)
    bals = np.ascontiguousarray(cases[:, 2], np.float32)
    outs = []
    for threads in (1, 4):
        h = L.pk_universes_build(n, N.addr(seeds), N.addr(tcs), N.addr(bals), threads)
        assert h, N.last_error()
        nT, nE, nR, foc = (np.zeros(n, np.int64) for _ in range(4))
        N.check(L.pk_universes_sizes(h, N.addr(nT), N.addr(nE), N.addr(nR), N.addr(foc)))
        er, rr = np.zeros(nE.sum(), np.int32), np.zeros(nR.sum(), np.int32)
        bh, bt, bg = (np.zeros((nT.sum(), 3), np.int32) for _ in range(3))
        lcg = np.zeros((n, 8), np.uint64)
        N.check(L.pk_universes_export(h, N.addr(bh), N.addr(bt), N.addr(bg), N.addr(er), N.addr(rr), None, None, N.addr(lcg)))
        L.pk_universes_free(h)
        outs.append((nT, nE, nR, er, rr, bh, bt, bg, lcg))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    nT, nE, nR, er, rr, bh, bt, bg, lcg = outs[0]
    eo, to = np.concatenate([[0], np.cumsum(nE)]), np.concatenate([[0], np.cumsum(nT)])
    libc = ctypes.CDLL("libc.so.6")
    for i in range(n):
        assert np.array_equal(er[eo[i]:eo[i + 1]], U["u%d_ent_remap" % i])
        assert np.array_equal(bg[to[i]:to[i + 1]], U["u%d_triples_global" % i])
        assert np.array_equal(bh[to[i]:to[i + 1]], U["u%d_triples_local" % i])
        libc.srand(ctypes.c_uint(int(seeds[i])))
        assert np.array_equal(lcg[i], np.array([libc.rand() for _ in range(8)], np.uint64))
        # (t,r,h) order of the tail index
        t = bt[to[i]:to[i + 1]]
        assert np.all(np.lexsort((t[:, 0], t[:, 1], t[:, 2])) == np.arange(t.shape[0]))


def test_builder_matches_oracle_on_many_seeds(wn18_dir):
    """The product's universe builder (per-relation entity lists precomputed at import, triple-id stamps,
    packed-key sorts) against the oracle's plain restatement of UniverseConstructor.h on 80 universes of the
    static script's ranges plus edge cases (tiny and huge focus subsets, walks that stall)."""
    import random
    from oracle import native as on
    L = N.lib()
    _load(L, wn18_dir)
    w = np.load(os.path.join(util.GOLDEN, "wn18.npz"))
    o = on.Oracle(threads=8, bern=0)
    o.import_train(w["train"], 40943, 18)
    rng = random.Random(7)
    cases = [(1000 + i, rng.randrange(500, 2000), rng.uniform(0.25, 0.5)) for i in range(72)]
    cases += [(5, 50, 0.02), (6, 3000, 0.9), (8, 700, 1.0), (9, 12000, 0.5), (10, 3, 0.5), (11, 2, 1.0), (12, 6000, 0.05)]
    n = len(cases)
    seeds = np.array([c[0] for c in cases], np.int64)
    tcs = np.array([c[1] for c in cases], np.int64)
    bals = np.array([c[2] for c in cases], np.float32)
    h = L.pk_universes_build(n, N.addr(seeds), N.addr(tcs), N.addr(bals), 3)
    assert h, N.last_error()
    nT, nE, nR, foc = (np.zeros(n, np.int64) for _ in range(4))
    N.check(L.pk_universes_sizes(h, N.addr(nT), N.addr(nE), N.addr(nR), N.addr(foc)))
    er, rr = np.zeros(nE.sum(), np.int32), np.zeros(nR.sum(), np.int32)
    bh, bt, bg = (np.zeros((nT.sum(), 3), np.int32) for _ in range(3))
    N.check(L.pk_universes_export(h, N.addr(bh), N.addr(bt), N.addr(bg), N.addr(er), N.addr(rr), None, None, None))
    L.pk_universes_free(h)
    eo, ro, to = (np.concatenate([[0], np.cumsum(x)]) for x in (nE, nR, nT))
    for i, (seed, tc, bal) in enumerate(cases):
        o.seed(seed)
        tri, oer, orr = o.universe(tc, float(np.float32(bal)))
        assert np.array_equal(bg[to[i]:to[i + 1]], tri), (i, seed, tc, bal)
        assert np.array_equal(er[eo[i]:eo[i + 1]], oer) and np.array_equal(rr[ro[i]:ro[i + 1]], orr)
        t = bt[to[i]:to[i + 1]]
        assert np.all(np.lexsort((t[:, 0], t[:, 1], t[:, 2])) == np.arange(t.shape[0]))
        hh = bh[to[i]:to[i + 1]]
        assert np.all(np.lexsort((hh[:, 2], hh[:, 1], hh[:, 0])) == np.arange(hh.shape[0]))
        assert sorted(map(tuple, hh.tolist())) == sorted(map(tuple, t.tolist()))


def test_universe_invariants_on_synthetic_graph(tmp_path):
    """The reference's opt-in Checks.h invariants (openke/base/UniverseSetting.h:203-269) as properties."""
    tr, va, te = util.synthetic_graph(3000, 12, 20000, 200, seed=7)
    path = util.write_dataset(str(tmp_path / "syn"), tr, va, te, 3000, 12)
    L = N.lib()
    _load(L, path, seed=3)
    train_set = set(map(tuple, tr[:, [0, 2, 1]].tolist()))
    for seed in range(20, 30):
        L.setRandomSeed(seed)
        L.randReset()
        L.getParallelUniverse(400 + 37 * (seed % 5), 0.3)
        nT, nE, nR = L.getTrainTotalUniverse(), L.getEntityTotalUniverse(), L.getRelationTotalUniverse()
        assert 0 < nT <= 400 + 37 * (seed % 5)
        tri = np.zeros((nT, 3), np.int32)
        N.check(L.pk_universe_triples(N.addr(tri)))
        assert all(tuple(x) in train_set for x in tri.tolist())          # universe is a subset of train
        assert len(set(map(tuple, tri.tolist()))) == nT                     # no duplicates
        er, rr = np.zeros(nE, np.int64), np.zeros(nR, np.int64)
        L.getEntityRemapping(N.addr(er))
        L.getRelationRemapping(N.addr(rr))
        assert len(set(er.tolist())) == nE and len(set(rr.tolist())) == nR  # remaps are injective
        L.swapHelpers()
        loc = np.zeros((nT, 3), np.int32)
        N.check(L.pk_train_index(N.addr(loc), None, None, None))
        assert loc[:, [0, 2]].max() == nE - 1 and loc[:, 1].max() == nR - 1  # max local id = count - 1
        back = np.stack([er[loc[:, 0]], rr[loc[:, 1]], er[loc[:, 2]]], 1)
        assert set(map(tuple, back.tolist())) == set(map(tuple, tri.tolist()))  # remap round-trips
        L.resetUniverse()


def test_eval_lists_and_filter_csr(wn18_dir, golden):
    L = N.lib()
    _load(L, wn18_dir)
    L.importTestFiles()
    assert (L.getTestTotal(), L.getValidTotal()) == (5000, 5000)
    tri = np.zeros((5000, 3), np.int32)
    N.check(L.pk_eval_triples(0, N.addr(tri)))
    assert np.array_equal(tri, golden["rank_transh_wn18"]["test_sorted"])
    g = golden["wn18"]
    allt = np.concatenate([g["train"], g["valid"], g["test"]])[:, [0, 2, 1]]   # -> (h,r,t)
    known = set(map(tuple, allt.tolist()))
    for side in (0, 1):
        cnt = ctypes.c_int64(0)
        off = np.zeros(5001, np.int64)
        N.check(L.pk_filter_csr(0, side, N.addr(off), None, ctypes.byref(cnt)))
        cand = np.zeros(cnt.value, np.int32)
        N.check(L.pk_filter_csr(0, side, N.addr(off), N.addr(cand), ctypes.byref(cnt)))
        assert off[-1] == cnt.value
        for i in list(range(0, 5000, 97)):
            h, r, t = tri[i].tolist()
            if side == 0:
                want = sorted(j for (j, rr, tt) in ((a, b, c) for (a, b, c) in known if b == r and c == t) if j != h)
            else:
                want = sorted(c for (a, b, c) in known if a == h and b == r and c != t)
            assert cand[off[i]:off[i + 1]].tolist() == want


def test_candidate_batches_follow_reference_order(wn18_dir):
    L = N.lib()
    _load(L, wn18_dir)
    L.importTestFiles()
    tri = np.zeros((5000, 3), np.int32)
    N.check(L.pk_eval_triples(0, N.addr(tri)))
    L.initTest()
    ph, pt, pr = (np.zeros(40943, np.int64) for _ in range(3))
    for i in range(3):
        L.getHeadBatch(N.addr(ph), N.addr(pt), N.addr(pr))
        h, r, t = tri[i].tolist()
        assert ph[0] == h and np.array_equal(ph[1:], np.delete(np.arange(40943), h))
        assert np.all(pt == t) and np.all(pr == r)
        L.getTailBatch(N.addr(ph), N.addr(pt), N.addr(pr))
        assert pt[0] == t and np.array_equal(pt[1:], np.delete(np.arange(40943), t)) and np.all(ph == h)


def test_hyper_draws_match_static_script(wn18_dir):
    """(tc, epochs, lr, margin) of universes 0-4 of the static WN18 script."""
    from openke.config import Parallel_Universe_Config
    pu = Parallel_Universe_Config.__new__(Parallel_Universe_Config)
    pu.min_triple_constraint, pu.max_triple_constraint = 500, 2000
    pu.min_balance, pu.max_balance = 0.25, 0.5
    pu.min_margin, pu.max_margin = 1, 4
    pu.const_num_epochs, pu.min_num_epochs, pu.max_num_epochs = None, 50, 200
    pu.min_lr, pu.max_lr = 0.001, 0.1
    want = [(983, 151, 0.048, 3), (1775, 185, 0.004, 2), (1675, 116, 0.005, 2), (1163, 62, 0.008, 2), (964, 82, 0.02, 2)]
    for u, w in enumerate(want):
        h = pu.draw_universe_hyper(4 + u)
        assert (h["tc"], h["epochs"], h["lr"], h["margin"]) == w


def test_native_hyper_draws_equal_python_random():
    """csrc/pyrandom_host.cpp restates CPython's random.seed(int) / randrange / uniform / round for the draws of
    Parallel_Universe_Config.py:211-212,232,237-240; the orchestrator uses it for whole chunks.  Every value must be
    the very object Python computes: 30 000 seeds (small, around 2^31/2^32, 2^40+, negative) x four range configurations
    (static WN18 script, WikidataEvolve script, constant epochs, wide ranges with more lr digits)."""
    from openke.config import Parallel_Universe_Config
    from openke import _native as N
    pu = Parallel_Universe_Config.__new__(Parallel_Universe_Config)
    pu.lib = N.lib()
    seeds = np.concatenate([np.arange(0, 20000), np.arange(2 ** 31 - 1000, 2 ** 31 + 1000), np.arange(2 ** 32 - 1000, 2 ** 32 + 1000),
                            2 ** 40 + np.arange(3000) * 7919, -np.arange(1, 3001)]).astype(np.int64)
    configs = [dict(tc=(500, 2000), bal=(0.25, 0.5), margin=(1, 4), const=None, ep=(50, 200), lr=(0.001, 0.1)),
               dict(tc=(500, 1500), bal=(0.25, 0.5), margin=(1, 5), const=None, ep=(50, 200), lr=(0.001, 0.1)),
               dict(tc=(1000, 1001), bal=(0.1, 0.9), margin=(2, 3), const=7, ep=(50, 200), lr=(0.01, 0.1)),
               dict(tc=(1, 100000), bal=(0.0, 1.0), margin=(1, 1000), const=None, ep=(1, 3), lr=(0.0001, 2.5))]
    for c in configs:
        pu.min_triple_constraint, pu.max_triple_constraint = c["tc"]
        pu.min_balance, pu.max_balance = c["bal"]
        pu.min_margin, pu.max_margin = c["margin"]
        pu.const_num_epochs, (pu.min_num_epochs, pu.max_num_epochs) = c["const"], c["ep"]
        pu.min_lr, pu.max_lr = c["lr"]
        got = pu.draw_hyper_batch(seeds)
        want = [pu.draw_universe_hyper(int(s_)) for s_ in seeds]
        assert got == want, next((s_, g, w) for s_, g, w in zip(seeds, got, want) if g != w)
        assert all(type(g[k]) is type(w[k]) for g, w in zip(got[:50], want[:50]) for k in w)
    # ranges the native generator refuses fall back to the interpreter (same values, same errors)
    pu.min_triple_constraint, pu.max_triple_constraint = 5, 5
    with pytest.raises(ValueError):
        pu.draw_hyper_batch(seeds[:3])


def test_link_metrics_accumulate_like_the_reference(golden):
    from openke.config.Tester import link_metrics
    g = golden["rank_transh_wn18"]
    for p in (1, 2):
        avg, _ = link_metrics(g["ranks_p%d" % p])
        assert np.array_equal(np.array(avg, np.float32), g["metrics_p%d" % p])
    pu = golden["putranse_wn18"]
    avg, _ = link_metrics(pu["ranks"])
    assert np.array_equal(np.array(avg, np.float32), pu["metrics"])   # rank sums here exceed 2**24


def test_sharded_evaluation_min_allreduce_gloo(tmp_path):
    """world_size-2 gloo run of the exchange step: per-rank energy tiles min-reduce to the tile a
    single rank holding every universe would have."""
    import subprocess
    import sys
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, numpy as np, torch, torch.distributed as dist\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "g = np.random.default_rng(0)\n"
        "full = g.random((6, 16, 50)).astype(np.float32)\n"       # 6 universes, 16 keys, 50 entities
        "full[g.random((6, 16, 50)) < 0.6] = np.inf\n"
        "mine = [u for u in range(6) if u % w == r]\n"
        "part = torch.from_numpy(full[mine].min(0)) if mine else torch.full((16, 50), float('inf'))\n"
        "dist.all_reduce(part, op=dist.ReduceOp.MIN)\n"
        "assert np.array_equal(part.numpy(), full.min(0))\n"
        "print('ok', r)\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


@pytest.mark.parametrize("cls_name,param", [("TransE", {"dim": 20, "p_norm": 1, "norm_flag": True}),
                                            ("TransH", {"dim": 20, "p_norm": 1, "norm_flag": True}),
                                            ("TransD", {"dim_e": 20, "dim_r": 20})])
def test_native_table_init_replays_torch_generator(cls_name, param):
    """pk_torch_init_tables must leave exactly what ``torch.manual_seed(seed); Model(nE, nR, ...)`` leaves
    (reference openke/module/model/TransE.py:17-22 + torch's nn.Embedding default init), for seeds beyond
    32 bits and table sizes on both sides of a multiple of 16."""
    import torch
    import openke.module.model as M
    from openke.config.Parallel_Universe_Config import Parallel_Universe_Config as PU
    cls = getattr(M, cls_name)
    pu = PU.__new__(PU)
    pu.embedding_model, pu.embedding_model_param, pu.sampler_threads = cls, param, 3
    pu.lib = N.lib()
    fused = pu._native_init_mode()
    assert fused in (0, 1), "neither rounding variant reproduces this torch build"
    nE, nR = np.array([583, 2632, 1084, 800]), np.array([11, 17, 13, 4])
    seeds = np.array([4, 5, (1 << 40) + 6, 7])
    specs = cls.table_specs(2, 1, **param)
    ent = set(cls._ent_tables)
    eo, ro = np.concatenate([[0], np.cumsum(nE)]), np.concatenate([[0], np.cumsum(nR)])
    host = {a: torch.full((int(eo[-1]) if a in ent else int(ro[-1]), d), float("nan")) for a, _, d in specs}
    offs = {a: (eo if a in ent else ro) for a in host}
    pu._native_init(cls, param, seeds, nE, nR, host, offs, fused)
    for i in range(4):
        torch.manual_seed(int(seeds[i]))
        m = cls(int(nE[i]), int(nR[i]), **param)
        for a in host:
            assert torch.equal(host[a][offs[a][i]:offs[a][i + 1]], getattr(m, a).weight.data), (cls_name, i, a)


@pytest.mark.skipif(not (os.path.isdir("/root/reference/openke") and os.path.exists(os.path.join(util.REPO, "oracle", "_ref", "Base.so"))),
                    reason="needs the reference tree (build container only)")
def test_reference_layout_checkpoint_round_trip(wn18_dir, tmp_path):
    """a checkpoint written by the reference (golden, minted by make_golden.py ref_ckpt) is
    imported by load_parameters, re-exported in the reference's layout, and read back by the UNMODIFIED reference:
    same modules, same id maps, same energies.  Host-only (conversion is not compute)."""
    import subprocess
    import sys
    import torch
    from openke.config import Parallel_Universe_Config
    from openke.data import TrainDataLoader, TestDataLoader
    from openke.module.model import TransE
    train = TrainDataLoader(in_path=wn18_dir, nbatches=20, threads=8, bern_flag=0, filter_flag=0, neg_ent=1, random_seed=123)
    test = TestDataLoader(wn18_dir, "link")
    pu = Parallel_Universe_Config(train_dataloader=train, test_dataloader=test, embedding_model=TransE,
                                  embedding_model_param={"dim": 20, "p_norm": 1, "norm_flag": 1},
                                  checkpoint_dir=util.GOLDEN + "/", valid_steps=10 ** 9, save_steps=None)
    pu.load_parameters("putranse_reference_layout.ckpt")
    assert pu.next_universe_id == 3 and sorted(pu._where) == [0, 1, 2]
    assert (pu.min_triple_constraint, pu.max_triple_constraint, pu.min_lr, pu.max_lr) == (500, 2000, 0.001, 0.1)
    ck = pu._chunks[0]
    assert ck.tables["ent_embeddings"].shape == (int(ck.nE.sum()), 20) and ck.ent_remap.min() >= 0
    g = np.load(os.path.join(util.GOLDEN, "putranse_wn18.npz"))      # the same three universes, minted earlier from the same seeds
    for u in range(3):
        assert np.array_equal(ck.ent_remap[ck.eoff[u]:ck.eoff[u + 1]], g["u%d_ent_remap" % u])
        assert np.array_equal(ck.rel_remap[ck.roff[u]:ck.roff[u + 1]], g["u%d_rel_remap" % u])
    out = str(tmp_path / "ours_reference_layout.ckpt")
    pu.save_parameters(out, layout="reference")
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    res = subprocess.run([sys.executable, os.path.join(util.REPO, "tests", "_refckpt_child.py"), str(tmp_path / "refpkg"),
                          os.path.join(util.REPO, "oracle", "_ref", "Base.so"), wn18_dir,
                          os.path.join(util.GOLDEN, "putranse_reference_layout.ckpt"), out],
                         capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert res.returncode == 0 and "REFCKPT-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-3000:]
    # and the flat layout holds the same ensemble
    pu.save_parameters(str(tmp_path / "flat.ckpt"))
    p2 = Parallel_Universe_Config(train_dataloader=train, test_dataloader=test, embedding_model=TransE,
                                  embedding_model_param={"dim": 20, "p_norm": 1, "norm_flag": 1},
                                  checkpoint_dir=str(tmp_path) + "/", valid_steps=10 ** 9, save_steps=None)
    p2.load_parameters("flat.ckpt")
    assert torch.equal(p2._chunks[0].tables["ent_embeddings"], ck.tables["ent_embeddings"]) and np.array_equal(p2._chunks[0].ent_remap, ck.ent_remap)


def test_builder_at_scale_matches_oracle_on_power_law_graph(tmp_path):
    """the universe builder on a large hub-dominated graph (the S1 generator of
    tools/synth.py scaled to 200 000 entities / 2 M triples / 400 relations: relation entity sets of tens of
    thousands, hubs with > 10^5 incident triples), bit-exact against the oracle's plain restatement of
    UniverseConstructor.h — whose std::set copies and std::advance picks take seconds per universe here while the
    builder takes about a millisecond."""
    import sys
    import time
    sys.path.insert(0, os.path.join(util.REPO, "tools"))
    import synth
    from oracle import native as on
    g = synth.power_law_graph(200_000, 400, 2_000_000 + 2000, seed=99)
    tr, va, te = g[:2_000_000], g[2_000_000:2_001_000], g[2_001_000:]
    path = synth.write_dataset(str(tmp_path / "big"), tr, va, te, 200_000, 400)
    L = N.lib()
    _load(L, path)
    assert L.getTrainTotal() == 2_000_000
    cases = [(31, 1999, 0.5), (32, 640, 0.25), (33, 1500, 0.37)]
    n = len(cases)
    seeds = np.array([c[0] for c in cases], np.int64)
    tcs = np.array([c[1] for c in cases], np.int64)
    bals = np.array([c[2] for c in cases], np.float32)
    t0 = time.perf_counter()
    h = L.pk_universes_build(n, N.addr(seeds), N.addr(tcs), N.addr(bals), 1)
    dt = time.perf_counter() - t0
    assert h, N.last_error()
    nT, nE, nR, foc = (np.zeros(n, np.int64) for _ in range(4))
    N.check(L.pk_universes_sizes(h, N.addr(nT), N.addr(nE), N.addr(nR), N.addr(foc)))
    er, rr = np.zeros(nE.sum(), np.int32), np.zeros(nR.sum(), np.int32)
    bg = np.zeros((nT.sum(), 3), np.int32)
    N.check(L.pk_universes_export(h, None, None, N.addr(bg), N.addr(er), N.addr(rr), None, None, None))
    L.pk_universes_free(h)
    assert dt < 1.0, "three universes took %.3f s" % dt
    o = on.Oracle(threads=8, bern=0)
    o.import_train(tr, 200_000, 400)
    eo, ro, to = (np.concatenate([[0], np.cumsum(x)]) for x in (nE, nR, nT))
    for i, (seed, tc, bal) in enumerate(cases):
        o.seed(seed)
        tri, oer, orr = o.universe(tc, float(np.float32(bal)))
        assert np.array_equal(bg[to[i]:to[i + 1]], tri), (i, seed)
        assert np.array_equal(er[eo[i]:eo[i + 1]], oer) and np.array_equal(rr[ro[i]:ro[i + 1]], orr)
    # leave the process-global state on a small graph again
    _load(L, util.write_dataset(str(tmp_path / "tiny"), [[0, 1, 0], [1, 2, 0]], [[0, 2, 0]], [[2, 0, 0]], 3, 1))


def _incremental_dataset(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(util.REPO, "tools"))
    import synth
    path = str(tmp_path / "evolve") + "/"
    states = synth.incremental_dataset(path)
    return path, states


def test_incremental_training_list_and_universes_bit_exact_with_reference(tmp_path, golden):
    """SURVEY.md 8(f) rank 3, training side.  IncrementalTrainDataLoader.load_snapshot (reference
    IncrementalTrainDataLoader.py:59-81 -> initializeTrainingOperations + evolveTrainList, Incremental.h:299-321,
    798-846) on the evolving graph of tools/synth.py: the training list with its duplicate records, the ORDER of the
    reference's three relation arrays (relations that vanish and come back move to the end), the freshly computed
    Bernoulli means, the loader's batch bookkeeping and deleted-triple set, and universes whose focus is drawn from
    the currently contained relations (UniverseConstructor.h:336-339) — all against vectors minted from the
    unmodified reference (make_golden.py incremental)."""
    from openke.data import IncrementalTrainDataLoader
    g = golden["incremental"]
    path, states = _incremental_dataset(tmp_path)
    dl = IncrementalTrainDataLoader(in_path=path, nbatches=20, threads=8, sampling_mode="normal", bern_flag=0, filter_flag=0,
                                    neg_ent=1, neg_rel=0, random_seed=4, incremental_setting=True, num_snapshots=3)
    L = dl.lib
    assert [dl.entTotal, dl.relTotal] == g["ent_rel_total"].tolist()

    def listing(which):
        n = L.pk_incremental_list(which, None)
        out = np.zeros(max(n, 1), np.int32)
        L.pk_incremental_list(which, N.addr(out))
        return out[:n]

    for s_ in (1, 2, 3):
        dl.load_snapshot(s_)
        n = L.getTrainTotal()
        by_head = np.zeros((n, 3), np.int32)
        lm, rm = np.zeros(dl.relTotal, np.float32), np.zeros(dl.relTotal, np.float32)
        N.check(L.pk_train_index(N.addr(by_head), None, N.addr(lm), N.addr(rm)))
        assert np.array_equal(by_head, g["s%d_train" % s_]), s_
        assert sorted(map(tuple, by_head[:, [0, 2, 1]].tolist())) == states[s_ - 1]          # and the generator's own replay
        assert np.array_equal(listing(0), g["s%d_rel_contained" % s_]), (s_, listing(0))
        assert np.array_equal(listing(1), g["s%d_rel_all" % s_])
        assert np.array_equal(listing(2), g["s%d_rel_deleted" % s_])
        present = np.unique(by_head[:, 1])
        assert np.array_equal(lm[present], g["s%d_left_mean" % s_][present]) and np.array_equal(rm[present], g["s%d_right_mean" % s_][present])
        assert [dl.tripleTotal, dl.batch_size, dl.nbatches, len(dl.deleted_triple_set)] == g["s%d_loader" % s_].tolist()
        for i, (seed, tc, bal) in enumerate(g["universe_cases"]):
            L.setRandomSeed(int(seed) + 10 * s_)
            L.randReset()
            dl.compile_universe_dataset(int(tc), float(bal))
            assert (L.getTrainTotalUniverse(), L.getEntityTotalUniverse(), L.getRelationTotalUniverse()) == tuple(g["s%d_u%d_sizes" % (s_, i)])
            er, rr = dl.get_universe_mappings()
            tri = np.zeros((L.getTrainTotalUniverse(), 3), np.int32)
            N.check(L.pk_universe_triples(N.addr(tri)))
            assert np.array_equal(er, g["s%d_u%d_ent_remap" % (s_, i)]) and np.array_equal(rr, g["s%d_u%d_rel_remap" % (s_, i)])
            assert np.array_equal(tri, g["s%d_u%d_triples_global" % (s_, i)])
            dl.reset_universe()
            assert dl.batch_size == g["s%d_loader" % s_][1]
    assert len(np.unique(g["s3_train"], axis=0)) < len(g["s3_train"])      # the duplicate record is there
    # the threaded builder sees the same incremental graph
    cases = g["universe_cases"]
    seeds = np.ascontiguousarray(cases[:, 0] + 30, np.int64)
    tcs, bals = np.ascontiguousarray(cases[:, 1], np.int64), np.ascontiguousarray(cases[:, 2], np.float32)
    h = L.pk_universes_build(len(cases), N.addr(seeds), N.addr(tcs), N.addr(bals), 3)
    assert h, N.last_error()
    nT, nE, nR, foc = (np.zeros(len(cases), np.int64) for _ in range(4))
    N.check(L.pk_universes_sizes(h, N.addr(nT), N.addr(nE), N.addr(nR), N.addr(foc)))
    er = np.zeros(nE.sum(), np.int32)
    N.check(L.pk_universes_export(h, None, None, None, N.addr(er), None, None, None, None))
    L.pk_universes_free(h)
    assert np.array_equal(er, np.concatenate([g["s3_u%d_ent_remap" % i] for i in range(len(cases))]))
    assert set(foc.tolist()) <= set(g["s3_rel_contained"].tolist())
    # back to the static setting: the next static loader starts from a clean slate
    L.pk_incremental_reset()
    _load(L, util.write_dataset(str(tmp_path / "tiny"), [[0, 1, 0], [1, 2, 0]], [[0, 2, 0]], [[2, 0, 0]], 3, 1))
    assert L.getTrainTotal() == 2
