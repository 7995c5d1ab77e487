"""Pins the oracle (oracle/) against vectors minted by the UNMODIFIED reference
(tests/golden/*.npz, see tests/golden/make_golden.py) and, when it was built, against the reference
library oracle/_ref/Base.so itself on fresh inputs.  CPU only."""
import ctypes
import os

import numpy as np
import pytest

import util
from oracle import native as on
from oracle.model_math import TorchOracle, closed_form_grads
from oracle import putranse_eval


def _wn18_oracle(golden, threads=8, bern=0):
    g = golden["wn18"]
    o = on.Oracle(threads=threads, bern=bern)
    o.import_train(g["train"], int(g["n_ent"]), int(g["n_rel"]))
    return o


def test_sampler_restatement_matches_reference_kats(golden):
    g = golden["sampler"]
    w = golden["wn18"]
    o = on.Oracle(threads=8, bern=0)
    # one oracle state imported 4 times = the golden process (import-count drift of the bern table)
    for name, bern, filt, k in (("b0f0k1", 0, 0, 1), ("b1f1k1", 1, 1, 1), ("b0f1k2", 0, 1, 2), ("b1f0k3", 1, 0, 3)):
        o.L.oracle_set(o.o, 8, bern)
        o.import_train(w["train"], 40943, 18)
        l, r = o.means()
        assert np.array_equal(l, g[name + "_left_mean"]) and np.array_equal(r, g[name + "_right_mean"])
        o.seed(4)
        for call in range(3):
            h, t, rr = o.sampling(1414, k, filt)
            assert np.array_equal(h, g[name][call, 0]), (name, call)
            assert np.array_equal(t, g[name][call, 1]), (name, call)
            assert np.array_equal(rr, g[name][call, 2]), (name, call)
    # SURVEY.md 8(c) known answers
    assert g["b0f0k1"][0, 0, :4].tolist() == [9340, 12095, 8467, 40541]
    assert g["b0f0k1"][0, 1, 1414:1418].tolist() == [36783, 32196, 29345, 19500]
    o3 = on.Oracle(threads=3, bern=0)
    o3.import_train(w["train"], 40943, 18)
    o3.seed(11)
    h, t, rr = o3.sampling(1001, 2, 1)
    assert np.array_equal(np.stack([h, t, rr]), g["odd_B1001_t3_seed11_f1k2"])


def test_universe_restatement_matches_reference(golden):
    U = golden["universe"]
    o = _wn18_oracle(golden)
    for i, (seed, tc, bal) in enumerate(U["cases"]):
        o.seed(int(seed))
        tri, er, rr = o.universe(int(tc), float(bal))
        assert (tri.shape[0], er.shape[0], rr.shape[0]) == tuple(U["u%d_sizes" % i])
        assert np.array_equal(tri, U["u%d_triples_global" % i])
        assert np.array_equal(er, U["u%d_ent_remap" % i]) and np.array_equal(rr, U["u%d_rel_remap" % i])
        o.swap()
        assert np.array_equal(o.train_list(), U["u%d_triples_local" % i])
        l, r = o.means()
        assert np.array_equal(l, U["u%d_left_mean" % i]) and np.array_equal(r, U["u%d_right_mean" % i])
        for call in range(2):   # sampler inside the universe (local ids)
            h, t, r2 = o.sampling(tri.shape[0] // 20, 1, 0)
            assert np.array_equal(np.stack([h, t, r2]), U["u%d_batches" % i][call])
        o.swap()


TRAIN = [("transe_l1_adagrad", "transe", 1, "adagrad"), ("transe_l2_sgd", "transe", 2, "sgd"),
         ("transh_l1_adagrad", "transh", 1, "adagrad"), ("transd_l1_adagrad", "transd", 1, "adagrad"),
         ("transe_l1_sgd_d50_k3", "transe", 1, "sgd")]


def _tables(g, name, which):
    out = {}
    for key in g.files:
        pre = "%s_%s_" % (name, which)
        if key.startswith(pre):
            out[key[len(pre):]] = g[key]
    return out


@pytest.mark.parametrize("name,model,p,opt", TRAIN)
def test_torch_oracle_reproduces_reference_trainer(golden, name, model, p, opt):
    """Same torch calls as the reference => same losses and tables (same machine: bit-equal)."""
    import torch
    torch.set_num_threads(1)
    g = golden["train"]
    lr, margin = g[name + "_hyper"]
    k = int(g[name + "_sizes"][3])
    o = TorchOracle(model, _tables(g, name, "init"), p_norm=p, opt=opt, lr=float(lr), margin=float(margin), k=k)
    losses = [o.step(*b) for b in g[name + "_batches"]]
    assert np.allclose(losses, g[name + "_losses"], rtol=1e-6, atol=1e-7)
    final = _tables(g, name, "final")
    for n, v in o.tables().items():
        assert np.allclose(v, final[n], rtol=1e-5, atol=1e-6), n


@pytest.mark.parametrize("name,model,p,opt", TRAIN)
def test_closed_form_gradients_match_autograd(golden, name, model, p, opt):
    g = golden["train"]
    _, margin = g[name + "_hyper"]
    k = int(g[name + "_sizes"][3])
    tabs = _tables(g, name, "init")
    o = TorchOracle(model, tabs, p_norm=p, opt="sgd", lr=1.0, margin=float(margin), k=k)
    bh, bt, br = g[name + "_batches"][0]
    o.opt.zero_grad()
    loss = o.loss(bh, bt, br)
    loss.backward()
    cl, G = closed_form_grads(model, tabs, bh, bt, br, k, float(margin), p)
    assert abs(cl - float(loss)) < 2e-6
    for n, t in o.t.items():
        assert np.allclose(t.grad.numpy(), G[n], rtol=2e-4, atol=2e-6), n


def test_rank_row_restatement_matches_reference_ranks(golden):
    """testHead/testTail restatement on the shipped TransH/WN18 checkpoint scores."""
    import torch
    torch.set_num_threads(4)
    g, w = golden["rank_transh_wn18"], golden["wn18"]
    o = _wn18_oracle(golden)
    o.import_test(w["test"], w["train"], w["valid"])
    assert np.array_equal(o.eval_list(0), g["test_sorted"])
    tabs = {n: g[n] for n in ("ent_embeddings", "rel_embeddings", "norm_vector")}
    E = 40943
    for p in (1, 2):
        m = TorchOracle("transh", tabs, p_norm=p)
        for idx in (0, 1, 2, 17, 333, 2500, 4999):
            h, r, t = g["test_sorted"][idx].tolist()
            with torch.no_grad():
                head = np.concatenate([[h], np.delete(np.arange(E), h)])
                sc = m.score(head, [t], [r], "head_batch").numpy()
                raw, filt = o.rank_row(0, sc, idx, True)
                tail = np.concatenate([[t], np.delete(np.arange(E), t)])
                sc2 = m.score([h], tail, [r], "tail_batch").numpy()
                raw2, filt2 = o.rank_row(0, sc2, idx, False)
            assert [raw, filt, raw2, filt2] == g["ranks_p%d" % p][idx].tolist(), (p, idx)


def test_putranse_aggregation_restatement_matches_reference(golden):
    g, w = golden["putranse_wn18"], golden["wn18"]
    n_univ = int(g["n_univ"])
    spaces = [dict(tables={"ent_embeddings": g["u%d_ent" % u], "rel_embeddings": g["u%d_rel" % u]},
                   ent_remap=g["u%d_ent_remap" % u], rel_remap=g["u%d_rel_remap" % u]) for u in range(n_univ)]
    o = _wn18_oracle(golden)
    o.import_test(w["test"], w["train"], w["valid"])
    tri = g["test_sorted"]
    allt = np.concatenate([w["train"], w["valid"], w["test"]])[:, [0, 2, 1]]
    # triples whose truth is embedded somewhere are the interesting ones; take those plus a few misses
    finite = np.nonzero((g["ranks"][:, 0] < 40943) | (g["ranks"][:, 2] < 40943))[0]
    picks = list(finite[:25]) + [1, 2, 3]
    for idx in picks:
        h, r, t = tri[idx].tolist()
        en = putranse_eval.universe_energies(spaces, 40943, t, r, 0)
        known = sorted(set(allt[(allt[:, 1] == r) & (allt[:, 2] == t)][:, 0].tolist()) - {h})
        raw, filt = putranse_eval.rank_from_energy(en, h, known)
        # the C restatement of testHead on the same row agrees with the numpy one
        assert (raw, filt) == o.rank_row(0, on.candidate_row(en, h), idx, True)
        en2 = putranse_eval.universe_energies(spaces, 40943, h, r, 1)
        known2 = sorted(set(allt[(allt[:, 0] == h) & (allt[:, 1] == r)][:, 2].tolist()) - {t})
        raw2, filt2 = putranse_eval.rank_from_energy(en2, t, known2)
        assert [raw, filt, raw2, filt2] == g["ranks"][idx].tolist(), idx


def test_null_vector_restatement_matches_reference(golden):
    """missing_embedding_handling='null_vector': per-triple ranks of the unmodified reference on the same
    eight universes (tests/golden/putranse_nullvec_wn18.npz, minted by make_golden.py putranse_nullvec)."""
    g, gn, w = golden["putranse_wn18"], golden["putranse_nullvec_wn18"], golden["wn18"]
    assert np.array_equal(g["test_sorted"], gn["test_sorted"])
    n_univ = int(g["n_univ"])
    spaces = [dict(tables={"ent_embeddings": g["u%d_ent" % u], "rel_embeddings": g["u%d_rel" % u]},
                   ent_remap=g["u%d_ent_remap" % u], rel_remap=g["u%d_rel_remap" % u]) for u in range(n_univ)]
    tri = g["test_sorted"]
    allt = np.concatenate([w["train"], w["valid"], w["test"]])[:, [0, 2, 1]]
    changed = np.nonzero((g["ranks"] != gn["ranks"]).any(1))[0]
    assert changed.size > 0, "the null-vector run must differ from the last-rank run somewhere"
    same = np.nonzero((g["ranks"] == gn["ranks"]).all(1))[0]
    for idx in list(changed[:20]) + list(same[:5]):
        h, r, t = tri[idx].tolist()
        en = putranse_eval.fill_missing(putranse_eval.universe_energies(spaces, 40943, t, r, 0), putranse_eval.tuple_score(spaces, t, r, 0))
        known = sorted(set(allt[(allt[:, 1] == r) & (allt[:, 2] == t)][:, 0].tolist()) - {h})
        raw, filt = putranse_eval.rank_from_energy(en, h, known)
        en2 = putranse_eval.fill_missing(putranse_eval.universe_energies(spaces, 40943, h, r, 1), putranse_eval.tuple_score(spaces, h, r, 1))
        known2 = sorted(set(allt[(allt[:, 0] == h) & (allt[:, 1] == r)][:, 2].tolist()) - {t})
        raw2, filt2 = putranse_eval.rank_from_energy(en2, t, known2)
        assert [raw, filt, raw2, filt2] == gn["ranks"][idx].tolist(), idx


@pytest.mark.skipif(not os.path.exists(on.REF_SO), reason="oracle/_ref/Base.so not built")
def test_restatement_against_live_reference_library(tmp_path):
    """Fresh synthetic graph, seeds never seen by the golden files: oracle == reference library."""
    tr, va, te = util.synthetic_graph(2000, 9, 15000, 100, seed=99)
    path = util.write_dataset(str(tmp_path / "syn"), tr, va, te, 2000, 9)
    R = on.load_reference()
    R.setInPath(ctypes.create_string_buffer(path.encode(), len(path) * 2))
    R.setBern(1)
    R.setWorkThreads(5)
    R.setRandomSeed(31)
    R.randReset()
    R.importTrainFiles()
    o = on.Oracle(threads=5, bern=1)
    o.import_train(tr, 2000, 9)
    o.seed(31)
    for B, k, filt in ((333, 2, 1), (100, 1, 0), (64, 3, 1)):
        n = B * (1 + k)
        h, t, r, y = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.float32)
        R.sampling(on._addr(h), on._addr(t), on._addr(r), on._addr(y), B, k, 0, 0, filt, 0, 0)
        oh, ot, orr = o.sampling(B, k, filt)
        assert np.array_equal(h, oh) and np.array_equal(t, ot) and np.array_equal(r, orr)
    for seed, tc, bal in ((41, 300, 0.3), (42, 777, 0.45)):
        R.setRandomSeed(seed)
        R.randReset()
        R.getParallelUniverse(tc, bal)
        ne, nr = R.getEntityTotalUniverse(), R.getRelationTotalUniverse()
        er, rr = np.zeros(ne, np.int64), np.zeros(nr, np.int64)
        R.getEntityRemapping(on._addr(er))
        R.getRelationRemapping(on._addr(rr))
        R.resetUniverse()
        o.seed(seed)
        tri, oer, orr2 = o.universe(tc, bal)
        assert np.array_equal(er, oer) and np.array_equal(rr, orr2)
