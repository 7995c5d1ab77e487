"""Parity at PRODUCTION length (BASELINE.json configs[1]) and the arithmetic of the batched-universe
kernel in isolation.

* tests/golden/putranse_full_wn18.npz was minted by the UNMODIFIED reference (make_golden.py
  putranse_full): the static PuTransE experiment on WN18, universes 0..23 (seeds 4..27) at their drawn
  50-199 epochs = 1 080-3 940 Adagrad steps each, per-step losses of the reference Trainer, final
  tables, and the reference's own link-prediction ranks and metrics of the 24-universe ensemble.
* the closed-form check runs 1, 2 and 3 steps of the batched-universe kernel (the FAST Adagrad path
  with its rsqrt.approx / rcp.approx update) against a float64 replay, so approximation error is
  separated from the chaotic divergence that long L1 trajectories show.

Tolerances (north_star: "within a stated fp32 tolerance") are stated where they are asserted.
"""
import ctypes
import os

import numpy as np
import pytest

import util

N = util.native()

STATIC = dict(min_margin=1, max_margin=4, min_lr=0.001, max_lr=0.1, min_num_epochs=50, max_num_epochs=200,
              min_triple_constraint=500, max_triple_constraint=2000, min_balance=0.25, max_balance=0.5)


def _static_putranse(wn18_dir, model_cls=None, param=None, nbatches=20, **over):
    from openke.config import Parallel_Universe_Config
    from openke.data import TrainDataLoader, TestDataLoader
    from openke.module.model import TransE
    train = TrainDataLoader(in_path=wn18_dir, nbatches=nbatches, threads=8, sampling_mode="normal", bern_flag=0, filter_flag=0,
                            neg_ent=1, neg_rel=0, random_seed=123)
    test = TestDataLoader(train.in_path, "link")
    kw = dict(STATIC)
    kw.update(over)
    pu = Parallel_Universe_Config(training_identifier="full", train_dataloader=train, test_dataloader=test,
                                  initial_num_universes=None, embedding_model=model_cls or TransE,
                                  embedding_model_param=param or {"dim": 20, "p_norm": 1, "norm_flag": 1},
                                  checkpoint_dir=None, valid_steps=10 ** 9, save_steps=None, training_setting="static",
                                  incremental_strategy=None, **kw)
    return pu


def _epoch_means(x, nb):
    n = len(x) // nb
    return np.asarray(x[:n * nb], dtype=np.float64).reshape(n, nb).mean(1)


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_full_length_universes_follow_the_reference(wn18_dir, golden):
    """24 universes x (50..199 epochs x 20 batches), exactly the reference's run.

    Integer work is bit-exact: subgraphs, id maps, steps per universe, which test triples no universe can score,
    and their ranks (the +inf branch of Test.h:181-206).
    Floating point: a universe's trajectory is a chaotic map (an L1 energy flips sign(s_i) when an element crosses
    zero, a hinge term switches on or off), so two correct fp32 implementations follow each other for a while and
    then stay statistically, not pointwise, equal.  Measured on the B200 (tools/parity_probe.py,
    profiles/r2_parity_probe.log): 18 of 24 universes never leave 1e-3 relative over their whole 1 000-4 000 steps,
    the first 100 steps of 23 universes agree to 1.3e-6; the high-learning-rate universes 10 (lr 0.094) and 19
    (lr 0.06, margin 1) separate after 178 and 15 steps.  Stated tolerances, as a function of the step s:
      s < 10         per-step loss rtol 2e-6, every universe
      s < 400        per-step loss rtol 1e-4 for at least 18 of the 24 universes (21 measured)
      every epoch    |mean loss of the epoch - reference's| <= 0.015 + 0.05 * reference's (largest measured: 0.0092),
                     and for at least 14 universes <= 1e-3 in every epoch (19 measured)
      last 10 epochs |mean loss - reference's| <= 3 % + 0.001 per universe (largest measured: 2 %), <= 1 % on average
      trained entity tables: median row norm within 3 % of the reference's
      ensemble       filtered MRR within 3 % relative (measured 0.8 %), MR within 1e-4 relative (1.7e-5), Hits@10/3/1
                     within 0.002 absolute (0.0007) of the reference Tester's values (Test.h:450-454); of the ranks
                     some universe scores, >= 55 % identical (66 %) and >= 70 % within +-3 (80 %).
    """
    g = golden["putranse_full_wn18"]
    n_univ, nb = int(g["n_univ"]), int(g["nbatches"])
    pu = _static_putranse(wn18_dir)
    pu.record_losses = True
    assert pu.initial_random_seed == int(g["initial_seed"]) == 4
    pu.train_parallel_universes(n_univ)
    tails, ref_tails, close_400, close_epochs = [], [], 0, 0
    for u in range(n_univ):
        er = np.array(sorted(pu.entity_id_mappings[u], key=pu.entity_id_mappings[u].get))
        rr = np.array(sorted(pu.relation_id_mappings[u], key=pu.relation_id_mappings[u].get))
        assert np.array_equal(er, g["u%d_ent_remap" % u]) and np.array_equal(rr, g["u%d_rel_remap" % u]), u
        got, want = pu.universe_losses[u], g["u%d_losses" % u]
        assert len(got) == len(want) == pu.universe_hyper[u]["epochs"] * nb, u
        assert np.allclose(got[:10], want[:10], rtol=2e-6), (u, got[:10], want[:10])
        close_400 += bool(np.allclose(got[:400], want[:400], rtol=1e-4, atol=1e-7))
        ge, we = _epoch_means(got, nb), _epoch_means(want, nb)
        assert np.all(np.abs(ge - we) <= 0.015 + 0.05 * we), (u, np.abs(ge - we).max(), int(np.argmax(np.abs(ge - we))))
        close_epochs += bool(np.all(np.abs(ge - we) <= 1e-3))
        gt, wt = ge[-10:].mean(), we[-10:].mean()
        assert abs(gt - wt) <= 0.03 * wt + 0.001, (u, gt, wt)
        tails.append(gt)
        ref_tails.append(wt)
        sp = pu.trained_embedding_spaces[u]
        ent = sp.ent_embeddings.weight.detach().cpu().numpy()
        assert ent.shape == g["u%d_ent_embeddings" % u].shape and np.isfinite(ent).all()
        n_got, n_ref = np.linalg.norm(ent, axis=1), np.linalg.norm(g["u%d_ent_embeddings" % u], axis=1)
        assert abs(np.median(n_got) / np.median(n_ref) - 1) < 0.03, (u, np.median(n_got), np.median(n_ref))
    assert close_400 >= 18, close_400
    assert close_epochs >= 14, close_epochs
    assert abs(np.mean(tails) / np.mean(ref_tails) - 1) < 0.01, (np.mean(tails), np.mean(ref_tails))

    mrr, mr, hit10, hit3, hit1 = pu.run_link_prediction()
    ranks, want = pu.last_ranks, g["ranks"]
    E = 40943
    missing, missing_t = want[:, 0] == E, want[:, 2] == E
    assert np.array_equal(ranks[:, 0] == E, missing) and np.array_equal(ranks[:, 2] == E, missing_t)
    assert np.array_equal(ranks[missing][:, [0, 1]], want[missing][:, [0, 1]])        # truth in no universe: integer logic
    assert np.array_equal(ranks[missing_t][:, [2, 3]], want[missing_t][:, [2, 3]])
    scored = np.concatenate([np.abs(ranks[~missing][:, 1] - want[~missing][:, 1]), np.abs(ranks[~missing_t][:, 3] - want[~missing_t][:, 3])])
    assert (scored == 0).mean() >= 0.55 and (scored <= 3).mean() >= 0.70, ((scored == 0).mean(), (scored <= 3).mean())
    ref = g["metrics"]   # mrr, mr, hit10, hit3, hit1
    assert abs(mrr - ref[0]) <= 0.03 * ref[0], (mrr, ref[0])
    assert abs(mr - ref[1]) <= 1e-4 * ref[1], (mr, ref[1])
    for got_h, want_h in ((hit10, ref[2]), (hit3, ref[3]), (hit1, ref[4])):
        assert abs(got_h - want_h) <= 0.002, (got_h, want_h)


def _fb15k_shape_dir(tmp_path):
    import sys
    sys.path.insert(0, os.path.join(util.REPO, "tools"))
    import synth
    tr, va, te, ne, nr = synth.fb15k_shape(n_valid=2000, n_test=2000)
    return synth.write_dataset(str(tmp_path / "fb15k_shape"), tr, va, te, ne, nr), synth.checksum(tr)


@pytest.mark.gpu
def test_full_length_fb15k_shape_follows_the_reference(tmp_path, golden):
    """BASELINE.json configs[3] shape at production length: eight universes of the FB15K-shaped synthetic graph
    (E = 14 951, R = 1 345; universes hold 360-550 relations, so the relation side of the batched kernel is what is
    exercised) at their drawn epochs, against the unmodified reference (make_golden.py putranse_full_fb15k).
    Same statement of tolerances as the WN18 test; measured values in profiles/r2_parity_probe_fb15k.log."""
    g = golden["putranse_full_fb15k"]
    path, checksum = _fb15k_shape_dir(tmp_path)
    assert checksum == int(g["train_checksum"]), "tools/synth.py generated a different graph than the one the reference was run on"
    n_univ, nb = int(g["n_univ"]), int(g["nbatches"])
    pu = _static_putranse(path)
    pu.record_losses = True
    pu.train_parallel_universes(n_univ)
    assert pu.universes_on_single_space_path == 0        # all of them fit the batched-universe kernel
    tails, ref_tails, close_400 = [], [], 0
    for u in range(n_univ):
        er = np.array(sorted(pu.entity_id_mappings[u], key=pu.entity_id_mappings[u].get))
        rr = np.array(sorted(pu.relation_id_mappings[u], key=pu.relation_id_mappings[u].get))
        assert np.array_equal(er, g["u%d_ent_remap" % u]) and np.array_equal(rr, g["u%d_rel_remap" % u]), u
        got, want = pu.universe_losses[u], g["u%d_losses" % u]
        assert len(got) == len(want), u
        assert np.allclose(got[:10], want[:10], rtol=2e-6), (u, got[:10], want[:10])
        close_400 += bool(np.allclose(got[:400], want[:400], rtol=1e-4, atol=1e-7))
        ge, we = _epoch_means(got, nb), _epoch_means(want, nb)
        assert np.all(np.abs(ge - we) <= 0.015 + 0.05 * we), (u, np.abs(ge - we).max(), int(np.argmax(np.abs(ge - we))))
        gt, wt = ge[-10:].mean(), we[-10:].mean()
        assert abs(gt - wt) <= 0.03 * wt + 0.001, (u, gt, wt)
        tails.append(gt)
        ref_tails.append(wt)
        ent = pu.trained_embedding_spaces[u].ent_embeddings.weight.detach().cpu().numpy()
        n_got, n_ref = np.linalg.norm(ent, axis=1), g["u%d_ent_embeddings_rownorm" % u]
        assert n_got.shape == n_ref.shape and abs(np.median(n_got) / np.median(n_ref) - 1) < 0.03, u
    assert close_400 >= 5, close_400
    assert abs(np.mean(tails) / np.mean(ref_tails) - 1) < 0.01
    mrr, mr, hit10, hit3, hit1 = pu.run_link_prediction()
    ranks, want = pu.last_ranks, g["ranks"]
    E = 14951
    missing, missing_t = want[:, 0] == E, want[:, 2] == E
    assert np.array_equal(ranks[:, 0] == E, missing) and np.array_equal(ranks[:, 2] == E, missing_t)
    assert np.array_equal(ranks[missing][:, [0, 1]], want[missing][:, [0, 1]]) and np.array_equal(ranks[missing_t][:, [2, 3]], want[missing_t][:, [2, 3]])
    scored = np.concatenate([np.abs(ranks[~missing][:, 1] - want[~missing][:, 1]), np.abs(ranks[~missing_t][:, 3] - want[~missing_t][:, 3])])
    # universes of ~800 entities with dense energies: a rank moves by tens of places for a last-bit change of a table,
    # so the rank-by-rank statement is loose here (measured: 54 % within 10, 93 % within 100) and the metrics carry it:
    # MRR within 5e-4 absolute (measured 3e-4 at 0.0044), MR within 1e-3 relative (1.7e-4), Hits within 0.002 (5e-4)
    assert (scored <= 100).mean() >= 0.85, (scored <= 100).mean()
    ref = g["metrics"]
    assert abs(mrr - ref[0]) <= 5e-4 and abs(mr - ref[1]) <= 1e-3 * ref[1], (mrr, mr, ref)
    for got_h, want_h in ((hit10, ref[2]), (hit3, ref[3]), (hit1, ref[4])):
        assert abs(got_h - want_h) <= 0.002, (got_h, want_h)


# ------------------------------------------------------------------------------------------------
def _one_universe_launch(wn18_dir, model_name, param, seed, tc, bal, B, steps, lr, margin):
    """`steps` steps of pk_train_universes on ONE universe with a hand-made descriptor (epochs = 1,
    nbatches = steps, batch_size = B), from the reference's initial tables."""
    import torch
    import openke.module.model as M
    from openke.data import TrainDataLoader
    dl = TrainDataLoader(in_path=wn18_dir, nbatches=20, threads=8, bern_flag=0, filter_flag=0, neg_ent=1, random_seed=4)
    L = dl.lib
    seeds, tcs, bals = np.array([seed], np.int64), np.array([tc], np.int64), np.array([bal], np.float32)
    h = L.pk_universes_build(1, N.addr(seeds), N.addr(tcs), N.addr(bals), 1)
    assert h
    nT, nE, nR, focus = (np.zeros(1, np.int64) for _ in range(4))
    N.check(L.pk_universes_sizes(h, N.addr(nT), N.addr(nE), N.addr(nR), N.addr(focus)))
    by_head = np.zeros((int(nT[0]), 3), np.int32)
    lcg = np.zeros((1, 8), np.uint64)
    N.check(L.pk_universes_export(h, N.addr(by_head), None, None, None, None, None, None, N.addr(lcg)))
    L.pk_universes_free(h)
    torch.manual_seed(seed)
    model = getattr(M, model_name)(int(nE[0]), int(nR[0]), **param)
    init = {n: getattr(model, n).weight.detach().numpy().copy() for n in model.table_names()}
    dev = torch.device("cuda", 0)
    tabs = {n: torch.from_numpy(v).to(dev) for n, v in init.items()}
    state = {n: torch.zeros_like(v) for n, v in tabs.items()}
    t = N.Tables()
    for j in range(2):
        t.ent[j] = t.rel[j] = t.ent_state[j] = t.rel_state[j] = None
    for j, n in enumerate(model._ent_tables):
        t.ent[j], t.ent_state[j] = tabs[n].data_ptr(), state[n].data_ptr()
    for j, n in enumerate(model._rel_tables):
        t.rel[j], t.rel_state[j] = tabs[n].data_ptr(), state[n].data_ptr()
    t.n_ent, t.n_rel = int(nE[0]), int(nR[0])
    darr = np.zeros(1, dtype=N.UNIVERSE_DESC_DTYPE)
    darr["n_tri"], darr["n_ent"], darr["n_rel"] = nT, nE, nR
    darr["batch_size"], darr["nbatches"], darr["epochs"] = B, steps, 1
    darr["margin"], darr["lr"] = np.float32(margin), np.float32(lr)
    darr["loss_off"] = 0
    darr["lcg"][:, :8] = lcg
    desc = (N.UniverseDesc * 1).from_buffer(darr)
    d_by_head = torch.from_numpy(by_head).to(dev)
    d_loss = torch.zeros(max(steps, 1), dtype=torch.float32, device=dev)
    cfg = model.native_cfg(opt=N.PK_ADAGRAD, neg_ent=1, bern=0, filt=0, work_threads=8)
    N.check(L.pk_train_universes(ctypes.byref(cfg), ctypes.byref(t), d_by_head.data_ptr(), None, None, None, desc, 1,
                                 d_loss.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "pk_train_universes")
    assert L.pk_last_launch_count() >= 1
    torch.cuda.synchronize()
    return init, {n: v.cpu().numpy() for n, v in tabs.items()}, {n: v.cpu().numpy() for n, v in state.items()}, d_loss.cpu().numpy()


def _float64_adagrad_replay(model_key, init, batches, lr, margin):
    from oracle.model_math import closed_form_grads
    W = {n: v.astype(np.float64) for n, v in init.items()}
    S = {n: np.zeros_like(v) for n, v in W.items()}
    losses = []
    for bh, bt, br in batches:
        loss, G = closed_form_grads(model_key, W, bh, bt, br, 1, margin, 1)
        losses.append(loss)
        for n in W:   # torch.optim.Adagrad, lr_decay = weight_decay = 0, eps = 1e-10 (reference Trainer.py:65-70)
            S[n] += G[n] * G[n]
            W[n] -= lr * G[n] / (np.sqrt(S[n]) + 1e-10)
    return W, S, losses


@pytest.mark.gpu
@pytest.mark.parametrize("model_name,param", [("TransE", {"dim": 20, "p_norm": 1, "norm_flag": 1}),
                                              ("TransH", {"dim": 20, "p_norm": 1, "norm_flag": 1}),
                                              ("TransD", {"dim_e": 20, "dim_r": 20, "p_norm": 1, "norm_flag": 1})])
@pytest.mark.parametrize("steps", [1, 2, 3])
def test_batched_universe_kernel_steps_against_float64_closed_form(wn18_dir, model_name, param, steps):
    """The FAST Adagrad path of k2_train_universes (approximate rsqrt/rcp in the update, Newton reciprocal in
    the normalisation, fixed-point relation sums, L2 reductions for repeated rows) against a float64
    replay of the reference's update rule on the same batches.  Tolerance: every table entry within 2e-6
    absolute after 1 step and 5e-6 after 3 (entries are 0.03-0.3 in magnitude, lr = 0.05), Adagrad sums
    within 1e-5 relative + 1e-9, per-step loss within 5e-6 relative."""
    from oracle import native as on
    seed, tc, bal, B, lr, margin = 9, 983, 0.37, 49, 0.05, 2.0
    init, got, got_state, losses = _one_universe_launch(wn18_dir, model_name, param, seed, tc, bal, B, steps, lr, margin)
    w = np.load(os.path.join(util.GOLDEN, "wn18.npz"))
    o = on.Oracle(threads=8, bern=0)
    o.import_train(w["train"], 40943, 18)
    o.seed(seed)
    o.universe(tc, bal)
    o.swap()
    batches = [o.sampling(B, 1, 0) for _ in range(steps)]
    o.swap()
    W, S, want_losses = _float64_adagrad_replay(model_name.lower(), init, batches, lr, margin)
    assert np.allclose(losses, want_losses, rtol=5e-6), (losses, want_losses)
    atol = 2e-6 if steps == 1 else 5e-6
    for n in W:
        err = np.abs(got[n] - W[n]).max()
        assert err <= atol, (model_name, steps, n, err)
        assert np.allclose(got_state[n], S[n], rtol=1e-5, atol=1e-9), (model_name, steps, n, np.abs(got_state[n] - S[n]).max())
        moved = np.abs(got[n] - init[n]).max()
        assert moved > 100 * atol, "the step must move the table by much more than the tolerance (%g)" % moved
