#!/usr/bin/env python
"""Mint golden vectors by running the UNMODIFIED reference in this container.

Run from the repo root (needs /root/reference and oracle/_ref/Base.so, i.e.
`make -C oracle ref` first):

    python tests/golden/make_golden.py dataset sampler universe train rank putranse

Every topic runs in a fresh interpreter because the reference keeps all state in
process-global C variables and its Bernoulli statistics depend on how many times
the process has imported the training files.

The reference Python package is imported from a scratch directory of symlinks
(/tmp/putranse_refpkg/openke -> /root/reference/openke/*, release/Base.so ->
oracle/_ref/Base.so); nothing is copied and nothing is written to /root/reference.
The outputs are small .npz files beside this script; they are what travels to the
GPU box (where /root/reference does not exist).
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
REF = "/root/reference"
WN18 = REF + "/benchmarks/WN18/"
OUT = os.path.dirname(os.path.abspath(__file__))
SCRATCH = "/tmp/putranse_refpkg"


def _setup_ref_import():
    pkg = os.path.join(SCRATCH, "openke")
    os.makedirs(os.path.join(pkg, "release"), exist_ok=True)
    for name in ("__init__.py", "base", "config", "data", "module"):
        dst = os.path.join(pkg, name)
        if not os.path.lexists(dst):
            os.symlink(os.path.join(REF, "openke", name), dst)
    so = os.path.join(pkg, "release", "Base.so")
    if os.path.lexists(so):
        os.remove(so)
    os.symlink(os.path.join(REPO, "oracle", "_ref", "Base.so"), so)
    sys.path.insert(0, SCRATCH)


class Triple(ctypes.Structure):
    _fields_ = [("h", ctypes.c_long), ("r", ctypes.c_long), ("t", ctypes.c_long)]


def _triples(lib, sym, n):
    """Read n records of the global `Triple *sym` straight out of the loaded reference .so."""
    ptr = ctypes.POINTER(Triple).in_dll(lib, sym)
    out = np.zeros((n, 3), dtype=np.int64)
    for i in range(n):
        out[i] = (ptr[i].h, ptr[i].r, ptr[i].t)
    return out


def _reals(lib, sym, n):
    ptr = ctypes.POINTER(ctypes.c_float).in_dll(lib, sym)
    return np.array([ptr[i] for i in range(n)], dtype=np.float32)


def _read_txt(path):
    return np.loadtxt(path, dtype=np.int64).reshape(-1, 3)


# --------------------------------------------------------------------------- dataset
def topic_dataset():
    """WN18 as shipped by the reference (public benchmark data, not source code), repacked."""
    tr, va, te = (_read_txt(WN18 + f) for f in ("train2id.txt", "valid2id.txt", "test2id.txt"))
    ne = sum(1 for _ in open(WN18 + "entity2id.txt"))
    nr = sum(1 for _ in open(WN18 + "relation2id.txt"))
    np.savez_compressed(os.path.join(OUT, "wn18.npz"), train=tr.astype(np.int32), valid=va.astype(np.int32),
                        test=te.astype(np.int32), n_ent=np.int64(ne), n_rel=np.int64(nr),
                        columns=np.array("h t r (file order of the header-less OpenKE-PuTransE format)"))
    print("wn18.npz", tr.shape, va.shape, te.shape, ne, nr)


# --------------------------------------------------------------------------- sampler
def topic_sampler():
    _setup_ref_import()
    from openke.data import TrainDataLoader, TestDataLoader
    res = {}
    cases = [  # name, bern, filter, k, extra imports before sampling
        ("b0f0k1", 0, 0, 1, 0), ("b1f1k1", 1, 1, 1, 0), ("b0f1k2", 0, 1, 2, 0), ("b1f0k3", 1, 0, 3, 0),
    ]
    # one process, several loaders => import count grows 1,2,3,4; record it with each case
    imports = 0
    for name, bern, filt, k, _ in cases:
        dl = TrainDataLoader(in_path=WN18, nbatches=100, threads=8, bern_flag=bern, filter_flag=filt, neg_ent=k,
                             random_seed=4)
        imports += 1
        rows = []
        for _ in range(3):
            d = dl.sampling()
            rows.append(np.stack([d["batch_h"], d["batch_t"], d["batch_r"]]).astype(np.int32))
        res[name] = np.stack(rows)  # [3 calls, 3 (h,t,r), B(1+k)]
        res[name + "_imports"] = np.int64(imports)
        res[name + "_left_mean"] = _reals(dl.lib, "left_mean", dl.relTotal)
        res[name + "_right_mean"] = _reals(dl.lib, "right_mean", dl.relTotal)
    # odd batch size (B % 8 != 0) and a different thread count
    dl = TrainDataLoader(in_path=WN18, batch_size=1001, threads=3, bern_flag=0, filter_flag=1, neg_ent=2,
                         random_seed=11)
    imports += 1
    d = dl.sampling()
    res["odd_B1001_t3_seed11_f1k2"] = np.stack([d["batch_h"], d["batch_t"], d["batch_r"]]).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), **res)
    print("sampler.npz", {k: getattr(v, "shape", v) for k, v in res.items()})


# --------------------------------------------------------------------------- universe
UNIVERSE_CASES = [(4, 1000, 0.5), (5, 1775, 0.31), (6, 1675, 0.44), (7, 600, 0.25), (11, 1999, 0.5), (12, 500, 0.37),
                  (8, 964, 0.26), (21, 300, 0.4)]


def topic_universe():
    _setup_ref_import()
    from openke.data import TrainDataLoader
    dl = TrainDataLoader(in_path=WN18, nbatches=20, threads=8, bern_flag=0, filter_flag=0, neg_ent=1, random_seed=4)
    lib = dl.lib
    res = {"cases": np.array(UNIVERSE_CASES, dtype=np.float64)}
    for i, (seed, tc, bal) in enumerate(UNIVERSE_CASES):
        lib.setRandomSeed(seed)
        lib.randReset()
        dl.compile_universe_dataset(tc, bal)
        nT, nE, nR = lib.getTrainTotalUniverse(), lib.getEntityTotalUniverse(), lib.getRelationTotalUniverse()
        er, rr = dl.get_universe_mappings()
        res[f"u{i}_sizes"] = np.array([nT, nE, nR], dtype=np.int64)
        res[f"u{i}_ent_remap"] = er.astype(np.int32)
        res[f"u{i}_rel_remap"] = rr.astype(np.int32)
        res[f"u{i}_triples_global"] = _triples(lib, "trainListUniverse", nT).astype(np.int32)      # collection order
        res[f"u{i}_triples_local"] = _triples(lib, "trainListUniverseEnum", nT).astype(np.int32)   # sorted (h,r,t)
        res[f"u{i}_left_mean"] = _reals(lib, "left_meanUniverse", nR)
        res[f"u{i}_right_mean"] = _reals(lib, "right_meanUniverse", nR)
        # two sampler calls inside the universe (local ids), k=1, unfiltered, as the PuTransE scripts do
        dl.swap_helpers()
        rows = []
        for _ in range(2):
            d = dl.sampling()
            rows.append(np.stack([d["batch_h"], d["batch_t"], d["batch_r"]]).astype(np.int32))
        res[f"u{i}_batches"] = np.stack(rows)
        dl.reset_universe()
    np.savez_compressed(os.path.join(OUT, "universe.npz"), **res)
    print("universe.npz", [res[f"u{i}_sizes"].tolist() for i in range(len(UNIVERSE_CASES))])


# --------------------------------------------------------------------------- train
TRAIN_CASES = [  # name, model, kwargs, opt, lr, margin, steps
    ("transe_l1_adagrad", "TransE", dict(dim=20, p_norm=1, norm_flag=True), "Adagrad", 0.05, 2, 60),
    ("transe_l2_sgd", "TransE", dict(dim=20, p_norm=2, norm_flag=True), "sgd", 0.5, 1, 60),
    ("transh_l1_adagrad", "TransH", dict(dim=20, p_norm=1, norm_flag=True), "Adagrad", 0.05, 3, 60),
    ("transd_l1_adagrad", "TransD", dict(dim_e=20, dim_r=20, p_norm=1, norm_flag=True), "Adagrad", 0.05, 2, 60),
    ("transe_l1_sgd_d50_k3", "TransE", dict(dim=50, p_norm=1, norm_flag=True), "sgd", 1.0, 5, 20),
]


def topic_train():
    """Reference Trainer on one WN18 universe (seed 7, tc 600, balance .25): batches, losses, tables."""
    _setup_ref_import()
    import torch
    import openke.module.model as M
    from openke.config import Trainer
    from openke.data import TrainDataLoader
    from openke.module.loss import MarginLoss
    from openke.module.strategy import NegativeSampling
    torch.set_num_threads(1)
    res = {}
    for name, model_name, kw, opt, lr, margin, steps in TRAIN_CASES:
        k = 3 if name.endswith("k3") else 1
        dl = TrainDataLoader(in_path=WN18, nbatches=20, threads=8, bern_flag=0, filter_flag=0, neg_ent=k,
                             random_seed=4)
        lib = dl.lib
        lib.setRandomSeed(7)
        lib.randReset()
        torch.manual_seed(7)
        dl.compile_universe_dataset(600, 0.25)
        nE, nR = lib.getEntityTotalUniverse(), lib.getRelationTotalUniverse()
        model = NegativeSampling(model=getattr(M, model_name)(nE, nR, **kw), loss=MarginLoss(margin=margin),
                                 batch_size=dl.batch_size)
        init = {k_: v.clone().numpy() for k_, v in model.model.state_dict().items()}
        trainer = Trainer(model=model, data_loader=dl, train_times=1, alpha=lr, use_gpu=False, opt_method=opt)
        dl.swap_helpers()
        # Trainer.run() builds the optimizer and loops; to record per-step data we drive its own
        # train_one_step() with its own optimizer factory (train_times=0 => run() only builds it).
        trainer.train_times = 0
        trainer.run()
        batches, losses = [], []
        for _ in range(steps):
            d = dl.sampling()
            batches.append(np.stack([d["batch_h"], d["batch_t"], d["batch_r"]]).astype(np.int32))
            losses.append(trainer.train_one_step(d))
        dl.reset_universe()
        res[name + "_sizes"] = np.array([nE, nR, dl.batch_size if False else batches[0].shape[1] // (1 + k), k],
                                        dtype=np.int64)
        res[name + "_hyper"] = np.array([lr, margin], dtype=np.float64)
        res[name + "_batches"] = np.stack(batches)
        res[name + "_losses"] = np.array(losses, dtype=np.float32)
        for k_, v in init.items():
            if k_.endswith(".weight"):
                res[name + "_init_" + k_[:-7]] = v
        for k_, v in model.model.state_dict().items():
            if k_.endswith(".weight"):
                res[name + "_final_" + k_[:-7]] = v.numpy()
        print(name, nE, nR, res[name + "_sizes"].tolist(), losses[:3], losses[-1])
    np.savez_compressed(os.path.join(OUT, "train.npz"), **res)


# --------------------------------------------------------------------------- rank
def _rank_all(lib, loader, predict, n, ent_tot):
    """Per-triple raw/filtered ranks through the reference's testHead/testTail (accumulators reset
    before every call so the float sums are exact integers)."""
    getf = lambda s: ctypes.c_float.in_dll(lib, s).value
    ranks = np.zeros((n, 4), dtype=np.int32)  # head raw, head filt, tail raw, tail filt
    loader.set_sampling_mode("link")
    last_head = ctypes.c_long.in_dll(lib, "lastHead")
    last_tail = ctypes.c_long.in_dll(lib, "lastTail")
    for idx in range(n):
        lib.initTest()               # zero the float accumulators (and the two cursors)
        last_head.value = idx        # put the reference's batch cursors back on triple idx
        last_tail.value = idx
        dh, dt = loader.sampling_lp()   # reference getHeadBatch / getTailBatch
        s = predict(dh)
        lib.testHead(s.__array_interface__["data"][0], idx, 0)
        s = predict(dt)
        lib.testTail(s.__array_interface__["data"][0], idx, 0)
        ranks[idx] = (getf("l_rank") - 1, getf("l_filter_rank") - 1, getf("r_rank") - 1, getf("r_filter_rank") - 1)
    return ranks


def topic_rank():
    """Known-answer ranking: shipped TransH/WN18 checkpoint through the reference Tester."""
    _setup_ref_import()
    import torch
    from openke.config import Tester
    from openke.data import TestDataLoader
    from openke.module.model import TransH
    sd = torch.load(REF + "/best_models/transH_WN18_optimal_model.ckpt", map_location="cpu")
    tl = TestDataLoader(WN18, "link")
    res = {k[:-7]: v.numpy() for k, v in sd.items() if k.endswith(".weight")}
    res["test_sorted"] = _triples(tl.lib, "testList", tl.testTotal).astype(np.int32)  # (h,r,t), sorted (r,h,t)
    for p in (1, 2):
        m = TransH(tl.entTotal, tl.relTotal, dim=20, p_norm=p, norm_flag=True)
        m.load_state_dict(sd)
        m.eval()
        tester = Tester(model=m, data_loader=tl, use_gpu=False)
        with torch.no_grad():
            out = tester.run_link_prediction()
        res[f"metrics_p{p}"] = np.array(out, dtype=np.float32)  # mrr, mr, hit10, hit3, hit1 (filtered, averaged)
        names = ["l_reci_rank", "l_rank", "l_tot", "l3_tot", "l1_tot", "r_reci_rank", "r_rank", "r_tot", "r3_tot",
                 "r1_tot", "l_filter_reci_rank", "l_filter_rank", "l_filter_tot", "l3_filter_tot", "l1_filter_tot",
                 "r_filter_reci_rank", "r_filter_rank", "r_filter_tot", "r3_filter_tot", "r1_filter_tot"]
        res[f"table_p{p}"] = np.array([ctypes.c_float.in_dll(tl.lib, s).value for s in names], dtype=np.float32)
        with torch.no_grad():
            res[f"ranks_p{p}"] = _rank_all(tl.lib, tl, tester.test_one_step, tl.testTotal, tl.entTotal)
        print("p", p, out, res[f"ranks_p{p}"][:3])
    res["table_names"] = np.array(" ".join(names))
    np.savez_compressed(os.path.join(OUT, "rank_transh_wn18.npz"), **res)


# --------------------------------------------------------------------------- putranse
def topic_putranse_nullvec():
    """Same run as topic_putranse but evaluated with missing_embedding_handling='null_vector'
    (reference Parallel_Universe_Config.py:378-388,494-514,634-640): candidates no universe can score
    get the key's tuple score ||0 + r^ - t^|| resp. ||h^ + r^ - 0|| instead of +inf."""
    topic_putranse(null_vector=True)


def topic_putranse(null_vector=False):
    """Static PuTransE script behaviour (seeds 4..), few epochs, then the reference evaluation."""
    _setup_ref_import()
    import torch
    from openke.config import Parallel_Universe_Config
    from openke.data import TrainDataLoader, TestDataLoader
    from openke.module.model import TransE
    torch.set_num_threads(8)
    n_univ, epochs = 8, 3
    train = TrainDataLoader(in_path=WN18, nbatches=20, threads=8, sampling_mode="normal", bern_flag=0, filter_flag=0,
                            neg_ent=1, neg_rel=0, random_seed=123)
    test = TestDataLoader(train.in_path, "link")
    pu = Parallel_Universe_Config(training_identifier="golden", train_dataloader=train, test_dataloader=test,
                                  initial_num_universes=None, min_margin=1, max_margin=4, min_lr=0.001, max_lr=0.1,
                                  min_num_epochs=50, max_num_epochs=200, const_num_epochs=epochs,
                                  min_triple_constraint=500, max_triple_constraint=2000, min_balance=0.25,
                                  max_balance=0.5, embedding_model=TransE,
                                  embedding_model_param={"dim": 20, "p_norm": 1, "norm_flag": 1},
                                  checkpoint_dir="/tmp/", valid_steps=10 ** 9, save_steps=10 ** 9,
                                  training_setting="static", incremental_strategy=None,
                                  missing_embedding_handling="null_vector" if null_vector else "last_rank")
    pu.use_gpu = False
    assert pu.initial_random_seed == 4
    pu.train_parallel_universes(n_univ)
    res = {"n_univ": np.int64(n_univ), "epochs": np.int64(epochs), "initial_seed": np.int64(pu.initial_random_seed)}
    for u in range(n_univ):
        sp = pu.trained_embedding_spaces[u]
        res[f"u{u}_ent"] = sp.ent_embeddings.weight.detach().numpy()
        res[f"u{u}_rel"] = sp.rel_embeddings.weight.detach().numpy()
        emap = pu.entity_id_mappings[u]
        rmap = pu.relation_id_mappings[u]
        er = np.zeros(len(emap), dtype=np.int32)
        for g, l in emap.items():
            er[l] = g
        rr = np.zeros(len(rmap), dtype=np.int32)
        for g, l in rmap.items():
            rr[l] = g
        res[f"u{u}_ent_remap"], res[f"u{u}_rel_remap"] = er, rr
    with torch.no_grad():
        pu.data_loader.set_sampling_mode("link")
        pu.eval_universes(eval_mode="test")
        res["ranks"] = _rank_all(pu.lib, pu.data_loader, pu.test_one_step, test.testTotal, test.entTotal)
        # and the reference's own aggregate through its public entry point (eval cache already filled)
        out = super(Parallel_Universe_Config, pu).run_link_prediction(False)
    res["metrics"] = np.array(out, dtype=np.float32)
    res["test_sorted"] = _triples(pu.lib, "testList", test.testTotal).astype(np.int32)
    print("putranse metrics", out)
    if null_vector:   # the universes are those of putranse_wn18.npz (same seeds): keep only the evaluation
        res = {k: v for k, v in res.items() if not k.startswith("u")}
        np.savez_compressed(os.path.join(OUT, "putranse_nullvec_wn18.npz"), **res)
        return
    np.savez_compressed(os.path.join(OUT, "putranse_wn18.npz"), **res)


# --------------------------------------------------------------------------- putranse at production length
STATIC_RANGES = dict(min_margin=1, max_margin=4, min_lr=0.001, max_lr=0.1, min_num_epochs=50, max_num_epochs=200,
                     min_triple_constraint=500, max_triple_constraint=2000, min_balance=0.25, max_balance=0.5)


def _record_step_losses():
    """Instrument (not modify) the reference Trainer: keep every value train_one_step returns."""
    from openke.config import Trainer
    log = []
    orig = Trainer.train_one_step

    def train_one_step(self, data):
        loss = orig(self, data)
        log.append(loss)
        return loss
    Trainer.train_one_step = train_one_step
    return log


def _run_full(in_path, out_name, n_univ, model_name, param, nbatches, ranges, torch_threads=8, store_tables=True, extra=None):
    """The static PuTrans* experiment exactly as experiments/static_experiment_PuTransE_on_WN18.py:43-88
    drives it (seeds 4.., drawn epochs 50-199, Adagrad), on the first n_univ universes: per-step losses,
    final tables, remaps, and the reference's own link-prediction ranks + metrics of the ensemble."""
    _setup_ref_import()
    import time
    import torch
    import openke.module.model as M
    from openke.config import Parallel_Universe_Config
    from openke.data import TrainDataLoader, TestDataLoader
    torch.set_num_threads(torch_threads)
    log = _record_step_losses()
    train = TrainDataLoader(in_path=in_path, nbatches=nbatches, threads=8, sampling_mode="normal", bern_flag=0,
                            filter_flag=0, neg_ent=1, neg_rel=0, random_seed=123)
    test = TestDataLoader(train.in_path, "link")
    pu = Parallel_Universe_Config(training_identifier="golden_full", train_dataloader=train, test_dataloader=test,
                                  initial_num_universes=None, embedding_model=getattr(M, model_name),
                                  embedding_model_param=param, checkpoint_dir="/tmp/", valid_steps=10 ** 9,
                                  save_steps=10 ** 9, training_setting="static", incremental_strategy=None, **ranges)
    pu.use_gpu = False
    assert pu.initial_random_seed == 4
    res = {"n_univ": np.int64(n_univ), "initial_seed": np.int64(pu.initial_random_seed), "nbatches": np.int64(nbatches),
           "model": np.array(model_name)}
    res.update(extra or {})
    t0 = time.time()
    positives = 0
    for u in range(n_univ):
        del log[:]
        pu.train_parallel_universes(1)
        res[f"u{u}_losses"] = np.array(log, dtype=np.float32)
        sp = pu.trained_embedding_spaces[u]
        for k_, v in sp.state_dict().items():
            if k_.endswith(".weight"):
                if store_tables:
                    res[f"u{u}_{k_[:-7]}"] = v.detach().numpy().copy()
                else:     # the row norms are what the full-length test compares
                    res[f"u{u}_{k_[:-7]}_rownorm"] = np.linalg.norm(v.detach().numpy(), axis=1).astype(np.float32)
        emap, rmap = pu.entity_id_mappings[u], pu.relation_id_mappings[u]
        er = np.zeros(len(emap), dtype=np.int32)
        for g, l in emap.items():
            er[l] = g
        rr = np.zeros(len(rmap), dtype=np.int32)
        for g, l in rmap.items():
            rr[l] = g
        res[f"u{u}_ent_remap"], res[f"u{u}_rel_remap"] = er, rr
        positives += len(log) * train.batch_size if False else 0
        print("universe", u, "steps", len(log), "nE", len(er), "nR", len(rr), "last loss", log[-1], flush=True)
    res["train_seconds"] = np.float64(time.time() - t0)
    res["torch_threads"] = np.int64(torch_threads)
    t0 = time.time()
    with torch.no_grad():
        pu.data_loader.set_sampling_mode("link")
        pu.eval_universes(eval_mode="test")
        res["ranks"] = _rank_all(pu.lib, pu.data_loader, pu.test_one_step, test.testTotal, test.entTotal)
        out = super(Parallel_Universe_Config, pu).run_link_prediction(False)
    res["eval_seconds"] = np.float64(time.time() - t0)
    res["metrics"] = np.array(out, dtype=np.float32)   # mrr, mr, hit10, hit3, hit1 (filtered, head/tail averaged)
    res["test_sorted"] = _triples(pu.lib, "testList", test.testTotal).astype(np.int32)
    print(out_name, "metrics", out, "train s", float(res["train_seconds"]), "eval s", float(res["eval_seconds"]))
    np.savez_compressed(os.path.join(OUT, out_name), **res)


def topic_putranse_full():
    """configs[1] at production length: 24 WN18 universes with their drawn epochs (1000-3980 steps each)."""
    _run_full(WN18, "putranse_full_wn18.npz", 24, "TransE", {"dim": 20, "p_norm": 1, "norm_flag": 1}, 20, STATIC_RANGES)


def topic_putranse_full_fb15k():
    """configs[3] shape at production length: the FB15K-shaped synthetic graph of tools/synth.py (E = 14 951,
    R = 1 345, 483 142 train triples; the reference's own FB15K train file is not shipped), eight universes at their
    drawn epochs, evaluated by the reference on 2 000 test triples (its evaluation is a Python loop)."""
    import tempfile
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import synth
    tr, va, te, ne, nr = synth.fb15k_shape(n_valid=2000, n_test=2000)
    path = synth.write_dataset(tempfile.mkdtemp(), tr, va, te, ne, nr)
    _run_full(path, "putranse_full_fb15k.npz", 8, "TransE", {"dim": 20, "p_norm": 1, "norm_flag": 1}, 20, STATIC_RANGES,
              store_tables=False, extra={"train_checksum": np.uint64(synth.checksum(tr)), "n_valid": np.int64(2000), "n_test": np.int64(2000)})


INCREMENTAL_UNIVERSES = [(41, 700, 0.3), (42, 1500, 0.5), (43, 400, 0.25), (44, 1100, 0.4)]


def topic_incremental():
    """The reference's incremental TRAINING path (openke/base/Incremental.h:798-846 evolveTrainList,
    UniverseConstructor.h:336-339 focus from the currently contained relations) on the evolving graph of
    tools/synth.py incremental_dataset: after every snapshot the training list, the ordered relation arrays, the
    Bernoulli means, the loader's bookkeeping, and four universes sampled from the evolved graph."""
    _setup_ref_import()
    import tempfile
    sys.path.insert(0, os.path.join(REPO, "tools"))
    import synth
    from openke.data import IncrementalTrainDataLoader
    path = tempfile.mkdtemp() + "/"
    synth.incremental_dataset(path)
    dl = IncrementalTrainDataLoader(in_path=path, nbatches=20, threads=8, sampling_mode="normal", bern_flag=0, filter_flag=0,
                                    neg_ent=1, neg_rel=0, random_seed=4, incremental_setting=True, num_snapshots=3)
    lib = dl.lib
    res = {"universe_cases": np.array(INCREMENTAL_UNIVERSES, dtype=np.float64), "ent_rel_total": np.array([dl.entTotal, dl.relTotal])}

    def int_array(sym, count_sym):
        n = ctypes.c_long.in_dll(lib, count_sym).value
        ptr = ctypes.POINTER(ctypes.c_long).in_dll(lib, sym)
        return np.array([ptr[i] for i in range(n)], dtype=np.int32)

    for s_ in (1, 2, 3):
        dl.load_snapshot(s_)
        n = lib.getTrainTotal()
        res[f"s{s_}_train"] = _triples(lib, "trainList", n).astype(np.int32)             # (h, r, t), sorted (h,r,t), duplicates kept
        res[f"s{s_}_rel_contained"] = int_array("currently_contained_train_relations", "num_currently_contained_train_relations")
        res[f"s{s_}_rel_all"] = int_array("all_train_relations", "num_all_train_relations")
        res[f"s{s_}_rel_deleted"] = int_array("deleted_train_relations", "num_deleted_train_relations")
        res[f"s{s_}_left_mean"] = _reals(lib, "left_mean", dl.relTotal)
        res[f"s{s_}_right_mean"] = _reals(lib, "right_mean", dl.relTotal)
        res[f"s{s_}_loader"] = np.array([dl.tripleTotal, dl.batch_size, dl.nbatches, len(dl.deleted_triple_set)], dtype=np.int64)
        for i, (seed, tc, bal) in enumerate(INCREMENTAL_UNIVERSES):
            lib.setRandomSeed(seed + 10 * s_)
            lib.randReset()
            dl.compile_universe_dataset(tc, bal)
            nT, nE, nR = lib.getTrainTotalUniverse(), lib.getEntityTotalUniverse(), lib.getRelationTotalUniverse()
            er, rr = dl.get_universe_mappings()
            res[f"s{s_}_u{i}_sizes"] = np.array([nT, nE, nR], dtype=np.int64)
            res[f"s{s_}_u{i}_ent_remap"] = er.astype(np.int32)
            res[f"s{s_}_u{i}_rel_remap"] = rr.astype(np.int32)
            res[f"s{s_}_u{i}_triples_global"] = _triples(lib, "trainListUniverse", nT).astype(np.int32)
            dl.swap_helpers()
            d = dl.sampling()
            res[f"s{s_}_u{i}_batch"] = np.stack([d["batch_h"], d["batch_t"], d["batch_r"]]).astype(np.int32)
            dl.reset_universe()
        print("snapshot", s_, res[f"s{s_}_loader"].tolist(), res[f"s{s_}_rel_contained"].tolist(), res[f"s{s_}_rel_deleted"].tolist())
    np.savez_compressed(os.path.join(OUT, "incremental.npz"), **res)


def topic_tc():
    """Triple classification (reference Test.h:573-599, Tester.py:120-191): the corrupted twins getTestBatch draws
    for the WN18 test set right after TestDataLoader.read (seed 4), and the reference Tester's accuracy + threshold
    for the shipped TransH checkpoint on exactly those pairs."""
    _setup_ref_import()
    import torch
    from openke.config import Tester
    from openke.data import TestDataLoader
    from openke.module.model import TransH
    tl = TestDataLoader(WN18, "classification")
    lib = tl.lib
    res = {}
    pos, neg = tl.sampling_tc()
    res["pos"] = np.stack([pos["batch_h"], pos["batch_t"], pos["batch_r"]]).astype(np.int32)
    res["neg"] = np.stack([neg["batch_h"], neg["batch_t"], neg["batch_r"]]).astype(np.int32)
    # which corruptions hit the reference's undefined read (entity without a record on that side: trainHead[-1])
    lef_head = ctypes.POINTER(ctypes.c_long).in_dll(lib, "lefHead")
    rig_head = ctypes.POINTER(ctypes.c_long).in_dll(lib, "rigHead")
    rig_tail = ctypes.POINTER(ctypes.c_long).in_dll(lib, "rigTail")
    res["head_has_records"] = np.array([rig_head[int(h)] >= 0 for h in pos["batch_h"]])
    res["tail_has_records"] = np.array([rig_tail[int(t)] >= 0 for t in pos["batch_t"]])
    sd = torch.load(REF + "/best_models/transH_WN18_optimal_model.ckpt", map_location="cpu")
    m = TransH(tl.entTotal, tl.relTotal, dim=20, p_norm=1, norm_flag=True)
    m.load_state_dict(sd)
    m.eval()
    tester = Tester(model=m, data_loader=tl, use_gpu=False)
    data = [(pos, neg)]
    with torch.no_grad():
        acc, thr = tester.run_triple_classification(data_iterator=data)
        acc2, _ = tester.run_triple_classification(threshlod=thr, data_iterator=data)
        res["scores_pos"] = tester.test_one_step(pos).astype(np.float32)
        res["scores_neg"] = tester.test_one_step(neg).astype(np.float32)
    res["acc"], res["threshold"], res["acc_given_threshold"] = np.float64(acc), np.float64(thr), np.float64(acc2)
    print("tc", acc, thr, acc2, int((~res["head_has_records"]).sum()), int((~res["tail_has_records"]).sum()))
    np.savez_compressed(os.path.join(OUT, "tc_wn18.npz"), **res)


def topic_ref_ckpt():
    """A checkpoint written BY THE REFERENCE (Parallel_Universe_Config.save_parameters, reference :929-931): three
    universes of two epochs, plus the reference's ranks for it; what load_parameters must import."""
    _setup_ref_import()
    import shutil
    import torch
    from openke.config import Parallel_Universe_Config
    from openke.data import TrainDataLoader, TestDataLoader
    from openke.module.model import TransE
    torch.set_num_threads(8)
    train = TrainDataLoader(in_path=WN18, nbatches=20, threads=8, sampling_mode="normal", bern_flag=0, filter_flag=0,
                            neg_ent=1, neg_rel=0, random_seed=123)
    test = TestDataLoader(train.in_path, "link")
    pu = Parallel_Universe_Config(training_identifier="golden", train_dataloader=train, test_dataloader=test,
                                  initial_num_universes=None, const_num_epochs=2, embedding_model=TransE,
                                  embedding_model_param={"dim": 20, "p_norm": 1, "norm_flag": 1}, checkpoint_dir="/tmp/",
                                  valid_steps=10 ** 9, save_steps=10 ** 9, training_setting="static", incremental_strategy=None,
                                  **STATIC_RANGES)
    pu.use_gpu = False
    pu.train_parallel_universes(3)
    pu.save_parameters("/tmp/putranse_reference_layout.ckpt")
    shutil.copy("/tmp/putranse_reference_layout.ckpt", os.path.join(OUT, "putranse_reference_layout.ckpt"))
    with torch.no_grad():
        pu.data_loader.set_sampling_mode("link")
        pu.eval_universes(eval_mode="test")
        ranks = _rank_all(pu.lib, pu.data_loader, pu.test_one_step, test.testTotal, test.entTotal)
    np.savez_compressed(os.path.join(OUT, "putranse_reference_layout_ranks.npz"), ranks=ranks,
                        test_sorted=_triples(pu.lib, "testList", test.testTotal).astype(np.int32))
    print("ref ckpt", os.path.getsize(os.path.join(OUT, "putranse_reference_layout.ckpt")), ranks[:3])


TOPICS = {"incremental": topic_incremental, "putranse_full_fb15k": topic_putranse_full_fb15k, "tc": topic_tc, "ref_ckpt": topic_ref_ckpt, "putranse_full": topic_putranse_full, "dataset": topic_dataset,"sampler": topic_sampler, "universe": topic_universe, "train": topic_train,
          "rank": topic_rank, "putranse": topic_putranse, "putranse_nullvec": topic_putranse_nullvec}

if __name__ == "__main__":
    args = sys.argv[1:] or list(TOPICS)
    if len(args) == 1:
        TOPICS[args[0]]()
    else:
        for a in args:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), a], cwd=REPO)
