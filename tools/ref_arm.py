#!/usr/bin/env python
"""The reference arm of bench.py: the UNMODIFIED reference Python package driving the UNMODIFIED
reference native core on this box's host cores.

    python tools/ref_arm.py --pkg baseline/_ref --data <dir> --model transe --nbatches 20 \
        --first-universe 0 --universes 2 [--budget-s 15] [--max-epochs N] [--torch-threads T] [--eval]

It runs in its own interpreter (the reference package is also called `openke`) with `--pkg` first on
sys.path: `--pkg/openke/{__init__.py,config,data,module}` are the reference's own files and
`--pkg/openke/release/Base.so` is the reference's native core compiled from its own sources
(populated by __graft_entry__.build(), git-ignored).  What is timed is exactly what the reference times
(reference openke/config/Parallel_Universe_Config.py:318-367): train_parallel_universes(1) per universe
with validation and checkpoints disabled, use_gpu = False.  Prints ONE JSON line.

Nothing of this repository's product is imported here.
"""
import argparse
import json
import os
import sys
import time


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pkg", required=True)
    ap.add_argument("--data", required=True)
    ap.add_argument("--model", default="transe")
    ap.add_argument("--nbatches", type=int, default=20)
    ap.add_argument("--first-universe", type=int, default=0)
    ap.add_argument("--universes", type=int, default=1, help="universes per step")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0, help="warm-up steps: same code path, universes of 2 epochs")
    ap.add_argument("--soft-limit-s", type=float, default=0.0,
                    help="after this many seconds the remaining steps run with epochs capped at 5 (and say so)")
    ap.add_argument("--budget-s", type=float, default=0.0, help="stop after this many seconds (checked between universes)")
    ap.add_argument("--max-epochs", type=int, default=0, help="cap const_num_epochs (0 = the drawn 50..199)")
    ap.add_argument("--torch-threads", type=int, default=0)
    ap.add_argument("--eval", action="store_true", help="also time run_link_prediction of what was trained")
    ap.add_argument("--eval-triples", type=int, default=0)
    args = ap.parse_args()

    sys.path.insert(0, os.path.abspath(args.pkg))
    real_stdout = os.dup(1)
    os.dup2(2, 1)                        # the reference prints and printf()s a lot; stdout carries the JSON only
    import torch
    import openke
    assert os.path.abspath(openke.__file__).startswith(os.path.abspath(args.pkg)), openke.__file__
    import openke.module.model as M
    from openke.config import Parallel_Universe_Config, Trainer
    from openke.data import TrainDataLoader, TestDataLoader
    cores = len(os.sched_getaffinity(0))
    if args.torch_threads:          # default: whatever the reference gets from torch (all cores)
        torch.set_num_threads(args.torch_threads)

    counter = {"positives": 0, "steps": 0}
    orig = Trainer.train_one_step

    def counted(self, data):     # instrumentation only: what one sampling() row count is worth
        counter["positives"] += self.data_loader.batch_size
        counter["steps"] += 1
        return orig(self, data)
    Trainer.train_one_step = counted

    model = {"transe": "TransE", "transh": "TransH", "transd": "TransD"}[args.model]
    param = {"dim_e": 20, "dim_r": 20, "p_norm": 1, "norm_flag": 1} if args.model == "transd" else {"dim": 20, "p_norm": 1, "norm_flag": 1}
    train = TrainDataLoader(in_path=args.data, nbatches=args.nbatches, threads=8, sampling_mode="normal", bern_flag=0,
                            filter_flag=0, neg_ent=1, neg_rel=0, random_seed=123)
    test = TestDataLoader(train.in_path, "link")
    pu = Parallel_Universe_Config(training_identifier="ref_arm", train_dataloader=train, test_dataloader=test,
                                  initial_num_universes=None, min_margin=1, max_margin=4, min_lr=0.001, max_lr=0.1,
                                  min_num_epochs=50, max_num_epochs=200, const_num_epochs=args.max_epochs or None,
                                  min_triple_constraint=500, max_triple_constraint=2000, min_balance=0.25, max_balance=0.5,
                                  embedding_model=getattr(M, model), embedding_model_param=param, checkpoint_dir="/tmp/",
                                  valid_steps=10 ** 9, save_steps=10 ** 9, training_setting="static", incremental_strategy=None)
    pu.use_gpu = False
    # warm-up steps (imports, allocator, thread pools): the same calls on far-away universe ids, two epochs each
    pu.next_universe_id = 10 ** 6
    pu.const_num_epochs = 2
    for _ in range(args.warmup):
        pu.train_parallel_universes(args.universes)
    # the reference's tensors are tiny (a batch of 25-99 rows of 20 floats): torch's intra-op pool usually costs more than
    # it gives.  Unless told otherwise, time the same 4-epoch universe with all threads and with one and keep the faster
    # setting, so that the baseline is the reference at its best on this host.
    calibration = {}
    if not args.torch_threads:
        pu.const_num_epochs = 4
        for t_ in sorted({cores, torch.get_num_threads(), 1}, reverse=True):
            torch.set_num_threads(t_)
            pu.next_universe_id = 10 ** 6 + 500
            t0 = time.perf_counter()
            pu.train_parallel_universes(1)
            calibration[t_] = time.perf_counter() - t0
        torch.set_num_threads(min(calibration, key=calibration.get))
    pu.const_num_epochs = args.max_epochs or None
    pu.next_universe_id = args.first_universe
    counter["positives"] = counter["steps"] = 0
    per_step, capped = [], 0
    t_all = time.perf_counter()
    for _ in range(args.steps):
        if args.soft_limit_s and time.perf_counter() - t_all > args.soft_limit_s and pu.const_num_epochs is None:
            pu.const_num_epochs = 5
        capped += pu.const_num_epochs == 5 and not args.max_epochs
        p0, s0, u0 = counter["positives"], counter["steps"], pu.next_universe_id
        t0 = time.perf_counter()
        # reference Parallel_Universe_Config.py:316-367; it prints its own "Time took for creation of embedding spaces"
        pu.train_parallel_universes(args.universes)
        dt = time.perf_counter() - t0
        per_step.append({"universes": [u0, pu.next_universe_id - 1], "seconds": dt, "positives": counter["positives"] - p0,
                         "steps": counter["steps"] - s0})
        if args.budget_s and time.perf_counter() - t_all > args.budget_s:
            break
    seconds = time.perf_counter() - t_all
    out = {"kind": "reference", "cores": cores, "torch_threads": torch.get_num_threads(), "sampler_threads": 8,
           "torch": torch.__version__, "positives": counter["positives"], "train_steps": counter["steps"], "seconds": seconds,
           "value": counter["positives"] / seconds, "per_step": per_step, "max_epochs": args.max_epochs,
           "steps_with_capped_epochs": int(capped), "steps_done": len(per_step),
           "thread_calibration_s": {str(k): round(v, 3) for k, v in calibration.items()}}
    if args.eval:
        if args.eval_triples:
            pu.data_loader.testTotal = min(pu.data_loader.testTotal, args.eval_triples)
        t0 = time.perf_counter()
        with torch.no_grad():
            res = pu.run_link_prediction()
        out["eval"] = {"seconds": time.perf_counter() - t0, "test_triples": int(pu.data_loader.testTotal), "result": repr(res)}
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(out) + "\n").encode())


if __name__ == "__main__":
    main()
