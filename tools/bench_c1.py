"""BASELINE.json configs[0] through the public API: TransE on WN18, one embedding space (d=50, L1,
margin 5, SGD lr 1.0, nbatches=100 => B=1414, k=1, Bernoulli + filtered negatives), Trainer.run for a
few epochs, then Tester.run_link_prediction.  Prints positive triples/s of Trainer.run (end to end,
including the per-epoch loss read-back) and the link-prediction time."""
import os
import sys
import tempfile
import time

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(REPO, "openke-putranse_b200"), REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import util  # noqa: E402


def main():
    epochs = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 50
    as_json = "--json" in sys.argv
    real_stdout = os.dup(1)
    if as_json:
        os.dup2(2, 1)
    from openke.config import Trainer, Tester
    from openke.data import TrainDataLoader, TestDataLoader
    from openke.module.loss import MarginLoss
    from openke.module.model import TransE
    from openke.module.strategy import NegativeSampling
    path = util.materialize_wn18(tempfile.mkdtemp())
    train = TrainDataLoader(in_path=path, nbatches=100, threads=8, sampling_mode="normal", bern_flag=1, filter_flag=1, neg_ent=1, neg_rel=0)
    test = TestDataLoader(path, "link")
    torch.manual_seed(0)
    transe = TransE(ent_tot=train.get_ent_tot(), rel_tot=train.get_rel_tot(), dim=50, p_norm=1, norm_flag=True)
    model = NegativeSampling(model=transe, loss=MarginLoss(margin=5.0), batch_size=train.get_batch_size())
    tr = Trainer(model=model, data_loader=train, train_times=3, alpha=1.0, use_gpu=True)
    tr.run(show_progress=False)      # warm-up (workspace, graph capture)
    torch.cuda.synchronize()
    tr.train_times = epochs
    t0 = time.perf_counter()
    tr.run(show_progress=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    pos = epochs * 100 * train.get_batch_size()
    print("Trainer.run: %d epochs x 100 batches x B=%d in %.3f s => %.1f M positive triples/s (%.1f us/step), loss %.4f -> %.4f"
          % (epochs, train.get_batch_size(), dt, pos / dt / 1e6, dt / (epochs * 100) * 1e6, tr.losses[0][0], tr.losses[-1][-1]))
    te = Tester(model=transe, data_loader=test, use_gpu=True)
    t0 = time.perf_counter()
    out = te.run_link_prediction()
    torch.cuda.synchronize()
    ev_s = time.perf_counter() - t0
    print("Tester.run_link_prediction: %.3f s, filtered (mrr, mr, hit10, hit3, hit1) = %s" % (ev_s, out))
    if as_json:
        import json
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps({
            "workload": "c1: TransE on WN18, one embedding space, d=50 L1 margin 5 SGD lr 1.0, nbatches=100 (B=%d), k=1, Bernoulli + "
                        "filtered negatives, Trainer.run (BASELINE.json configs[0])" % train.get_batch_size(),
            "value": pos / dt, "unit": "positive triples/s", "epochs": epochs, "us_per_step": dt / (epochs * 100) * 1e6,
            "first_loss": float(tr.losses[0][0]), "last_loss": float(tr.losses[-1][-1]),
            "link_prediction_seconds": ev_s, "filtered": dict(zip(("mrr", "mr", "hits10", "hits3", "hits1"), (float(x) for x in out))),
        }) + "\n").encode())


if __name__ == "__main__":
    main()
