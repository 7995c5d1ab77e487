"""The static WN18 experiment flow (reference experiments/static_experiment_PuTransE_on_WN18.py): train N universes with a
validation on the valid split every 100, then the test evaluation; with and without validation behind the next launch."""
import contextlib
import io
import os
import sys
import tempfile
import time

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(REPO, "openke-putranse_b200"), REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import util  # noqa: E402
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    path = util.materialize_wn18(tempfile.mkdtemp())
    for rep in range(2):
        for pipelined in (False, True):
            pu = bench.make_pu(path)
            pu.valid_steps, pu.early_stopping_patience = 100, 10 ** 9
            pu.pipeline_validation = pipelined
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                pu.train_parallel_universes(n)
                t1 = time.perf_counter()
                m = pu.run_link_prediction()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print("rep %d pipelined=%d: %d universes, %d validations: %.3f s (%.1f M positives/s including validation), test evaluation %.3f s, "
                  "best valid hits@10 %.4f, test filtered MRR %.4f hits@10 %.4f" % (rep, pipelined, n, n // 100, t1 - t0,
                  pu.positive_triples / (t1 - t0) / 1e6, t2 - t1, pu.best_hit10, m[0], m[2]), flush=True)
            del pu


if __name__ == "__main__":
    main()
