"""Deviation of the CUDA path from the reference's full-length golden run (diagnostics behind the
tolerances stated in tests/test_parity_full.py)."""
import os
import sys
import tempfile

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(REPO, "openke-putranse_b200"), REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import util  # noqa: E402
import test_parity_full as T  # noqa: E402


def main():
    fb = len(sys.argv) > 1 and sys.argv[1] == "fb15k"
    g = np.load(os.path.join(util.GOLDEN, "putranse_full_fb15k.npz" if fb else "putranse_full_wn18.npz"))
    if fb:
        import pathlib
        path, _ = T._fb15k_shape_dir(pathlib.Path(tempfile.mkdtemp()))
    else:
        path = util.materialize_wn18(tempfile.mkdtemp())
    pu = T._static_putranse(path)
    pu.record_losses = True
    n, nb = int(g["n_univ"]), int(g["nbatches"])
    pu.train_parallel_universes(n)
    print(" u  steps  lr     m | first step with rel dev > 1e-4 / 1e-3 / 1e-2 / 1e-1 | max dev [:10] [:40] [:100] | epoch-mean dev max (rel, abs) | tail10 ours ref")
    for u in range(n):
        got, want = pu.universe_losses[u].astype(np.float64), g["u%d_losses" % u].astype(np.float64)
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-3)
        first = [int(np.argmax(rel > t)) if (rel > t).any() else -1 for t in (1e-4, 1e-3, 1e-2, 1e-1)]
        ge, we = T._epoch_means(got, nb), T._epoch_means(want, nb)
        hy = pu.universe_hyper[u]
        print("%2d %6d %.3f %d | %5d %5d %5d %5d | %.2e %.2e %.2e | %.3f %.4f | %.4f %.4f" % (
            u, len(got), hy["lr"], hy["margin"], *first, rel[:10].max(), rel[:40].max(), rel[:100].max(),
            (np.abs(ge - we) / we).max(), np.abs(ge - we).max(), ge[-10:].mean(), we[-10:].mean()))
    out = pu.run_link_prediction()
    ranks, want = pu.last_ranks, g["ranks"]
    E = 14951 if fb else 40943
    mh, mt = want[:, 0] == E, want[:, 2] == E
    d = np.concatenate([np.abs(ranks[~mh][:, 1] - want[~mh][:, 1]), np.abs(ranks[~mt][:, 3] - want[~mt][:, 3])])
    print("ours", out)
    print("ref ", g["metrics"].tolist())
    print("missing sets equal:", np.array_equal(ranks[:, 0] == E, mh), np.array_equal(ranks[:, 2] == E, mt),
          "inf-branch ranks equal:", np.array_equal(ranks[mh][:, :2], want[mh][:, :2]), np.array_equal(ranks[mt][:, 2:], want[mt][:, 2:]))
    for t in (0, 1, 3, 10, 100):
        print("scored filtered ranks within %3d of the reference's: %.4f" % (t, (d <= t).mean()))
    rr = lambda x: 1.0 / (x + 1.0)
    print("scored: mean |d(1/rank)|", np.abs(np.concatenate([rr(ranks[~mh][:, 1]) - rr(want[~mh][:, 1]), rr(ranks[~mt][:, 3]) - rr(want[~mt][:, 3])])).mean())


if __name__ == "__main__":
    main()
