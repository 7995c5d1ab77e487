"""Probe for the batched-universe kernel: trains N universes with a fixed number of epochs and
prints microseconds per training step of the slowest universe.  Used under ncu."""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    path = util.materialize_wn18(tempfile.mkdtemp())
    pu = bench.make_pu(path)
    pu.const_num_epochs = epochs
    dev = torch.device("cuda", 0)
    ids = [int(x) for x in os.environ["K2_PROBE_IDS"].split(",")] if os.environ.get("K2_PROBE_IDS") else list(range(n))
    n = len(ids)
    launch = bench.prepare_resident_launch(pu, ids, dev)["sets"][0]
    ts = []
    for _ in range(reps):
        launch["reset"]()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        launch["run"]()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    steps = epochs * bench.NBATCHES
    sizes = [(pu.universe_hyper[u]["nE"], pu.universe_hyper[u]["batch_size"]) for u in ids]
    print("universes=%d epochs=%d steps/universe=%d  kernel ms=%s  => %.2f us/step (slowest universe)  sizes(nE,B)=%s"
          % (n, epochs, steps, ["%.3f" % t for t in ts], min(ts) * 1e3 / steps, sizes))


if __name__ == "__main__":
    main()
