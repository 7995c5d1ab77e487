"""Per-universe time inside one K2 launch of the benchmark workload (pk_debug_universe_timer):
which universes bound the launch, and their microseconds per step."""
import os
import sys
import tempfile

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(REPO, "openke-putranse_b200"), REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import util  # noqa: E402
import bench  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    if len(sys.argv) > 2:
        bench.set_model(sys.argv[2])
    path = util.materialize_wn18(tempfile.mkdtemp())
    pu = bench.make_pu(path)
    dev = torch.device("cuda", 0)
    launch = bench.prepare_resident_launch(pu, list(range(n)), dev)["sets"][0]
    buf = torch.zeros(2 * n, dtype=torch.int64, device=dev)
    pu.lib.pk_debug_universe_timer(buf.data_ptr())
    best = None
    for _ in range(3):
        launch["reset"]()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        launch["run"]()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if best is None or ms < best[0]:
            best = (ms, buf.cpu().numpy().reshape(n, 2).copy())
    pu.lib.pk_debug_universe_timer(None)
    ms, tm = best
    off = 0
    by_off = {}
    for u in range(n):
        hy = pu.universe_hyper[u]
        by_off[off] = u
        off += hy["epochs"] * bench.NBATCHES
    rows = []
    for lo, ns in tm:
        u = by_off[int(lo)]
        hy = pu.universe_hyper[u]
        steps = hy["epochs"] * bench.NBATCHES
        rows.append((ns / 1e6, u, hy["nE"], hy["batch_size"], hy["epochs"], steps, ns / 1e3 / steps))
    rows.sort(reverse=True)
    print("launch %.3f ms, %d universes" % (ms, n))
    print("   ms     u    nE    B  epochs steps  us/step")
    for r in rows:
        print("%7.3f %4d %5d %4d %5d %6d %7.2f" % r)


if __name__ == "__main__":
    main()
