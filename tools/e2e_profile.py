"""Where the launching thread's time goes in the end-to-end loop (cProfile over train_parallel_universes calls)."""
import cProfile
import os
import pstats
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    path = util.materialize_wn18(tempfile.mkdtemp())
    pu = bench.make_pu(path)
    pu.record_losses = True
    pu.async_training = True
    for _ in range(10):
        pu.train_parallel_universes(n)
    pu.synchronize()
    torch.cuda.synchronize()
    pu.timings.clear()
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    for _ in range(steps):
        pu.train_parallel_universes(n)
    pu.synchronize()
    pr.disable()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("n=%d: %.2f ms/step under the profiler; host: %s" % (n, dt / steps * 1e3, {k: round(v / steps * 1e3, 2) for k, v in pu.timings.items()}))
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
    main()
