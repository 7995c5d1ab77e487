"""How much the SMs gain from keeping several universe launches in flight: train_parallel_universes(n)
called `steps` times, synchronous (the reference's semantics) against asynchronous launch slots."""
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


def run(path, n, steps, async_, slots, losses=True):
    pu = bench.make_pu(path)
    pu.record_losses = losses
    pu.async_training = async_
    pu.launch_slots = slots
    for _ in range(slots + 5):       # every slot's buffers exist before the timed loop
        pu.train_parallel_universes(n)
    pu.synchronize()
    torch.cuda.synchronize()
    pu.timings.clear()
    if pu._walker is not None:
        pu._walker.time_walks = True
    p0 = pu.positive_triples
    t0 = time.perf_counter()
    for _ in range(steps):
        pu.train_parallel_universes(n)
    pu.synchronize()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    pos = pu.positive_triples - p0
    if pu._walker is not None and pu._walker.walk_events:
        w = [a.elapsed_time(b) for a, b in pu._walker.walk_events]
        print("      walks on the device, queue to finish: mean %.2f ms, max %.2f ms (%d walks)" % (sum(w) / len(w), max(w), len(w)))
    print("n=%4d steps=%2d async=%d slots=%d losses=%d: %.2f ms/step  %.1f M positives/s   host: %s" % (
        n, steps, async_, slots, losses, dt / steps * 1e3, pos / dt / 1e6,
        {k: round(v / steps * 1e3, 2) for k, v in pu.timings.items()}), flush=True)


def main():
    path = util.materialize_wn18(tempfile.mkdtemp())
    import io, contextlib
    combos = ((0, 1), (1, 2), (1, 3), (1, 4))
    if len(sys.argv) > 1:   # e.g. "4,6,8": asynchronous slot counts only, longer runs
        combos = tuple((1, int(x)) for x in sys.argv[1].split(","))
    for n, steps in ((100, 40 if len(sys.argv) > 1 else 20), (1000, 3)):
        if len(sys.argv) > 1 and n != 100:
            continue
        for async_, slots in combos:
            run(path, n, steps, async_, slots)


if __name__ == "__main__":
    main()
