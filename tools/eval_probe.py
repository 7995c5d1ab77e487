"""Where the time of run_link_prediction goes (400 universes on WN18): host preparation against kernels."""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    path = util.materialize_wn18(tempfile.mkdtemp())
    pu = bench.make_pu(path)
    pu.const_num_epochs = 1
    pu.async_training = True
    for _ in range(n // 100):
        pu.train_parallel_universes(100)
    pu.synchronize()
    for tile in (1024, 4096):
        pu.eval_tile_rows = tile
        for rep in range(3):
            pu._rank_cache.clear()
            pu.timings.clear()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ranks = pu._rank_split(pu.data_loader)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            from openke.config.Tester import link_metrics
            m = link_metrics(ranks)
            t2 = time.perf_counter()
            print("universes %d tile %d rep %d: _rank_split %.1f ms (host prep %.1f ms), link_metrics %.2f ms -> %.0f test triples/s; launches %d"
                  % (n, tile, rep, (t1 - t0) * 1e3, pu.timings["eval_host_prep"] * 1e3, (t2 - t1) * 1e3, ranks.shape[0] / (t2 - t0), pu.gpu_launches))
    import cProfile
    import pstats
    pu._rank_cache.clear()
    pr = cProfile.Profile()
    pr.enable()
    pu._rank_split(pu.data_loader)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)


if __name__ == "__main__":
    main()
