"""Device universe walk: time per chunk (CUDA events on the walk stream) beside the host builder."""
import os, sys, time, tempfile
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(REPO, "openke-putranse_b200"), os.path.join(REPO, "tests"), REPO]
import torch
import util
from random import Random
from openke import _native as N
from openke.data import TrainDataLoader
from openke.universe_walk import DeviceWalker

import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "m2"
path = bench.workload_dataset(wl)[0] if wl != "m2" else util.materialize_wn18(tempfile.mkdtemp())
print("workload", wl)
dl = TrainDataLoader(in_path=path, nbatches=20, threads=8, bern_flag=0, filter_flag=0, neg_ent=1, random_seed=4)
L = dl.lib
w = DeviceWalker(L, torch.device("cuda", 0))
w.streams = w.streams[:1]      # one stream: the events below bracket the launch
for n in (1, 100, 125, 1000):
    seeds = np.arange(4, 4 + n, dtype=np.int64)
    tcs, bals = np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.float32)
    for i in range(n):
        rnd = Random(int(seeds[i])); tcs[i] = rnd.randrange(500, 2000); bals[i] = round(rnd.uniform(0.25, 0.5), 2)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(w.stream):
            e0.record(w.stream)
        res = w.submit(seeds, tcs, bals, 8)
        t1 = time.perf_counter()
        with torch.cuda.stream(w.stream):
            e1.record(w.stream)
        s = res.wait()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        res.release()
    print("n=%5d device walk %.3f ms (submit call %.3f ms, until sizes on host %.3f ms); rounds max %d mean %.1f, draws mean %.0f, nT mean %.0f, statuses %s" % (
        n, e0.elapsed_time(e1), (t1 - t0) * 1e3, (t2 - t0) * 1e3, int(s[:, 6].max()), s[:, 6].mean(), s[:, 4].mean(), s[:, 0].mean(), np.unique(s[:, 5]).tolist()))
    t0 = time.perf_counter()
    h = L.pk_universes_build_lean(n, N.addr(seeds), N.addr(tcs), N.addr(bals), 8)
    t1 = time.perf_counter()
    L.pk_universes_free(h)
    print("        host builder, 8 threads: %.3f ms" % ((t1 - t0) * 1e3))
