import cProfile, pstats, contextlib, io, os, sys, tempfile, time
REPO = "/root/repo"
for p in (os.path.join(REPO, "openke-putranse_b200"), REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import torch, util, bench
path = util.materialize_wn18(tempfile.mkdtemp())
for rep in range(2):
    pu = bench.make_pu(path)
    pu.valid_steps, pu.early_stopping_patience = 100, 10 ** 9
    pr = cProfile.Profile()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        if rep: pr.enable()
        pu.train_parallel_universes(2000)
        if rep: pr.disable()
    torch.cuda.synchronize()
    print("rep", rep, time.perf_counter() - t0)
    if rep:
        print({k: round(v * 1e3, 1) for k, v in pu.timings.items()})
        pstats.Stats(pr).sort_stats("tottime").print_stats(22)
