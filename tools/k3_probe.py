"""Probe for the single-space ranking kernels (pk_rank_space): WN18 test set (5000 triples, both
sides, raw + filtered) against random tables of a given dimension; prints ms and useful GFLOP/s
(4*T*E*d flops)."""
import os
import sys
import tempfile

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(REPO, "openke-putranse_b200"), REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import util  # noqa: E402


def main():
    from openke.config import Tester
    from openke.data import TestDataLoader
    import openke.module.model as M
    path = util.materialize_wn18(tempfile.mkdtemp())
    test = TestDataLoader(path, "link")
    for cls, kw in (("TransE", {"dim": 20}), ("TransE", {"dim": 50}), ("TransH", {"dim": 20}), ("TransD", {"dim_e": 20, "dim_r": 20}),
                    ("TransE", {"dim": 50, "p_norm": 2})):
        torch.manual_seed(1)
        model = getattr(M, cls)(40943, 18, **kw)
        t = Tester(model=model, data_loader=test, use_gpu=True)
        t.rank_all()
        torch.cuda.synchronize()
        ms = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            t.rank_all()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        d = kw.get("dim", kw.get("dim_e"))
        flops = 4.0 * 5000 * 40943 * d
        print("%s %s: %s ms  => %.1f k test triples/s, %.1f useful GFLOP/s" % (cls, kw, ["%.1f" % x for x in ms], 5000 / min(ms), flops / min(ms) / 1e6))


if __name__ == "__main__":
    main()
