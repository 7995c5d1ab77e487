"""Single-space train step (K0+K1) on the synthetic S1 graph of SURVEY.md 8(d): E entities, T distinct
triples (Zipf relation frequencies, power-law entity degrees, seed 1234), TransE d, k=1,
B = T // nbatches.  This is the HBM-bound configuration: the tables (E*d*4 bytes, plus the Adagrad
state) exceed the 126 MB L2, so every gathered row comes from DRAM.

    python tools/bench_k1.py [--entities 1000000 --triples 10000000 --dim 64 --opt sgd|adagrad --steps 100]

Prints one JSON line with positive triples/s and the algorithmic GB/s (SURVEY.md 8(d) bytes per
positive) against the measured HBM peak."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (os.path.join(REPO, "openke-putranse_b200"), REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)


def synthetic_graph(E, R, T, seed=1234, degree="power"):
    """Distinct (h,r,t): relation ~ Zipf(1.0) over R, entity ~ power law (alpha ~ 2) over E, or uniform entity
    degrees (degree="uniform": no hub rows for the L2 to hold, every gathered row is a DRAM access)."""
    rng = np.random.default_rng(seed)
    pr = 1.0 / np.arange(1, R + 1)
    pr /= pr.sum()
    keys = np.zeros(0, dtype=np.int64)                                   # (h * R + r) * E + t
    while keys.shape[0] < T:
        m = int((T - keys.shape[0]) * 1.3) + 1000
        u = rng.random((m, 2))
        ent = np.minimum((E * (u ** (2.0 if degree == "power" else 1.0))).astype(np.int64), E - 1)   # power: density ~ x^-1/2
        ent = (ent * 2654435761) % E                                      # scatter the popular ids over the table
        rel = rng.choice(R, size=m, p=pr)
        ok = ent[:, 0] != ent[:, 1]
        keys = np.sort(np.concatenate([keys, ((ent[ok, 0] * R + rel[ok]) * E + ent[ok, 1])]))
        keys = keys[np.concatenate([[True], keys[1:] != keys[:-1]])]
    if keys.shape[0] > T:
        keys = keys[np.sort(rng.choice(keys.shape[0], size=T, replace=False))]
    out = np.stack([keys // (R * E), (keys // E) % R, keys % E], axis=1)
    return out  # sorted (h, r, t), distinct


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entities", type=int, default=1000000)
    ap.add_argument("--relations", type=int, default=1000)
    ap.add_argument("--triples", type=int, default=10000000)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--nbatches", type=int, default=100)
    ap.add_argument("--opt", default="sgd")
    ap.add_argument("--model", default="transe")
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--filter", type=int, default=0)
    ap.add_argument("--bern", type=int, default=0)
    ap.add_argument("--degree", default="power", choices=["power", "uniform"])
    args = ap.parse_args()
    import torch
    import bench
    from openke import _native as N
    N.require_cuda()
    L = N.lib()
    dev = torch.device("cuda", 0)
    E, R, T, d, k = args.entities, args.relations, args.triples, args.dim, 1
    t0 = time.time()
    tri = synthetic_graph(E, R, T, degree=args.degree)
    by_head = tri.astype(np.int32)
    order = np.argsort((tri[:, 2] * R + tri[:, 1]) * E + tri[:, 0], kind="stable")
    by_tail = by_head[order]
    print("graph: %d entities, %d relations, %d triples (%.1f s)" % (E, R, tri.shape[0], time.time() - t0), file=sys.stderr)
    d_bh, d_bt = torch.from_numpy(by_head).to(dev), torch.from_numpy(by_tail).to(dev)
    lm = rm = None
    if args.bern:   # reference Reader.h:148-166
        freq = np.bincount(tri[:, 1], minlength=R).astype(np.float64)
        nh = np.bincount(np.unique(tri[:, [1, 0]], axis=0)[:, 0], minlength=R)
        nt = np.bincount(np.unique(tri[:, [1, 2]], axis=0)[:, 0], minlength=R)
        lm = torch.from_numpy((freq / np.maximum(nh, 1)).astype(np.float32)).to(dev)
        rm = torch.from_numpy((freq / np.maximum(nt, 1)).astype(np.float32)).to(dev)
    model = {"transe": N.PK_TRANSE, "transh": N.PK_TRANSH, "transd": N.PK_TRANSD}[args.model]
    opt = N.PK_ADAGRAD if args.opt == "adagrad" else N.PK_SGD
    cfg = N.ModelCfg(model=model, dim=d, p_norm=1, norm_flag=1, opt=opt, neg_ent=k, bern=args.bern, filter=args.filter,
                     work_threads=8, reserved=0)
    ntE, ntR = (2 if model == N.PK_TRANSD else 1), (1 if model == N.PK_TRANSE else 2)
    g = torch.Generator(device="cpu").manual_seed(1)
    a = (6.0 / (E + d)) ** 0.5
    ent = [((torch.rand(E, d, generator=g) * 2 - 1) * a).to(dev) for _ in range(ntE)]
    rel = [((torch.rand(R, d, generator=g) * 2 - 1) * a).to(dev) for _ in range(ntR)]
    ent_s = [torch.zeros_like(t) for t in ent]
    rel_s = [torch.zeros_like(t) for t in rel]
    tab = N.Tables()
    for i in range(2):
        tab.ent[i] = ent[i].data_ptr() if i < ntE else None
        tab.rel[i] = rel[i].data_ptr() if i < ntR else None
        tab.ent_state[i] = ent_s[i].data_ptr() if (i < ntE and opt == N.PK_ADAGRAD) else None
        tab.rel_state[i] = rel_s[i].data_ptr() if (i < ntR and opt == N.PK_ADAGRAD) else None
    tab.n_ent, tab.n_rel = E, R
    B = tri.shape[0] // args.nbatches
    lcg = torch.from_numpy(np.arange(1, 9, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)).to(dev)
    smp = N.Sampler(by_head=d_bh.data_ptr(), by_tail=d_bt.data_ptr(), left_mean=lm.data_ptr() if lm is not None else None,
                    right_mean=rm.data_ptr() if rm is not None else None, lcg=lcg.data_ptr(), n_tri=tri.shape[0], n_ent=E, n_rel=R)
    ws = L.pk_workspace_create(ctypes.byref(cfg), E, R, B)
    assert ws, N.last_error()
    loss = torch.zeros(args.steps, dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    ms = []
    with torch.cuda.stream(stream):
        for rep in range(args.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            N.check(L.pk_train_steps(ctypes.byref(cfg), ctypes.byref(tab), ctypes.byref(smp), ws, B, args.steps, 1.0, 0.01,
                                     loss.data_ptr(), stream.cuda_stream), "pk_train_steps")
            e1.record(stream)
            stream.synchronize()
            if rep:
                ms.append(e0.elapsed_time(e1))
    launches = L.pk_last_launch_count()
    N.check(L.pk_workspace_check(ws, stream.cuda_stream))
    L.pk_workspace_free(ws)
    best = float(np.mean(ms))
    pos = B * args.steps
    bpp = bench.algorithmic_bytes_per_positive(args.model, d, k, args.opt)
    peak, src = bench.measured_peaks()
    gbs = pos * bpp / (best * 1e-3) / 1e9
    print(json.dumps({"workload": "s1: single space %s d=%d k=1 %s, E=%d R=%d T=%d (%s entity degrees), B=%d, %d steps/call" %
                      (args.model, d, args.opt, E, R, tri.shape[0], args.degree, B, args.steps),
                      "positive_triples_per_s": pos / (best * 1e-3), "ms_per_call": ms, "us_per_step": best * 1e3 / args.steps,
                      "launches_per_call": launches, "algorithmic_bytes_per_positive": bpp,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": src},
                      "loss_first_last": [float(loss[0]), float(loss[-1])]}))


if __name__ == "__main__":
    main()
