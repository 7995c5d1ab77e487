#!/usr/bin/env python
"""bench.py — PuTransE positive triples/sec on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path, same box

Workload `m2` (default; BASELINE.json configs[1]): the static PuTransE experiment on WN18 —
universes seeded 4,5,6,... with the hyper-parameter ranges of the reference's
experiments/static_experiment_PuTransE_on_WN18.py:43-88 (TransE d=20, L1, Adagrad, nbatches=20,
k=1, unfiltered uniform negatives, tc in [500,2000), epochs in [50,200)).  One STEP = training one
chunk of `--universes` universes from scratch (all epochs x 20 batches of every universe).
positive triples = sum over universes of epochs * nbatches * batch_size.

`value`   : device-resident leg — subgraphs sampled, tables initialised and uploaded BEFORE the
            timed region; the timed region is the batched-universe kernel (K2) only.
`e2e`     : the same job through the public API, Parallel_Universe_Config.train_parallel_universes:
            host subgraph sampling + torch table init + H2D + K2 + D2H of the per-step losses.
`roofline`: K2's algorithmic bytes / its CUDA-event duration against the measured HBM peak.
`cpu_baseline` / `--impl reference`: reference Base.so (oracle/_ref) sampling + the op-for-op torch
            CPU port of the reference's Python training loop (oracle/model_math.py) on host cores.
With N > 1 (torchrun) every rank runs its own `--universes` universes (weak scaling; universes are
independent, no training-time communication).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(REPO, "openke-putranse_b200"), REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

STATIC_RANGES = dict(min_margin=1, max_margin=4, min_lr=0.001, max_lr=0.1, min_num_epochs=50, max_num_epochs=200,
                     min_triple_constraint=500, max_triple_constraint=2000, min_balance=0.25, max_balance=0.5)
MODEL_PARAM = {"dim": 20, "p_norm": 1, "norm_flag": 1}
NBATCHES, K_NEG = 20, 1
MODEL = "transe"     # --model transh / transd = BASELINE.json configs[2] (PuTransH / PuTransD on WN18)


def set_model(name):
    """configs[2]: experiments/static_experiment_PuTransH_on_WN18.py (nbatches 20) and ...PuTransD... (nbatches 10,
    dim_e = dim_r = 20)."""
    global MODEL, MODEL_PARAM, NBATCHES
    MODEL = name
    if name == "transd":
        MODEL_PARAM = {"dim_e": 20, "dim_r": 20, "p_norm": 1, "norm_flag": 1}
        NBATCHES = 10


def model_class():
    import openke.module.model as M
    return {"transe": M.TransE, "transh": M.TransH, "transd": M.TransD}[MODEL]


def algorithmic_bytes_per_positive(model="transe", d=20, k=1, opt="adagrad"):
    """rows(model,k) * d * 4 * rw(opt) + idx."""
    rows = {"transe": 3 + k, "transh": 4 + k, "transd": 6 + 2 * k}[model]
    rw = 4 if opt == "adagrad" else 2
    return rows * d * 4 * rw + 12 + 4 * k


def k2_traffic(model):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant (staged) K2 launch of THIS workload, from
    the committed `ncu --set full` capture (profiles/r2_k2_traffic.json, from the raw page of the capture);
    None when no capture of this model's launch is committed."""
    try:
        d = json.load(open(os.path.join(REPO, "profiles", "r2_k2_traffic.json")))
        return float(d["traffic_bytes"]) if model == "transe" else None
    except Exception:
        return None


def k1_traffic(opt, degree):
    try:
        d = json.load(open(os.path.join(REPO, "profiles", "r2_k1_traffic.json")))
        return float(d["%s_%s" % (opt, degree)]["traffic_bytes"])
    except Exception:
        return None


def k1_step_traffic(opt, degree):
    try:
        d = json.load(open(os.path.join(REPO, "profiles", "r2_k1_traffic.json")))
        return float(d["%s_%s" % (opt, degree)]["step_traffic_bytes_steady_state"])
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_ev = index, [], threading.Event()

    def _nvml_loop(self):
        """NVML in-process: a sample every 5 ms (an nvidia-smi process takes longer than a whole step)."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self._stop_ev.is_set():
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "0"] +
                             ["Active" if r & b else "Not Active" for _, b in bits])
            self._stop_ev.wait(0.005)

    def run(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            pass
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


STRONG = ("m4", "s1pu")      # workloads whose universes per step are a TOTAL sharded over the ranks
WORKLOADS = {
    "s1pu": "PuTransE on the synthetic 1M-entity / 10M-triple graph (BASELINE.json configs[4], SURVEY 8(d) S1-PU; E=1000000, R=1000, "
            "tools/synth.py seed 1234): %(n)d universes per step IN TOTAL, sharded u %% n_gpus, %(model)s d=20 L1 Adagrad, nbatches=%(nb)d, k=1",
    "m2": "PuTransE static WN18 (BASELINE.json configs[1]): %(n)d universes per GPU per step (seeds 4..), %(model)s d=20 L1 Adagrad, "
          "nbatches=%(nb)d, k=1, tc in [500,2000), epochs in [50,200)",
    "m4": "PuTransE on an FB15K-shaped graph (BASELINE.json configs[3]; E=14951, R=1345, 483142 synthetic train triples, tools/synth.py "
          "seed 1234): %(n)d universes per step IN TOTAL, sharded u %% n_gpus, %(model)s d=20 L1 Adagrad, nbatches=%(nb)d, k=1",
}


def workload_dataset(name):
    """Materialise the workload's graph as the header-less files the loaders read; returns (path, data text)."""
    import tempfile
    import util
    if name in ("m4", "s1pu"):
        sys.path.insert(0, os.path.join(REPO, "tools"))
        import synth
        tr, va, te, ne, nr = synth.fb15k_shape() if name == "m4" else synth.s1_shape()
        path = synth.write_dataset(tempfile.mkdtemp(), tr, va, te, ne, nr)
        return path, "synthetic %s graph (tools/synth.py, seed 1234, checksum %d); universes sampled from it" % (
            "FB15K-shaped" if name == "m4" else "1M-entity / 10M-triple", synth.checksum(tr))
    return (util.materialize_wn18(tempfile.mkdtemp()),
            "WN18 graph (repacked reference benchmark files, tests/golden/wn18.npz); universes sampled from it")


def make_pu(path, seed_offset=0):
    from openke.config import Parallel_Universe_Config
    from openke.data import TrainDataLoader, TestDataLoader
    train = TrainDataLoader(in_path=path, nbatches=NBATCHES, threads=8, sampling_mode="normal", bern_flag=0, filter_flag=0,
                            neg_ent=K_NEG, neg_rel=0, random_seed=123)
    test = TestDataLoader(path, "link")          # re-seeds the shared state with 4, like the static script
    pu = Parallel_Universe_Config(training_identifier="bench", train_dataloader=train, test_dataloader=test,
                                  initial_num_universes=None, embedding_model=model_class(), embedding_model_param=MODEL_PARAM,
                                  checkpoint_dir=None, valid_steps=10 ** 9, save_steps=None, training_setting="static",
                                  incremental_strategy=None, **STATIC_RANGES)
    pu.initial_random_seed += seed_offset
    if os.environ.get("PK_NO_PREFETCH"):
        pu.prefetch_sampling = False
    if os.environ.get("PK_SLOTS"):
        pu.launch_slots = int(os.environ["PK_SLOTS"])
    return pu


N_SETS = 4   # resident input sets of the device-resident leg (together larger than the 126 MB L2)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from openke import _native as N
    rank, world, local = dist_env()
    N.require_cuda()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    path, data_text = workload_dataset(args.workload)
    strong = args.workload in STRONG
    nU = args.universes
    # m2: every rank trains its own nU universes per step (weak scaling); m4: the step's nU universes are
    # sharded over the ranks (strong scaling)
    my_ids = [u for u in range(nU) if u % world == rank] if strong else list(range(nU))

    # ---- device-resident leg: N_SETS resident copies of the step's inputs, launched round robin on N_SETS
    # streams with no host synchronisation between steps (a job of many chunks keeps the SMs full this way:
    # a launch is as long as its slowest universe and 100 universes occupy 100 of 148 SMs)
    pu = make_pu(path, seed_offset=0 if strong else rank * nU)
    pu.record_losses = True
    launch = prepare_resident_launch(pu, my_ids, dev, N_SETS)
    evs = []

    def one_step(i, timed):
        s = launch["sets"][i % N_SETS]
        with torch.cuda.stream(s["stream"]):
            s["reset"]()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(s["stream"])
            s["run"]()
            ev1.record(s["stream"])
        if timed:
            evs.append((ev0, ev1))

    # one launch alone on an idle GPU: the latency of a single chunk (what round 1 reported as the step)
    torch.cuda.synchronize()
    single_ms = []
    for i in range(max(3, args.warmup)):
        one_step(i, True)
        torch.cuda.synchronize()
        single_ms.append(evs[-1][0].elapsed_time(evs[-1][1]))
    del evs[:]
    # untimed: the library's descriptor/scratch buffers (one per launch in flight, up to 16) are allocated at first use, and
    # an allocation waits behind the launches in flight — queue as many launches as the timed region will before timing
    for i in range(max(args.warmup, min(args.steps, 16))):
        one_step(i, False)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    clocks = ClockSampler(local)
    clocks.start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        one_step(i, True)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    elapsed = time.perf_counter() - t0
    clk = clocks.stop()
    launch_ms = [a.elapsed_time(b) for a, b in evs]
    region_ms = evs[0][0].elapsed_time(evs[-1][1]) if len(evs) > 1 else launch_ms[0]
    for a, b in evs:   # with several streams the last launch issued need not be the last to finish
        region_ms = max(region_ms, evs[0][0].elapsed_time(b))
    launches = launch["launches"] * args.steps
    positives = launch["positives"]
    final_loss = float(launch["sets"][0]["loss"][-1].item())
    util_ = sm_time_util(pu, launch, dev, args.steps)
    if dist:
        t = torch.tensor([elapsed, region_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed, region_ms = float(t[0]), float(t[1])
        tot = torch.tensor([float(positives)], device=dev, dtype=torch.float64)
        dist.all_reduce(tot)
        positives_all = float(tot[0])
    else:
        positives_all = float(positives)
    value = positives_all * args.steps / elapsed

    # ---- end-to-end leg through the public API: one Parallel_Universe_Config (loaders built once,
    # as a user would), every step = train_parallel_universes(...) on the NEXT universes of the seed
    # sequence: subgraph sampling (bit-exact walk on the GPU, launched two chunks ahead; host threads + H2D of the triple
    # index when filter/Bernoulli sampling needs the full universe helpers) + table init + descriptors + K2 + D2H of the
    # universe sizes and of the per-step losses into pinned memory.  async_training: a call returns when its launch is queued, so
    # consecutive chunks overlap on the device exactly like the resident leg; the final synchronize() is
    # inside the timed region.
    e2e_steps = args.e2e_steps if args.e2e_steps > 0 else max(20, args.steps)
    per_call = nU if strong else nU * world    # the orchestrator shards universe ids over the ranks itself
    p2 = make_pu(path)
    p2.record_losses = True
    p2.async_training = True
    # warm-up: slabs, the scratch / pinned buffers of every launch slot and of every walk in flight (each is allocated at
    # its first use, and an allocation waits for the launches in flight)
    for _ in range(max(args.warmup, p2.launch_slots + p2.device_walk_depth + 3)):
        p2.train_parallel_universes(per_call)
    p2.synchronize()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    p2.timings.clear()
    pos0 = p2.positive_triples
    h2d0, d2h0 = p2.h2d_bytes, p2.d2h_bytes
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        p2.train_parallel_universes(per_call)
    p2.synchronize()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    e2e_elapsed = time.perf_counter() - t0
    e2e_pos = p2.positive_triples - pos0
    h2d = (p2.h2d_bytes - h2d0) // e2e_steps
    d2h = (p2.d2h_bytes - d2h0) // e2e_steps
    timings = {k_: v_ / e2e_steps for k_, v_ in p2.timings.items()}
    loss_check = float(np.mean([p2.universe_losses[u][-1] for u in list(p2.universe_losses)[-5:]]))
    if dist:
        t = torch.tensor([e2e_elapsed], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(t[0])
        tot = torch.tensor([float(e2e_pos)], device=dev, dtype=torch.float64)
        dist.all_reduce(tot)
        e2e_pos = float(tot[0])
    e2e_value = e2e_pos / e2e_elapsed

    # ---- evaluation leg: link prediction of the universes trained by the e2e leg over the test set (both sides,
    # raw + filtered; min-energy aggregation, NCCL min all-reduce when the universes are sharded over ranks).
    # Reported beside the throughput, not part of `value`.
    ev = None
    if not args.no_eval:
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        mrr, mr, hit10, hit3, hit1 = p2.run_link_prediction()
        torch.cuda.synchronize()
        ev_s = time.perf_counter() - t0
        p2._rank_cache.clear()                    # the ranks again, from the energies: what a validation loop pays for chunks it
        t0 = time.perf_counter()                  # has seen before (their index and work items are cached)
        p2.run_link_prediction()
        torch.cuda.synchronize()
        ev_s2 = time.perf_counter() - t0
        ev = {"universes": int(p2.next_universe_id), "test_triples": int(p2.last_ranks.shape[0]), "seconds": ev_s,
              "test_triples_per_s": p2.last_ranks.shape[0] / ev_s, "repeat_seconds": ev_s2,
              "repeat_test_triples_per_s": p2.last_ranks.shape[0] / ev_s2,
              "filtered": {"mrr": float(mrr), "mr": float(mr), "hits10": float(hit10), "hits3": float(hit3), "hits1": float(hit1)}}
        if args.workload == "m2" and not strong:  # the ensemble size VERDICT r1 quoted: 400 universes in total
            p4 = make_pu(path)
            p4.train_parallel_universes(400)
            p4.synchronize()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m4_ = p4.run_link_prediction()
            torch.cuda.synchronize()
            e1 = time.perf_counter() - t0
            p4._rank_cache.clear()
            t0 = time.perf_counter()
            p4.run_link_prediction()
            torch.cuda.synchronize()
            e2 = time.perf_counter() - t0
            ev["ensemble_of_400"] = {"universes": 400, "seconds": e1, "test_triples_per_s": p4.last_ranks.shape[0] / e1, "repeat_seconds": e2,
                                     "repeat_test_triples_per_s": p4.last_ranks.shape[0] / e2, "filtered_mrr": float(m4_[0]),
                                     "filtered_hits10": float(m4_[2])}
            del p4

    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    bpp = algorithmic_bytes_per_positive(MODEL, 20, K_NEG, "adagrad")
    eff_launch_ms = region_ms / args.steps          # launches overlap: device time of the region per launch
    achieved = positives * bpp / (eff_launch_ms * 1e-3) / 1e9
    names = {"transe": "TransE", "transh": "TransH", "transd": "TransD"}
    line = {
        "metric": "PuTransE positive triples/sec (all universes)", "value": value, "unit": "positive triples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
        "data": data_text,
        "config": bench_config(args, world),
        "e2e": {"value": e2e_value, "unit": "positive triples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e2e_elapsed / e2e_steps * 1e3,
                "host_breakdown_s_per_step": {k: v for k, v in timings.items()},
                "mean_final_loss_last_universes": loss_check,
                "host": {"cores": len(os.sched_getaffinity(0)), "ranks_on_host": world,
                         "universes_built_on": "gpu" if getattr(p2, "_walker", None) is not None and p2._walker.launches else "host threads",
                         "note": "subgraphs of the next chunks are sampled beside the launch (bit-exact glibc rand() walk): on the GPU "
                                 "(k_walk_universes) for unfiltered non-Bernoulli training, else on host threads"}},
        "gpu_launches": launches,
        "positive_triples_per_step_per_gpu": positives, "final_loss_last_universe": final_loss,
        "launch_ms": {"alone_on_the_gpu": float(np.median(single_ms)), "inside_the_pipelined_region": float(np.mean(launch_ms)),
                      "region_ms_per_launch": eff_launch_ms},
        "single_chunk_value": positives / (float(np.median(single_ms)) * 1e-3),
        "sm_time_util": util_,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": k2_traffic(MODEL), "peak_source": peak_src, "algorithmic_bytes_per_positive": bpp,
                     "units_per_launch": positives, "launch_duration_ms": eff_launch_ms,
                     "note": "k2_train_universes; launches of consecutive steps overlap, so the duration is the timed region's device "
                             "time per launch (CUDA events on the launch streams).  The kernel is latency-bound, not HBM-bound: "
                             "universe tables are shared-memory/L2 resident (traffic << algorithmic bytes), see DESIGN.md and "
                             "sm_time_util"},
        "clocks": clk,
    }
    if ev is not None:
        line["eval"] = ev
    if world == 1 and not args.no_extras and args.workload == "m2" and MODEL == "transe":
        line["extras"] = run_extras(args, path, dev)
    if world == 1 and not args.no_s1:
        line["roofline_s1"] = s1_roofline(args)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = reference_arm(path, budget_s=args.cpu_budget) or cpu_reference_leg(path, budget_s=args.cpu_budget)
    _OUT.emit(json.dumps(line))


def bench_config(args, world):
    names = {"transe": "TransE", "transh": "TransH", "transd": "TransD"}
    return {"workload": args.workload + ": " + WORKLOADS[args.workload] % dict(n=args.universes, model=names[MODEL], nb=NBATCHES),
            "universes_per_step": args.universes, "l2": "inputs larger than L2: %d resident input sets (tables + optimizer state + "
            "triple index + initial-table copies, ~50 MB each at 100 universes) used round robin; no flush" % N_SETS,
            "pipelining": "steps are launched on %d streams without host synchronisation between them" % N_SETS}


def sm_time_util(pu, launch, dev, steps):
    """Sum over universes of the time their thread block ran / (SMs x device time of the region), for one
    isolated launch and for a pipelined round of launches (pk_debug_universe_timer: %globaltimer per block)."""
    import torch
    lib = pu.lib
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    n = launch["n"]
    out = {"sms": sms}
    try:
        bufs = [torch.zeros(2 * n, dtype=torch.int64, device=dev) for _ in range(N_SETS)]

        def round_(k):
            torch.cuda.synchronize()
            evs = []
            for i in range(k):
                s = launch["sets"][i % N_SETS]
                with torch.cuda.stream(s["stream"]):
                    s["reset"]()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(s["stream"])
                    lib.pk_debug_universe_timer(bufs[i % N_SETS].data_ptr())
                    s["run"]()
                    b.record(s["stream"])
                    evs.append((a, b))
            torch.cuda.synchronize()
            lib.pk_debug_universe_timer(None)
            region = max(evs[0][0].elapsed_time(b) for _, b in evs)
            busy = sum(float(bufs[i % N_SETS][1::2].sum().item()) for i in range(k)) / 1e6
            return busy, region
        busy, region = round_(1)
        out["single_launch"] = busy / (sms * region)
        out["slowest_universe_ms"] = float(bufs[0][1::2].max().item()) / 1e6
        busy, region = round_(N_SETS)
        out["pipelined_round_of_%d" % N_SETS] = busy / (sms * region)
    except Exception as e:   # diagnostics only
        out["error"] = "%s: %s" % (type(e).__name__, e)
    return out


def _child_json(cmd, timeout=900):
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    if not lines:
        raise RuntimeError("no JSON line from %s: %s" % (" ".join(cmd[-6:]), out.stderr[-400:]))
    return json.loads(lines[-1])


def run_extras(args, path, dev):
    """Other configurations BASELINE.json names, measured in the same run (N = 1): ten times the universes per
    GPU, PuTransH / PuTransD (configs[2]), the single-space Trainer (configs[0]) and the FB15K-shaped graph
    (configs[3]).  Each is guarded: a failing extra must not cost the headline line."""
    import torch
    ex = {}
    try:   # 1000 universes per GPU through the public API, one call
        p3 = make_pu(path)
        p3.async_training = True
        for _ in range(p3.launch_slots + p3.device_walk_depth + 2):   # every slot's buffers at this size
            p3.train_parallel_universes(1000)
        p3.synchronize()
        pos0 = p3.positive_triples
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            p3.train_parallel_universes(1000)
        p3.synchronize()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ex["m2_1000_universes_per_call"] = {"e2e_value": (p3.positive_triples - pos0) / dt, "unit": "positive triples/s",
                                            "calls": 4, "ms_per_call": dt / 4 * 1e3}
        del p3
    except Exception as e:
        ex["m2_1000_universes_per_call"] = {"error": "%s: %s" % (type(e).__name__, e)}
    try:   # the static WN18 experiment as its script runs it (experiments/static_experiment_PuTransE_on_WN18.py:43-110), 1/3 length:
        # 2 000 universes, validation on the valid split every 100 universes (early stopping off), then the test evaluation
        import contextlib
        import io
        p5 = make_pu(path)
        p5.valid_steps, p5.early_stopping_patience = 100, 10 ** 9
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            p5.train_parallel_universes(2000)
            t1 = time.perf_counter()
            m5 = p5.run_link_prediction()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        ex["static_experiment_2000_universes_valid_every_100"] = {
            "seconds_training_with_20_validations": t1 - t0, "seconds_test_evaluation": t2 - t1, "positive_triples": int(p5.positive_triples),
            "positive_triples_per_s_including_validation": p5.positive_triples / (t1 - t0), "best_valid_hits10": float(p5.best_hit10),
            "test_filtered": {"mrr": float(m5[0]), "mr": float(m5[1]), "hits10": float(m5[2])},
            "note": "a validation folds only the 100 new universes into the resident min-energy matrix of the valid split and runs behind "
                    "the next chunk's launch; first use of a fresh object: includes slab / matrix allocation"}
        del p5
    except Exception as e:
        ex["static_experiment_2000_universes_valid_every_100"] = {"error": "%s: %s" % (type(e).__name__, e)}
    me = [sys.executable, os.path.abspath(__file__), "--no-extras", "--no-s1", "--no-cpu-baseline", "--steps", "8", "--warmup", "3"]
    for key, extra in (("putransh", ["--model", "transh", "--no-eval"]), ("putransd", ["--model", "transd", "--no-eval"]),
                       ("m4_fb15k_shape_1000_universes", ["--workload", "m4", "--universes", "1000", "--steps", "3", "--e2e-steps", "6"])):
        try:
            d = _child_json(me + extra)
            ex[key] = {"value": d["value"], "e2e_value": d["e2e"]["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"],
                       "e2e_ms_per_step": d["e2e"]["ms_per_step"], "workload": d["config"]["workload"]}
            if "eval" in d:
                ex[key]["eval"] = d["eval"]
        except Exception as e:
            ex[key] = {"error": "%s: %s" % (type(e).__name__, e)}
    for key, degree, opt in (("s1_uniform_degree_adagrad", "uniform", "adagrad"), ("s1_uniform_degree_sgd", "uniform", "sgd"),
                             ("s1_power_law_sgd", "power", "sgd")):
        ex[key] = s1_roofline(args, degree=degree, opt=opt)
    try:
        ex["c1_transe_wn18_trainer_run"] = _child_json([sys.executable, os.path.join(REPO, "tools", "bench_c1.py"), "30", "--json"])
    except Exception as e:
        ex["c1_transe_wn18_trainer_run"] = {"error": "%s: %s" % (type(e).__name__, e)}
    return ex


def s1_roofline(args, degree="power", opt="adagrad"):
    """The HBM-bound configuration of the same train step (SURVEY.md 8(d) "S1"): one embedding space
    with 1 M entities / 10 M triples, TransE d=64, k=1, B=100 000, Adagrad — tables far beyond L2, so
    every gathered row comes from DRAM.  Run as a child process (tools/bench_k1.py) so that its 1.5 GB
    of tables do not stay resident; its JSON line is embedded."""
    try:
        d = _child_json([sys.executable, os.path.join(REPO, "tools", "bench_k1.py"), "--opt", opt, "--steps", "100", "--reps", "3",
                         "--degree", degree], timeout=600)
        r = d["roofline"]
        tr = k1_traffic(opt, degree)
        if tr is not None:    # dram__bytes_read + dram__bytes_write of k1_grad from the committed ncu capture, per launch
            r["traffic"] = tr
        st_ = k1_step_traffic(opt, degree)
        if st_ is not None:   # all kernels of a step, caches not flushed between them (profiles/r2_k1_steady_state_dram_*.csv)
            r["traffic_whole_step_steady_state"] = st_
            r["frac_measured_dram"] = st_ / (d["us_per_step"] * 1e-6) / 1e9 / r["peak"]
        r.update(workload=d["workload"], kernel="k1_prepare + k1_grad + k1_apply (whole step)", us_per_step=d["us_per_step"],
                 positive_triples_per_s=d["positive_triples_per_s"], algorithmic_bytes_per_positive=d["algorithmic_bytes_per_positive"])
        return r
    except Exception as e:   # the headline line must not depend on this extra
        return {"error": "%s: %s" % (type(e).__name__, e)}


def prepare_resident_launch(pu, ids, dev, n_sets=1):
    """Everything _train_chunk does before the kernel, kept resident in n_sets independent copies (own
    tables, optimizer state, loss buffer, triple index and stream each); returns per set closures that reset
    the tables to their initial values and relaunch K2."""
    import torch
    from openke import _native as N
    ck = pu._train_chunk(ids)           # first (untimed) training: samples the subgraphs, leaves the index resident
    by_head0 = ck.train_inputs[0]
    pu.synchronize()
    torch.cuda.synchronize()
    by_head0 = by_head0.clone()
    # initial tables again (same seeds): rebuild them on the host exactly as the reference does
    init = {}
    for name in ck.proto.table_names():
        init[name] = torch.empty_like(ck.tables[name])
    seeds = [pu.initial_random_seed + u for u in ids]
    for i, u in enumerate(ids):
        torch.manual_seed(seeds[i])
        sp = pu.embedding_model(int(ck.nE[i]), int(ck.nR[i]), **pu.embedding_model_param)
        for name in sp.table_names():
            o = ck.eoff if name in sp._ent_tables else ck.roff
            init[name][o[i]:o[i + 1]].copy_(getattr(sp, name).weight.data)
    nb = pu.train_dataloader.nbatches
    n = len(ids)
    W = pu.train_dataloader.work_threads
    lib = pu.lib
    # re-derive the descriptors (the library call copied them; we need our own for replays)
    s = np.array(seeds, dtype=np.int64)
    tcs = np.array([pu.universe_hyper[u]["tc"] for u in ids], dtype=np.int64)
    bals = np.array([pu.universe_hyper[u]["balance"] for u in ids], dtype=np.float32)
    h = lib.pk_universes_build(n, N.addr(s), N.addr(tcs), N.addr(bals), 0)
    lcg = np.zeros((n, W), dtype=np.uint64)
    N.check(lib.pk_universes_export(h, None, None, None, None, None, None, None, N.addr(lcg)))
    lib.pk_universes_free(h)
    hy = [pu.universe_hyper[u] for u in ids]
    epochs = np.array([h_["epochs"] for h_ in hy], dtype=np.int64)
    Bs = np.array([h_["batch_size"] for h_ in hy], dtype=np.int64)
    steps = epochs * nb
    darr = np.zeros(n, dtype=N.UNIVERSE_DESC_DTYPE)
    darr["tri_off"], darr["ent_off"], darr["rel_off"] = ck.tri_off[:n], ck.eoff[:n], ck.roff[:n]
    darr["n_tri"], darr["n_ent"], darr["n_rel"] = ck.nT, ck.nE, ck.nR
    darr["batch_size"], darr["nbatches"], darr["epochs"] = Bs, nb, epochs
    darr["margin"] = np.array([h_["margin"] for h_ in hy], dtype=np.float32)
    darr["lr"] = np.array([h_["lr"] for h_ in hy], dtype=np.float32)
    darr["loss_off"] = np.cumsum(steps) - steps
    darr["lcg"][:, :min(W, 8)] = lcg[:, :min(W, 8)]
    desc = (N.UniverseDesc * n).from_buffer(darr)
    loss_total, positives = int(steps.sum()), int((steps * Bs).sum())
    cfg = ck.proto.native_cfg(opt=N.PK_ADAGRAD, neg_ent=K_NEG, bern=0, filt=0, work_threads=W)
    state = {"launches": 0}
    sets = []
    for j in range(n_sets):
        tables = {name: t.clone() for name, t in init.items()}
        opt_state = {name: torch.zeros_like(t) for name, t in init.items()}
        by_head = by_head0.clone()
        d_loss = torch.zeros(loss_total, dtype=torch.float32, device=dev)
        stream = torch.cuda.Stream(device=dev)
        tab = N.Tables()
        for i in range(2):
            tab.ent[i] = tab.rel[i] = tab.ent_state[i] = tab.rel_state[i] = None
        for i, name in enumerate(ck.proto._ent_tables):
            tab.ent[i], tab.ent_state[i] = tables[name].data_ptr(), opt_state[name].data_ptr()
        for i, name in enumerate(ck.proto._rel_tables):
            tab.rel[i], tab.rel_state[i] = tables[name].data_ptr(), opt_state[name].data_ptr()
        tab.n_ent, tab.n_rel = int(ck.eoff[-1]), int(ck.roff[-1])

        def reset(tables=tables, opt_state=opt_state):
            for name, t in tables.items():
                t.copy_(init[name])
                opt_state[name].zero_()

        def run(tab=tab, by_head=by_head, d_loss=d_loss, stream=stream):
            N.check(lib.pk_train_universes(ctypes.byref(cfg), ctypes.byref(tab), by_head.data_ptr(), None, None, None, desc, n,
                                           d_loss.data_ptr(), stream.cuda_stream), "pk_train_universes")
            state["launches"] = lib.pk_last_launch_count()

        sets.append({"reset": reset, "run": run, "loss": d_loss, "stream": stream, "tables": tables, "keep": (tab, by_head, opt_state)})
    torch.cuda.synchronize()
    with torch.cuda.stream(sets[0]["stream"]):
        sets[0]["reset"]()
        sets[0]["run"]()
    torch.cuda.synchronize()
    return {"sets": sets, "launches": state["launches"], "positives": positives, "n": n, "keep": (darr, desc, init)}


# ------------------------------------------------------------------------------------------------
_TUNED = {}


def cpu_reference_leg(path, budget_s=20.0, universes=None, max_epochs=None):
    """The reference's CPU path for the same workload on this box's host cores: reference Base.so
    (oracle/_ref, compiled unmodified from /root/reference) for subgraph + batch sampling when it is
    present — else the oracle's C++ restatement — and the op-for-op torch port of the reference's
    Python training loop (oracle/model_math.py).  Thread counts (torch intra-op threads and the
    reference sampler's pthreads) are auto-tuned first on a few untimed steps so that the baseline
    is the best this box's host cores can do: at universe scale (B~50, d=20) more threads are
    usually slower.  Bounded sample: universes 0,1,... of the same seed sequence until `budget_s`
    seconds of CPU work are spent (or exactly `universes` of them)."""
    import torch
    from oracle import native as on
    from oracle.model_math import TorchOracle
    from openke.config import Parallel_Universe_Config
    TransE = model_class()
    cores = os.cpu_count() or 1
    w = np.load(os.path.join(REPO, "tests", "golden", "wn18.npz"))
    R = on.load_reference()
    kind = "port"
    devnull = saved = None
    if R is not None:
        R.setInPath(ctypes.create_string_buffer(path.encode(), len(path) * 2))
        R.setBern(0)
        R.setWorkThreads(8)
        R.setRandomSeed(4)
        R.randReset()
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        sys.stdout.flush()
        os.dup2(devnull, 1)          # the reference printf()s on every call
    else:
        o = on.Oracle(threads=8, bern=0)
        o.import_train(w["train"], 40943, 18)
    pu = Parallel_Universe_Config.__new__(Parallel_Universe_Config)
    for k_, v_ in STATIC_RANGES.items():
        setattr(pu, k_, v_)
    pu.const_num_epochs = max_epochs

    def run_universe(u, step_cap=None):
        seed = 4 + u
        hy = pu.draw_universe_hyper(seed)
        torch.manual_seed(seed)
        if R is not None:
            R.setRandomSeed(seed)
            R.randReset()
            R.getParallelUniverse(hy["tc"], ctypes.c_float(hy["balance"]))
            nT, nE, nR = R.getTrainTotalUniverse(), R.getEntityTotalUniverse(), R.getRelationTotalUniverse()
            R.swapHelpers()
        else:
            o.seed(seed)
            tri, er, rr = o.universe(hy["tc"], hy["balance"])
            nT, nE, nR = tri.shape[0], er.shape[0], rr.shape[0]
            o.swap()
        B = nT // NBATCHES
        ref = TransE(nE, nR, **MODEL_PARAM)
        orc = TorchOracle(MODEL, {n: getattr(ref, n).weight.detach().numpy() for n in ref.table_names()}, p_norm=1,
                          opt="adagrad", lr=hy["lr"], margin=hy["margin"], k=K_NEG)
        n = B * (1 + K_NEG)
        bh, bt, br, by = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.float32)
        steps = hy["epochs"] * NBATCHES if step_cap is None else min(step_cap, hy["epochs"] * NBATCHES)
        for _ in range(steps):
            if R is not None:
                R.sampling(on._addr(bh), on._addr(bt), on._addr(br), on._addr(by), B, K_NEG, 0, 0, 0, 0, 0)
                orc.step(bh, bt, br)
            else:
                orc.step(*o.sampling(B, K_NEG, 0))
        if R is not None:
            R.resetUniverse()
        else:
            o.swap()
        return steps, steps * B

    try:
        if R is not None:
            R.importTrainFiles()
        if not _TUNED:
            best = None
            for tt in sorted({1, 2, 4, cores}):
                for stt in ((1, 8) if R is not None else (8,)):
                    torch.set_num_threads(tt)
                    if R is not None:
                        R.setWorkThreads(stt)
                    run_universe(0, step_cap=5)
                    t0 = time.perf_counter()
                    run_universe(0, step_cap=40)
                    dt = time.perf_counter() - t0
                    if best is None or dt < best[0]:
                        best = (dt, tt, stt)
            _TUNED.update(torch_threads=best[1], sampler_threads=best[2])
        torch.set_num_threads(_TUNED["torch_threads"])
        if R is not None:
            R.setWorkThreads(_TUNED["sampler_threads"])
        t0 = time.perf_counter()
        positives, u, steps_total = 0, 0, 0
        while True:
            st_, pos_ = run_universe(u)
            steps_total += st_
            positives += pos_
            u += 1
            if (universes is not None and u >= universes) or (universes is None and time.perf_counter() - t0 > budget_s):
                break
        dt = time.perf_counter() - t0
    finally:
        if saved is not None:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(devnull)
            os.close(saved)
    used = max(_TUNED["torch_threads"], _TUNED["sampler_threads"])
    return {"value": positives / dt, "unit": "positive triples/s", "cores": used, "kind": kind,
            "sample": "universes 0..%d of the same seed sequence (%d steps, %d positive triples, %.1f s): %s sampling + torch %s CPU "
                      "port of the reference training loop; auto-tuned for best throughput on this host (%d cores): torch threads=%d, "
                      "sampler threads=%d" % (u - 1, steps_total, positives, dt,
                                              "reference Base.so (oracle/_ref)" if R is not None else "oracle C++", torch.__version__,
                                              cores, _TUNED["torch_threads"], _TUNED["sampler_threads"])}


REF_PKG = os.path.join(REPO, "baseline", "_ref")


def reference_arm(path, budget_s=0.0, steps=1, warmup=0, universes_per_step=1, soft_limit_s=0.0, raw=False):
    """The reference ITSELF on this box's host cores: its unmodified Python package
    (Parallel_Universe_Config.train_parallel_universes, reference :316-367, use_gpu False) on its own native core,
    from baseline/_ref (installed by __graft_entry__.build(), git-ignored), in a child interpreter
    (tools/ref_arm.py).  None when baseline/_ref is not there."""
    if not os.path.exists(os.path.join(REF_PKG, "openke", "release", "Base.so")):
        return None
    cmd = [sys.executable, os.path.join(REPO, "tools", "ref_arm.py"), "--pkg", REF_PKG, "--data", path, "--model", MODEL,
           "--nbatches", str(NBATCHES), "--universes", str(universes_per_step), "--steps", str(steps), "--warmup", str(warmup)]
    if budget_s:
        cmd += ["--budget-s", str(budget_s), "--steps", "1000"]
    if soft_limit_s:
        cmd += ["--soft-limit-s", str(soft_limit_s)]
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    out = subprocess.run(cmd, capture_output=True, text=True, env=env)
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    if not lines:
        sys.stderr.write("reference arm failed: %s\n" % out.stderr[-2000:])
        return None
    d = json.loads(lines[-1])
    if raw:
        return d
    us = [x["universes"] for x in d["per_step"]]
    return {"value": d["value"], "unit": "positive triples/s", "cores": d["cores"], "kind": "reference",
            "sample": "the unmodified reference package (baseline/_ref: its Python + its Base.so) on universes %d..%d of the same seed "
                      "sequence at their drawn epochs: %d train steps, %d positive triples in %.1f s; torch %s with %d intra-op threads, "
                      "8 sampler pthreads, host has %d cores (the faster of all-threads / one thread on a 4-epoch universe: %s s)%s" % (
                          us[0][0], us[-1][1], d["train_steps"], d["positives"], d["seconds"], d["torch"], d["torch_threads"], d["cores"],
                          d.get("thread_calibration_s"),
                                                                  "; %d late steps ran with epochs capped at 5" % d["steps_with_capped_epochs"]
                                                                  if d["steps_with_capped_epochs"] else "")}


def run_reference(args):
    """`--impl reference`: one step = ONE universe of the workload trained by the unmodified reference at its drawn
    epochs (8-27 s of host time); step i trains universe i of the same seed sequence as our arm."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    path, data_text = workload_dataset(args.workload)
    t0 = time.perf_counter()
    d = reference_arm(path, steps=args.steps, warmup=args.warmup, universes_per_step=args.ref_universes, soft_limit_s=args.ref_soft_limit,
                      raw=True)
    if d is None:     # no baseline/_ref on this box: the CPU port (oracle/) of the same loop
        res = [cpu_reference_leg(path, universes=args.ref_universes) for _ in range(args.steps)]
        v = float(np.mean([r["value"] for r in res]))
        base = dict(res[-1])
        base["value"] = v
        ms = (time.perf_counter() - t0) / args.steps * 1e3
    else:
        v = d["value"]
        ms = d["seconds"] / max(1, d["steps_done"]) * 1e3
        base = {"value": v, "unit": "positive triples/s", "cores": d["cores"], "kind": "reference",
                "sample": "step i = universe i (seeds 4..) trained by the unmodified reference package from baseline/_ref "
                          "(Parallel_Universe_Config.train_parallel_universes(%d), use_gpu False) at its drawn 50-199 epochs: %d train steps, "
                          "%d positive triples in %.1f s; torch %s with %d intra-op threads (the faster of all-threads / one thread on a "
                          "4-epoch universe: %s s) + 8 sampler pthreads on %d host cores; warm-up "
                          "steps are 2-epoch universes%s" % (args.ref_universes, d["train_steps"], d["positives"], d["seconds"], d["torch"],
                                                             d["torch_threads"], d.get("thread_calibration_s"), d["cores"],
                                                             "; %d late steps ran with epochs capped at 5 (soft time limit)" %
                                                             d["steps_with_capped_epochs"] if d["steps_with_capped_epochs"] else "")}
    line = {"impl": "reference", "metric": "PuTransE positive triples/sec (all universes)", "value": v,
            "unit": "positive triples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None,
            "dtype": "f32", "data": data_text, "config": bench_config(args, world),
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "positive triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _OUT.emit(json.dumps(line))


class _StdoutToStderr(object):
    """Everything the run prints (Python and the library's C printf) goes to stderr; stdout carries
    exactly ONE line, the JSON result, written by emit()."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


_OUT = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--universes", type=int, default=100)
    ap.add_argument("--workload", default="m2", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=0, help="end-to-end steps (0 = as many as --steps, at least 20)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--ref-universes", type=int, default=1, help="reference arm: universes per step")
    ap.add_argument("--ref-soft-limit", type=float, default=1200.0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--model", default="transe", choices=["transe", "transh", "transd"])
    ap.add_argument("--no-s1", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    set_model(args.model)
    global _OUT
    with _StdoutToStderr() as _OUT:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
