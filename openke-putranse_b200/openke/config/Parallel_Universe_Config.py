"""PuTransE orchestrator: many small embedding spaces ("universes") trained on random-walk
subgraphs, combined at evaluation time by a minimum over universes.

Constructor kwargs, public attributes and method names follow the reference's
``openke/config/Parallel_Universe_Config.py:62-946``.  What changed is where the work happens:

reference (per universe, sequential)                  here (per CHUNK of universes)
-------------------------------------------------     ---------------------------------------------
set_random_seed; Python draws tc, balance             same draws, same order, private Random(seed)
getParallelUniverse (single-threaded, libc rand)      pk_universes_build: all universes of the chunk
                                                      on host threads, bit-identical subgraphs
torch.manual_seed + model init                        same (torch CPU generator), packed + one H2D
Trainer.run: epochs x nbatches Python steps           pk_train_universes: ONE launch, a block/universe
eval_universes: Python loop triple x universe with    pk_universe_energies + (NCCL min all-reduce when
  per-entity .item() min                                universes are sharded over ranks) +
testHead/testTail in C on the host                    pk_rank_from_energy

Universes are independent given ``initial_random_seed + universe_id`` (reference :324), so with
``torch.distributed`` initialised each rank trains the universes ``u % world_size == rank`` with
no communication; evaluation min-all-reduces the energy tiles.
"""
import ctypes
import os
import time
from collections import defaultdict
from random import Random

import numpy as np
import torch

from .. import _native as N
from ..module.loss import MarginLoss
from ..module.model.Model import Model
from ..module.strategy import NegativeSampling
from .Tester import Tester, link_metrics
from ._universe_maps import SpaceMap, LocalIdMaps, MembershipMap
from ..data.TestDataLoader import TestDataLoader


def get_string_key(entity, relation):
    return "{},{}".format(entity, relation)


def defaultdict_int(innerfactory=int):
    return defaultdict(innerfactory)


def float_default():
    """Module-level factory the reference pickles inside its evaluation caches (reference :51-52)."""
    return float("inf")


def _host_cores():
    """Cores this process may run on (a container's CPU set can be smaller than os.cpu_count())."""
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 2


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


class _Chunk(object):
    """One pk_train_universes launch worth of universes: packed device tables + host-side maps."""
    d_ent_remap = d_rel_remap = None      # device copies of the packed remaps (device-built universes)
    _ent_remap = _rel_remap = None

    # local -> global ids of all universes of the chunk, back to back; universes built on the GPU leave them on the
    # device (evaluation reads them there) and the host copy is fetched when something first asks for it
    @property
    def ent_remap(self):
        if self._ent_remap is None and self.d_ent_remap is not None:
            self._ent_remap = self.d_ent_remap.cpu().numpy()
        return self._ent_remap

    @ent_remap.setter
    def ent_remap(self, v):
        self._ent_remap = v

    @property
    def rel_remap(self):
        if self._rel_remap is None and self.d_rel_remap is not None:
            self._rel_remap = self.d_rel_remap.cpu().numpy()
        return self._rel_remap

    @rel_remap.setter
    def rel_remap(self, v):
        self._rel_remap = v


class Parallel_Universe_Config(Tester):
    def __init__(self, train_dataloader=None, training_identifier="", valid_dataloader=None, test_dataloader=None,
                 initial_num_universes=5000, min_margin=1, max_margin=4, min_lr=0.01, max_lr=0.1, min_num_epochs=50,
                 max_num_epochs=200, const_num_epochs=None, min_triple_constraint=500, max_triple_constraint=2000,
                 min_balance=0.25, max_balance=0.5, embedding_model=None, embedding_model_param=None,
                 missing_embedding_handling="last_rank", save_steps=5, checkpoint_dir="./checkpoint/", valid_steps=5,
                 early_stopping_patience=5, training_setting="static", incremental_strategy="normal"):
        super().__init__(data_loader=test_dataloader, use_gpu=torch.cuda.is_available())
        if training_setting not in ("static", "incremental"):
            raise ValueError("training_setting must be 'static' or 'incremental'")
        if missing_embedding_handling not in ("last_rank", "null_vector"):
            raise ValueError("missing_embedding_handling must be 'last_rank' or 'null_vector'")
        self.train_dataloader = train_dataloader
        self.ent_tot = train_dataloader.entTotal
        self.rel_tot = train_dataloader.relTotal
        self.training_identifier = training_identifier
        self.embedding_model = embedding_model
        self.embedding_model_param = embedding_model_param or {}

        self.initial_num_universes = initial_num_universes
        self.next_universe_id = 0
        # the reference's containers (reference :80-90), materialised lazily from the packed chunks
        self._chunks = []
        self._where = {}                                            # universe id -> (chunk, index in chunk)
        self._maps_version = 0
        self.trained_embedding_spaces = SpaceMap(self)              # universe id -> embedding space
        self.entity_id_mappings = LocalIdMaps(self, "ent")          # universe id -> global entity -> local
        self.relation_id_mappings = LocalIdMaps(self, "rel")
        self.entity_universes = MembershipMap(self, "ent")          # global entity -> universe ids
        self.relation_universes = MembershipMap(self, "rel")
        # the loaders share one library state: this is whatever seed was set last (reference :97)
        self.initial_random_seed = self.train_dataloader.lib.getRandomSeed()

        self.min_margin, self.max_margin = min_margin, max_margin
        self.min_lr, self.max_lr = min_lr, max_lr
        self.min_num_epochs, self.max_num_epochs = min_num_epochs, max_num_epochs
        self.const_num_epochs = const_num_epochs
        self.min_triple_constraint, self.max_triple_constraint = min_triple_constraint, max_triple_constraint
        self.min_balance, self.max_balance = min_balance, max_balance
        self.save_steps, self.checkpoint_dir = save_steps, checkpoint_dir
        self.missing_embedding_handling = missing_embedding_handling

        if valid_dataloader is None and training_setting == "incremental":
            raise ValueError("the incremental setting needs an explicit valid_dataloader (IncrementalTestDataLoader, mode='valid'): "
                             "the static default would re-import train2id.txt over the evolving training list")
        self.valid_dataloader = valid_dataloader if valid_dataloader is not None else TestDataLoader(
            train_dataloader.in_path, sampling_mode="link", mode="valid")
        self.valid_steps = valid_steps
        self.early_stopping_patience = early_stopping_patience
        self.early_stopping_patience_const = early_stopping_patience
        self.bad_counts = 0
        self.best_hit10 = 0
        self.best_state = None
        self.training_setting = training_setting
        self.incremental_strategy = incremental_strategy      # "normal" | "deprecate" (reference :146-148)
        self.deprecated_embeddingspaces = set()
        self._deprecated_version = 0

        # B200 build state
        self.universe_hyper = {}          # universe id -> dict(tc, balance, margin, epochs, lr, nT, nE, nR, focus)
        self._universe_losses = {}        # universe id -> np.float32 [epochs*nbatches] (if record_losses)
        self.record_losses = False
        self.sampler_threads = 0          # host threads for universe construction (0 = all cores)
        self.max_chunk = 1024             # universes per train_parallel_universes chunk
        # Launch slots.  Every chunk is trained on the stream of the next slot (round robin); a slot owns its
        # per-launch device buffers (triple index, means, Adagrad state, losses), so the launches of consecutive
        # chunks may be in flight together: chunk i+1's thread blocks fill the SMs that chunk i's short universes
        # have already left (a launch is as long as its slowest universe, and 100 universes cover 100 of 148 SMs).
        # Nothing waits for a launch until its results are needed (evaluation, checkpoints, the reference's
        # containers, per-step losses) or its slot comes round again.
        self.launch_slots = 4
        # False: train_parallel_universes returns when its universes are trained (the reference's behaviour, and
        # what its "Time took for creation of embedding spaces" line measures).  True: it returns when they
        # are LAUNCHED, so a caller that trains chunk after chunk keeps several launches in flight; anything
        # that reads results synchronises by itself (synchronize()).
        self.async_training = False
        # A validation that cannot end the training (bad_counts + 1 < early_stopping_patience) is run BEHIND the launch of
        # the next chunk, on the ensemble as it was when the validation fell due (`_view`): the GPU trains chunk i+1 while
        # chunk i's universes are folded into the valid split's energy matrix and ranked.  Same validations, same order,
        # same checkpoints; a validation that can stop the training runs before anything else is launched.
        self.pipeline_validation = True
        # chunks launched ahead of the oldest outstanding validation.  More than one does not pay: the training blocks are
        # long-running and own their SMs, so with three launches in flight the evaluation kernels starve (2 000 universes with
        # 20 validations: 0.27 s at lag 1, 0.58 s at lag 3, 0.48 s synchronous)
        self.validation_lag = 1
        self._view = None                 # (chunks visible, next_universe_id) while a deferred validation runs
        self._slots = None
        self._launch_index = 0
        self._pinned = {}
        self._pinned_busy = None
        self._arena, self._arena_used = None, 0
        # trained tables are carved from slabs (a cudaMalloc waits for the launches in flight: tens of ms when four are);
        # 1 GiB = the tables of ~80 chunks of 100 WN18 universes.  Untouched slab memory costs address space only.
        self.arena_slab_floats = 256 << 20
        # sample the next chunk's subgraphs on a host thread while the GPU trains this one (the universe
        # ids of the next call are predictable: they continue the sequence).  Measured on the B200 box:
        # 38.5 -> 27.5 ms per 100 universes end to end, i.e. host sampling disappears behind the kernel.
        self.prefetch_sampling = True
        self.prefetch_depth = 1           # chunks sampled ahead of the one being launched (deeper = the sampler's Python parts
                                          # compete with the launching thread for the GIL: launch time 1.3 -> 13 ms per chunk)
        self._prefetched = {}             # sampling key -> sample in flight (("host", future) or ("device", hyper, WalkResult))
        self._pool = None
        # Subgraphs on the GPU (csrc/walk_device.cu, openke/universe_walk.py) whenever training reads only the lean
        # universe (no filter, no Bernoulli): no host cores, no H2D copy of the triple index; bit-identical to the host
        # builder.  The walks of the next `device_walk_depth` chunks run beside the training launches.
        self.device_walk = True
        self.device_walk_depth = 2
        self._walker = None
        # Validation every `valid_steps` universes re-evaluates an ensemble that only GROWS: the [keys, E] min-energy matrix of
        # a key set (valid split, test split) stays resident on the device and an evaluation folds in only the chunks trained
        # since the previous one (min is associative and idempotent, so the ranks are those of a full evaluation).  1.6 GB
        # for a WN18 split; key sets above the budget are evaluated tile by tile from scratch as before.
        self.energy_cache = True
        self.energy_cache_bytes = 24 << 30
        self._energy_cache = {}
        self.max_energy_bytes = 8 << 30   # upper bound for the [keys, E] energy tile buffers of an evaluation (three of them)
        self.eval_tile_rows = 4096        # key rows per energy tile (a tile is the unit of the NCCL min all-reduce); 671 MB on WN18
        self.training_duration = 0.0
        self.positive_triples = 0         # sum over universes of epochs * nbatches * batch_size
        self.gpu_launches = 0
        self.universes_on_single_space_path = 0   # universes that did not fit the batched kernel (relation-rich graphs)
        self.h2d_bytes = 0                # bytes copied host -> device by training (index, descriptors, tables if host-initialised)
        self.d2h_bytes = 0                # bytes copied device -> host by training (per-step losses)
        self._rank_cache = {}
        self.timings = defaultdict(float)

    # ------------------------------------------------------------------ small reference-API helpers
    def get_default_value_list(self):
        return [float("inf") for _ in range(self.ent_tot)]

    def set_min_max_triple_constraint(self, min, max):
        self.min_triple_constraint, self.max_triple_constraint = min, max

    def set_random_seed(self, rand_seed):
        import random
        self.train_dataloader.lib.setRandomSeed(rand_seed)
        self.train_dataloader.lib.randReset()
        random.seed(rand_seed)
        torch.manual_seed(rand_seed)

    def set_valid_dataloader(self, valid_dataloader):
        self.valid_dataloader = valid_dataloader

    def set_test_dataloader(self, test_dataloader):
        self.data_loader = test_dataloader

    def embedding_model_factory(self, ent_tot, rel_tot, margin):
        return NegativeSampling(model=self.embedding_model(ent_tot, rel_tot, **self.embedding_model_param),
                                loss=MarginLoss(margin=margin), batch_size=self.train_dataloader.batch_size)

    def reset_valid_variables(self):
        self.early_stopping_patience = self.early_stopping_patience_const
        self.best_state = {}
        self.best_hit10 = 0
        self.bad_counts = 0

    def gather_embedding_spaces(self, entity_1, rel, entity_2=None):
        ids = self.entity_universes[entity_1].intersection(self.relation_universes[rel])
        if entity_2 is not None:
            ids = ids.intersection(self.entity_universes[entity_2])
        return ids

    def calculate_unembedded_ratio(self, mode="examine_entities"):
        mapping = self.entity_universes if mode == "examine_entities" else self.relation_universes
        total = self.train_dataloader.entTotal if mode == "examine_entities" else self.train_dataloader.relTotal
        return sum(1 for i in range(total) if len(mapping[i]) == 0) / total

    # ------------------------------------------------------------------ hyper-parameter draws
    def draw_universe_hyper(self, seed):
        """The reference's Python draws for one universe, in its order (reference :211-212,232,237-240),
        from a generator seeded like ``random.seed(seed)``."""
        rnd = Random(seed)
        tc = rnd.randrange(self.min_triple_constraint, self.max_triple_constraint)
        balance = round(rnd.uniform(self.min_balance, self.max_balance), 2)
        margin = rnd.randrange(self.min_margin, self.max_margin)
        epochs = self.const_num_epochs if self.const_num_epochs is not None else \
            rnd.randrange(self.min_num_epochs, self.max_num_epochs)
        lr = round(rnd.uniform(self.min_lr, self.max_lr), len(str(self.min_lr).split(".")[1]))
        return dict(tc=tc, balance=balance, margin=margin, epochs=epochs, lr=lr)

    def draw_hyper_batch(self, seeds):
        """draw_universe_hyper for many universes in one native call (CPython's generator restated in
        csrc/pyrandom_host.cpp: the interpreter needs 8-17 us per universe on the launching thread); ranges the
        native generator does not cover go through random.Random."""
        n = len(seeds)
        try:
            ints = [int(v) for v in (self.min_triple_constraint, self.max_triple_constraint, self.min_margin, self.max_margin)]
            exact = all(a == b for a, b in zip(ints, (self.min_triple_constraint, self.max_triple_constraint, self.min_margin, self.max_margin)))
            const = self.const_num_epochs is not None
            ep = (int(self.const_num_epochs), int(self.const_num_epochs)) if const else (int(self.min_num_epochs), int(self.max_num_epochs))
            digits = len(str(self.min_lr).split(".")[1])
            s64 = np.ascontiguousarray(seeds, dtype=np.int64)
            tc, margin, epochs = (np.zeros(n, dtype=np.int64) for _ in range(3))
            bal, lr = np.zeros(n, dtype=np.float64), np.zeros(n, dtype=np.float64)
            rc = -1 if not exact else self.lib.pk_python_hyper_draws(
                n, N.addr(s64), ints[0], ints[1], float(self.min_balance), float(self.max_balance), ints[2], ints[3], ep[0], ep[1],
                0 if const else 1, float(self.min_lr), float(self.max_lr), digits, N.addr(tc), N.addr(bal), N.addr(margin),
                N.addr(epochs), N.addr(lr))
        except (TypeError, ValueError, IndexError, OverflowError):
            rc = -1
        if rc != 0:
            return [self.draw_universe_hyper(int(s_)) for s_ in seeds]
        ep_out = [self.const_num_epochs] * n if const else epochs.tolist()
        return [dict(tc=a, balance=b, margin=c, epochs=d, lr=e)
                for a, b, c, d, e in zip(tc.tolist(), bal.tolist(), margin.tolist(), ep_out, lr.tolist())]

    # ------------------------------------------------------------------ training
    def _device(self):
        if not self.use_gpu:
            raise N.NativeError("PuTransE on the B200 path needs a CUDA device: there is no CPU implementation")
        N.require_cuda()
        return torch.device("cuda", torch.cuda.current_device())

    class _Slot(object):
        """One in-flight launch: its stream, its per-launch device buffers and the event that ends it."""

        def __init__(self, dev):
            self.stream = torch.cuda.Stream(device=dev)
            self.done = torch.cuda.Event(blocking=True)   # waiting on it sleeps (the cores go to the sampler threads)
            self.busy = False
            self.scratch, self.state, self.pinned = {}, {}, {}
            self.chunk = None
            self.host_loss = None

    def _train_chunk(self, universe_ids, prefetch_ids=None):
        """Train `universe_ids` with ONE launch on the next launch slot's stream and return without waiting
        for it (see launch_slots)."""
        dev = self._device()
        if self._slots is None:
            self._slots = [Parallel_Universe_Config._Slot(dev) for _ in range(max(1, int(self.launch_slots)))]
        slot = self._slots[self._launch_index % len(self._slots)]
        self._launch_index += 1
        self._retire(slot)                              # its previous launch must have left its buffers
        slot.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(slot.stream):
            ck = self._train_piece(universe_ids, slot, prefetch_ids=prefetch_ids)
            if ck.d_loss is not None:                   # per-step losses follow the launch into pinned memory
                if slot.host_loss is None or slot.host_loss.numel() < ck.d_loss.numel():
                    slot.host_loss = torch.empty(int(ck.d_loss.numel() * 1.5) + 1024, dtype=torch.float32, pin_memory=True)
                slot.host_loss[:ck.d_loss.numel()].copy_(ck.d_loss, non_blocking=True)
            slot.done.record(slot.stream)
        slot.busy, slot.chunk = True, ck
        return ck

    def _retire(self, slot):
        """Wait for the slot's launch (if any) and take its host-side results."""
        if not slot.busy:
            return
        t0 = time.perf_counter()
        slot.done.synchronize()
        self.timings["wait_for_launch"] += time.perf_counter() - t0
        ck, slot.busy, slot.chunk = slot.chunk, False, None
        if ck.d_loss is not None:
            host = slot.host_loss[:ck.d_loss.numel()].numpy().copy()   # one copy out of the pinned buffer; universes get views
            self.d2h_bytes += host.nbytes
            o = 0
            for u in ck.ids:
                steps = self.universe_hyper[u]["epochs"] * self.universe_hyper[u]["nbatches"]
                self._universe_losses[u] = host[o:o + steps]
                o += steps
            ck.d_loss = None
        ck.train_inputs = None
        if getattr(ck, "walk", None) is not None:
            ck.walk.release()
            ck.walk = None

    @property
    def universe_losses(self):
        self.synchronize()
        return self._universe_losses

    def _visible_chunks(self):
        return self._chunks if self._view is None else self._chunks[:self._view[0]]

    def _visible_next_id(self):
        return self.next_universe_id if self._view is None else self._view[1]

    def synchronize(self):
        """Wait for every launch in flight (training is asynchronous between calls; everything that reads
        trained tables goes through here first).  Inside a deferred validation only the launches of the chunks
        it looks at are waited for."""
        if self._slots:
            seen = None if self._view is None else {id(ck) for ck in self._visible_chunks()}
            slots = [s_ for s_ in self._slots if seen is None or s_.chunk is None or id(s_.chunk) in seen]
            for slot in slots:
                self._retire(slot)
            dev = self._device()
            for slot in slots:   # later work on the caller's stream is ordered after the launches
                torch.cuda.current_stream(dev).wait_stream(slot.stream)

    def _sampling_key(self, universe_ids):
        dl = self.train_dataloader
        return (tuple(universe_ids), self.initial_random_seed, self.min_triple_constraint, self.max_triple_constraint,
                self.min_balance, self.max_balance, self.min_margin, self.max_margin, self.min_lr, self.max_lr,
                self.min_num_epochs, self.max_num_epochs, self.const_num_epochs, dl.work_threads, bool(dl.filter),
                self.lib.pk_import_count(), self.lib.getTrainTotal(), getattr(dl, "tripleTotal", 0))

    def _sample_universes(self, universe_ids, threads=None):
        """Hyper-parameter draws + subgraphs of a set of universes (host only; pk_universes_build is
        re-entrant and the ctypes call releases the GIL, so this also runs on a worker thread)."""
        lib, dl = self.lib, self.train_dataloader
        background = threads is not None
        t_begin = time.perf_counter()
        threads = int(self.sampler_threads) if threads is None else int(threads)
        n = len(universe_ids)
        seeds = np.array([self.initial_random_seed + u for u in universe_ids], dtype=np.int64)
        hyper = self.draw_hyper_batch(seeds)
        tcs = np.array([h["tc"] for h in hyper], dtype=np.int64)
        bals = np.array([h["balance"] for h in hyper], dtype=np.float32)
        # -- subgraphs: bit-identical to the reference's getParallelUniverse, on host threads
        lean = not dl.filter and not dl.bern     # then training reads only the (h,r,t) list and the remaps
        build = lib.pk_universes_build_lean if lean else lib.pk_universes_build
        handle = build(n, N.addr(seeds), N.addr(tcs), N.addr(bals), threads)
        if not handle:
            raise N.NativeError("pk_universes_build: %s" % N.last_error())
        try:
            nT, nE, nR, focus = (np.zeros(n, dtype=np.int64) for _ in range(4))
            N.check(lib.pk_universes_sizes(handle, N.addr(nT), N.addr(nE), N.addr(nR), N.addr(focus)), "pk_universes_sizes")
            sT, sE, sR = int(nT.sum()), int(nE.sum()), int(nR.sum())
            W = dl.work_threads
            by_head = np.zeros((sT, 3), dtype=np.int32)
            by_tail = np.zeros((sT, 3), dtype=np.int32) if dl.filter else None
            ent_remap, rel_remap = np.zeros(sE, dtype=np.int32), np.zeros(sR, dtype=np.int32)
            lm, rm = np.zeros(sR, dtype=np.float32), np.zeros(sR, dtype=np.float32)
            lcg = np.zeros((n, W), dtype=np.uint64)
            N.check(lib.pk_universes_export(handle, N.addr(by_head), N.addr(by_tail) if by_tail is not None else None, None,
                                            N.addr(ent_remap), N.addr(rel_remap), None if lean else N.addr(lm),
                                            None if lean else N.addr(rm), N.addr(lcg)),
                    "pk_universes_export")
        finally:
            lib.pk_universes_free(handle)
        if background:   # runs beside the GPU: reported separately from the launching thread's time
            self.timings["universe_sampling_background"] += time.perf_counter() - t_begin
        return dict(hyper=hyper, seeds=seeds, nT=nT, nE=nE, nR=nR, focus=focus, by_head=by_head, by_tail=by_tail,
                    ent_remap=ent_remap, rel_remap=rel_remap, lm=lm, rm=rm, lcg=lcg)

    def _train_piece(self, universe_ids, slot, prefetch_ids=None):
        lib, dl = self.lib, self.train_dataloader
        dev = self._device()
        n = len(universe_ids)
        t0 = time.perf_counter()
        lib.setWorkThreads(dl.work_threads)
        # subgraphs of these universes may already have been sampled beside the previous launch
        key = self._sampling_key(universe_ids)
        entry = self._prefetched.pop(key, None)
        for k_old in [k_ for k_ in self._prefetched if k_[0][0] <= universe_ids[0] or k_[1:] != key[1:]]:
            self._drop_sample(self._prefetched.pop(k_old))   # samples nobody will ask for any more (passed, or of another graph/config)
        if entry is None:
            entry = self._submit_sample(universe_ids, background=False)
        smp = self._resolve_sample(entry, universe_ids)
        walk = smp.get("walk")
        hyper, seeds, nT, nE, nR, focus = smp["hyper"], smp["seeds"], smp["nT"], smp["nE"], smp["nR"], smp["focus"]
        by_head, by_tail, ent_remap, rel_remap, lm, rm, lcg = (smp[k_] for k_ in ("by_head", "by_tail", "ent_remap", "rel_remap",
                                                                                   "lm", "rm", "lcg"))
        sT, sE, sR = int(nT.sum()), int(nE.sum()), int(nR.sum())
        W = dl.work_threads
        t1 = time.perf_counter()
        self.timings["universe_sampling"] += t1 - t0
        toff = np.concatenate([[0], np.cumsum(nT)]).astype(np.int64)
        eoff = np.concatenate([[0], np.cumsum(nE)]).astype(np.int64)
        roff = np.concatenate([[0], np.cumsum(nR)]).astype(np.int64)

        # -- initial tables: the reference's torch CPU initialisation (same generator, same draws,
        #    same order: torch.manual_seed(seed + u) then the model constructor's init), written
        #    straight into one packed pinned buffer per table
        model_cls, param = self.embedding_model, self.embedding_model_param
        specs0 = model_cls.table_specs(2, 1, **param)
        ent_names = set(model_cls._ent_tables)
        offs = {attr: (eoff if attr in ent_names else roff) for attr, _, _ in specs0}
        fused = self._native_init_mode()
        native_ok = (fused is not None and param.get("margin") is None
                     and min(int(nR.min()), int(nE.min())) * min(s_[2] for s_ in specs0) >= 16)
        tables = None
        ta = time.perf_counter()
        self.timings["init_mode_checks"] += ta - t1
        if native_ok and self._device_init_ok(fused):
            # the torch generator is replayed on the GPU: no host RNG time, no H2D copy of the tables
            tables = {attr: self._arena_rows(dev, sE if attr in ent_names else sR, dim) for attr, _, dim in specs0}
            self.timings["init_alloc"] += time.perf_counter() - ta
            self._native_init(model_cls, param, seeds, nE, nR, tables, offs, fused, device_stream=slot.stream.cuda_stream)
        else:
            packed_host = {attr: self._pinned_rows(attr, sE if attr in ent_names else sR, dim) for attr, _, dim in specs0}
            if native_ok:
                # host threads, bit-identical to torch's generator (verified once per process in _native_init_mode)
                self._native_init(model_cls, param, seeds, nE, nR, packed_host, offs, fused)
            else:
                for i, u in enumerate(universe_ids):
                    torch.manual_seed(int(seeds[i]))
                    views = {attr: packed_host[attr][offs[attr][i]:offs[attr][i + 1]] for attr in packed_host}
                    model_cls.initial_tables_into(int(nE[i]), int(nR[i]), views, **param)
        proto = self._proto()
        t2 = time.perf_counter()
        self.timings["table_init"] += t2 - t1

        ck = _Chunk()
        ck.items_cache = {}
        ck.ids = list(universe_ids)
        ck.nT, ck.nE, ck.nR, ck.eoff, ck.roff, ck.toff = nT, nE, nR, eoff, roff, toff
        ck.ent_remap, ck.rel_remap = ent_remap, rel_remap
        ck.proto = proto
        if tables is None:
            tables = {name: self._arena_rows(dev, t.shape[0], t.shape[1]).copy_(t, non_blocking=True) for name, t in packed_host.items()}
            self.h2d_bytes += sum(t.numel() * 4 for t in packed_host.values())
            self._pinned_busy = torch.cuda.Event()
            self._pinned_busy.record(slot.stream)
        ck.tables = tables
        adagrad = True  # reference :241-242 hard-codes opt_method='Adagrad' for universes
        ck.state = {name: self._state_rows(slot, name, t) for name, t in ck.tables.items()} if adagrad else None
        if walk is not None:      # the walk wrote the local triple lists where the training kernel reads them
            slot.stream.wait_event(walk.event)
            # from the table slab (a fresh allocation per chunk costs a cudaMalloc every few chunks, with launches in flight)
            ck.d_ent_remap = self._arena_rows(dev, max(sE, 1), 1).view(torch.int32).view(-1)[:sE]
            ck.d_rel_remap = self._arena_rows(dev, max(sR, 1), 1).view(torch.int32).view(-1)[:sR]
            N.check(lib.pk_walk_pack_remaps(n, walk.bufs["sizes"].data_ptr(), walk.bufs["ent_remap"].data_ptr(),
                                            walk.bufs["rel_remap"].data_ptr(), ck.d_ent_remap.data_ptr(), ck.d_rel_remap.data_ptr(),
                                            slot.stream.cuda_stream), "pk_walk_pack_remaps")
            d_by_head = walk.bufs["tri"].view(-1, 3)
            tri_off = np.arange(n, dtype=np.int64) * self._walker.cap
        else:
            d_by_head = self._dev_scratch(slot, "by_head", dev, by_head)
            tri_off = toff[:n]
        d_by_tail = self._dev_scratch(slot, "by_tail", dev, by_tail) if by_tail is not None else None
        d_lm = self._dev_scratch(slot, "lm", dev, lm) if dl.bern else None
        d_rm = self._dev_scratch(slot, "rm", dev, rm) if dl.bern else None

        nb = dl.nbatches
        Bs = np.asarray(nT, dtype=np.int64) // nb
        if Bs.min() < 1:
            i = int(np.argmin(Bs))
            raise N.NativeError("universe %d has %d triples: fewer than nbatches=%d" % (universe_ids[i], nT[i], nb))
        epochs = np.array([h["epochs"] for h in hyper], dtype=np.int64)
        steps = epochs * nb
        # the descriptor array column by column (a Python loop over ctypes fields costs ~1 ms per 100 universes,
        # all of it with the GPU idle)
        darr = np.zeros(n, dtype=N.UNIVERSE_DESC_DTYPE)
        darr["tri_off"], darr["ent_off"], darr["rel_off"] = tri_off, eoff[:n], roff[:n]
        darr["n_tri"], darr["n_ent"], darr["n_rel"] = nT, nE, nR
        darr["batch_size"], darr["nbatches"], darr["epochs"] = Bs, nb, epochs
        darr["margin"] = np.array([h["margin"] for h in hyper], dtype=np.float32)
        darr["lr"] = np.array([h["lr"] for h in hyper], dtype=np.float32)
        darr["loss_off"] = (np.cumsum(steps) - steps) if self.record_losses else -1
        darr["lcg"][:, :min(W, 8)] = lcg[:, :min(W, 8)]
        desc = (N.UniverseDesc * n).from_buffer(darr)
        loss_total = int(steps.sum())
        self.positive_triples += int((steps * Bs).sum())
        for i in range(n):
            h = hyper[i]
            h.update(nT=int(nT[i]), nE=int(nE[i]), nR=int(nR[i]), focus=int(focus[i]), batch_size=int(Bs[i]), nbatches=nb)
            self.universe_hyper[universe_ids[i]] = h
        d_loss = self._dev_scratch(slot, "loss", dev, None, numel=max(loss_total, 1), dtype=torch.float32) if self.record_losses else None

        cfg = proto.native_cfg(opt=N.PK_ADAGRAD, neg_ent=dl.negative_ent, bern=1 if dl.bern else 0,
                               filt=1 if dl.filter else 0, work_threads=W)
        tab = self._packed_tables(ck, with_state=True)
        st = slot.stream.cuda_stream
        # universes whose relation tables / batch scratch do not fit the universe kernel (relation-rich
        # graphs such as FB15K) are trained one by one with the single-space kernels on their slice
        # (the class is monotone in every size: if the element-wise largest shape fits, all of them do)
        worst = lib.pk_universe_kernel_class(ctypes.byref(cfg), int(np.max(nE)), int(np.max(nR)), int(Bs.max()))
        if worst in (0, 1):
            klass = [worst] * n
        else:
            klass = [lib.pk_universe_kernel_class(ctypes.byref(cfg), int(nE[i]), int(nR[i]), int(Bs[i])) for i in range(n)]
        if min(klass) < 0:
            raise N.NativeError("pk_universe_kernel_class: %s" % N.last_error())
        big = [i for i in range(n) if klass[i] == 2]
        if big:
            small = [i for i in range(n) if klass[i] != 2]
            desc_k2 = (N.UniverseDesc * max(len(small), 1))()
            for j, i in enumerate(small):
                ctypes.memmove(ctypes.byref(desc_k2[j]), ctypes.byref(desc[i]), ctypes.sizeof(N.UniverseDesc))
        else:
            small, desc_k2 = list(range(n)), desc
        if small:
            N.check(lib.pk_train_universes(ctypes.byref(cfg), ctypes.byref(tab), d_by_head.data_ptr(),
                                           d_by_tail.data_ptr() if d_by_tail is not None else None,
                                           d_lm.data_ptr() if d_lm is not None else None,
                                           d_rm.data_ptr() if d_rm is not None else None,
                                           desc_k2, len(small), d_loss.data_ptr() if d_loss is not None else None, st),
                    "pk_train_universes")
            self.gpu_launches += lib.pk_last_launch_count()
        for i in big:
            self._train_single_space(ck, cfg, i, desc[i], d_by_head, d_by_tail, d_lm, d_rm, lcg[i], d_loss, st, dev)
        self.universes_on_single_space_path += len(big)
        self._queue_sampling(universe_ids, prefetch_ids)
        ck.train_inputs = (d_by_head, d_by_tail, d_lm, d_rm)  # keep alive until the stream is done
        ck.walk = walk
        ck.tri_off = np.asarray(tri_off, dtype=np.int64)   # row offsets of the universes' triple lists inside train_inputs[0]
        if walk is not None:   # walk inputs (40 B per universe) in, sizes + remaps out; the triple index never leaves the device
            self.h2d_bytes += n * 40 + ctypes.sizeof(desc) + n * (8 + len(ck.tables) * 20)
            self.d2h_bytes += n * 32
            self.gpu_launches += 3
        else:
            self.h2d_bytes += sum(t.numel() * t.element_size() for t in ck.train_inputs if t is not None) + ctypes.sizeof(desc) \
                + n * (8 + len(ck.tables) * 20)   # + seeds / rows / offsets / bounds of the device initialiser
        ck.d_loss = d_loss
        t3 = time.perf_counter()
        self.timings["launch"] += t3 - t2

        # -- the reference's per-universe containers are views of this chunk, built on first use
        for i, u in enumerate(universe_ids):
            self._where[u] = (ck, i)
        self._maps_version += 1
        self._chunks.append(ck)
        self._rank_cache.clear()
        return ck

    def _use_device_walk(self):
        dl = self.train_dataloader
        if not self.device_walk or dl.filter or dl.bern:
            return False
        if self._walker is None:
            from ..universe_walk import DeviceWalker
            self._walker = DeviceWalker(self.lib, self._device())
        cap = self._walker.cap
        tc_max = int(self.max_triple_constraint) - 1
        if tc_max > cap or int(np.float32(self.max_balance) * np.float32(tc_max)) > cap:
            return False
        return self._walker.usable()

    def _submit_sample(self, universe_ids, background):
        """Start sampling the subgraphs of `universe_ids`: a walk launch on the GPU, or the host builder (on the sampler
        thread when `background`)."""
        self.lib.setWorkThreads(self.train_dataloader.work_threads)
        if self._use_device_walk():
            t0 = time.perf_counter()
            seeds = np.array([self.initial_random_seed + u for u in universe_ids], dtype=np.int64)
            hyper = self.draw_hyper_batch(seeds)
            tcs = np.array([h["tc"] for h in hyper], dtype=np.int64)
            bals = np.array([h["balance"] for h in hyper], dtype=np.float32)
            t1 = time.perf_counter()
            res = self._walker.submit(seeds, tcs, bals, self.train_dataloader.work_threads, copy_remaps=False)
            self.timings["hyper_draws"] += t1 - t0
            self.timings["walk_submit"] += time.perf_counter() - t1
            return ("device", hyper, res)
        if not background:
            return ("host", self._sample_universes(universe_ids))
        if self._pool is None:
            from concurrent.futures import ThreadPoolExecutor
            self._pool = ThreadPoolExecutor(max_workers=1)
        # this rank's share of the host's cores, minus one for the launching thread and the CUDA driver's threads
        _, _, world_ = _dist()
        bg_threads = int(self.sampler_threads) or max(2, _host_cores() // max(world_, 1) - 1)
        return ("host", self._pool.submit(self._sample_universes, universe_ids, bg_threads))

    def _resolve_sample(self, entry, universe_ids):
        if entry[0] == "host":
            return entry[1] if isinstance(entry[1], dict) else entry[1].result()
        _, hyper, res = entry
        t0 = time.perf_counter()
        sizes = res.wait()
        self.timings["wait_for_walk"] += time.perf_counter() - t0
        if not res.ok:      # beyond the kernel (capacity, isolated entity, empty walk): the host builder decides
            res.release()
            return self._sample_universes(universe_ids)
        return dict(hyper=hyper, seeds=res.seeds, nT=sizes[:, 0].copy(), nE=sizes[:, 1].copy(), nR=sizes[:, 2].copy(),
                    focus=sizes[:, 3].copy(), by_head=None, by_tail=None, ent_remap=None, rel_remap=None, lm=None, rm=None,
                    lcg=res.lcg, walk=res)

    def _drop_sample(self, entry):
        if entry[0] == "host":
            if not isinstance(entry[1], dict):
                entry[1].cancel()
        else:
            entry[2].event.synchronize()
            entry[2].release()

    def _queue_sampling(self, universe_ids, prefetch_ids):
        """Queue the subgraphs of the chunks after this one.  Called AFTER this chunk's launch has been issued.  Host
        builder: its Python parts (hyper-parameter draws, array allocation) hold the GIL, and the launching thread must
        not compete for it while the GPU waits for its launch (measured: 1.3 ms -> 13 ms of launch time per chunk when
        the sampler ran beside the launch), so it works one chunk ahead.  Device walk: launches on the walk stream."""
        if not (prefetch_ids and self.prefetch_sampling):
            return
        step = prefetch_ids[0] - universe_ids[0]
        depth = int(self.device_walk_depth) if self._use_device_walk() else int(self.prefetch_depth)
        for ahead in range(1, max(1, depth) + 1):
            ids_next = [u + ahead * step for u in universe_ids]
            key_next = self._sampling_key(ids_next)
            if key_next not in self._prefetched:
                self._prefetched[key_next] = self._submit_sample(ids_next, background=True)

    def _arena_rows(self, dev, rows, dim):
        """[rows, dim] fp32 device view carved from a slab.  The trained tables of every chunk stay
        alive for evaluation, so they are sub-allocated from large slabs instead of one cudaMalloc per
        table per chunk (which costs up to tens of milliseconds when the allocator has to grow)."""
        need = rows * dim
        need_al = (need + 63) // 64 * 64
        if self._arena is None or self._arena_used + need_al > self._arena.numel():
            self._arena = torch.empty(max(need_al, int(self.arena_slab_floats)), dtype=torch.float32, device=dev)
            self._arena_used = 0
        view = self._arena[self._arena_used:self._arena_used + need].view(rows, dim)
        self._arena_used += need_al
        return view

    def _dev_scratch(self, slot, name, dev, host=None, numel=None, dtype=None):
        """Per-launch device inputs/outputs (triple index, means, per-step losses) live in grow-only
        buffers owned by the launch slot (launches of different slots are in flight together and must not
        share them; a slot is reused only after its previous launch has finished): asking the allocator for
        slightly different sizes every chunk costs a cudaMalloc now and then, which stalls the launch by
        tens of ms."""
        if host is not None:
            t = torch.from_numpy(np.ascontiguousarray(host))
            numel, dtype = t.numel(), t.dtype
        buf = slot.scratch.get(name)
        if buf is None or buf.numel() < numel or buf.dtype != dtype or buf.device != dev:
            buf = torch.empty(int(numel * 1.5) + 1024, dtype=dtype, device=dev)
            slot.scratch[name] = buf
        view = buf[:numel]
        if host is not None:
            # through the slot's own pinned staging buffer: a copy from pageable memory blocks the launching thread
            # until the device has taken it, behind whatever the other slots have in flight
            pin = slot.pinned.get(name)
            if pin is None or pin.numel() < numel or pin.dtype != dtype:
                pin = torch.empty(int(numel * 1.5) + 1024, dtype=dtype, pin_memory=True)
                slot.pinned[name] = pin
            np.copyto(pin[:numel].numpy(), t.reshape(-1).numpy())   # one memcpy on this thread (torch's CPU copy_ wakes the intra-op pool)
            view = view.view(t.shape)
            view.copy_(pin[:numel].view(t.shape), non_blocking=True)
        else:
            view.zero_()
        return view

    def _state_rows(self, slot, name, like):
        """Zeroed optimizer state for one launch: a grow-only scratch per table and launch slot
        (universes are trained once; their Adagrad sums are not needed afterwards)."""
        buf = slot.state.get(name)
        if buf is None or buf.numel() < like.numel() or buf.device != like.device:
            buf = torch.empty(int(like.numel() * 1.25) + 1024, dtype=torch.float32, device=like.device)
            slot.state[name] = buf
        view = buf[:like.numel()].view_as(like)
        view.zero_()
        return view

    def _pinned_rows(self, name, rows, dim):
        """[rows, dim] view of a grow-only pinned staging buffer per table (cudaHostAlloc costs
        milliseconds, so it is not repeated per chunk).  The previous chunk's H2D copy must have
        finished before the buffer is overwritten."""
        if self._pinned_busy is not None:
            self._pinned_busy.synchronize()
            self._pinned_busy = None
        buf = self._pinned.get(name)
        need = rows * dim
        if buf is None or buf.numel() < need:
            buf = torch.empty(int(need * 1.25) + 1024, dtype=torch.float32, pin_memory=True)
            self._pinned[name] = buf
        return buf[:need].view(rows, dim)

    def _native_init(self, model_cls, param, seeds, nE, nR, packed_host, offs, fused, device_stream=None):
        import math
        specs0 = model_cls.table_specs(2, 1, **param)
        ent_names = set(model_cls._ent_tables)
        n, T = len(seeds), len(specs0)
        rows = np.stack([(nE if attr in ent_names else nR) for attr, _, _ in specs0], axis=1).astype(np.int64)
        dims = np.array([dim for _, _, dim in specs0], dtype=np.int32)
        row_off = np.stack([offs[attr][:-1] for attr, _, _ in specs0], axis=1).astype(np.int64)
        # nn.init.xavier_uniform_: a = sqrt(3) * gain * sqrt(2 / (fan_in + fan_out)), in Python doubles
        bounds = np.array([[math.sqrt(3.0) * (1.0 * math.sqrt(2.0 / float(int(rows[i, t]) + int(dims[t])))) for t in range(T)]
                           for i in range(n)], dtype=np.float64)
        ptrs = (ctypes.c_void_p * T)(*[packed_host[attr].data_ptr() for attr, _, _ in specs0])
        seeds64, rows, row_off = np.ascontiguousarray(seeds, dtype=np.int64), np.ascontiguousarray(rows), np.ascontiguousarray(row_off)
        if device_stream is not None:
            N.check(self.lib.pk_init_tables_device(n, N.addr(seeds64), T, N.addr(rows), N.addr(dims), ptrs, N.addr(row_off),
                                                   N.addr(bounds), int(fused), device_stream), "pk_init_tables_device")
            self.gpu_launches += self.lib.pk_last_launch_count()
        else:
            N.check(self.lib.pk_torch_init_tables(n, N.addr(seeds64), T, N.addr(rows), N.addr(dims), ptrs, N.addr(row_off),
                                                  N.addr(bounds), int(fused), int(self.sampler_threads)), "pk_torch_init_tables")

    _INIT_MODE = {}
    _DEVICE_INIT = {}

    def _device_init_ok(self, fused):
        """Whether pk_init_tables_device reproduces pk_torch_init_tables bit-for-bit for this model on
        this GPU (checked once per process on three small spaces that straddle the generator's
        624-word refill and the 16-element tail rule)."""
        key = (self.embedding_model, tuple(sorted((k, str(v)) for k, v in self.embedding_model_param.items())), int(fused))
        if key not in Parallel_Universe_Config._DEVICE_INIT:
            ok = False
            try:
                model_cls, param = self.embedding_model, self.embedding_model_param
                specs0 = model_cls.table_specs(2, 1, **param)
                ent_names = set(model_cls._ent_tables)
                nE, nR, seeds = np.array([37, 64, 700]), np.array([5, 16, 9]), np.array([12345, (1 << 33) + 7, 99])
                eo, ro = np.concatenate([[0], np.cumsum(nE)]), np.concatenate([[0], np.cumsum(nR)])
                host = {a: torch.empty((int(eo[-1]) if a in ent_names else int(ro[-1]), dim)) for a, _, dim in specs0}
                offs = {a: (eo if a in ent_names else ro) for a in host}
                self._native_init(model_cls, param, seeds, nE, nR, host, offs, fused)
                dev = self._device()
                devt = {a: torch.full_like(t, float("nan"), device=dev) for a, t in host.items()}
                self._native_init(model_cls, param, seeds, nE, nR, devt, offs, fused,
                                  device_stream=torch.cuda.current_stream(dev).cuda_stream)
                ok = all(torch.equal(devt[a].cpu(), host[a]) for a in host)
            except Exception:
                ok = False
            Parallel_Universe_Config._DEVICE_INIT[key] = ok
        return Parallel_Universe_Config._DEVICE_INIT[key]

    def _native_init_mode(self):
        """Which rounding variant of pk_torch_init_tables reproduces THIS torch build's CPU generator
        bit-for-bit for this model (1 fused, 0 unfused), or None if neither does (then torch itself
        initialises every universe).  Checked once per process on two small spaces."""
        key = (self.embedding_model, tuple(sorted((k, str(v)) for k, v in self.embedding_model_param.items())))
        if key not in Parallel_Universe_Config._INIT_MODE:
            model_cls, param = self.embedding_model, self.embedding_model_param
            mode = None
            try:
                specs0 = model_cls.table_specs(2, 1, **param)
                ent_names = set(model_cls._ent_tables)
                nE, nR, seeds = np.array([37, 64]), np.array([5, 16]), np.array([12345, (1 << 33) + 7])
                want = []
                with torch.random.fork_rng(devices=[]):
                    for i in range(2):
                        torch.manual_seed(int(seeds[i]))
                        m = model_cls(int(nE[i]), int(nR[i]), **param)
                        want.append({a: getattr(m, a).weight.data.clone() for a, _, _ in specs0})
                eo, ro = np.array([0, 37, 101]), np.array([0, 5, 21])
                for fused in (1, 0):
                    host = {a: torch.empty((101 if a in ent_names else 21, dim)) for a, _, dim in specs0}
                    offs = {a: (eo if a in ent_names else ro) for a in host}
                    self._native_init(model_cls, param, seeds, nE, nR, host, offs, fused)
                    if all(torch.equal(host[a][offs[a][i]:offs[a][i + 1]], want[i][a]) for i in range(2) for a in host):
                        mode = fused
                        break
            except Exception:
                mode = None
            Parallel_Universe_Config._INIT_MODE[key] = mode
        return Parallel_Universe_Config._INIT_MODE[key]

    def _train_single_space(self, ck, cfg, i, dd, d_by_head, d_by_tail, d_lm, d_rm, lcg_i, d_loss, st, dev):
        """One universe through K1 (pk_train_steps) on its row range of the packed tables: the path for
        universes that do not fit the batched-universe kernel.  Same sampler streams, same arithmetic."""
        lib = self.lib
        e0, r0, t0 = int(dd.ent_off), int(dd.rel_off), int(dd.tri_off)
        nE, nR, nT, B = int(dd.n_ent), int(dd.n_rel), int(dd.n_tri), int(dd.batch_size)
        d = ck.proto.dim_native
        t = N.Tables()
        for j in range(2):
            t.ent[j] = t.rel[j] = t.ent_state[j] = t.rel_state[j] = None
        for j, name in enumerate(ck.proto._ent_tables):
            t.ent[j] = ck.tables[name].data_ptr() + e0 * d * 4
            t.ent_state[j] = ck.state[name].data_ptr() + e0 * d * 4
        for j, name in enumerate(ck.proto._rel_tables):
            t.rel[j] = ck.tables[name].data_ptr() + r0 * d * 4
            t.rel_state[j] = ck.state[name].data_ptr() + r0 * d * 4
        t.n_ent, t.n_rel = nE, nR
        d_lcg = torch.from_numpy(np.ascontiguousarray(lcg_i).view(np.int64).copy()).to(dev)
        smp = N.Sampler(by_head=d_by_head.data_ptr() + t0 * 12, by_tail=(d_by_tail.data_ptr() + t0 * 12) if d_by_tail is not None else None,
                        left_mean=(d_lm.data_ptr() + r0 * 4) if d_lm is not None else None,
                        right_mean=(d_rm.data_ptr() + r0 * 4) if d_rm is not None else None,
                        lcg=d_lcg.data_ptr(), n_tri=nT, n_ent=nE, n_rel=nR, head_off=None, tail_off=None)
        ws = lib.pk_workspace_create(ctypes.byref(cfg), nE, nR, B)
        if not ws:
            raise N.NativeError("pk_workspace_create: %s" % N.last_error())
        try:
            steps = int(dd.epochs) * int(dd.nbatches)
            loss = d_loss if d_loss is not None else torch.zeros(max(steps, 1), dtype=torch.float32, device=dev)
            off = int(dd.loss_off) if d_loss is not None else 0
            N.check(lib.pk_train_steps(ctypes.byref(cfg), ctypes.byref(t), ctypes.byref(smp), ws, B, steps, float(dd.margin), float(dd.lr),
                                       loss.data_ptr() + off * 4, st), "pk_train_steps")
            self.gpu_launches += lib.pk_last_launch_count()
            torch.cuda.current_stream(dev).synchronize() if st == torch.cuda.current_stream(dev).cuda_stream else torch.cuda.synchronize()
        finally:
            lib.pk_workspace_free(ws)

    def _proto(self):
        """A 2-entity instance of the embedding model: carries dim / p_norm / norm_flag / table names."""
        if getattr(self, "_proto_space", None) is None:
            with torch.random.fork_rng(devices=[]):
                self._proto_space = self.embedding_model(2, 1, **self.embedding_model_param)
        return self._proto_space

    def _packed_tables(self, ck, with_state=False):
        t = N.Tables()
        for i in range(2):
            t.ent[i] = t.rel[i] = t.ent_state[i] = t.rel_state[i] = None
        for i, name in enumerate(ck.proto._ent_tables):
            t.ent[i] = ck.tables[name].data_ptr()
            if with_state and ck.state is not None:
                t.ent_state[i] = ck.state[name].data_ptr()
        for i, name in enumerate(ck.proto._rel_tables):
            t.rel[i] = ck.tables[name].data_ptr()
            if with_state and ck.state is not None:
                t.rel_state[i] = ck.state[name].data_ptr()
        t.n_ent, t.n_rel = int(ck.eoff[-1]), int(ck.roff[-1])
        return t

    def _validate(self, view):
        """The validation branch of reference :330-356 on the ensemble `view` = (chunks, next_universe_id) describes.
        Returns True when early stopping ends the training."""
        self._view = view
        try:
            print("Universe %d has finished, validating..." % (view[1] - 1))
            self.eval_universes(eval_mode="valid")
            hit10 = self.valid()
            print("Current hit@10: {}".format(hit10))
            if hit10 > self.best_hit10:
                self.best_hit10 = hit10
                print("Best model | hit@10 of valid set is %f" % self.best_hit10)
                if self.checkpoint_dir:
                    self.save_model("Best_model_Pu{}_{}.ckpt".format(self.embedding_model.__name__, self.training_identifier))
                self.bad_counts = 0
            else:
                print("Hit@10 of valid set is %f | bad count is %d" % (hit10, self.bad_counts))
                self.bad_counts += 1
            if self.bad_counts == self.early_stopping_patience:
                print("Early stopping at universe {}".format(view[1] - 1))
                return True
            return False
        finally:
            self._view = None

    def train_parallel_universes(self, num_of_embedding_spaces):
        """Reference :316-367.  Universes are trained in chunks that end where the reference would
        validate (every ``valid_steps`` universes); with torch.distributed each rank trains its share.
        The chunks of one call are launched without waiting for each other; a validation that cannot end the
        training runs behind the next chunk's launch (pipeline_validation)."""
        dist, rank, world = _dist()
        t_call, t_valid = time.time(), 0.0
        done = 0
        pending = []            # validations deferred behind later launches: (chunks visible, next_universe_id), oldest first
        stop = False

        def flush(keep=0):
            nonlocal t_valid, stop
            while len(pending) > keep:
                t0 = time.time()
                stop = self._validate(pending.pop(0)) or stop
                t_valid += time.time() - t0

        while done < num_of_embedding_spaces and not stop:
            c = min(num_of_embedding_spaces - done, self.valid_steps - (done % self.valid_steps), self.max_chunk * world)
            ids = [self.next_universe_id + j for j in range(c)]
            mine = [u for u in ids if u % world == rank]
            if mine:
                self._train_chunk(mine, prefetch_ids=[u + c for u in mine])
            self.next_universe_id += c
            done += c
            # validations that fell due `validation_lag` chunks ago (none of them can stop the training)
            flush(max(0, min(int(self.validation_lag), int(self.launch_slots)) - 1))
            if done % self.valid_steps == 0:
                view = (len(self._chunks), self.next_universe_id)
                # even if every outstanding validation and this one fail to improve, patience is not exhausted
                harmless = self.bad_counts + len(pending) + 1 < self.early_stopping_patience
                if (self.pipeline_validation and harmless and done < num_of_embedding_spaces
                        and self.training_setting == "static"):
                    pending.append(view)
                else:
                    flush()
                    t0 = time.time()
                    stop = self._validate(view)
                    t_valid += time.time() - t0
            if self.save_steps and self.checkpoint_dir and (done // self.save_steps) > ((done - c) // self.save_steps):
                flush()         # the checkpoint carries best_hit10 / bad_counts: nothing may be outstanding
                if not stop:
                    self.save_model()
        flush()
        if not self.async_training:
            self.synchronize()
        training_duration = time.time() - t_call - t_valid
        self.training_duration += training_duration
        print("Time took for creation of embedding spaces: {:5.3f}s".format(training_duration))

    def add_embedding_space(self, embedding_space):
        """Reference :260-264.  Externally trained spaces are kept for inspection only: evaluation on
        this path reads the packed chunks that train_parallel_universes produces."""
        for p in embedding_space.parameters():
            p.requires_grad = False
        self.trained_embedding_spaces[self.next_universe_id] = embedding_space

    # ------------------------------------------------------------------ evaluation
    def _fold_keys(self, k_fixed, k_rel, k_side, consume, want_tuple=False, keys_token=None):
        """energy[key, e] = min over the universes that hold the key's fixed entity and relation of that universe's
        energy of candidate e (+inf where no universe speaks): reference eval_universes :556-603 +
        transmit_max_scores :446-465, for ALL keys, tile by tile.

        Keys (fixed entity, relation, side) must be sorted by nothing in particular; they are processed in tiles
        of `eval_tile_rows` rows, three tile buffers deep.  The host-side index work (items per chunk) is done once,
        vectorised, and uploaded once; per tile only kernels are launched.  With torch.distributed every rank folds
        ITS universes into the tile and the tile is min-all-reduced over NCCL asynchronously: the reduction of tile
        t runs beside the energy kernels of tile t+1.  `consume(ti, k0, k1, energy, tuple_score)` is called, in
        tile order, on the current stream after the tile's reduction has been ordered before it; `energy` is the
        [k1-k0, E] tile (with null-vector filling applied when that handling is configured)."""
        lib = self.lib
        dev = self._device()
        self.synchronize()
        dist, rank, world = _dist()
        sharded = dist is not None and world > 1
        t_host = time.perf_counter()
        if self.incremental_strategy == "deprecate":
            self.determine_deprecated_embedding_spaces()
        E = self.ent_tot
        K = int(k_fixed.shape[0])
        st = torch.cuda.current_stream(dev).cuda_stream
        rows_per_tile = max(1, min(K, int(self.eval_tile_rows), int(self.max_energy_bytes // (3 * 4 * E))))
        null_vector = self.missing_embedding_handling == "null_vector"
        tuples = null_vector or want_tuple
        # (key row, universe, local fixed entity, local relation, side) for every chunk, all keys at once; the
        # items come out sorted by key row, so a tile's items are a contiguous range
        # (a key set that comes back — the valid split every valid_steps universes, the test split — finds the items
        # of the chunks it has already seen cached on the chunk: `keys_token` names the key set)
        cache = None
        if (self.energy_cache and keys_token is not None and not tuples and self.incremental_strategy != "deprecate"
                and K * E * 4 <= int(self.energy_cache_bytes)):
            ckey = (keys_token, K, E, dev)
            cache = self._energy_cache.get(ckey)
            if cache is None:
                for old_key in [k_ for k_ in self._energy_cache if k_[0] == keys_token]:
                    del self._energy_cache[old_key]
                while len(self._energy_cache) >= 3:          # key sets that went away (snapshots of the incremental setting)
                    del self._energy_cache[next(iter(self._energy_cache))]
                energy_all = torch.empty((K, E), dtype=torch.float32, device=dev)
                N.check(lib.pk_fill_inf(energy_all.data_ptr(), energy_all.numel(), st), "pk_fill_inf")
                cache = self._energy_cache[ckey] = {"energy": energy_all, "folded": set()}
        per_chunk = []
        for ck in self._visible_chunks():
            if cache is not None and id(ck) in cache["folded"]:
                continue                   # its universes have spoken in this matrix already
            ix = self._chunk_index(ck)
            dep = self._deprecated_version if self.incremental_strategy == "deprecate" else 0
            cached = ck.items_cache.get((keys_token, rows_per_tile, dep)) if keys_token is not None else None
            if cached is None:
                items = self._energy_items(ix, k_fixed, k_rel, k_side)
                if dep and self.deprecated_embeddingspaces:   # reference :568-571: deprecated universes do not speak
                    gone = np.isin(np.asarray(ck.ids, dtype=np.int64)[items["universe"]],
                                   np.fromiter(self.deprecated_embeddingspaces, dtype=np.int64))
                    items = items[~gone]
                bounds = np.searchsorted(items["key_row"], np.arange(0, K + rows_per_tile, rows_per_tile))
                d_items = torch.from_numpy(items.view(np.int32).reshape(-1, 6)).to(dev) if items.shape[0] else None
                cached = (d_items, bounds)
                if keys_token is not None:
                    ck.items_cache[(keys_token, rows_per_tile, dep)] = cached
            d_items, bounds = cached
            if d_items is None:
                continue
            cfg_e, tab_e = self._eval_operands(ck, dev)
            per_chunk.append((ck, ix, d_items, bounds, cfg_e, tab_e))
        self.timings["eval_host_prep"] += time.perf_counter() - t_host
        n_tiles = (K + rows_per_tile - 1) // rows_per_tile
        if cache is not None:
            # resident matrix: fold the new chunks in, tile by tile (a tile is still the unit of the all-reduce and of the
            # rank counting), then hand every tile to the consumer
            for ti in range(n_tiles):
                k0 = ti * rows_per_tile
                k1 = min(K, k0 + rows_per_tile)
                energy = cache["energy"][k0:k1]
                for ck, ix, d_items, bounds, cfg, tab in per_chunk:
                    i0, i1 = int(bounds[ti]), int(bounds[ti + 1])
                    if i1 == i0:
                        continue
                    N.check(lib.pk_universe_energies(ctypes.byref(cfg), ctypes.byref(tab), ix["d_eoff"].data_ptr(),
                                                     ix["d_roff"].data_ptr(), ix["d_nE"].data_ptr(), ix["d_remap"].data_ptr(),
                                                     d_items.data_ptr() + i0 * 24, i1 - i0, energy.data_ptr() - k0 * E * 4, E, st),
                            "pk_universe_energies")
                    self.gpu_launches += lib.pk_last_launch_count()
                if sharded:    # in place: the minimum over the ranks is as good a starting point for later folds as the local one
                    dist.all_reduce(energy, op=dist.ReduceOp.MIN)
                consume(ti, k0, k1, energy, None)
            for ck in self._visible_chunks():
                cache["folded"].add(id(ck))
            cache["chunks"] = list(self._chunks)      # keeps the ids alive and unique
            return rows_per_tile
        bufs = [torch.empty((rows_per_tile, E), dtype=torch.float32, device=dev) for _ in range(min(3, n_tiles))]
        tbufs = [torch.empty(rows_per_tile, dtype=torch.float32, device=dev) for _ in bufs] if tuples else None

        def fold(ti):
            """energy kernels of tile ti (+ the asynchronous min-all-reduce when the universes are sharded)"""
            k0 = ti * rows_per_tile
            k1 = min(K, k0 + rows_per_tile)
            energy = bufs[ti % len(bufs)][:k1 - k0]
            N.check(lib.pk_fill_inf(energy.data_ptr(), energy.numel(), st), "pk_fill_inf")
            tuple_score = None
            if tuples:   # reference :494-514: per key, min over universes of the (zero vector, r, fixed) score
                tuple_score = tbufs[ti % len(bufs)][:k1 - k0]
                N.check(lib.pk_fill_inf(tuple_score.data_ptr(), tuple_score.numel(), st), "pk_fill_inf")
            for ck, ix, d_items, bounds, cfg, tab in per_chunk:
                i0, i1 = int(bounds[ti]), int(bounds[ti + 1])
                if i1 == i0:
                    continue
                # key rows in the items are global: hand the kernels the address row 0 WOULD have
                N.check(lib.pk_universe_energies(ctypes.byref(cfg), ctypes.byref(tab), ix["d_eoff"].data_ptr(),
                                                 ix["d_roff"].data_ptr(), ix["d_nE"].data_ptr(), ix["d_remap"].data_ptr(),
                                                 d_items.data_ptr() + i0 * 24, i1 - i0, energy.data_ptr() - k0 * E * 4, E, st),
                        "pk_universe_energies")
                self.gpu_launches += lib.pk_last_launch_count()
                if tuples:
                    N.check(lib.pk_universe_tuple_scores(ctypes.byref(cfg), ctypes.byref(tab), ix["d_eoff"].data_ptr(),
                                                         ix["d_roff"].data_ptr(), d_items.data_ptr() + i0 * 24, i1 - i0,
                                                         tuple_score.data_ptr() - k0 * 4, st), "pk_universe_tuple_scores")
                    self.gpu_launches += lib.pk_last_launch_count()
            works = []
            if sharded:       # NCCL min over NVLink: the one exchange step of the whole path
                works.append(dist.all_reduce(energy, op=dist.ReduceOp.MIN, async_op=True))
                if tuples:
                    works.append(dist.all_reduce(tuple_score, op=dist.ReduceOp.MIN, async_op=True))
            return k0, k1, energy, tuple_score, works

        def finish(ti, k0, k1, energy, tuple_score, works):
            for w in works:
                w.wait()          # orders this stream after the reduction; the host does not block
            if null_vector:   # reference :634-640: candidates no universe scored get the key's tuple score
                for r0 in range(0, k1 - k0, 65535):
                    r1 = min(k1 - k0, r0 + 65535)
                    N.check(lib.pk_fill_missing_energies(energy[r0:r1].data_ptr(), r1 - r0, E, tuple_score[r0:r1].data_ptr(), st),
                            "pk_fill_missing_energies")
                    self.gpu_launches += lib.pk_last_launch_count()
            consume(ti, k0, k1, energy, tuple_score)

        pending = []
        for ti in range(n_tiles):
            pending.append((ti,) + fold(ti))
            if len(pending) == len(bufs):      # the buffer that tile ti+1 will reuse must have been consumed
                finish(*pending.pop(0))
        while pending:
            finish(*pending.pop(0))
        return rows_per_tile

    def _eval_operands(self, ck, dev):
        """(model configuration, tables) the energy kernels score a chunk with.  TransE with norm_flag: the operand of a
        candidate is its normalised row whatever the key, and an ensemble evaluation visits a universe once per key that
        it can answer — the chunk's tables are normalised ONCE (same operations, same order: bit-identical energies) and
        scored with norm_flag = 0; 45 % of the energy kernel's instructions were that division."""
        cfg = ck.proto.native_cfg()
        if ck.proto._pk_model != N.PK_TRANSE or not cfg.norm_flag:
            return cfg, self._packed_tables(ck)
        cache = getattr(ck, "norm_tables", None)
        if cache is None:
            st = torch.cuda.current_stream(dev).cuda_stream
            cache = {}
            for name, t in ck.tables.items():
                out = self._arena_rows(dev, t.shape[0], t.shape[1])      # from the table slab: no cudaMalloc per chunk
                N.check(self.lib.pk_normalise_rows(t.data_ptr(), out.data_ptr(), t.shape[0], t.shape[1], st), "pk_normalise_rows")
                self.gpu_launches += self.lib.pk_last_launch_count()
                cache[name] = out
            ck.norm_tables = cache
        cfg.norm_flag = 0
        t = N.Tables()
        for i in range(2):
            t.ent[i] = t.rel[i] = t.ent_state[i] = t.rel_state[i] = None
        for i, name in enumerate(ck.proto._ent_tables):
            t.ent[i] = cache[name].data_ptr()
        for i, name in enumerate(ck.proto._rel_tables):
            t.rel[i] = cache[name].data_ptr()
        t.n_ent, t.n_rel = int(ck.eoff[-1]), int(ck.roff[-1])
        return cfg, t

    def _rank_split(self, loader):
        """Raw/filtered ranks [n,4] of every triple of `loader` under the min-over-universes energy
        (reference eval_universes :556-603 + global_energy_estimation :605-642 + Test.h ranking).  Keys = distinct
        (side, fixed entity, relation) of the queries; under torch.distributed the ranks of a tile's queries are
        counted by the rank that owns them (a contiguous share of the tile), then summed."""
        lib = self.lib
        dev = self._device()
        dist, rank, world = _dist()
        sharded = dist is not None and world > 1
        E = self.ent_tot
        q = self._query_arrays(loader, dev)
        n, K, order = q["n"], q["K"], q["order"]
        k_fixed, k_rel, k_side, sorted_keys = q["k_fixed"], q["k_rel"], q["k_side"], q["sorted_keys"]
        d_off, d_cand, d_truth, d_keyrow = q["d_off"], q["d_cand"], q["d_truth"], q["d_keyrow"]
        ranks_sorted = torch.zeros((2 * n, 2), dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        tile = max(1, min(K, int(self.eval_tile_rows), int(self.max_energy_bytes // (3 * 4 * E))))
        q_bounds = np.searchsorted(sorted_keys, np.arange(0, K + tile, tile))

        # incremental setting: the candidates are the entities the snapshot currently contains (Test.h:181-206)
        d_mask, n_cand = None, 0
        if hasattr(loader, "candidate_mask"):
            mask = loader.candidate_mask()
            d_mask, n_cand = torch.from_numpy(mask).to(dev), int(mask.sum())

        def count(ti, k0, k1, energy, tuple_score):
            lo, hi = int(q_bounds[ti]), int(q_bounds[ti + 1])
            if sharded:       # this rank's share of the tile's queries (the others add zeros)
                per = (hi - lo + world - 1) // world
                lo, hi = min(hi, lo + rank * per), min(hi, lo + (rank + 1) * per)
            if hi > lo and d_mask is not None:
                N.check(lib.pk_rank_from_energy_masked(energy.data_ptr() - k0 * E * 4, E, hi - lo, d_keyrow.data_ptr() + lo * 4,
                                                       d_truth.data_ptr() + lo * 4, d_off.data_ptr() + lo * 8, d_cand.data_ptr(),
                                                       ranks_sorted.data_ptr() + lo * 8, d_mask.data_ptr(), n_cand, st),
                        "pk_rank_from_energy_masked")
                self.gpu_launches += lib.pk_last_launch_count()
            elif hi > lo:
                N.check(lib.pk_rank_from_energy(energy.data_ptr() - k0 * E * 4, E, hi - lo, d_keyrow.data_ptr() + lo * 4,
                                                d_truth.data_ptr() + lo * 4, d_off.data_ptr() + lo * 8, d_cand.data_ptr(),
                                                ranks_sorted.data_ptr() + lo * 8, st), "pk_rank_from_energy")
                self.gpu_launches += lib.pk_last_launch_count()

        used = self._fold_keys(k_fixed, k_rel, k_side, count, keys_token=q["token"])
        assert used == tile
        if sharded:
            dist.all_reduce(ranks_sorted)      # every query was ranked by exactly one rank
        r_sorted = ranks_sorted.cpu().numpy()
        r = np.empty_like(r_sorted)
        r[order] = r_sorted
        return np.concatenate([r[:n], r[n:]], axis=1)   # head raw, head filt, tail raw, tail filt

    _TOKENS = [0]

    def _query_arrays(self, loader, dev):
        """Everything about a loader's queries that does not depend on the universes — keys, key order, the filter
        lists gathered in key order — built once per loader (vectorised numpy) and kept on the device."""
        q = getattr(loader, "_pu_query_arrays", None)
        if q is not None and q["device"] == dev and q["E"] == self.ent_tot and q["arrays"] is loader.eval_arrays():
            return q
        arrays = loader.eval_arrays()
        tri, filt = arrays
        n = tri.shape[0]
        E = self.ent_tot
        # keys: head side fixes (t, r); tail side fixes (h, r)
        fixed = np.concatenate([tri[:, 2], tri[:, 0]]).astype(np.int64)
        rel = np.concatenate([tri[:, 1], tri[:, 1]]).astype(np.int64)
        side = np.concatenate([np.zeros(n, np.int64), np.ones(n, np.int64)])
        truth = np.concatenate([tri[:, 0], tri[:, 2]]).astype(np.int32)
        code = (side * E + fixed) * self.rel_tot + rel
        ucode, key_of_query = np.unique(code, return_inverse=True)
        foff = np.concatenate([filt[0][0], filt[0][0][-1] + filt[1][0][1:]]).astype(np.int64)
        fcand = np.concatenate([filt[0][1], filt[1][1]]).astype(np.int32)
        # queries in key order, with their filter lists gathered in that order (one vectorised pass)
        order = np.argsort(key_of_query, kind="stable")
        sorted_keys = key_of_query[order]
        cnt = foff[order + 1] - foff[order]
        s_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        take = np.repeat(foff[order] - s_off[:-1], cnt) + np.arange(int(s_off[-1]), dtype=np.int64)
        s_cand = fcand[take] if take.size else np.zeros(1, np.int32)
        Parallel_Universe_Config._TOKENS[0] += 1
        q = {"device": dev, "E": E, "arrays": arrays, "token": Parallel_Universe_Config._TOKENS[0], "n": n, "K": int(ucode.shape[0]),
             "order": order, "sorted_keys": sorted_keys, "k_rel": ucode % self.rel_tot, "k_fixed": (ucode // self.rel_tot) % E,
             "k_side": ucode // (self.rel_tot * E), "d_off": torch.from_numpy(s_off).to(dev),
             "d_cand": torch.from_numpy(np.ascontiguousarray(s_cand, dtype=np.int32)).to(dev),
             "d_truth": torch.from_numpy(np.ascontiguousarray(truth[order])).to(dev),
             "d_keyrow": torch.from_numpy(sorted_keys.astype(np.int32)).to(dev)}
        loader._pu_query_arrays = q
        return q

    def triple_energies(self, batch_h, batch_t, batch_r):
        """Global energy of explicit triples, min over the universes that hold head, relation AND tail, +inf when
        there is none (reference predict_triple :413-444); with missing_embedding_handling='null_vector' an
        unscored triple gets min(tuple score of (h, r), tuple score of (r, t)) (reference test_one_step :720-729).

        Note on the reference: its test_one_step unpacks `head, tail, rel = batch_h[i], batch_r[i], batch_t[i]`
        (:716) and so looks relations up under tail ids and vice versa; this implements what predict_triple's own
        signature says (head, relation, tail) — see DESIGN.md, decisions on reference defects."""
        dev = self._device()
        h = np.asarray(batch_h, dtype=np.int64).reshape(-1)
        t = np.asarray(batch_t, dtype=np.int64).reshape(-1)
        r = np.asarray(batch_r, dtype=np.int64).reshape(-1)
        n = h.shape[0]
        if n == 0:
            return np.zeros(0, np.float32)
        E, R = self.ent_tot, self.rel_tot
        null_vector = self.missing_embedding_handling == "null_vector"
        # the tail-side key (h, r) carries the triple's energy in column t; the head-side key (t, r) is only needed
        # for its tuple score
        fixed = np.concatenate([h, t]) if null_vector else h
        rel = np.concatenate([r, r]) if null_vector else r
        side = np.concatenate([np.ones(n, np.int64), np.zeros(n, np.int64)]) if null_vector else np.ones(n, np.int64)
        code = (side * E + fixed) * R + rel
        ucode, key_of = np.unique(code, return_inverse=True)
        k_rel, k_fixed, k_side = ucode % R, (ucode // R) % E, ucode // (R * E)
        d_key = torch.from_numpy(key_of[:n].astype(np.int64)).to(dev)
        d_t = torch.from_numpy(t).to(dev)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        tuple_all = torch.empty(ucode.shape[0], dtype=torch.float32, device=dev) if null_vector else None
        order = torch.argsort(d_key)
        sorted_key = d_key[order]

        def gather(ti, k0, k1, energy, tuple_score):
            lo = int(torch.searchsorted(sorted_key, k0).item())
            hi = int(torch.searchsorted(sorted_key, k1).item())
            if hi > lo:
                q = order[lo:hi]
                out[q] = energy[d_key[q] - k0, d_t[q]]
            if null_vector:
                tuple_all[k0:k1] = tuple_score

        # the raw per-universe energies are wanted here: null-vector FILLING of whole rows is for ranking only
        saved = self.missing_embedding_handling
        self.missing_embedding_handling = "last_rank"
        try:
            self._fold_keys(k_fixed, k_rel, k_side, gather, want_tuple=null_vector)
        finally:
            self.missing_embedding_handling = saved
        if null_vector:
            d_key_head = torch.from_numpy(key_of[n:].astype(np.int64)).to(dev)
            fill = torch.minimum(tuple_all[d_key], tuple_all[d_key_head])
            out = torch.where(torch.isinf(out), fill, out)
        return out.cpu().numpy()

    def _chunk_index(self, ck):
        """Inverted index of one chunk: which universes hold a global entity / relation, and under
        which local id (replaces the reference's entity_universes / *_id_mappings dict lookups)."""
        if getattr(ck, "index", None) is not None:
            return ck.index
        dev = self._device()
        nU = len(ck.ids)
        univ_of_row = np.repeat(np.arange(nU, dtype=np.int32), ck.nE)
        local = (np.arange(ck.ent_remap.shape[0], dtype=np.int64) - np.repeat(ck.eoff[:-1], ck.nE)).astype(np.int32)
        # stable order of the chunk's ~1 500 entities per universe by global id: numpy sorts 16-bit keys by radix (1 ms per
        # 100 universes) and 32-bit keys by merging (7 ms, more than the chunk's energy kernels); larger id spaces are
        # sorted on the device
        if self.ent_tot <= 65535:
            order = np.argsort(ck.ent_remap.astype(np.uint16), kind="stable")
        elif torch.cuda.is_available() and self.use_gpu:
            d_ids = ck.d_ent_remap if ck.d_ent_remap is not None else torch.from_numpy(ck.ent_remap).to(dev)
            order = torch.argsort(d_ids, stable=True).cpu().numpy()
        else:
            order = np.argsort(ck.ent_remap, kind="stable")
        ix = {"ent_sorted": ck.ent_remap[order], "ent_univ": univ_of_row[order], "ent_local": local[order]}
        rel_local = np.full((self.rel_tot, nU), -1, dtype=np.int32)
        runiv = np.repeat(np.arange(nU, dtype=np.int32), ck.nR)
        rloc = (np.arange(ck.rel_remap.shape[0], dtype=np.int64) - np.repeat(ck.roff[:-1], ck.nR)).astype(np.int32)
        rel_local[ck.rel_remap, runiv] = rloc
        ix["rel_local"] = rel_local
        ix["d_eoff"] = torch.from_numpy(ck.eoff[:-1].copy()).to(dev)
        ix["d_roff"] = torch.from_numpy(ck.roff[:-1].copy()).to(dev)
        ix["d_nE"] = torch.from_numpy(ck.nE.astype(np.int32)).to(dev)
        ix["d_remap"] = ck.d_ent_remap if ck.d_ent_remap is not None else torch.from_numpy(ck.ent_remap).to(dev)
        ck.index = ix
        return ix

    @staticmethod
    def _energy_items(ix, k_fixed, k_rel, k_side):
        """(key row, universe, local fixed entity, local relation, side) for every universe of the
        chunk that contains both the key's fixed entity and its relation."""
        lo = np.searchsorted(ix["ent_sorted"], k_fixed, side="left")
        hi = np.searchsorted(ix["ent_sorted"], k_fixed, side="right")
        cnt = hi - lo
        tot = int(cnt.sum())
        items = np.zeros(tot, dtype=N.ENERGY_ITEM_DTYPE)
        if tot == 0:
            return items
        key_row = np.repeat(np.arange(k_fixed.shape[0], dtype=np.int64), cnt)
        pos = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt) + np.repeat(lo, cnt)
        u = ix["ent_univ"][pos]
        rl = ix["rel_local"][k_rel[key_row], u]
        keep = rl >= 0
        items = items[:int(keep.sum())]
        items["key_row"] = key_row[keep]
        items["universe"] = u[keep]
        items["fixed_local"] = ix["ent_local"][pos][keep]
        items["rel_local"] = rl[keep]
        items["side"] = k_side[key_row][keep]
        return items

    def _rank_state(self, eval_mode):
        loader = self.data_loader if eval_mode == "test" else self.valid_dataloader
        return (self._visible_next_id(), self.incremental_strategy, self.missing_embedding_handling, id(loader.eval_arrays()),
                len(getattr(self.train_dataloader, "deleted_triple_set", ())))

    def eval_universes(self, eval_mode):
        loader = self.data_loader if eval_mode == "test" else self.valid_dataloader
        self._rank_cache[eval_mode] = (self._rank_state(eval_mode), self._rank_split(loader))

    def reset_evaluation_helpers(self):
        """Reference :545-554."""
        self._rank_cache.clear()
        self._energy_cache.clear()
        self.incremental_strategy = "normal"

    # ---- incremental setting: universes that hold a since-deleted triple (reference :817-823)
    def _universes_holding(self, h, t, r):
        """Ids of the universes that contain head, relation AND tail of at least one of the given triples."""
        out = []
        h, t, r = (np.asarray(x, dtype=np.int64) for x in (h, t, r))
        for ck in self._chunks:
            ix = self._chunk_index(ck)
            lo = np.searchsorted(ix["ent_sorted"], h, side="left")
            hi = np.searchsorted(ix["ent_sorted"], h, side="right")
            cnt = hi - lo
            tot = int(cnt.sum())
            if tot == 0:
                continue
            tri = np.repeat(np.arange(h.shape[0], dtype=np.int64), cnt)
            pos = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt) + np.repeat(lo, cnt)
            u = ix["ent_univ"][pos].astype(np.int64)
            keep = ix["rel_local"][r[tri], u] >= 0
            tri, u = tri[keep], u[keep]
            # (universe, tail) membership through the sorted (entity, universe) pairs
            pair = ix.get("pair_code")
            if pair is None:
                pair = np.sort(ix["ent_sorted"].astype(np.int64) * len(ck.ids) + ix["ent_univ"].astype(np.int64))
                ix["pair_code"] = pair
            want = t[tri] * len(ck.ids) + u
            at = np.searchsorted(pair, want)
            at[at >= pair.shape[0]] = pair.shape[0] - 1
            hit = pair[at] == want
            out.append(np.asarray(ck.ids, dtype=np.int64)[np.unique(u[hit])])
        return np.unique(np.concatenate(out)) if out else np.zeros(0, np.int64)

    def determine_deprecated_embedding_spaces(self):
        deleted = getattr(self.train_dataloader, "deleted_triple_set", None) or ()
        old = set(self.deprecated_embeddingspaces)
        self.deprecated_embeddingspaces.clear()
        if deleted:
            arr = np.array([[int(a), int(b), int(c)] for a, b, c in deleted], dtype=np.int64)   # (head, tail, relation)
            self.deprecated_embeddingspaces.update(int(u) for u in self._universes_holding(arr[:, 0], arr[:, 1], arr[:, 2]))
        if old != self.deprecated_embeddingspaces or self._deprecated_version == 0:
            self._deprecated_version += 1

    def load_triple_classification_file(self, file):
        """Reference :767-789: lines "head tail relation truth"."""
        cols = [[], [], [], [], [], []]
        with open(str(file), mode="rt", encoding="UTF-8") as f:
            for line in f:
                head, tail, rel, truth_value = line.split()
                o = 0 if truth_value == "1" else 3
                cols[o].append(int(head)); cols[o + 1].append(int(tail)); cols[o + 2].append(int(rel))
        return tuple(cols)

    def tc_datastructure_adapter(self, pos_h, pos_t, pos_r, neg_h, neg_t, neg_r):
        """Reference :751-765."""
        a = lambda x: np.asarray(x, dtype=np.int64) if len(x) else np.empty(0, dtype=np.int64)
        return [({"batch_h": a(pos_h), "batch_t": a(pos_t), "batch_r": a(pos_r), "mode": "normal"},
                 {"batch_h": a(neg_h), "batch_t": a(neg_t), "batch_r": a(neg_r), "mode": "normal"})]

    def run_classification_of_deleted_triples(self, snapshot_idx, threshlod):
        """Reference :791-800: every tc_* file of the snapshot, classified with the given threshold."""
        folder = os.path.join(self.data_loader.in_path, "incremental", str(snapshot_idx))
        res = {}
        for name in sorted(os.listdir(folder)):
            if name.startswith("tc_"):
                data = self.tc_datastructure_adapter(*self.load_triple_classification_file(os.path.join(folder, name)))
                acc, _ = Tester.run_triple_classification(self, threshlod, data_iterator=data)
                print("Accuracy for {} is: {}".format(name, acc))
                res[name] = acc
        return res

    def run_triple_classification_from_files(self, snapshot):
        """Reference :802-815."""
        folder = os.path.join(self.data_loader.in_path, "incremental", str(snapshot))
        name = "triple_classification_prepared_test_examples.txt"
        data = self.tc_datastructure_adapter(*self.load_triple_classification_file(os.path.join(folder, name)))
        acc, threshlod = Tester.run_triple_classification(self, data_iterator=data)
        print("Accuracy for {} is: {}".format(name, acc))
        print("Determined threshold: {}".format(threshlod))
        print("Run negative triple classification...")
        return acc, threshlod, self.run_classification_of_deleted_triples(snapshot, threshlod)

    def _ranks(self, eval_mode):
        c = self._rank_cache.get(eval_mode)
        if c is None or c[0] != self._rank_state(eval_mode):
            self.eval_universes(eval_mode)
            c = self._rank_cache[eval_mode]
        return c[1]

    def valid(self):
        ranks = self._ranks("valid")
        n = np.float32(ranks.shape[0])
        return float((np.float32((ranks[:, 1] < 10).sum()) / n + np.float32((ranks[:, 3] < 10).sum()) / n) / np.float32(2))

    def run_link_prediction(self, type_constrain=False):
        if type_constrain:
            raise NotImplementedError("type-constrained ranking is not on the PuTransE hot path")
        self.data_loader.set_sampling_mode("link")
        self.last_ranks = self._ranks("test")
        (mrr, mr, hit10, hit3, hit1), self.last_table = link_metrics(self.last_ranks)
        print("Mean Reciprocal Rank: {}".format(mrr))
        print("Mean Rank: {}".format(mr))
        print("Hits@10: {}".format(hit10))
        print("Hits@3: {}".format(hit3))
        print("Hits@1: {}".format(hit1))
        return mrr, mr, hit10, hit3, hit1

    def global_energy_estimation(self, data):
        """Energy row of ONE candidate batch in the reference's candidate order (reference :605-642)."""
        mode = data["mode"]
        h, t, r = np.asarray(data["batch_h"]), np.asarray(data["batch_t"]), np.asarray(data["batch_r"])
        fixed = int(t[0]) if mode == "head_batch" else int(h[0])
        cands = h if mode == "head_batch" else t
        side = 0 if mode == "head_batch" else 1
        dev = self._device()
        self.synchronize()
        st = torch.cuda.current_stream(dev).cuda_stream
        energy = torch.empty((1, self.ent_tot), dtype=torch.float32, device=dev)
        N.check(self.lib.pk_fill_inf(energy.data_ptr(), energy.numel(), st), "pk_fill_inf")
        null_vector = self.missing_embedding_handling == "null_vector"
        tuple_score = torch.full((1,), float("inf"), dtype=torch.float32, device=dev) if null_vector else None
        for ck in self._chunks:
            ix = self._chunk_index(ck)
            items = self._energy_items(ix, np.array([fixed]), np.array([int(r[0])]), np.array([side]))
            if items.shape[0]:
                d_items = torch.from_numpy(items.view(np.int32).reshape(-1, 6)).to(dev)
                cfg, tab = ck.proto.native_cfg(), self._packed_tables(ck)
                N.check(self.lib.pk_universe_energies(ctypes.byref(cfg), ctypes.byref(tab), ix["d_eoff"].data_ptr(),
                                                      ix["d_roff"].data_ptr(), ix["d_nE"].data_ptr(), ix["d_remap"].data_ptr(),
                                                      d_items.data_ptr(), items.shape[0], energy.data_ptr(), self.ent_tot, st),
                        "pk_universe_energies")
                if null_vector:
                    N.check(self.lib.pk_universe_tuple_scores(ctypes.byref(cfg), ctypes.byref(tab), ix["d_eoff"].data_ptr(),
                                                              ix["d_roff"].data_ptr(), d_items.data_ptr(), items.shape[0],
                                                              tuple_score.data_ptr(), st), "pk_universe_tuple_scores")
        if null_vector:
            N.check(self.lib.pk_fill_missing_energies(energy.data_ptr(), 1, self.ent_tot, tuple_score.data_ptr(), st),
                    "pk_fill_missing_energies")
        return energy[0].cpu().numpy()[cands.astype(np.int64)]

    def test_one_step(self, data):
        """Reference :704-733: candidate rows for the two link-prediction modes, triple energies for 'normal'."""
        if data["mode"] == "normal":
            return self.triple_energies(data["batch_h"], data["batch_t"], data["batch_r"])
        return self.global_energy_estimation(data)

    def run_triple_classification(self, threshlod=None, data_iterator=None):
        """Reference :745-749 (Tester.run_triple_classification on the global energies)."""
        acc, threshlod = super().run_triple_classification(threshlod, data_iterator)
        print("Accuracy is: {}".format(acc))
        return acc, threshlod

    # ------------------------------------------------------------------ checkpoints
    # Two layouts.  "flat" (default): the packed chunk tensors + offsets + remaps — what thousands of universes
    # need.  "reference": the dict of reference :852-888 — one pickled nn.Module per
    # universe plus the nested id-map dicts — which the reference's own load_parameters reads (class and factory
    # paths are the same in both packages) and which load_parameters here imports.
    def save_model(self, filename=None):
        if not filename:
            filename = "Pu{}_learned_spaces-{}_{}.ckpt".format(self.embedding_model.__name__, self._visible_next_id(),
                                                               self.training_identifier)
        self.save_parameters(os.path.join("{}{}".format(self.checkpoint_dir, filename)))

    def extend_state_dict(self):
        self.synchronize()
        chunks = []
        for ck in self._visible_chunks():
            chunks.append({"ids": ck.ids, "nT": ck.nT, "nE": ck.nE, "nR": ck.nR, "ent_remap": ck.ent_remap,
                           "rel_remap": ck.rel_remap, "tables": {k: v.cpu() for k, v in ck.tables.items()}})
        upto = self._visible_next_id()
        return {"format": "putranse-b200/flat-1", "next_universe_id": upto, "chunks": chunks,
                "universe_hyper": {u: h for u, h in self.universe_hyper.items() if u < upto}, "embedding_model": self.embedding_model.__name__,
                "embedding_model_param": self.embedding_model_param, "best_hit10": self.best_hit10,
                "bad_counts": self.bad_counts, "initial_random_seed": self.initial_random_seed,
                "ent_tot": self.ent_tot, "rel_tot": self.rel_tot}

    def export_reference_state(self):
        """The reference's checkpoint dict (reference extend_state_dict :852-888): CPU nn.Modules and the four
        nested id maps, materialised from the packed chunks."""
        self.synchronize()
        spaces, ent_maps, rel_maps = {}, defaultdict(defaultdict_int), defaultdict(defaultdict_int)
        ent_univ, rel_univ = defaultdict(set), defaultdict(set)
        for ck in self._visible_chunks():
            for i, u in enumerate(ck.ids):
                with torch.random.fork_rng(devices=[]):
                    sp = self.embedding_model(int(ck.nE[i]), int(ck.nR[i]), **self.embedding_model_param)
                for name in sp.table_names():
                    off = ck.eoff if name in sp._ent_tables else ck.roff
                    getattr(sp, name).weight.data.copy_(ck.tables[name][off[i]:off[i + 1]])
                for q in sp.parameters():
                    q.requires_grad = False
                spaces[u] = sp
                for l, g in enumerate(ck.ent_remap[ck.eoff[i]:ck.eoff[i + 1]].tolist()):
                    ent_maps[u][g] = l
                    ent_univ[g].add(u)
                for l, g in enumerate(ck.rel_remap[ck.roff[i]:ck.roff[i + 1]].tolist()):
                    rel_maps[u][g] = l
                    rel_univ[g].add(u)
        for u, sp in self.trained_embedding_spaces._extra.items():
            spaces[u] = sp
        state = {"initial_num_universes": self.initial_num_universes, "next_universe_id": self._visible_next_id(),
                 "trained_embedding_spaces": spaces, "entity_id_mappings": ent_maps, "relation_id_mappings": rel_maps,
                 "entity_universes": ent_univ, "relation_universes": rel_univ}
        for k_ in ("min_margin", "max_margin", "min_lr", "max_lr", "min_num_epochs", "max_num_epochs", "min_triple_constraint",
                   "max_triple_constraint", "min_balance", "max_balance", "embedding_model", "embedding_model_param", "best_hit10",
                   "bad_counts"):
            state[k_] = getattr(self, k_)
        # the reference stores its evaluation caches too and never reads them back (its `if '' in state_dict`, :921)
        state.update(current_tested_universes=0, current_validated_universes=0, evaluation_head2tail_triple_score_dict={},
                     evaluation_tail2head_triple_score_dict={}, evaluation_head2rel_tuple_score_dict={},
                     evaluation_tail2rel_tuple_score_dict={})
        return state

    def save_parameters(self, path, layout="flat"):
        """Collective under torch.distributed: every rank holds only the universes u % world == rank, so rank 0
        writes `path` (with the list of its sibling shards) and rank r > 0 writes `path.rank<r>of<world>`."""
        dist, rank, world = _dist()
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        if layout == "reference":
            if world > 1:
                raise NotImplementedError("the reference layout is a single-process pickle: gather with layout='flat' and re-export")
            torch.save(self.export_reference_state(), path)
            return
        state = self.extend_state_dict()
        if world > 1:
            state["shard"] = (rank, world)
            state["shards"] = ["%s.rank%dof%d" % (os.path.basename(path), r, world) for r in range(1, world)]
            torch.save(state, path if rank == 0 else "%s.rank%dof%d" % (path, rank, world))
            dist.barrier()      # a checkpoint is complete when every shard is on disk
        else:
            torch.save(state, path)

    def _adopt_chunk(self, ids, nT, nE, nR, ent_remap, rel_remap, tables):
        # without a CUDA device a checkpoint can still be read, inspected and re-exported (conversion between the two
        # layouts is not compute); every compute entry point asks for the device itself and fails loudly
        dev = self._device() if (self.use_gpu and torch.cuda.is_available()) else torch.device("cpu")
        ck = _Chunk()
        ck.items_cache = {}
        ck.ids = list(ids)
        ck.nT, ck.nE, ck.nR = (np.asarray(x, dtype=np.int64) for x in (nT, nE, nR))
        ck.ent_remap, ck.rel_remap = np.ascontiguousarray(ent_remap, dtype=np.int32), np.ascontiguousarray(rel_remap, dtype=np.int32)
        ck.eoff = np.concatenate([[0], np.cumsum(ck.nE)]).astype(np.int64)
        ck.roff = np.concatenate([[0], np.cumsum(ck.nR)]).astype(np.int64)
        ck.toff = np.concatenate([[0], np.cumsum(ck.nT)]).astype(np.int64)
        ck.tables = {k: v.to(dev).contiguous() for k, v in tables.items()}
        ck.state = None
        ck.d_loss = ck.train_inputs = None
        ck.proto = self._proto()
        for i, u in enumerate(ck.ids):
            self._where[u] = (ck, i)
        self._chunks.append(ck)
        return ck

    def _reset_containers(self):
        self.synchronize()
        self._chunks = []
        self._where = {}
        for m in (self.trained_embedding_spaces, self.entity_id_mappings, self.relation_id_mappings,
                  self.entity_universes, self.relation_universes):
            m.clear()
        self._rank_cache.clear()
        self._energy_cache.clear()

    def process_state_dict(self, state, shard_states=()):
        if "format" not in state and "trained_embedding_spaces" in state:
            return self.import_reference_state(state)
        dist, rank, world = _dist()
        self._reset_containers()
        chunks = [c for st_ in (state,) + tuple(shard_states) for c in st_["chunks"]]
        have = sorted(u for c in chunks for u in c["ids"])
        if have != list(range(state["next_universe_id"])):
            raise N.NativeError("checkpoint holds %d of %d universes (a shard of a multi-GPU checkpoint is missing)"
                                % (len(have), state["next_universe_id"]))
        for j, c in enumerate(chunks):
            if world > 1 and j % world != rank:   # evaluation min-all-reduces over ranks: every chunk lives on ONE rank
                continue
            self._adopt_chunk(c["ids"], c["nT"], c["nE"], c["nR"], c["ent_remap"], c["rel_remap"], c["tables"])
        self._maps_version += 1
        self.next_universe_id = state["next_universe_id"]
        self.universe_hyper = dict(state.get("universe_hyper", {}))
        for st_ in shard_states:
            self.universe_hyper.update(st_.get("universe_hyper", {}))
        self.best_hit10 = state.get("best_hit10", 0)
        self.bad_counts = state.get("bad_counts", 0)

    def import_reference_state(self, state):
        """A checkpoint written by the reference (reference :852-888,929-931): one nn.Module per universe and
        global->local dicts; packed into one chunk.  Hyper-parameter ranges are restored as the reference
        restores them (:890-919)."""
        dist, rank, world = _dist()
        self._reset_containers()
        spaces = state["trained_embedding_spaces"]
        ids = sorted(spaces)
        names = self._proto().table_names()
        ent_names = set(self._proto()._ent_tables)
        nE, nR, er, rr = [], [], [], []
        tabs = {n_: [] for n_ in names}
        mine = [u for j, u in enumerate(ids) if world == 1 or j % world == rank]
        for u in mine:
            sd = spaces[u].state_dict()
            for n_ in names:
                tabs[n_].append(sd[n_ + ".weight"].detach().float().cpu())
            for maps, sizes, remaps, key in ((state["entity_id_mappings"], nE, er, next(iter(ent_names))),
                                             (state["relation_id_mappings"], nR, rr, [n_ for n_ in names if n_ not in ent_names][0])):
                rows = sd[key + ".weight"].shape[0]
                remap = np.full(rows, -1, dtype=np.int64)
                for g, l in maps[u].items():
                    remap[int(l)] = int(g)
                if (remap < 0).any():
                    raise N.NativeError("reference checkpoint: universe %d has local ids without a global id" % u)
                sizes.append(rows)
                remaps.append(remap)
        if mine:
            self._adopt_chunk(mine, np.zeros(len(mine), np.int64), nE, nR, np.concatenate(er), np.concatenate(rr),
                              {n_: torch.cat(tabs[n_], 0) for n_ in names})
        self._maps_version += 1
        self.next_universe_id = state["next_universe_id"]
        for k_ in ("initial_num_universes", "min_margin", "max_margin", "min_lr", "max_lr", "min_num_epochs", "max_num_epochs",
                   "min_triple_constraint", "max_triple_constraint", "min_balance", "max_balance", "best_hit10", "bad_counts"):
            if k_ in state:
                setattr(self, k_, state[k_])

    def load_parameters(self, filename):
        path = (self.checkpoint_dir or "") + filename
        state = torch.load(path, map_location="cpu", weights_only=False)
        shards = [torch.load(os.path.join(os.path.dirname(path), s_), map_location="cpu", weights_only=False)
                  for s_ in state.get("shards", [])] if isinstance(state, dict) else []
        self.process_state_dict(state, shards)

    def extend_parallel_universe(self, other):
        """Append another instance's universes after this one's (reference :797-823)."""
        self.synchronize()
        other.synchronize()
        shift = self.next_universe_id
        for ck in other._chunks:
            ck.ids = [u + shift for u in ck.ids]
            ck.index = None
            ck.items_cache = {}
            for i, u in enumerate(ck.ids):
                self._where[u] = (ck, i)
            self._chunks.append(ck)
        for u, space in other.trained_embedding_spaces._extra.items():
            self.trained_embedding_spaces[u + shift] = space
        for u, h in other.universe_hyper.items():
            self.universe_hyper[u + shift] = h
        self.next_universe_id += other.next_universe_id
        self._maps_version += 1
        self._rank_cache.clear()
