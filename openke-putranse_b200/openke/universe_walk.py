"""Universe subgraphs built on the GPU (csrc/walk_device.cu) for the orchestrator.

Replaces the host threads of ``pk_universes_build_lean`` for the reference's ``getParallelUniverse``
(openke/base/UniverseConstructor.h:327-397, called through ``TrainDataLoader.compile_universe_dataset``,
openke/data/TrainDataLoader.py:117-130) when training needs only the lean universe (no filter, no Bernoulli):
one launch per chunk on its own high-priority stream, results copied into pinned memory behind it.
The output buffers are strided (PK_WALK_CAP triples per universe) and stay on the device: the training kernel
reads the local triple lists where the walk wrote them."""
import ctypes

import numpy as np
import torch

from . import _native as N


class WalkResult(object):
    """One chunk's walk in flight.  ``wait()`` blocks until the sizes are on the host."""

    def __init__(self, walker, bufs, n, seeds, tcs, bals, lcg, focus):
        self.walker, self.bufs, self.n = walker, bufs, n
        self.seeds, self.tcs, self.bals, self.lcg, self.focus = seeds, tcs, bals, lcg, focus
        self.event = torch.cuda.Event(blocking=True)
        self.sizes = None
        self.copied_remaps = False

    def wait(self):
        if self.sizes is None:
            self.event.synchronize()
            self.sizes = self.bufs["h_sizes"][:self.n].numpy().astype(np.int64)
        return self.sizes

    @property
    def ok(self):
        return bool((self.wait()[:, 5] == 0).all())

    def packed_remaps(self):
        """(ent_remap [sum nE], rel_remap [sum nR]) int32 numpy, universes back to back (what evaluation indexes)."""
        s = self.wait()
        if self.bufs["h_ent_remap"] is not None and self.copied_remaps:
            er = self.bufs["h_ent_remap"][:self.n].numpy()
            rr = self.bufs["h_rel_remap"][:self.n].numpy()
        else:       # not copied behind the launch: fetch them now
            er = self.bufs["ent_remap"][:self.n].cpu().numpy()
            rr = self.bufs["rel_remap"][:self.n].cpu().numpy()
        nE, nR = s[:, 1].tolist(), s[:, 2].tolist()
        return (np.concatenate([er[i, :nE[i]] for i in range(self.n)]), np.concatenate([rr[i, :nR[i]] for i in range(self.n)]))

    def release(self):
        """The training launch that read d_tri has finished: the buffers may carry another chunk."""
        if self.bufs is not None:
            self.walker._free.append(self.bufs)
            self.bufs = None


class DeviceWalker(object):
    def __init__(self, lib, device):
        self.lib, self.device = lib, device
        self.cap = int(lib.pk_walk_cap())
        lo, hi = torch.cuda.Stream.priority_range()
        # two streams, used alternately: under training load a walk takes 4.9 ms from queue to finish (its blocks wait for
        # SMs the training kernel vacates) against 5.3 ms per training step — one stream would nearly be the bottleneck.
        # High priority: its few blocks go first when an SM frees up.
        self.streams = [torch.cuda.Stream(device=device, priority=hi) for _ in range(2)]
        self.stream = self.streams[0]
        self._next_stream = 0
        self._free = []
        self.launches = 0
        self.d2h_bytes = 0
        self.time_walks = False       # profiling aid: CUDA events around every walk (queue-to-finish on the device)
        self.walk_events = []

    def usable(self):
        return self.lib.pk_walk_device_check() == 0

    def _buffers(self, n):
        key = self._graph_key()
        self._free = [b for b in self._free if b["graph"] == key]    # buffers sized for a graph that is gone
        for i, b in enumerate(self._free):
            if b["n"] >= n:
                return self._free.pop(i)
        cap, dev = self.cap, self.device
        need = np.zeros(3, dtype=np.int64)
        m = int(n * 1.25) + 8
        N.check(self.lib.pk_walk_scratch_bytes(m, N.addr(need)), "pk_walk_scratch_bytes")
        b = dict(n=m, graph=self._graph_key(),
                 bitmaps=torch.empty(max(int(need[0]) // 4, 1), dtype=torch.int32, device=dev),
                 got=torch.empty((m, cap, 3), dtype=torch.int32, device=dev),
                 trees=torch.empty(max(int(need[2]) // 4, 1), dtype=torch.int32, device=dev) if need[2] else None,
                 tri=torch.empty((m, cap, 3), dtype=torch.int32, device=dev),
                 ent_remap=torch.empty((m, 2 * cap), dtype=torch.int32, device=dev),
                 rel_remap=torch.empty((m, cap), dtype=torch.int32, device=dev),
                 sizes=torch.empty((m, 8), dtype=torch.int32, device=dev),
                 h_sizes=torch.empty((m, 8), dtype=torch.int32, pin_memory=True),
                 h_ent_remap=None, h_rel_remap=None)     # pinned copies of the strided remaps: only on request
        return b

    def _graph_key(self):
        return (int(self.lib.pk_import_count()), int(self.lib.getEntityTotal()), int(self.lib.getTrainTotal()))

    def submit(self, seeds, tcs, bals, work_threads, copy_remaps=True):
        """Launch the walk of len(seeds) universes; returns a WalkResult at once."""
        n = len(seeds)
        seeds = np.ascontiguousarray(seeds, dtype=np.int64)
        tcs = np.ascontiguousarray(tcs, dtype=np.int64)
        bals = np.ascontiguousarray(bals, dtype=np.float32)
        lcg = np.zeros((n, int(work_threads)), dtype=np.uint64)
        focus = np.zeros(n, dtype=np.int64)
        b = self._buffers(n)
        res = WalkResult(self, b, n, seeds, tcs, bals, lcg, focus)
        self.stream = self.streams[self._next_stream % len(self.streams)]
        self._next_stream += 1
        with torch.cuda.stream(self.stream):
            if self.time_walks:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(self.stream)
            N.check(self.lib.pk_universes_walk_device(
                n, N.addr(seeds), N.addr(tcs), N.addr(bals), N.addr(lcg), N.addr(focus),
                b["bitmaps"].data_ptr(), b["got"].data_ptr(), b["trees"].data_ptr() if b["trees"] is not None else None,
                b["tri"].data_ptr(), b["ent_remap"].data_ptr(), b["rel_remap"].data_ptr(), b["sizes"].data_ptr(),
                ctypes.c_void_p(self.stream.cuda_stream)), "pk_universes_walk_device")
            self.launches += self.lib.pk_last_launch_count()
            b["h_sizes"][:n].copy_(b["sizes"][:n], non_blocking=True)
            self.d2h_bytes += n * 32
            if copy_remaps:
                if b["h_ent_remap"] is None:
                    b["h_ent_remap"] = torch.empty((b["n"], 2 * self.cap), dtype=torch.int32, pin_memory=True)
                    b["h_rel_remap"] = torch.empty((b["n"], self.cap), dtype=torch.int32, pin_memory=True)
                b["h_ent_remap"][:n].copy_(b["ent_remap"][:n], non_blocking=True)
                b["h_rel_remap"][:n].copy_(b["rel_remap"][:n], non_blocking=True)
                self.d2h_bytes += n * 3 * self.cap * 4
                res.copied_remaps = True
            res.event.record(self.stream)
            if self.time_walks:
                e1.record(self.stream)
                self.walk_events.append((e0, e1))
        return res
