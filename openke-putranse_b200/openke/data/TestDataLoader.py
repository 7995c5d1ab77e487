"""Evaluation-triple source with the interface of the reference's
``openke/data/TestDataLoader.py:27-236`` (``mode='test'|'valid'``, link prediction only).

Iterating yields the reference's ``[head_batch_dict, tail_batch_dict]`` pairs (slot 0 = the true
triple, then every other entity ascending).  The rankers in ``openke.config`` do not iterate: they
take ``eval_arrays()`` — all triples plus the filter CSR — and rank everything in one device pass.
"""
import numpy as np

from .. import _native as N


class TestDataSampler(object):
    def __init__(self, data_total, data_sampler):
        self.data_total, self.data_sampler, self.total = data_total, data_sampler, 0

    def __iter__(self):
        return self

    def __next__(self):
        self.total += 1
        if self.total > self.data_total:
            raise StopIteration()
        return self.data_sampler()

    def __len__(self):
        return self.data_total


class TestDataLoader(object):
    __test__ = False  # not a pytest class

    def __init__(self, in_path="./", sampling_mode="link", random_seed=4, mode="test", setting="static",
                 load_all_triples=False):
        if setting != "static":
            raise NotImplementedError("only the static setting is on the B200 hot path (SURVEY.md 8(f))")
        if mode not in ("test", "valid"):
            raise ValueError("mode must be 'test' or 'valid'")
        self.lib = N.lib()
        self.setting, self.mode, self.load_all_triples = setting, mode, load_all_triples
        self.in_path, self.sampling_mode, self.random_seed = in_path, sampling_mode, random_seed
        self._arrays = None
        self.read()

    def set_path(self, in_path):
        self.lib.setInPath(in_path.encode())

    # reference TestDataLoader.py:86-146: note that this re-seeds the shared library state and
    # imports the training files once more (which is what drifts the reference's Bernoulli table).
    def read(self):
        self.set_path(self.in_path)
        self.lib.setRandomSeed(self.random_seed)
        self.lib.randReset()
        self.lib.importTrainFiles()
        if self.load_all_triples:   # the filter set comes from triple2id.txt (reference Reader.h:295-308)
            self.lib.activateLoadOfAllTriples(1)
        self.lib.importTestFiles()
        self.relTotal = self.lib.getRelationTotal()
        self.entTotal = self.lib.getEntityTotal()
        self.testTotal = self.lib.getTestTotal()
        self.validTotal = self.lib.getValidTotal()
        if self.testTotal == 0 and self.validTotal == 0:
            raise N.NativeError("importTestFiles: %s" % N.last_error())
        self._h = np.zeros(self.entTotal, dtype=np.int64)
        self._t = np.zeros(self.entTotal, dtype=np.int64)
        self._r = np.zeros(self.entTotal, dtype=np.int64)

    def eval_arrays(self):
        """(triples int32 [n,3] (h,r,t) sorted (r,h,t), [(offsets, candidates) for head side, tail side])."""
        if self._arrays is None:
            which = 0 if self.mode == "test" else 1
            n = self.testTotal if which == 0 else self.validTotal
            tri = np.zeros((n, 3), dtype=np.int32)
            N.check(self.lib.pk_eval_triples(which, N.addr(tri)), "pk_eval_triples")
            filt = []
            for side in (0, 1):
                import ctypes
                cnt = ctypes.c_int64(0)
                off = np.zeros(n + 1, dtype=np.int64)
                N.check(self.lib.pk_filter_csr(which, side, N.addr(off), None, ctypes.byref(cnt)), "pk_filter_csr")
                cand = np.zeros(max(cnt.value, 1), dtype=np.int32)
                N.check(self.lib.pk_filter_csr(which, side, N.addr(off), N.addr(cand), ctypes.byref(cnt)), "pk_filter_csr")
                filt.append((off, cand[:cnt.value] if cnt.value else cand[:0]))
            self._arrays = (tri, filt)
        return self._arrays

    def sampling_lp(self):
        head, tail = (self.lib.getHeadBatch, self.lib.getTailBatch) if self.mode == "test" else \
                     (self.lib.getValidHeadBatch, self.lib.getValidTailBatch)
        res = []
        head(N.addr(self._h), N.addr(self._t), N.addr(self._r))
        res.append({"batch_h": self._h.copy(), "batch_t": self._t[:1].copy(), "batch_r": self._r[:1].copy(),
                    "mode": "head_batch"})
        tail(N.addr(self._h), N.addr(self._t), N.addr(self._r))
        res.append({"batch_h": self._h[:1].copy(), "batch_t": self._t.copy(), "batch_r": self._r[:1].copy(),
                    "mode": "tail_batch"})
        return res

    def sampling_tc(self):
        """Triple classification inputs (reference TestDataLoader.py:183-206, Test.h:573-599): every test triple
        and its corrupted twin, drawn by the device kernel behind getTestBatch."""
        n = self.testTotal
        buf = [np.zeros(n, dtype=np.int64) for _ in range(6)]
        self.lib.getTestBatch(*[N.addr(b) for b in buf])
        return [{"batch_h": buf[0], "batch_t": buf[1], "batch_r": buf[2], "mode": "normal"},
                {"batch_h": buf[3], "batch_t": buf[4], "batch_r": buf[5], "mode": "normal"}]

    def get_ent_tot(self):
        return self.entTotal

    def get_rel_tot(self):
        return self.relTotal

    def get_triple_tot(self):
        return self.testTotal

    def set_sampling_mode(self, sampling_mode):
        self.sampling_mode = sampling_mode

    def __len__(self):
        return self.testTotal if self.mode == "test" else self.validTotal

    def __iter__(self):
        if self.sampling_mode != "link":      # reference TestDataLoader.py:234-236
            self.lib.initTest()
            return TestDataSampler(1, self.sampling_tc)
        if self.mode == "test":
            self.lib.initTest()
            return TestDataSampler(self.testTotal, self.sampling_lp)
        self.lib.validInit()
        return TestDataSampler(self.validTotal, self.sampling_lp)
