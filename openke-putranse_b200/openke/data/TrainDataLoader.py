"""Training-batch source with the constructor, attributes and iteration protocol of the reference's
``openke/data/TrainDataLoader.py:27-329``.

Differences that matter:
* ``self.lib`` is libputranse.so.  ``sampling()`` returns the same numpy dict as the reference, but
  the batch is drawn by the CUDA sampler kernel (bit-identical to the reference's pthread sampler).
* ``device_sampler()`` exposes the current id space (global graph, or the universe after
  ``swap_helpers()``) as device-resident index arrays, so ``Trainer.run`` can keep sampling on the
  GPU and never materialise a host batch.
* cross sampling (``sampling_mode != 'normal'``) and relation negatives are not on the PuTransE hot path and
  are refused.  ``incremental_setting=True`` (reference :129: no training files are imported; the list is evolved
  by IncrementalTrainDataLoader.load_snapshot) is supported.
"""
import ctypes

import numpy as np

from .. import _native as N


class TrainDataSampler(object):
    def __init__(self, nbatches, datasampler):
        self.nbatches, self.datasampler, self.batch = nbatches, datasampler, 0

    def __iter__(self):
        return self

    def __next__(self):
        self.batch += 1
        if self.batch > self.nbatches:
            raise StopIteration()
        return self.datasampler()

    def __len__(self):
        return self.nbatches


class TrainDataLoader(object):
    def __init__(self, in_path="./", batch_size=None, nbatches=None, threads=8, sampling_mode="normal", bern_flag=0,
                 filter_flag=1, neg_ent=1, neg_rel=0, random_seed=2, incremental_setting=False):
        if neg_rel != 0:
            raise NotImplementedError("relation negatives (neg_rel) are not on the PuTransE hot path")
        if sampling_mode != "normal":
            raise NotImplementedError("cross sampling is not on the PuTransE hot path")
        self.lib = N.lib()
        self.in_path = in_path
        self.work_threads = threads
        self.nbatches = nbatches
        self.batch_size = batch_size
        self.bern = bern_flag
        self.filter = filter_flag
        self.negative_ent = neg_ent
        self.negative_rel = neg_rel
        self.sampling_mode = sampling_mode
        self.random_seed = random_seed
        self.incremental_setting = incremental_setting
        self._dev = None  # (index epoch token, dict of device tensors)
        self.read()

    # reference TrainDataLoader.py:117-136
    def read(self):
        self.lib.setInPath(self.in_path.encode())
        self.lib.setBern(self.bern)
        self.lib.setWorkThreads(self.work_threads)
        self.lib.setRandomSeed(self.random_seed)
        self.lib.randReset()
        if self.incremental_setting:     # reference :129: the training list starts empty and is evolved per snapshot
            self.lib.pk_incremental_reset()
            return
        self.lib.pk_incremental_reset()
        self.lib.importTrainFiles()
        self.relTotal = self.lib.getRelationTotal()
        self.entTotal = self.lib.getEntityTotal()
        self.tripleTotal = self.lib.getTrainTotal()
        if self.tripleTotal == 0:
            raise N.NativeError("importTrainFiles: %s" % N.last_error())
        if self.batch_size is None:
            self.batch_size = self.tripleTotal // self.nbatches
        if self.nbatches is None:
            self.nbatches = self.tripleTotal // self.batch_size
        self.update_batch_arrays()

    def update_batch_arrays(self):
        self.batch_seq_size = self.batch_size * (1 + self.negative_ent + self.negative_rel)
        self.batch_h = np.zeros(self.batch_seq_size, dtype=np.int64)
        self.batch_t = np.zeros(self.batch_seq_size, dtype=np.int64)
        self.batch_r = np.zeros(self.batch_seq_size, dtype=np.int64)
        self.batch_y = np.zeros(self.batch_seq_size, dtype=np.float32)
        self.batch_h_addr, self.batch_t_addr = N.addr(self.batch_h), N.addr(self.batch_t)
        self.batch_r_addr, self.batch_y_addr = N.addr(self.batch_r), N.addr(self.batch_y)

    # ---- universe hooks (reference TrainDataLoader.py:152-175)
    def swap_helpers(self):
        self.lib.swapHelpers()
        self._dev = None

    def reset_universe(self):
        self.lib.resetUniverse()
        self._dev = None
        self.set_nbatches(self.lib.getTrainTotal(), self.nbatches)

    def get_universe_mappings(self):
        ne, nr = self.lib.getEntityTotalUniverse(), self.lib.getRelationTotalUniverse()
        ent, rel = np.zeros(ne, dtype=np.int64), np.zeros(nr, dtype=np.int64)
        self.lib.getEntityRemapping(N.addr(ent))
        self.lib.getRelationRemapping(N.addr(rel))
        return ent, rel

    def compile_universe_dataset(self, triple_constraint, balance_param):
        self.lib.getParallelUniverse(triple_constraint, balance_param)
        if self.lib.getTrainTotalUniverse() == 0:
            raise N.NativeError("getParallelUniverse: %s" % N.last_error())
        self.set_nbatches(self.lib.getTrainTotalUniverse(), self.nbatches)

    # ---- host-visible batch, drawn on the device
    def sampling(self):
        N.require_cuda()
        self.lib.sampling(self.batch_h_addr, self.batch_t_addr, self.batch_r_addr, self.batch_y_addr, self.batch_size,
                          self.negative_ent, self.negative_rel, 0, self.filter, 0, 0)
        return {"batch_h": self.batch_h, "batch_t": self.batch_t, "batch_r": self.batch_r, "batch_y": self.batch_y,
                "mode": "normal"}

    # ---- device-resident view of the current id space, for Trainer.run
    def device_sampler(self, device):
        import torch
        if self._dev is not None and self._dev["device"] == device:
            return self._dev
        nT, nE, nR = self.lib.getTrainTotal(), self.lib.getEntityTotal(), self.lib.getRelationTotal()
        by_head, by_tail = np.zeros((nT, 3), dtype=np.int32), np.zeros((nT, 3), dtype=np.int32)
        lm, rm = np.zeros(nR, dtype=np.float32), np.zeros(nR, dtype=np.float32)
        N.check(self.lib.pk_train_index(N.addr(by_head), N.addr(by_tail), N.addr(lm), N.addr(rm)), "pk_train_index")
        lcg = np.zeros(max(self.work_threads, 1), dtype=np.uint64)
        N.check(self.lib.pk_get_lcg(N.addr(lcg), int(lcg.shape[0])), "pk_get_lcg")
        # first record of every head / tail: filtered corruption then searches one entity's records only
        head_off = np.searchsorted(by_head[:, 0], np.arange(nE + 1)).astype(np.int64)
        tail_off = np.searchsorted(by_tail[:, 2], np.arange(nE + 1)).astype(np.int64)
        dev = {"device": device, "n_tri": nT, "n_ent": nE, "n_rel": nR,
               "head_off": torch.from_numpy(head_off).to(device), "tail_off": torch.from_numpy(tail_off).to(device),
               "by_head": torch.from_numpy(by_head).to(device), "by_tail": torch.from_numpy(by_tail).to(device),
               "left_mean": torch.from_numpy(lm).to(device), "right_mean": torch.from_numpy(rm).to(device),
               "lcg": torch.from_numpy(lcg.view(np.int64)).to(device)}
        s = N.Sampler()
        s.by_head, s.by_tail = dev["by_head"].data_ptr(), dev["by_tail"].data_ptr()
        s.left_mean, s.right_mean = dev["left_mean"].data_ptr(), dev["right_mean"].data_ptr()
        s.lcg = dev["lcg"].data_ptr()
        s.n_tri, s.n_ent, s.n_rel = nT, nE, nR
        s.head_off, s.tail_off = dev["head_off"].data_ptr(), dev["tail_off"].data_ptr()
        dev["struct"] = s
        self._dev = dev
        return dev

    def sync_lcg_from_device(self):
        """After device-side sampling, hand the advanced stream states back to the library so that a
        later host-visible ``sampling()`` continues the same sequence the reference would."""
        if self._dev is not None:
            lcg = self._dev["lcg"].cpu().numpy().view(np.uint64).copy()
            N.check(self.lib.pk_set_lcg(N.addr(lcg), int(lcg.shape[0])), "pk_set_lcg")

    # ---- setters/getters of the reference (TrainDataLoader.py:278-326)
    def set_work_threads(self, work_threads):
        self.work_threads = work_threads

    def set_in_path(self, in_path):
        self.in_path = in_path

    def set_nbatches(self, triple_total, nbatches):
        self.nbatches = nbatches
        self.batch_size = triple_total // nbatches
        self.update_batch_arrays()

    def set_batch_size(self, triple_total, batch_size):
        self.nbatches = triple_total // batch_size
        self.batch_size = batch_size
        self.update_batch_arrays()

    def set_ent_neg_rate(self, rate):
        self.negative_ent = rate

    def set_rel_neg_rate(self, rate):
        if rate != 0:
            raise NotImplementedError("relation negatives are not on the PuTransE hot path")

    def set_bern_flag(self, bern):
        self.bern = bern
        self.lib.setBern(bern)

    def set_filter_flag(self, filter):
        self.filter = filter

    def get_batch_size(self):
        return self.batch_size

    def get_ent_tot(self):
        return self.entTotal

    def get_rel_tot(self):
        return self.relTotal

    def get_triple_tot(self):
        return self.tripleTotal

    def __iter__(self):
        return TrainDataSampler(self.nbatches, self.sampling)

    def __len__(self):
        return self.nbatches
