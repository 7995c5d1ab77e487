"""Training loader of the incremental setting: the training list is not read from train2id.txt but EVOLVED,
snapshot by snapshot, by replaying ``<in_path>/incremental/<s>/train-op2id.txt`` ("h t r +|-" per line).
Interface of the reference's ``openke/data/IncrementalTrainDataLoader.py:28-81``; the replay and the rebuild of every
training index happen in the library (``evolveTrainList``, reference openke/base/Incremental.h:798-846)."""
import os

from .. import _native as N
from .TrainDataLoader import TrainDataLoader


class IncrementalTrainDataLoader(TrainDataLoader):
    def __init__(self, in_path="./benchmarks/Wikidata/datasets/incremental", batch_size=None, nbatches=None, threads=8,
                 sampling_mode="normal", bern_flag=0, filter_flag=1, neg_ent=1, neg_rel=0, random_seed=2,
                 incremental_setting=False, num_snapshots=None):
        super().__init__(in_path=in_path, batch_size=batch_size, nbatches=nbatches, threads=threads,
                         sampling_mode=sampling_mode, bern_flag=bern_flag, filter_flag=filter_flag, neg_ent=neg_ent,
                         neg_rel=neg_rel, random_seed=random_seed, incremental_setting=incremental_setting)
        self.num_snapshots = num_snapshots
        self.initialize_incremental_loading()
        self.deleted_triple_set = set()

    def initialize_incremental_loading(self):
        """Constant along all snapshots: the global id space (reference :50-57)."""
        self.lib.activateIncrementalSetting()
        self.lib.readGlobalNumEntities()
        self.lib.readGlobalNumRelations()
        self.relTotal = self.lib.getRelationTotal()
        self.entTotal = self.lib.getEntityTotal()
        if self.relTotal == 0 or self.entTotal == 0:
            raise N.NativeError("readGlobalNumEntities/Relations: %s" % N.last_error())
        self.lib.setNumSnapshots(self.num_snapshots or 0)

    def load_snapshot(self, snapshot_idx):
        """Reference :59-69: replay the snapshot's operations, then size the batches for the evolved list."""
        self.lib.initializeTrainingOperations(int(snapshot_idx))
        self.lib.evolveTrainList()
        self._dev = None
        self.tripleTotal = self.lib.getTrainTotal()
        if self.tripleTotal == 0:
            raise N.NativeError("evolveTrainList: %s" % N.last_error())
        self.batch_size = self.tripleTotal // self.nbatches
        self.nbatches = self.tripleTotal // self.batch_size
        self.update_batch_arrays()
        self.track_deleted_triples(snapshot_idx)

    def track_deleted_triples(self, snapshot_idx):
        """Reference :71-81: triples deleted so far and not re-inserted, as (head, tail, relation) STRINGS in file
        order of the columns — what the 'deprecate' strategy looks universes up with."""
        path = os.path.join(self.in_path, "incremental", str(snapshot_idx), "train-op2id.txt")
        with open(path, mode="rt", encoding="utf-8") as f:
            for line in f:
                head, tail, rel, op_type = line.split()
                triple = (head, tail, rel)
                if op_type == "-":
                    self.deleted_triple_set.add(triple)
                elif op_type == "+":
                    self.deleted_triple_set.discard(triple)
