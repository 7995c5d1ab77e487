"""Evaluation loader of the incremental setting (interface of the reference's
``openke/data/IncrementalTestDataLoader.py:28-117``): per snapshot, the evaluation list comes from
``incremental/<s>/{test,valid}2id.txt`` (file order, reference openke/base/Incremental.h:248-296) and the filter
set and candidate entities from ``incremental/<s>/global_triple2id.txt`` (``evolveTripleList2`` →
``loadSnapshotTriples``, Incremental.h:891-924).

Candidates of a link-prediction query are the entities the snapshot currently contains.  The reference's C loops
for this case fill only every other candidate slot (stray ``i++``, openke/base/Test.h:49-59,84-95) and overwrite the
first entries of the entity array with relation ids (Incremental.h:882-887); this loader implements what those
loops evidently mean — truth in slot 0, then every other contained entity ascending (reference defects)."""
import numpy as np

from .. import _native as N
from .TestDataLoader import TestDataLoader, TestDataSampler


class IncrementalTestDataLoader(TestDataLoader):
    def __init__(self, in_path="./benchmarks/Wikidata/datasets/incremental", sampling_mode="link", random_seed=4,
                 mode="test", setting="static", num_snapshots=None):
        if mode not in ("test", "valid"):
            raise ValueError("mode must be 'test' or 'valid'")
        self.lib = N.lib()
        self.setting, self.mode, self.load_all_triples = setting, mode, False
        self.in_path, self.sampling_mode, self.random_seed = in_path, sampling_mode, random_seed
        self._arrays = None
        self.num_snapshots = num_snapshots
        self.testTotal = self.validTotal = 0
        self.currently_contained_entTotal = 0
        self.contained_entities = np.zeros(0, dtype=np.int32)
        self.set_path(in_path)
        self.lib.setRandomSeed(random_seed)        # reference TestDataLoader.read :86-89
        self.lib.randReset()
        self.initialize_incremental_loading()

    def initialize_incremental_loading(self):
        self.lib.activateIncrementalSetting()
        self.lib.readGlobalNumEntities()
        self.lib.readGlobalNumRelations()
        self.relTotal = self.lib.getRelationTotal()
        self.entTotal = self.lib.getEntityTotal()
        self.lib.setNumSnapshots(self.num_snapshots or 0)

    def evolveTripleList(self, snapshot_idx):
        raise NotImplementedError("triple-op replay for the filter list was superseded in the reference by "
                                  "evolveTripleList2 (experiments/incremental_experiment_PuTransE_on_WikidataEvolve.py:45-46)")

    def evolveTripleList2(self, snapshot_idx):
        self.set_path(self.in_path)
        self.lib.loadSnapshotTriples(int(snapshot_idx))
        self._arrays = None

    def load_snapshot(self, snapshot_idx):
        self.set_path(self.in_path)
        if self.mode == "test":
            self.lib.loadTestData(int(snapshot_idx))
            self.testTotal = self.lib.getTestTotal()
        else:
            self.lib.loadValidData(int(snapshot_idx))
            self.validTotal = self.lib.getValidTotal()
        n = self.lib.pk_incremental_list(3, None)
        self.contained_entities = np.zeros(max(int(n), 0), dtype=np.int32)
        if n > 0:
            self.lib.pk_incremental_list(3, N.addr(self.contained_entities))
        self.currently_contained_entTotal = int(self.lib.getNumCurrentlyContainedEntities())
        self._arrays = None

    def candidate_mask(self):
        """uint8 [entTotal]: 1 for the entities the snapshot currently contains."""
        mask = np.zeros(self.entTotal, dtype=np.uint8)
        mask[self.contained_entities] = 1
        return mask

    def sampling_lp(self):
        """[head batch, tail batch] of the next evaluation triple: truth first, then every other currently contained
        entity ascending."""
        tri, _ = self.eval_arrays()
        i = self._cursor
        self._cursor += 1
        h, r, t = (int(x) for x in tri[i])
        ents = self.contained_entities.astype(np.int64)
        heads = np.concatenate([[h], ents[ents != h]])
        tails = np.concatenate([[t], ents[ents != t]])
        return [{"batch_h": heads, "batch_t": np.array([t], dtype=np.int64), "batch_r": np.array([r], dtype=np.int64), "mode": "head_batch"},
                {"batch_h": np.array([h], dtype=np.int64), "batch_t": tails, "batch_r": np.array([r], dtype=np.int64), "mode": "tail_batch"}]

    def __iter__(self):
        self._cursor = 0
        return TestDataSampler(len(self), self.sampling_lp)
