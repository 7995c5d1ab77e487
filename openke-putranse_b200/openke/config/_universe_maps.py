"""Lazy views of the reference's per-universe Python containers.

The reference fills, universe by universe, ``trained_embedding_spaces`` (an nn.Module per universe),
``entity_id_mappings`` / ``relation_id_mappings`` (dict global id -> local id per universe) and
``entity_universes`` / ``relation_universes`` (set of universe ids per global id)
(reference openke/config/Parallel_Universe_Config.py:195-207,260-264).  Training and evaluation on
the B200 path never read them — they work on the packed chunk arrays — so building ~10^5 dict/set
entries per chunk eagerly would only sit on the end-to-end critical path.  These mappings keep the
reference's attribute names and indexing behaviour and materialise an entry the first time it is read.
"""
from collections.abc import Mapping

import numpy as np
import torch


class _ChunkBacked(Mapping):
    def __init__(self, owner):
        self._o = owner
        self._cache = {}
        self._version = -1

    def _sync(self):
        v = self._o._maps_version
        if v != self._version:
            self._cache = {}
            self._version = v
            self._rebuild()

    def _rebuild(self):
        pass

    def clear(self):
        self._cache = {}
        self._version = -1


class SpaceMap(_ChunkBacked):
    """universe id -> embedding-space module whose tables are views of the packed device tables."""

    def __init__(self, owner):
        super().__init__(owner)
        self._extra = {}

    def __setitem__(self, u, space):
        self._extra[u] = space

    def __getitem__(self, u):
        self._sync()
        if u in self._extra:
            return self._extra[u]
        if u in self._cache:
            return self._cache[u]
        ck, i = self._o._where[u]
        o = self._o
        o.synchronize()   # training is asynchronous: the tables are final only after their launch
        with torch.random.fork_rng(devices=[]):   # building the module must not disturb the caller's RNG
            space = o.embedding_model(int(ck.nE[i]), int(ck.nR[i]), **o.embedding_model_param)
        for name in space.table_names():
            off = ck.eoff if name in space._ent_tables else ck.roff
            getattr(space, name).weight = torch.nn.Parameter(ck.tables[name][off[i]:off[i + 1]], requires_grad=False)
        dev = next(iter(ck.tables.values())).device
        for p in (space.zero_const, space.pi_const):
            p.data = p.data.to(dev)
        space.eval()
        self._cache[u] = space
        return space

    def __iter__(self):
        self._sync()
        yield from sorted(set(self._o._where) | set(self._extra))

    def __len__(self):
        return len(set(self._o._where) | set(self._extra))

    def clear(self):
        super().clear()
        self._extra = {}


class LocalIdMaps(_ChunkBacked):
    """universe id -> {global id: local id}."""

    def __init__(self, owner, kind):
        super().__init__(owner)
        self._kind = kind

    def __getitem__(self, u):
        self._sync()
        if u not in self._cache:
            if u not in self._o._where:
                self._cache[u] = {}
            else:
                ck, i = self._o._where[u]
                remap, off = (ck.ent_remap, ck.eoff) if self._kind == "ent" else (ck.rel_remap, ck.roff)
                g = remap[off[i]:off[i + 1]].tolist()
                self._cache[u] = dict(zip(g, range(len(g))))
        return self._cache[u]

    def __iter__(self):
        yield from sorted(self._o._where)

    def __len__(self):
        return len(self._o._where)


class MembershipMap(_ChunkBacked):
    """global id -> set of universe ids that contain it (empty set when none does)."""

    def __init__(self, owner, kind):
        super().__init__(owner)
        self._kind = kind
        self._ids = self._univ = self._start = None

    def _rebuild(self):
        ids, univ = [], []
        for ck in self._o._chunks:
            remap, n = (ck.ent_remap, ck.nE) if self._kind == "ent" else (ck.rel_remap, ck.nR)
            ids.append(remap.astype(np.int64))
            univ.append(np.repeat(np.asarray(ck.ids, dtype=np.int64), n))
        total = self._o.ent_tot if self._kind == "ent" else self._o.rel_tot
        if ids:
            ids, univ = np.concatenate(ids), np.concatenate(univ)
            order = np.argsort(ids, kind="stable")
            self._univ = univ[order]
            self._start = np.searchsorted(ids[order], np.arange(total + 1))
        else:
            self._univ = np.zeros(0, np.int64)
            self._start = np.zeros(total + 1, np.int64)

    def __getitem__(self, g):
        self._sync()
        if g not in self._cache:
            if 0 <= g < len(self._start) - 1:
                self._cache[g] = set(self._univ[self._start[g]:self._start[g + 1]].tolist())
            else:
                self._cache[g] = set()
        return self._cache[g]

    def __iter__(self):
        self._sync()
        yield from (g for g in range(len(self._start) - 1) if self._start[g + 1] > self._start[g])

    def __len__(self):
        self._sync()
        return int((np.diff(self._start) > 0).sum())
