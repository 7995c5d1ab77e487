"""ORACLE — test infrastructure only.  ctypes wrapper of liboracle_putranse.so (oracle_native.cpp),
the plain-C++ restatement of the reference's reader / sampler / universe walk / rank counting, and
of oracle/_ref/Base.so, the UNMODIFIED reference library compiled from /root/reference."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle_putranse.so")
REF_SO = os.path.join(HERE, "_ref", "Base.so")
I, F, VP = ctypes.c_long, ctypes.c_float, ctypes.c_void_p


def _addr(a):
    return a.__array_interface__["data"][0]


class Oracle(object):
    """One independent oracle state (the restatement is instance-based; only libc rand() is global)."""

    def __init__(self, threads=8, bern=0):
        if not os.path.exists(ORACLE_SO):
            raise RuntimeError("oracle not built: run `make -C oracle`")
        L = ctypes.CDLL(ORACLE_SO)
        L.oracle_new.restype = VP
        for name, args, res in [
            ("oracle_free", [VP], None), ("oracle_set", [VP, I, I], None),
            ("oracle_import_train", [VP, VP, I, I, I], None), ("oracle_train_total", [VP], I),
            ("oracle_ent_total", [VP], I), ("oracle_rel_total", [VP], I), ("oracle_means", [VP, VP, VP], None),
            ("oracle_train_list", [VP, VP], None), ("oracle_seed", [VP, I], None), ("oracle_get_lcg", [VP, VP], None),
            ("oracle_sampling", [VP, VP, VP, VP, VP, I, I, ctypes.c_bool], None),
            ("oracle_universe", [VP, I, F], I), ("oracle_universe_export", [VP, VP, VP, VP], None),
            ("oracle_universe_ent", [VP], I), ("oracle_universe_rel", [VP], I), ("oracle_swap", [VP], None),
            ("oracle_import_test", [VP, VP, I, VP, I, VP, I], None), ("oracle_eval_list", [VP, ctypes.c_int, VP], None),
            ("oracle_rank_row", [VP, ctypes.c_int, VP, I, ctypes.c_int, I, VP], None)]:
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, res
        self.L, self.o, self.threads = L, L.oracle_new(), threads
        L.oracle_set(self.o, threads, bern)
        self._keep = []

    def __del__(self):
        try:
            self.L.oracle_free(self.o)
        except Exception:
            pass

    def import_train(self, htr, n_ent, n_rel):
        a = np.ascontiguousarray(htr, dtype=np.int64)
        self.L.oracle_import_train(self.o, _addr(a), a.shape[0], n_ent, n_rel)
        self.n_ent_global = n_ent

    def import_test(self, test, train, valid):
        t, tr, v = (np.ascontiguousarray(x, dtype=np.int64) for x in (test, train, valid))
        self.L.oracle_import_test(self.o, _addr(t), t.shape[0], _addr(tr), tr.shape[0], _addr(v), v.shape[0])
        self._n_eval = (t.shape[0], v.shape[0])

    def totals(self):
        return self.L.oracle_train_total(self.o), self.L.oracle_ent_total(self.o), self.L.oracle_rel_total(self.o)

    def means(self):
        nr = self.L.oracle_rel_total(self.o)
        l, r = np.zeros(nr, np.float32), np.zeros(nr, np.float32)
        self.L.oracle_means(self.o, _addr(l), _addr(r))
        return l, r

    def train_list(self):
        out = np.zeros((self.L.oracle_train_total(self.o), 3), np.int64)
        self.L.oracle_train_list(self.o, _addr(out))
        return out

    def seed(self, s):
        self.L.oracle_seed(self.o, s)

    def lcg(self):
        s = np.zeros(self.threads, np.uint64)
        self.L.oracle_get_lcg(self.o, _addr(s))
        return s

    def sampling(self, batch_size, neg, filter_flag):
        n = batch_size * (1 + neg)
        h, t, r = (np.zeros(n, np.int64) for _ in range(3))
        self.L.oracle_sampling(self.o, _addr(h), _addr(t), _addr(r), None, batch_size, neg, bool(filter_flag))
        return h, t, r

    def universe(self, tc, balance):
        nt = self.L.oracle_universe(self.o, tc, balance)
        ne, nr = self.L.oracle_universe_ent(self.o), self.L.oracle_universe_rel(self.o)
        g, er, rr = np.zeros((nt, 3), np.int64), np.zeros(ne, np.int64), np.zeros(nr, np.int64)
        self.L.oracle_universe_export(self.o, _addr(g), _addr(er), _addr(rr))
        return g, er, rr

    def swap(self):
        self.L.oracle_swap(self.o)

    def eval_list(self, which):
        out = np.zeros((self._n_eval[which], 3), np.int64)
        self.L.oracle_eval_list(self.o, which, _addr(out))
        return out

    def rank_row(self, which, con, index, head):
        con = np.ascontiguousarray(con, dtype=np.float32)
        out = np.zeros(2, np.int64)
        self.L.oracle_rank_row(self.o, which, _addr(con), index, 1 if head else 0, con.shape[0], _addr(out))
        return int(out[0]), int(out[1])


def candidate_row(energy_by_entity, truth):
    """Entity-indexed energies -> the reference's candidate order (slot 0 = truth, rest ascending)."""
    e = np.asarray(energy_by_entity, dtype=np.float32)
    return np.concatenate([e[truth:truth + 1], e[:truth], e[truth + 1:]])


def load_reference():
    """The unmodified reference library (oracle/_ref/Base.so), or None when it was not built."""
    if not os.path.exists(REF_SO):
        return None
    L = ctypes.CDLL(REF_SO)
    L.sampling.argtypes = [VP, VP, VP, VP] + [ctypes.c_int64] * 7
    L.getParallelUniverse.argtypes = [ctypes.c_int64, ctypes.c_float]
    L.getEntityRemapping.argtypes = [VP]
    L.getRelationRemapping.argtypes = [VP]
    L.testHead.argtypes = [VP, ctypes.c_int64, ctypes.c_int64]
    L.testTail.argtypes = [VP, ctypes.c_int64, ctypes.c_int64]
    L.setRandomSeed.argtypes = [ctypes.c_int64]
    L.setWorkThreads.argtypes = [ctypes.c_int64]
    L.setBern.argtypes = [ctypes.c_int64]
    return L
