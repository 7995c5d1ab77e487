"""Synthetic knowledge graphs of named shapes: the reference's FB15K training
file and the 1 M-entity graph are not available, so bench.py, the parity tests and the golden-vector
script all build them here, identically (seeded numpy PCG64 double stream + integer arithmetic only).

  fb15k_shape()  E = 14 951, R = 1 345, 483 142 train / 50 000 valid / 59 071 test triples
  s1_shape()     E = 1 000 000, R = 1 000, 10 000 000 train triples (+ 5 000 / 5 000)

Relation frequencies are Zipf(1.0); entity endpoint frequencies follow rank^-1 (degree power law with
alpha ~ 2) under a seeded permutation of the ids; triples are distinct, h != t, and the three splits are
disjoint.  Columns are (h, t, r) — the file order of this fork's header-less format
(reference openke/base/Reader.h:176-197).
"""
import os

import numpy as np


def _zipf_cdf(n, s):
    p = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    c = np.cumsum(p)
    return c / c[-1]


def power_law_graph(n_ent, n_rel, n_total, seed, ent_exponent=1.0, rel_exponent=1.0):
    """n_total distinct (h, t, r) rows in generation order."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ent_perm = rng.permutation(n_ent)
    rel_perm = rng.permutation(n_rel)
    ecdf, rcdf = _zipf_cdf(n_ent, ent_exponent), _zipf_cdf(n_rel, rel_exponent)
    keys = np.zeros(0, dtype=np.int64)
    while keys.shape[0] < n_total:
        m = int((n_total - keys.shape[0]) * 1.3) + 1024
        u = rng.random((3, m))
        h = ent_perm[np.minimum(np.searchsorted(ecdf, u[0]), n_ent - 1)]
        t = ent_perm[np.minimum(np.searchsorted(ecdf, u[1]), n_ent - 1)]
        r = rel_perm[np.minimum(np.searchsorted(rcdf, u[2]), n_rel - 1)]
        ok = h != t
        k = (h[ok].astype(np.int64) * n_rel + r[ok]) * n_ent + t[ok]
        keys = np.concatenate([keys, k])
        _, first = np.unique(keys, return_index=True)     # distinct, first occurrence kept, generation order
        keys = keys[np.sort(first)]
    keys = keys[:n_total]
    t = keys % n_ent
    r = (keys // n_ent) % n_rel
    h = keys // (n_ent * n_rel)
    return np.stack([h, t, r], axis=1).astype(np.int64)


def fb15k_shape(seed=1234, n_train=483142, n_valid=50000, n_test=59071, n_ent=14951, n_rel=1345):
    g = power_law_graph(n_ent, n_rel, n_train + n_valid + n_test, seed)
    return g[:n_train], g[n_train:n_train + n_valid], g[n_train + n_valid:], n_ent, n_rel


def s1_shape(seed=1234, n_train=10_000_000, n_eval=5000, n_ent=1_000_000, n_rel=1000):
    g = power_law_graph(n_ent, n_rel, n_train + 2 * n_eval, seed)
    return g[:n_train], g[n_train:n_train + n_eval], g[n_train + n_eval:], n_ent, n_rel


def _write_ids(path, arr):
    """`h t r` lines, fast (np.savetxt takes minutes for 10 M rows)."""
    a = np.ascontiguousarray(arr, dtype=np.int64)
    try:
        import pandas as pd
        pd.DataFrame(a).to_csv(path, sep=" ", header=False, index=False)
    except Exception:
        np.savetxt(path, a, fmt="%d")


def write_dataset(path, train, valid, test, n_ent, n_rel):
    os.makedirs(path, exist_ok=True)
    for name, arr in (("train2id.txt", train), ("valid2id.txt", valid), ("test2id.txt", test)):
        _write_ids(os.path.join(path, name), arr)
    with open(os.path.join(path, "entity2id.txt"), "w") as f:
        f.write("".join("e%d\t%d\n" % (i, i) for i in range(n_ent)))
    with open(os.path.join(path, "relation2id.txt"), "w") as f:
        f.write("".join("r%d\t%d\n" % (i, i) for i in range(n_rel)))
    return path if path.endswith("/") else path + "/"


def checksum(arr):
    """Order-sensitive 64-bit checksum of an id array (pins the generator across machines)."""
    a = np.ascontiguousarray(arr, dtype=np.uint64).ravel()
    w = (np.arange(a.shape[0], dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    return int((a * w).sum(dtype=np.uint64))


def incremental_dataset(path, seed=7, n_ent=3000, n_rel=24, n_snapshots=3, n_first=30000, n_add=6000, n_del=4000, n_eval=300):
    """A small evolving graph in the layout of the reference's WikidataEvolve benchmark
    (benchmarks/Wikidata/WikidataEvolve/incremental; reference openke/base/Incremental.h:248-321,891-924):

        <path>/incremental/{entity2id,relation2id}.txt            the global id space
        <path>/incremental/<s>/train-op2id.txt                    "h t r +|-" operations that turn snapshot s-1 into s
        <path>/incremental/<s>/{valid,test}2id.txt                evaluation triples of snapshot s (not in train)
        <path>/incremental/<s>/global_triple2id.txt               train + valid + test of snapshot s
        <path>/incremental/<s>/triple_classification_prepared_test_examples.txt, tc_negative_deleted_test_triples.txt

    Later snapshots delete triples (whole relations among them, so that the set of currently contained relations
    shrinks and re-grows), re-insert some deleted ones, insert one triple twice and delete one that never existed —
    the cases the reference's replay loop distinguishes.  Returns per snapshot the training list as a multiset."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pool = power_law_graph(n_ent, n_rel, n_first + n_snapshots * (n_add + 4 * n_eval), seed, ent_exponent=0.8, rel_exponent=0.9)
    root = os.path.join(path, "incremental")
    os.makedirs(root, exist_ok=True)
    for name, n in (("entity2id.txt", n_ent), ("relation2id.txt", n_rel)):
        for where in (root, path):
            with open(os.path.join(where, name), "w") as f:
                f.write("".join("x%d\t%d\n" % (i, i) for i in range(n)))
    cursor = 0
    count = {}          # (h, t, r) -> multiplicity: the training list as the reference keeps it (duplicates stay)
    deleted_pool = []
    states = []
    dead_rel = None

    def current():
        return [k for k, c in count.items() for _ in range(c)]

    for s in range(1, n_snapshots + 1):
        ops = []
        if s == 1:
            new = pool[cursor:cursor + n_first]
            cursor += n_first
            ops += [(int(h), int(t), int(r), "+") for h, t, r in new]
        else:
            cur = np.array(current(), dtype=np.int64)
            # drop three whole relations (the rarest present ones) and a random share of the rest
            rels, counts = np.unique(cur[:, 2], return_counts=True)
            gone = rels[np.argsort(counts, kind="stable")[:3]]
            whole = np.isin(cur[:, 2], gone)
            first = cur[whole]            # replayed FIRST: the two relations vanish (and come back later in the snapshot, at the
                                          # end of the reference's array of currently contained relations)
            dels = cur[~whole & (rng.random(cur.shape[0]) < n_del / max(1, cur.shape[0]))]
            new = pool[cursor:cursor + n_add]
            cursor += n_add
            back = np.array(deleted_pool, dtype=np.int64).reshape(-1, 3)
            back = back[rng.random(back.shape[0]) < 0.3] if back.shape[0] else back
            mixed = [(int(h), int(t), int(r), "-") for h, t, r in dels] + [(int(h), int(t), int(r), "+") for h, t, r in new] + \
                    [(int(h), int(t), int(r), "+") for h, t, r in back]
            order = rng.permutation(len(mixed))
            ops = [(int(h), int(t), int(r), "-") for h, t, r in first] + [mixed[i] for i in order]
            # one relation leaves for good in snapshot 2 (the reference aborts with "out of memory" when its array of
            # deleted relations becomes EMPTY again: realloc(…, 0) in Utilities.h:60-71 via Incremental.h:169-180), the
            # other vanished relations come back later in the same snapshot
            if dead_rel is None:
                dead_rel = int(gone[0])
            ops = [o for o in ops if not (o[2] == dead_rel and o[3] == "+")]
            h0, t0, r0 = (int(x) for x in new[0])
            ops.append((h0, t0, r0, "+"))                       # the same triple a second time: a duplicate record
            ops.append((n_ent - 1, n_ent - 2, int(r0), "-"))    # never existed: reported and skipped
        os.makedirs(os.path.join(root, str(s)), exist_ok=True)
        with open(os.path.join(root, str(s), "train-op2id.txt"), "w") as f:
            f.write("".join("%d %d %d %s\n" % o for o in ops))
        for h, t, r, op in ops:     # the reference's semantics: append / remove the first equal record
            k = (h, t, r)
            if op == "+":
                count[k] = count.get(k, 0) + 1
            elif count.get(k, 0) > 0:
                count[k] -= 1
                if count[k] == 0:
                    del count[k]
                deleted_pool.append(k)
        train = np.array(current(), dtype=np.int64)
        ev = pool[cursor:cursor + 2 * n_eval]
        cursor += 2 * n_eval
        present = np.unique(train[:, :2])
        rel_present = np.unique(train[:, 2])
        ev = ev[np.isin(ev[:, 0], present) & np.isin(ev[:, 1], present) & np.isin(ev[:, 2], rel_present)]
        va, te = ev[:len(ev) // 2], ev[len(ev) // 2:]
        _write_ids(os.path.join(root, str(s), "valid2id.txt"), va)
        _write_ids(os.path.join(root, str(s), "test2id.txt"), te)
        _write_ids(os.path.join(root, str(s), "global_triple2id.txt"), np.concatenate([train, va, te]))
        neg = te.copy()
        neg[:, 1] = present[rng.integers(0, present.shape[0], neg.shape[0])]
        with open(os.path.join(root, str(s), "triple_classification_prepared_test_examples.txt"), "w") as f:
            f.write("".join("%d %d %d 1\n" % tuple(x) for x in te.tolist()))
            f.write("".join("%d %d %d 0\n" % tuple(x) for x in neg.tolist()))
        if s > 1:
            with open(os.path.join(root, str(s), "tc_negative_deleted_test_triples.txt"), "w") as f:
                f.write("".join("%d %d %d 0\n" % x for x in deleted_pool[-200:]))
        states.append(sorted(current()))
    return states
