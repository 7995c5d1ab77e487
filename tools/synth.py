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
