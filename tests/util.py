"""Shared helpers of the test-suite: golden fixtures, dataset materialisation, synthetic graphs."""
import os

import numpy as np

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLDEN = os.path.join(REPO, "tests", "golden")


class Golden(object):
    def __init__(self):
        self._c = {}

    def __getitem__(self, name):
        if name not in self._c:
            self._c[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
        return self._c[name]


def write_dataset(path, train, valid, test, n_ent, n_rel):
    """Header-less `h t r` files of this fork (reference openke/base/Reader.h:176-197)."""
    os.makedirs(path, exist_ok=True)
    for name, arr in (("train2id.txt", train), ("valid2id.txt", valid), ("test2id.txt", test)):
        np.savetxt(os.path.join(path, name), np.asarray(arr, dtype=np.int64), fmt="%d")
    with open(os.path.join(path, "entity2id.txt"), "w") as f:
        f.write("".join("e%d\t%d\n" % (i, i) for i in range(n_ent)))
    with open(os.path.join(path, "relation2id.txt"), "w") as f:
        f.write("".join("r%d\t%d\n" % (i, i) for i in range(n_rel)))
    return path if path.endswith("/") else path + "/"


def materialize_wn18(path):
    g = np.load(os.path.join(GOLDEN, "wn18.npz"))
    return write_dataset(path, g["train"], g["valid"], g["test"], int(g["n_ent"]), int(g["n_rel"]))


def synthetic_graph(n_ent, n_rel, n_train, n_eval, seed):
    """Random graph with skewed relation frequencies and entity degrees; columns h, t, r; train,
    valid and test are disjoint and duplicate-free."""
    rng = np.random.default_rng(seed)
    total = n_train + 2 * n_eval
    seen, rows = set(), []
    pr = 1.0 / np.arange(1, n_rel + 1)
    pr /= pr.sum()
    pe = 1.0 / np.sqrt(np.arange(1, n_ent + 1))
    pe /= pe.sum()
    while len(rows) < total:
        m = (total - len(rows)) * 2
        h, t = rng.choice(n_ent, m, p=pe), rng.choice(n_ent, m, p=pe)
        r = rng.choice(n_rel, m, p=pr)
        for a, b, c in zip(h.tolist(), t.tolist(), r.tolist()):
            if a != b and (a, b, c) not in seen:
                seen.add((a, b, c))
                rows.append((a, b, c))
                if len(rows) == total:
                    break
    arr = np.array(rows, dtype=np.int64)
    # every entity and relation must occur in train so that universes can reach them
    return arr[:n_train], arr[n_train:n_train + n_eval], arr[n_train + n_eval:]


def native():
    from openke import _native
    return _native
