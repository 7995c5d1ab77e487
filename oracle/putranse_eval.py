"""ORACLE — test infrastructure only.  PuTransE global energy estimation and ranking semantics
restated with numpy (reference openke/config/Parallel_Universe_Config.py:446-465,516-642 and
openke/base/Test.h:118-238).

energy(key, e) = min over universes u containing the key's fixed entity and relation of E_u(e)
(+inf if none).  Ranks: strict `<` against the true entity's energy; +inf truth => last rank.
Parity status: PINNED by tests/test_oracle.py against tests/golden/putranse_wn18.npz (per-triple
ranks produced by the unmodified reference's eval_universes + testHead/testTail).
"""
import numpy as np

from .model_math import TorchOracle


def universe_energies(spaces, n_ent_global, fixed, rel, side, model="transe", p_norm=1):
    """spaces: list of dict(tables=..., ent_remap, rel_remap).  Returns float32 [n_ent_global]."""
    import torch
    out = np.full(n_ent_global, np.inf, dtype=np.float32)
    for sp in spaces:
        er, rr = sp["ent_remap"], sp["rel_remap"]
        fl, rl = np.nonzero(er == fixed)[0], np.nonzero(rr == rel)[0]
        if fl.size == 0 or rl.size == 0:
            continue
        o = sp.setdefault("_oracle", TorchOracle(model, sp["tables"], p_norm=p_norm))
        cand = np.arange(er.shape[0])
        with torch.no_grad():
            if side == 0:   # head batch: all local entities as heads, fixed tail
                s = o.score(cand, fl[:1], rl[:1], "head_batch").numpy()
            else:
                s = o.score(fl[:1], cand, rl[:1], "tail_batch").numpy()
        out[er] = np.minimum(out[er], s)
    return out


def tuple_score(spaces, fixed, rel, side, p_norm=1):
    """missing_embedding_handling='null_vector' (reference Parallel_Universe_Config.py:378-388,494-514):
    min over universes holding (fixed, rel) of _calc() with the missing side replaced by a zero vector,
    on the RAW embedding rows (the reference does not apply the TransH/TransD projection here)."""
    import torch
    import torch.nn.functional as F
    best = np.float32(np.inf)
    for sp in spaces:
        er, rr = sp["ent_remap"], sp["rel_remap"]
        fl, rl = np.nonzero(er == fixed)[0], np.nonzero(rr == rel)[0]
        if fl.size == 0 or rl.size == 0:
            continue
        e = F.normalize(torch.from_numpy(np.asarray(sp["tables"]["ent_embeddings"][fl[0]], dtype=np.float32)), 2, -1)
        r = F.normalize(torch.from_numpy(np.asarray(sp["tables"]["rel_embeddings"][rl[0]], dtype=np.float32)), 2, -1)
        zero = torch.zeros(1)
        s = zero + (r - e) if side == 0 else (e + r) - zero
        best = min(best, np.float32(torch.norm(s, p_norm, -1).item()))
    return best


def fill_missing(energy, tuple_sc):
    """reference global_energy_estimation :634-640"""
    out = energy.copy()
    if np.isfinite(tuple_sc):
        out[np.isinf(out)] = tuple_sc
    return out


def rank_from_energy(energy, truth, known):
    """(raw, filtered) 0-based ranks; `known` = known-true candidates other than the truth."""
    E = energy.shape[0]
    tgt = energy[truth]
    if np.isinf(tgt):
        return E, E - len(known)
    better = energy < tgt
    raw = int(better.sum())
    return raw, raw - int(better[np.asarray(known, dtype=np.int64)].sum()) if len(known) else raw
