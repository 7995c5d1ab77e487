"""Child process of tests/test_host.py::test_reference_layout_checkpoint_round_trip: the UNMODIFIED reference
package reads (a) a checkpoint it wrote itself (tests/golden/putranse_reference_layout.ckpt) and (b) the
reference-layout export this repository made from (a), and must find the same ensemble in both.

usage: _refckpt_child.py <scratch> <reference Base.so> <wn18 dir> <golden ckpt> <our export>
"""
import os
import sys

import numpy as np

scratch, ref_so, wn18, golden_ckpt, ours_ckpt = sys.argv[1:6]
REF = "/root/reference"
pkg = os.path.join(scratch, "openke")
os.makedirs(os.path.join(pkg, "release"), exist_ok=True)
for name in ("__init__.py", "base", "config", "data", "module"):
    dst = os.path.join(pkg, name)
    if not os.path.lexists(dst):
        os.symlink(os.path.join(REF, "openke", name), dst)
so = os.path.join(pkg, "release", "Base.so")
if os.path.lexists(so):
    os.remove(so)
os.symlink(ref_so, so)
sys.path.insert(0, scratch)

import torch                                                        # noqa: E402
from openke.config import Parallel_Universe_Config                  # noqa: E402  (reference classes)
from openke.data import TrainDataLoader, TestDataLoader             # noqa: E402
from openke.module.model import TransE                              # noqa: E402


def load(path):
    train = TrainDataLoader(in_path=wn18, nbatches=20, threads=8, sampling_mode="normal", bern_flag=0, filter_flag=0,
                            neg_ent=1, neg_rel=0, random_seed=123)
    test = TestDataLoader(train.in_path, "link")
    pu = Parallel_Universe_Config(training_identifier="x", train_dataloader=train, test_dataloader=test,
                                  initial_num_universes=None, min_margin=1, max_margin=4, min_lr=0.001, max_lr=0.1,
                                  min_num_epochs=50, max_num_epochs=200, const_num_epochs=2, min_triple_constraint=500,
                                  max_triple_constraint=2000, min_balance=0.25, max_balance=0.5, embedding_model=TransE,
                                  embedding_model_param={"dim": 20, "p_norm": 1, "norm_flag": 1}, checkpoint_dir="/tmp/",
                                  valid_steps=10 ** 9, save_steps=10 ** 9, training_setting="static", incremental_strategy=None)
    pu.use_gpu = False
    # reference load_parameters is a bare torch.load (:933-935), which torch >= 2.6 refuses for pickled modules
    pu.process_state_dict(torch.load(path, map_location="cpu", weights_only=False))
    return pu


a, b = load(golden_ckpt), load(ours_ckpt)
assert a.next_universe_id == b.next_universe_id == 3
assert type(b.trained_embedding_spaces[0]).__module__ == "openke.module.model.TransE"
for u in range(3):
    sa, sb = a.trained_embedding_spaces[u].state_dict(), b.trained_embedding_spaces[u].state_dict()
    assert sorted(sa) == sorted(sb), (sorted(sa), sorted(sb))
    for k in sa:
        assert torch.equal(sa[k], sb[k]), (u, k)
    assert dict(a.entity_id_mappings[u]) == dict(b.entity_id_mappings[u])
    assert dict(a.relation_id_mappings[u]) == dict(b.relation_id_mappings[u])
ents = sorted(a.entity_id_mappings[1])[:40]
for e in ents:
    assert a.entity_universes[e] == b.entity_universes[e]
for r in range(18):
    assert a.relation_universes[r] == b.relation_universes[r]
for key in ("min_margin", "max_margin", "min_lr", "max_lr", "min_triple_constraint", "max_triple_constraint", "min_balance", "max_balance"):
    assert getattr(a, key) == getattr(b, key), key
# the reference's own scoring on both: min-energy of triples held by universe 1
with torch.no_grad():
    g2l = a.entity_id_mappings[1]
    r_glob = sorted(a.relation_id_mappings[1])[0]
    checked = 0
    for h in ents[:6]:
        for t in ents[6:12]:
            ea, eb = a.predict_triple(h, r_glob, t), b.predict_triple(h, r_glob, t)
            assert float(ea) == float(eb) and np.isfinite(float(ea)), (h, r_glob, t, ea, eb)
            checked += 1
print("REFCKPT-OK", checked)
