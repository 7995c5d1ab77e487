"""Link-prediction driver with the interface of the reference's ``openke/config/Tester.py:17-93``.

The reference scores one [E]-long candidate batch per test triple and side, copies the scores to
the host and ranks them in C (``testHead``/``testTail``).  Here all test triples are ranked in one
device pass (``pk_rank_space``); the five returned numbers are accumulated in float32 in test order,
as the reference's C globals are (openke/base/Test.h:14-16,213-223,398-454).
"""
import ctypes

import numpy as np
import torch

from .. import _native as N


def _f32_running_sum(values):
    """acc = (float)(acc + v) for v in order, v and the sum in double: how a C `float += double` behaves."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    return np.float32(N.lib().pk_f32_running_sum(N.addr(v), int(v.shape[0])))


def link_metrics(ranks):
    """ranks int [n,4] = head raw, head filtered, tail raw, tail filtered (0-based count of better
    candidates).  Returns ((mrr, mr, hit10, hit3, hit1) filtered and head/tail-averaged — what the
    reference's getTestLink* return — and the full raw/filtered table), accumulated like the
    reference's float globals: sequentially, in test order (openke/base/Test.h:213-223,398-454)."""
    n = ranks.shape[0]
    f32 = np.float32
    table = {}
    for name, col in (("l_raw", 0), ("l_filter", 1), ("r_raw", 2), ("r_filter", 3)):
        r = ranks[:, col].astype(np.int64)
        rank_sum = _f32_running_sum((r + 1).astype(np.float64))
        reci_sum = _f32_running_sum(1.0 / (r + 1))
        table[name] = (reci_sum / f32(n), rank_sum / f32(n), f32((r < 10).sum()) / f32(n), f32((r < 3).sum()) / f32(n),
                       f32((r < 1).sum()) / f32(n))
    avg = tuple(float((f32(a) + f32(b)) / f32(2)) for a, b in zip(table["l_filter"], table["r_filter"]))
    return avg, table


class Tester(object):
    def __init__(self, model=None, data_loader=None, use_gpu=True):
        self.lib = N.lib()
        self.model = model
        self.data_loader = data_loader
        self.use_gpu = use_gpu
        self.last_ranks = None
        self.last_table = None
        if self.use_gpu and self.model is not None:
            self.model.cuda()

    def set_model(self, model):
        self.model = model

    def set_data_loader(self, data_loader):
        self.data_loader = data_loader

    def set_use_gpu(self, use_gpu):
        self.use_gpu = use_gpu
        if self.use_gpu and self.model is not None:
            self.model.cuda()

    def to_var(self, x, use_gpu):
        t = torch.from_numpy(x)
        return t.cuda() if use_gpu else t

    def test_one_step(self, data):
        return self.model.predict(data)

    def rank_all(self, loader=None, model=None):
        """int32 [n,4] raw/filtered ranks of every triple of the loader, both sides."""
        loader = loader or self.data_loader
        model = model or self.model
        if not self.use_gpu:
            raise N.NativeError("use_gpu=False: link prediction on the B200 path has no CPU implementation")
        N.require_cuda()
        model.cuda()
        dev = model.ent_embeddings.weight.device
        tri, filt = loader.eval_arrays()
        d_tri = torch.from_numpy(tri).to(dev)
        d_f = [(torch.from_numpy(o).to(dev), torch.from_numpy(np.ascontiguousarray(c) if c.size else np.zeros(1, np.int32)).to(dev))
               for o, c in filt]
        ranks = torch.zeros((tri.shape[0], 4), dtype=torch.int32, device=dev)
        cfg, tab = model.native_cfg(), model.native_tables()
        N.check(self.lib.pk_rank_space(ctypes.byref(cfg), ctypes.byref(tab), tri.shape[0], d_tri.data_ptr(),
                                       d_f[0][0].data_ptr(), d_f[0][1].data_ptr(), d_f[1][0].data_ptr(),
                                       d_f[1][1].data_ptr(), ranks.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                "pk_rank_space")
        self.gpu_launches = self.lib.pk_last_launch_count()
        return ranks.cpu().numpy()

    def run_link_prediction(self, type_constrain=False):
        if type_constrain:
            raise NotImplementedError("type-constrained ranking is not on the PuTransE hot path")
        self.data_loader.set_sampling_mode("link")
        self.last_ranks = self.rank_all()
        (mrr, mr, hit10, hit3, hit1), self.last_table = link_metrics(self.last_ranks)
        return mrr, mr, hit10, hit3, hit1

    # ---- triple classification (reference Tester.py:95-191); the loops over sorted (answer, score) pairs are
    #      evaluated as prefix sums, with the reference's tie order (np.argsort of the scores) and first-maximum rule
    @staticmethod
    def _sorted_answers(score, ans):
        order = np.argsort(score)
        return np.asarray(ans)[order], np.asarray(score)[order]

    def get_best_threshlod(self, score, ans):
        a, sc = self._sorted_answers(score, ans)
        total_all = float(len(sc))
        total_false = total_all - float(np.sum(a))
        cur = np.cumsum(a == 1).astype(np.float64)
        res = (2 * cur + total_false - np.arange(len(sc)) - 1) / total_all
        if len(sc) == 0 or not (res > 0.0).any():
            return None, 0.0
        i = int(np.argmax(res))               # first index of the maximum, as `if res_current > res_mx` keeps it
        return sc[i], float(res[i])

    def determine_classification_cross_table_values(self, res, threshold):
        ans, score = res[:, 0], res[:, 1]
        pred = score < threshold
        table = {"tp": int((pred & (ans == 1)).sum()), "fp": int((pred & (ans == 0)).sum()),
                 "tn": int((~pred & (ans == 0)).sum()), "fn": int((~pred & (ans == 1)).sum())}
        print("True Positives :{}".format(table["tp"]))
        print("True Negatives :{}".format(table["tn"]))
        print("False Positives :{}".format(table["fp"]))
        print("False Negatives :{}".format(table["fn"]))
        self.last_cross_table = table
        return table

    def run_triple_classification(self, threshlod=None, data_iterator=None):
        self.lib.initTest()
        score, ans = [], []
        if data_iterator is None:
            self.data_loader.set_sampling_mode("classification")
            data_iterator = self.data_loader
        for pos_ins, neg_ins in data_iterator:
            res_pos = np.asarray(self.test_one_step(pos_ins), dtype=np.float32).reshape(-1)
            ans += [1] * len(res_pos)
            score.append(res_pos)
            res_neg = np.asarray(self.test_one_step(neg_ins), dtype=np.float32).reshape(-1)
            ans += [0] * len(res_neg)
            score.append(res_neg)
        score = np.concatenate(score, axis=-1)
        ans = np.array(ans)
        if threshlod is None:
            threshlod, _ = self.get_best_threshlod(score, ans)
        a, sc = self._sorted_answers(score, ans)
        total_all = float(len(sc))
        total_true = float(np.sum(a))
        total_false = total_all - total_true
        # Parallel-universe scores: nothing was scored at all (reference :169-176)
        if threshlod == float("inf") and np.isinf(sc).all():
            if total_true == 0:
                return 1.0, threshlod
            if total_false == 0:
                return 0.0, threshlod
            if total_false == total_true:
                return 0.5, threshlod
        acc = 0
        if threshlod is not None:
            above = np.nonzero(sc > threshlod)[0]
            if above.size:                     # the reference's loop stops at the first score above the threshold
                i = int(above[0])
                acc = (2 * float((a[:i] == 1).sum()) + total_false - i) / total_all
            self.determine_classification_cross_table_values(np.stack([a.astype(np.float64), sc.astype(np.float64)], axis=1), threshlod)
        return acc, threshlod
