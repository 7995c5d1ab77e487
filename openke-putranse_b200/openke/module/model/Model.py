"""Base class of the translational models: owns the embedding tables, hands their device pointers
to the CUDA kernels, and scores index batches through ``pk_score_batch``.

Interface of reference ``openke/module/model/Model.py:6-16`` plus what the three model files of the
reference repeat verbatim (``forward``/``predict``, TransE.py:62-74,88-94).  The arithmetic itself
(normalise, project, translate, p-norm) lives in csrc/rank.cu and csrc/kge_device.cuh.
"""
import ctypes

import torch
import torch.nn as nn

from ..BaseModule import BaseModule
from ... import _native as N


class Model(BaseModule):
    #: (PK_* model id, entity tables, relation tables) — filled by subclasses
    _pk_model = None
    _ent_tables = ("ent_embeddings",)
    _rel_tables = ("rel_embeddings",)

    def __init__(self, ent_tot, rel_tot):
        super().__init__()
        self.ent_tot = ent_tot
        self.rel_tot = rel_tot

    # ---- shared construction (reference TransE.py:17-43 and twins)
    @staticmethod
    def _fill_tables(specs, margin, epsilon, views, generator=None):
        """Initial values of the tables in `specs` order, written into the [rows, dim] float32 CPU
        tensors `views[attr]`, consuming the torch CPU generator exactly like the reference's
        constructor: every nn.Embedding draws a normal_() at creation (torch's default init), then
        all tables are re-drawn with xavier_uniform_ (or uniform_(+-(margin+epsilon)/dim)).
        `generator`: a torch.Generator seeded like torch.manual_seed(seed) yields the same stream as
        the global one, which lets several universes be initialised on different threads."""
        for attr, _, _ in specs:
            views[attr].normal_(generator=generator)
        for attr, _, dim in specs:
            if margin is None or epsilon is None:
                nn.init.xavier_uniform_(views[attr], generator=generator)
            else:
                rng = (margin + epsilon) / dim
                nn.init.uniform_(tensor=views[attr], a=-rng, b=rng, generator=generator)

    @classmethod
    def table_specs(cls, ent_tot, rel_tot, **param):
        """[(attr, rows, dim)] in the reference's creation order (it fixes the RNG stream of the init)."""
        raise NotImplementedError

    @classmethod
    def initial_tables_into(cls, ent_tot, rel_tot, views, generator=None, **param):
        """What ``cls(ent_tot, rel_tot, **param)`` would hold right after construction, written
        straight into `views` (e.g. slices of one packed pinned buffer) without building a module."""
        cls._fill_tables(cls.table_specs(ent_tot, rel_tot, **param), param.get("margin"), param.get("epsilon"), views,
                         generator=generator)

    def _init_tables(self, specs, margin, epsilon, ranges):
        """specs: [(attr, rows, dim)] in the reference's creation order; xavier unless margin+epsilon."""
        for attr, rows, dim in specs:
            emb = nn.Embedding(rows, dim, _weight=torch.empty(rows, dim))   # no draw here: _fill_tables replays it
            setattr(self, attr, emb)
        self._fill_tables(specs, margin, epsilon, {attr: getattr(self, attr).weight.data for attr, _, _ in specs})
        if not (margin is None or epsilon is None):
            for name, value in ranges.items():
                setattr(self, name, nn.Parameter(torch.Tensor([value]), requires_grad=False))
        if margin is not None:
            self.margin = nn.Parameter(torch.Tensor([margin]), requires_grad=False)
            self.margin_flag = True
        else:
            self.margin_flag = False

    # ---- native views
    @property
    def dim_native(self):
        return self.ent_embeddings.weight.shape[1]

    def native_cfg(self, opt=N.PK_SGD, neg_ent=1, bern=0, filt=0, work_threads=8):
        return N.ModelCfg(model=self._pk_model, dim=self.dim_native, p_norm=int(self.p_norm),
                          norm_flag=1 if self.norm_flag else 0, opt=opt, neg_ent=neg_ent, bern=bern, filter=filt,
                          work_threads=work_threads, reserved=0)

    def native_tables(self, states=None):
        """pk_tables over the live ``nn.Embedding.weight`` storages (updated in place by the kernels)."""
        t = N.Tables()
        for i in range(2):
            t.ent[i] = t.rel[i] = t.ent_state[i] = t.rel_state[i] = None
        for i, name in enumerate(self._ent_tables):
            w = getattr(self, name).weight
            self._check_weight(w, name)
            t.ent[i] = w.data_ptr()
            if states is not None:
                t.ent_state[i] = states[name].data_ptr()
        for i, name in enumerate(self._rel_tables):
            w = getattr(self, name).weight
            self._check_weight(w, name)
            t.rel[i] = w.data_ptr()
            if states is not None:
                t.rel_state[i] = states[name].data_ptr()
        t.n_ent, t.n_rel = self.ent_tot, self.rel_tot
        return t

    @staticmethod
    def _check_weight(w, name):
        if not w.is_cuda:
            raise N.NativeError("%s is on %s: the PuTransE hot path runs on CUDA only (call model.cuda())" % (name, w.device))
        if w.dtype != torch.float32 or not w.is_contiguous():
            raise N.NativeError("%s must be a contiguous float32 table" % name)

    def table_names(self):
        return tuple(self._ent_tables) + tuple(self._rel_tables)

    # ---- scoring
    def _score(self, batch_h, batch_t, batch_r, mode):
        dev = self.ent_embeddings.weight.device
        h = torch.as_tensor(batch_h, dtype=torch.int64, device=dev).reshape(-1).contiguous()
        t = torch.as_tensor(batch_t, dtype=torch.int64, device=dev).reshape(-1).contiguous()
        r = torch.as_tensor(batch_r, dtype=torch.int64, device=dev).reshape(-1).contiguous()
        n = max(h.numel(), t.numel(), r.numel())
        out = torch.empty(n, dtype=torch.float32, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        cfg, tab = self.native_cfg(), self.native_tables()
        N.check(N.lib().pk_score_batch(ctypes.byref(cfg), ctypes.byref(tab), h.data_ptr(), h.numel(), t.data_ptr(),
                                       t.numel(), r.data_ptr(), r.numel(), 1 if mode == "head_batch" else 0,
                                       out.data_ptr(), bad.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                "pk_score_batch")
        self._last_bad = bad
        return out

    def forward(self, data):
        score = self._score(data["batch_h"], data["batch_t"], data["batch_r"], data["mode"])
        return self.margin - score if self.margin_flag else score

    def predict(self, data):
        score = self.forward(data)
        if self.margin_flag:
            score = self.margin - score
        res = score.cpu().numpy()
        if int(self._last_bad.item()) != 0:
            raise IndexError("predict: entity or relation id out of range")
        return res

    def _calc(self, h, t, r, mode):
        """Energy of explicit embedding rows (reference TransE.py:46-60), for callers that bring their
        own vectors.  Plain torch on whatever device the rows live on; not used by the train/rank path."""
        import torch.nn.functional as F
        if self.norm_flag:
            h, r, t = F.normalize(h, 2, -1), F.normalize(r, 2, -1), F.normalize(t, 2, -1)
        if mode != "normal":
            h = h.view(-1, r.shape[0], h.shape[-1])
            t = t.view(-1, r.shape[0], t.shape[-1])
            r = r.view(-1, r.shape[0], r.shape[-1])
        score = h + (r - t) if mode == "head_batch" else (h + r) - t
        return torch.norm(score, self.p_norm, -1).flatten()
