"""ORACLE — test infrastructure only.  Floating-point half of the reference's hot path.

The reference's arithmetic is PyTorch itself (third-party; requirements.txt pins torch==1.5.0, this
image has 2.11): nn.Embedding gather, F.normalize, broadcasting add/sub, torch.norm, torch.max,
mean, autograd, torch.optim.SGD / Adagrad.  ``TorchOracle`` issues the same torch calls in the same
order as reference openke/module/model/{TransE,TransH,TransD}.py, module/strategy/NegativeSampling.py:
13-33, module/loss/MarginLoss.py:28 and config/Trainer.py:44-56,65-88 — on the CPU, in fp32.
``closed_form_grads`` is an independent float64 numpy derivation of the same gradients (SURVEY.md
section 3.1 / appendix D), used to check the analytic backward that the CUDA kernels implement.

Parity status: PINNED by tests/test_oracle.py against tests/golden/train.npz (losses and final
tables produced by the unmodified reference Trainer on the same batches).
"""
import numpy as np
import torch
import torch.nn.functional as F


class TorchOracle(object):
    def __init__(self, model, tables, p_norm=1, norm_flag=True, opt="adagrad", lr=0.1, margin=1.0, k=1):
        """model in {'transe','transh','transd'}; tables: dict name -> float32 ndarray (copied)."""
        self.model, self.p, self.norm_flag, self.k, self.margin = model, p_norm, norm_flag, k, float(margin)
        self.t = {n: torch.tensor(np.array(v, dtype=np.float32), requires_grad=True) for n, v in tables.items()}
        params = list(self.t.values())
        if opt.lower() == "adagrad":
            self.opt = torch.optim.Adagrad(params, lr=lr, lr_decay=0, weight_decay=0)
        else:
            self.opt = torch.optim.SGD(params, lr=lr, weight_decay=0)

    # reference TransE.py:46-60
    def _calc(self, h, t, r, mode):
        if self.norm_flag:
            h, r, t = F.normalize(h, 2, -1), F.normalize(r, 2, -1), F.normalize(t, 2, -1)
        if mode != "normal":
            h = h.view(-1, r.shape[0], h.shape[-1])
            t = t.view(-1, r.shape[0], t.shape[-1])
            r = r.view(-1, r.shape[0], r.shape[-1])
        score = h + (r - t) if mode == "head_batch" else (h + r) - t
        return torch.norm(score, self.p, -1).flatten()

    def score(self, bh, bt, br, mode="normal"):
        bh, bt, br = (torch.as_tensor(np.asarray(x), dtype=torch.int64) for x in (bh, bt, br))
        E, R = self.t["ent_embeddings"], self.t["rel_embeddings"]
        emb = F.embedding   # nn.Embedding.forward: its backward sums duplicate rows like the reference's
        h, t, r = emb(bh, E), emb(bt, E), emb(br, R)
        if self.model == "transh":       # TransH.py:68-76,83-88
            def transfer(e, norm):
                norm = F.normalize(norm, p=2, dim=-1)
                if e.shape[0] != norm.shape[0]:
                    e = e.view(-1, norm.shape[0], e.shape[-1])
                    norm = norm.view(-1, norm.shape[0], norm.shape[-1])
                    e = e - torch.sum(e * norm, -1, True) * norm
                    return e.view(-1, e.shape[-1])
                return e - torch.sum(e * norm, -1, True) * norm
            w = emb(br, self.t["norm_vector"])
            h, t = transfer(h, w), transfer(t, w)
        elif self.model == "transd":     # TransD.py:94-110,117-125 (dim_e == dim_r)
            def transfer(e, ep, rp):
                if e.shape[0] != rp.shape[0]:
                    e = e.view(-1, rp.shape[0], e.shape[-1])
                    ep = ep.view(-1, rp.shape[0], ep.shape[-1])
                    rp = rp.view(-1, rp.shape[0], rp.shape[-1])
                    e = F.normalize(e + torch.sum(e * ep, -1, True) * rp, p=2, dim=-1)
                    return e.view(-1, e.shape[-1])
                return F.normalize(e + torch.sum(e * ep, -1, True) * rp, p=2, dim=-1)
            ET, RT = self.t["ent_transfer"], self.t["rel_transfer"]
            rt = emb(br, RT)
            h, t = transfer(h, emb(bh, ET), rt), transfer(t, emb(bt, ET), rt)
        return self._calc(h, t, r, mode)

    def loss(self, bh, bt, br):
        score = self.score(bh, bt, br)
        B = score.shape[0] // (1 + self.k)
        p = score[:B].view(-1, B).permute(1, 0)          # NegativeSampling.py:13-21
        n = score[B:].view(-1, B).permute(1, 0)
        m = torch.tensor([self.margin])
        return torch.max(p - n, -m).mean() + m           # MarginLoss.py:28

    def step(self, bh, bt, br):                          # Trainer.py:44-56
        self.opt.zero_grad()
        loss = self.loss(bh, bt, br)
        loss.backward()
        self.opt.step()
        return float(loss.item())

    def tables(self):
        return {n: v.detach().numpy().copy() for n, v in self.t.items()}


def closed_form_grads(model, tables, bh, bt, br, k, margin, p_norm=1):
    """float64 loss and dense table gradients from the closed forms of SURVEY.md 3.1 / appendix D."""
    T = {n: np.asarray(v, dtype=np.float64) for n, v in tables.items()}
    G = {n: np.zeros_like(v) for n, v in T.items()}
    bh, bt, br = (np.asarray(x, dtype=np.int64) for x in (bh, bt, br))
    n_all = bh.shape[0]
    B = n_all // (1 + k)
    eps = 1e-12

    def N(x):
        n = max(np.sqrt((x * x).sum()), eps)
        return x / n, n

    def Nb(y, n, g):
        return (g - y * (y @ g)) / n

    fw = []
    for i in range(n_all):
        h, t, r = T["ent_embeddings"][bh[i]], T["ent_embeddings"][bt[i]], T["rel_embeddings"][br[i]]
        c = {}
        if model == "transh":
            wh, nw = N(T["norm_vector"][br[i]])
            ah, at = h @ wh, t @ wh
            hp, tp = h - ah * wh, t - at * wh
            c.update(wh=wh, nw=nw, ah=ah, at=at, h=h, t=t)
        elif model == "transd":
            he, te, rp = T["ent_transfer"][bh[i]], T["ent_transfer"][bt[i]], T["rel_transfer"][br[i]]
            ah, at = h @ he, t @ te
            h1, n1h = N(h + ah * rp)
            t1, n1t = N(t + at * rp)
            hp, tp = h1, t1
            c.update(he=he, te=te, rp=rp, ah=ah, at=at, h1=h1, t1=t1, n1h=n1h, n1t=n1t, h=h, t=t)
        else:
            hp, tp = h, t
        hh, nh = N(hp)
        th, nt = N(tp)
        rh, nr = N(r)
        s = hh + rh - th
        sc = np.abs(s).sum() if p_norm == 1 else np.sqrt((s * s).sum())
        c.update(hh=hh, nh=nh, th=th, nt=nt, rh=rh, nr=nr, s=s, sc=sc)
        fw.append(c)
    gscore = np.zeros(n_all)
    loss = 0.0
    for i in range(B):
        for j in range(k):
            o = B + j * B + i
            diff = fw[i]["sc"] - fw[o]["sc"]
            loss += max(diff, -margin)
            g = (1.0 if diff > -margin else (0.5 if diff == -margin else 0.0)) / (B * k)
            gscore[i] += g
            gscore[o] -= g
    loss = loss / (B * k) + margin
    for i in range(n_all):
        c = fw[i]
        if gscore[i] == 0.0:
            continue
        if p_norm == 1:
            gs = gscore[i] * np.sign(c["s"])
        else:
            gs = gscore[i] * (c["s"] / c["sc"] if c["sc"] > 0 else 0 * c["s"])
        g_hp, g_tp = Nb(c["hh"], c["nh"], gs), Nb(c["th"], c["nt"], -gs)
        G["rel_embeddings"][br[i]] += Nb(c["rh"], c["nr"], gs)
        if model == "transe":
            G["ent_embeddings"][bh[i]] += g_hp
            G["ent_embeddings"][bt[i]] += g_tp
        elif model == "transh":
            wh = c["wh"]
            G["ent_embeddings"][bh[i]] += g_hp - (g_hp @ wh) * wh
            G["ent_embeddings"][bt[i]] += g_tp - (g_tp @ wh) * wh
            gw = -(c["ah"] * g_hp + (g_hp @ wh) * c["h"]) - (c["at"] * g_tp + (g_tp @ wh) * c["t"])
            G["norm_vector"][br[i]] += Nb(wh, c["nw"], gw)
        else:
            gu_h, gu_t = Nb(c["h1"], c["n1h"], g_hp), Nb(c["t1"], c["n1t"], g_tp)
            ch, ct = gu_h @ c["rp"], gu_t @ c["rp"]
            G["ent_embeddings"][bh[i]] += gu_h + ch * c["he"]
            G["ent_embeddings"][bt[i]] += gu_t + ct * c["te"]
            G["ent_transfer"][bh[i]] += ch * c["h"]
            G["ent_transfer"][bt[i]] += ct * c["t"]
            G["rel_transfer"][br[i]] += c["ah"] * gu_h + c["at"] * gu_t
    return loss, G
