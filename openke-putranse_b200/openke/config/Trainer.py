"""Epoch x batch training driver with the constructor and methods of the reference's
``openke/config/Trainer.py:16-138``.

``run()`` keeps the whole loop on the GPU: for every epoch one ``pk_train_steps`` call performs
``nbatches`` x (device sampling -> fused forward/backward/update), captured as a CUDA graph; the only
host traffic is the per-epoch read-back of the step losses.  ``train_one_step(data)`` is the
reference's per-batch entry point for callers that bring their own host batch.
"""
import ctypes
import os

import numpy as np
import torch
from tqdm import tqdm

from .. import _native as N


class Trainer(object):
    def __init__(self, model=None, data_loader=None, valid_dataloader=None, train_times=1000, alpha=0.5, use_gpu=True,
                 opt_method="sgd", save_steps=None, checkpoint_dir=None):
        self.work_threads = 8
        self.train_times = train_times
        self.opt_method = opt_method
        self.optimizer = None      # kept for API compatibility; the optimizer lives inside the kernel
        self.lr_decay = 0
        self.weight_decay = 0
        self.alpha = alpha
        self.model = model
        self.data_loader = data_loader
        self.use_gpu = use_gpu
        self.save_steps = save_steps
        self.checkpoint_dir = checkpoint_dir
        self.losses = []           # per-step losses of the last run(), epoch-major
        self.gpu_launches = 0
        self._state = None         # Adagrad accumulators, one tensor per table
        self._ws = None
        self._ws_key = None

    # ---- plumbing
    def _opt_id(self):
        m = (self.opt_method or "sgd").lower()
        if m == "adagrad":
            return N.PK_ADAGRAD
        if m == "sgd":
            return N.PK_SGD
        raise NotImplementedError("opt_method %r: the fused train step implements SGD and Adagrad "
                                  "(the optimizers the PuTransE path uses)" % self.opt_method)

    def _prepare(self):
        if not self.use_gpu:
            raise N.NativeError("use_gpu=False: the B200 train step has no CPU implementation")
        N.require_cuda()
        if self.lr_decay != 0 or self.weight_decay != 0:
            raise NotImplementedError("lr_decay / weight_decay are not part of the fused train step")
        self.model.cuda()
        emb = self.model.model
        dl = self.data_loader
        opt = self._opt_id()
        if opt == N.PK_ADAGRAD and self._state is None:
            self._state = {n: torch.zeros_like(getattr(emb, n).weight) for n in emb.table_names()}
        cfg = emb.native_cfg(opt=opt, neg_ent=dl.negative_ent, bern=1 if dl.bern else 0, filt=1 if dl.filter else 0,
                             work_threads=dl.work_threads)
        key = (emb._pk_model, emb.dim_native, dl.negative_ent, emb.ent_tot, emb.rel_tot, dl.batch_size)
        if self._ws is None or self._ws_key != key:
            self._free_ws()
            self._ws = N.lib().pk_workspace_create(ctypes.byref(cfg), emb.ent_tot, emb.rel_tot, dl.batch_size)
            if not self._ws:
                raise N.NativeError("pk_workspace_create: %s" % N.last_error())
            self._ws_key = key
        tab = emb.native_tables(self._state if opt == N.PK_ADAGRAD else None)
        return emb, cfg, tab

    def _free_ws(self):
        if self._ws:
            N.lib().pk_workspace_free(self._ws)
            self._ws = None

    def __del__(self):
        try:
            self._free_ws()
        except Exception:
            pass

    def _margin(self):
        return float(self.model.loss.margin.item())

    # ---- reference Trainer.train_one_step (Trainer.py:44-56)
    def train_one_step(self, data):
        emb, cfg, tab = self._prepare()
        dev = emb.ent_embeddings.weight.device
        ids = [torch.from_numpy(np.ascontiguousarray(data[k]).astype(np.int32, copy=False)).to(dev, non_blocking=True)
               for k in ("batch_h", "batch_t", "batch_r")]
        k = self.data_loader.negative_ent
        bsz = ids[0].numel() // (1 + k)
        loss = torch.zeros(1, dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().pk_train_step(ctypes.byref(cfg), ctypes.byref(tab), self._ws, bsz, ids[0].data_ptr(),
                                      ids[1].data_ptr(), ids[2].data_ptr(), self._margin(), float(self.alpha),
                                      loss.data_ptr(), st), "pk_train_step")
        self.gpu_launches += N.lib().pk_last_launch_count()
        out = loss.item()
        N.check(N.lib().pk_workspace_check(self._ws, st), "pk_train_step")
        return out

    # ---- reference Trainer.run (Trainer.py:58-104)
    def run(self, show_progress=True):
        emb, cfg, tab = self._prepare()
        dl = self.data_loader
        dev = emb.ent_embeddings.weight.device
        smp = dl.device_sampler(dev)
        if smp["n_ent"] != emb.ent_tot or smp["n_rel"] != emb.rel_tot:
            raise N.NativeError("model tables (%d entities, %d relations) do not match the loader's id space (%d, %d)"
                                % (emb.ent_tot, emb.rel_tot, smp["n_ent"], smp["n_rel"]))
        nb = dl.nbatches
        st = torch.cuda.current_stream(dev).cuda_stream
        margin, lr = self._margin(), float(self.alpha)
        self.losses = []
        loss_buf = torch.zeros(max(nb, 1), dtype=torch.float32, device=dev)
        rng = tqdm(range(self.train_times)) if show_progress else range(self.train_times)
        for epoch in rng:
            N.check(N.lib().pk_train_steps(ctypes.byref(cfg), ctypes.byref(tab), ctypes.byref(smp["struct"]), self._ws,
                                           dl.batch_size, nb, margin, lr, loss_buf.data_ptr(), st), "pk_train_steps")
            self.gpu_launches += N.lib().pk_last_launch_count()
            ep = loss_buf.cpu().numpy().copy()   # the one D2H per epoch (the reference syncs every step)
            self.losses.append(ep)
            if show_progress:
                rng.set_description("Epoch %d | loss: %f" % (epoch, ep[-1]))
            if self.save_steps and self.checkpoint_dir and (epoch + 1) % self.save_steps == 0:
                self.model.save_checkpoint(os.path.join(self.checkpoint_dir + "-" + str(epoch) + ".ckpt"))
        dl.sync_lcg_from_device()

    # ---- setters of the reference (Trainer.py:106-138)
    def set_model(self, model):
        self.model = model

    def to_var(self, x, use_gpu):
        t = torch.from_numpy(x)
        return t.cuda() if use_gpu else t

    def set_use_gpu(self, use_gpu):
        self.use_gpu = use_gpu

    def set_alpha(self, alpha):
        self.alpha = alpha

    def set_lr_decay(self, lr_decay):
        self.lr_decay = lr_decay

    def set_weight_decay(self, weight_decay):
        self.weight_decay = weight_decay

    def set_opt_method(self, opt_method):
        self.opt_method = opt_method

    def set_train_times(self, train_times):
        self.train_times = train_times

    def set_save_steps(self, save_steps, checkpoint_dir=None):
        self.save_steps = save_steps
        if not self.checkpoint_dir:
            self.set_checkpoint_dir(checkpoint_dir)

    def set_checkpoint_dir(self, checkpoint_dir):
        self.checkpoint_dir = checkpoint_dir
