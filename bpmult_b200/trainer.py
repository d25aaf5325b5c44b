"""Graph-captured, data-parallel training step for the drop-in model (train.py:341-448 semantics: forward, BCEWithLogits,
backward, Adam; `grad_accum` micro-batches per optimizer step as train.py:390-398; the learning rate lives in device memory so
a ReduceLROnPlateau scheduler (train.py:128-136, 408) can change it between replays of the captured step -- see `set_lr`).

  * one process per GPU; gradients live in ONE flat fp32 buffer laid out in the order the encoders finish their backward
    (wave-2 first), so each encoder's bucket is all-reduced (NCCL, side stream) as soon as its last wgrad has landed, while the
    remaining backward keeps the SMs busy; parameters that never receive a gradient (transfm_*, an unused proj_*) are left out.
  * parameters are re-pointed into one flat fp32 buffer as well => a single fused Adam launch (bpm_adam_step).
  * the whole step (H2D of the batch from pinned memory, weight staging, forward, loss, backward, all-reduces, Adam, D2H of
    the loss) is captured once into a CUDA graph and replayed: shapes are static by construction (everything is padded to 512
    time steps) and the dropout seed lives in device memory."""
import os

import torch
import torch.distributed as dist

from .modules import MultiprojectionMMTransformer3DGMUClf, MultiprojectionMMTransformerGMUClf


class _Loss:
    """handle of an enqueued step: item() waits for that step and returns its loss"""

    def __init__(self, host, event):
        self.host, self.event = host, event

    def item(self):
        if self.event is not None:
            self.event.synchronize()
        return float(self.host[0])


class Trainer:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, pos_weight=None, seed=1234, use_graph=None, grad_accum=1):
        assert isinstance(model, (MultiprojectionMMTransformer3DGMUClf, MultiprojectionMMTransformerGMUClf))
        assert grad_accum >= 1
        self.model = model
        self.device = next(model.parameters()).device
        self.lr, self.betas, self.eps = lr, betas, eps
        self.grad_accum, self.micro = int(grad_accum), 0
        self.lr_t = torch.tensor([lr], dtype=torch.float32, device=self.device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.eng = model.engine(self.device)
        self.ops = self.eng.ops
        if use_graph is None:
            use_graph = os.environ.get("BPM_NO_GRAPH", "0") != "1"
        self.use_graph = use_graph and self.device.type == "cuda"
        self.pos_weight = None if pos_weight is None else pos_weight.to(self.device, torch.float32)
        self._flatten()
        dev = self.device
        self.seed_t = torch.tensor([seed + 7919 * self.rank], dtype=torch.int64, device=dev)     # per-rank Philox seed
        self.step_t = torch.zeros(1, dtype=torch.int64, device=dev)
        self.on_gpu = dev.type == "cuda"                     # (CPU only in the gloo host-logic tests, with the ops emulation)
        self.comm = torch.cuda.Stream(device=dev) if (self.world > 1 and self.on_gpu) else None
        self.graphs, self.static, self.shapes = {}, None, None
        self.loss_slots = [torch.zeros(1, dtype=torch.float32) for _ in range(2)]     # two steps may be in flight (step_async)
        if self.device.type == "cuda":
            self.loss_slots = [t.pin_memory() for t in self.loss_slots]
        self.loss_host = self.loss_slots[0]
        self._h2d_done = self._taken = None
        self._copy_stream = None
        self.steps_done = 0

    # ---------------------------------------------------------------- flat parameter / gradient buffers
    def _flatten(self):
        eng, model = self.eng, self.model
        pmap = dict(model.trunk_named_parameters())
        used = [n for n in eng.param_shapes() if n not in set(eng.unused_params())]
        order = []
        self.buckets = []                                   # (name, start, end) in flat-buffer elements
        off = 0
        for enc in eng.backward_order():
            names = [n for n in used if n.startswith("trans_%s." % enc)]
            n_el = sum(pmap[n].numel() for n in names)
            self.buckets.append((enc, off, off + n_el))
            order += names
            off += n_el
        misc = [n for n in used if not n.startswith("trans_")]
        self.buckets.append(("misc", off, off + sum(pmap[n].numel() for n in misc)))
        order += misc
        total = sum(pmap[n].numel() for n in order)
        dev = self.device
        self.flat_p = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.params, self.grads = {}, {}
        off = 0
        for n in order:
            p = pmap[n]
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            self.params[n], self.grads[n] = p.data, p.grad
            off += k
        for n in eng.unused_params():                       # packed too (never read by a kernel), keeps pack() uniform
            if n in pmap:
                self.params[n] = pmap[n].data
        self.n_params = total
        self.order = order                                  # flat-buffer layout: parameter names in bucket order
        self.sync_replicas()

    def sync_replicas(self):
        """Data parallelism only all-reduces GRADIENTS, so the replicas must start identical: rank 0's parameters and optimizer state
        are broadcast to every rank (one call each: the buffers are flat).  nn.DataParallel (train.py:354-356) has a single parameter
        copy and DDP broadcasts at construction; ranks built under different RNG state, or a checkpoint loaded on rank 0 only, would
        otherwise silently train different models."""
        if self.world > 1:
            for t in (self.flat_p, self.flat_m, self.flat_v, self.flat_g):
                dist.broadcast(t, 0)
            if hasattr(self, "step_t"):
                dist.broadcast(self.step_t, 0)

    # ---------------------------------------------------------------- one step (enqueue only)
    def _enqueue(self, *batch, apply=True):
        """one micro-batch: forward, loss, backward, gradients into the flat buffer; with `apply` also the all-reduces and Adam.
        With grad_accum > 1 the flat gradient buffer ACCUMULATES over the micro-batches (it is cleared after Adam) and only the
        last micro-batch reduces it across ranks -- the 1/grad_accum of train.py:391 rides on Adam's gradient scale."""
        eng, o = self.eng, self.ops
        acc = self.grad_accum > 1
        if apply:
            self.step_t += 1
        self.seed_t += 1
        eng.pack(self.params)
        *feats, tgt = batch                             # (txt, img, audio[, poster]), targets
        logits, _ = eng.forward(*feats, training=True, seed=0, seed_ptr=self.seed_t)
        loss, dlogits = eng.loss(logits, tgt, self.pos_weight, 1.0)
        eng.zero_grads()
        cur = torch.cuda.current_stream(self.device) if self.on_gpu else None
        bucket_of = {b[0]: b for b in self.buckets}

        def reduce_bucket(name):
            if self.world == 1 or not apply:
                return
            _, s, e = bucket_of[name]
            if not self.on_gpu:
                dist.all_reduce(self.flat_g[s:e])
                return
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))      # the lane stream this encoder's backward ran on
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                dist.all_reduce(self.flat_g[s:e])

        def on_done(enc):
            eng.enc[enc].unpack_grads(self.grads, "trans_%s." % enc, accumulate=acc)
            reduce_bucket(enc)
        eng.backward(dlogits, None, on_done)
        eng.unpack_misc(self.grads, accumulate=acc)              # everything but the encoders (gated units, head, projections)
        if not apply:
            return loss
        reduce_bucket("misc")
        if self.world > 1 and self.on_gpu:
            cur.wait_stream(self.comm)
        o.adam_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.lr, self.betas[0], self.betas[1], self.eps,
                    1.0 / (self.world * self.grad_accum), self.step_t, self.lr_t)
        if acc:
            self.flat_g.zero_()
        return loss

    # ---------------------------------------------------------------- learning rate / optimizer state (train.py:361-379,408,417-425)
    def set_lr(self, lr):
        """takes effect at the next optimizer step, captured graph or not (the kernel reads the rate from device memory)."""
        self.lr = float(lr)
        self.lr_t.fill_(self.lr)

    def get_lr(self):
        return self.lr

    def optimizer_state_dict(self):
        """native (flat) format; `names` pins the flat layout so an engine change cannot silently re-map a saved state"""
        return dict(step=int(self.step_t.item()), lr=self.lr, betas=self.betas, eps=self.eps, exp_avg=self.flat_m.clone(),
                    exp_avg_sq=self.flat_v.clone(), seed=int(self.seed_t.item()), rank=self.rank, micro=self.micro, grad=self.flat_g.clone(),
                    names=list(self.order))

    def load_optimizer_state_dict(self, sd):
        """accepts the native flat format or torch.optim.Adam's state_dict (the reference checkpoint's "optimizer" entry)"""
        if "param_groups" in sd:
            return self.load_torch_optimizer_state_dict(sd)
        if "names" in sd and list(sd["names"]) != list(self.order):
            raise ValueError("optimizer state was saved with a different flat parameter layout (%d vs %d tensors)" % (len(sd["names"]), len(self.order)))
        self.step_t.fill_(sd["step"])
        self.seed_t.fill_(sd["seed"] + 7919 * (self.rank - sd.get("rank", 0)))          # the seed stream stays per-rank
        self.flat_m.copy_(sd["exp_avg"])
        self.flat_v.copy_(sd["exp_avg_sq"])
        self.flat_g.copy_(sd["grad"])
        self.micro = sd["micro"]
        self.set_lr(sd["lr"])
        self.sync_replicas()

    # torch.optim.Adam layout (train.py:123-125 `optim.Adam(model.parameters(), lr)`, saved at train.py:422, loaded at :378):
    #   {"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [{"lr", "betas", "eps", "weight_decay", "amsgrad", ..., "params": [0..]}]}
    # with i = position in model.parameters(); parameters that never received a gradient have no state entry.
    def _param_index(self):
        return {n: i for i, (n, _) in enumerate(self.model.named_parameters())}

    def torch_optimizer_state_dict(self):
        idx = self._param_index()
        step = float(self.step_t.item())
        state = {}
        off = 0
        for n in self.order:
            k = self.params[n].numel()
            if step > 0:
                state[idx[n]] = {"step": torch.tensor(step), "exp_avg": self.flat_m[off:off + k].view_as(self.params[n]).clone(),
                                 "exp_avg_sq": self.flat_v[off:off + k].view_as(self.params[n]).clone()}
            off += k
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
                 "params": list(range(len(idx)))}
        return {"state": state, "param_groups": [group]}

    def load_torch_optimizer_state_dict(self, sd):
        names = [n for n, _ in self.model.named_parameters()]
        g = sd["param_groups"][0]
        assert len(sd["param_groups"]) == 1 and len(g["params"]) == len(names), "optimizer state does not match model.parameters()"
        assert not g.get("amsgrad", False) and not g.get("weight_decay", 0), "the fused Adam has no amsgrad / weight decay (train.py:124 uses neither)"
        pos = {n: i for i, n in enumerate(names)}
        self.flat_m.zero_()
        self.flat_v.zero_()
        step, off = 0, 0
        for n in self.order:
            k = self.params[n].numel()
            st = sd["state"].get(g["params"][pos[n]])
            if st is not None:
                self.flat_m[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.flat_v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                step = max(step, int(float(st["step"])))
            off += k
        self.step_t.fill_(step)
        self.betas, self.eps = tuple(g["betas"]), g["eps"]
        if self.graphs:                                      # betas / eps are baked into a captured step
            self.close()
        self.set_lr(g["lr"])
        self.sync_replicas()

    @property
    def param_groups(self):
        """lets torch.optim.lr_scheduler.ReduceLROnPlateau-style code (train.py:128-136, 408) read and write `param_groups[0]["lr"]`"""
        tr = self

        class _Group(dict):
            def __setitem__(self, k, v):
                dict.__setitem__(self, k, v)
                if k == "lr":
                    tr.set_lr(v)
        return [_Group(lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=0)]

    # ---------------------------------------------------------------- reference checkpoint format (utils/utils.py:21-30, train.py:372-379,419-430)
    def checkpoint(self, epoch=0, scheduler=None, n_no_improve=0, best_metric=float("-inf")):
        """the dict train.py:419-430 hands to save_checkpoint: loadable by the reference (`model.load_state_dict(ck["state_dict"])`,
        `optim.Adam.load_state_dict(ck["optimizer"])`)"""
        sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        return {"epoch": epoch, "state_dict": sd, "optimizer": self.torch_optimizer_state_dict(),
                "scheduler": scheduler.state_dict() if scheduler is not None else {}, "n_no_improve": n_no_improve, "best_metric": best_metric}

    def load_checkpoint(self, ck, strict=False):
        """`ck`: a reference-format checkpoint dict (or a path to one).  Keys saved under nn.DataParallel carry a "module." prefix
        (train.py:354-356,422); parameters of modules outside the trunk (enc.bert.*, audio_enc.*) are ignored unless strict."""
        if isinstance(ck, (str, bytes, os.PathLike)):
            ck = torch.load(ck, map_location="cpu", weights_only=False)
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in ck["state_dict"].items()}
        own = self.model.state_dict()
        missing = [k for k in own if k not in sd and not k.endswith(("_float_tensor", "version"))]
        unexpected = [k for k in sd if k not in own]
        if strict and (missing or unexpected):
            raise KeyError("checkpoint mismatch: missing %s unexpected %s" % (missing[:4], unexpected[:4]))
        with torch.no_grad():
            for k, v in sd.items():
                if k in own and not k.endswith("_float_tensor"):
                    own[k].copy_(v)                      # parameters are views into the flat buffer: copy in place, never re-point
        if ck.get("optimizer"):
            self.load_optimizer_state_dict(ck["optimizer"])
        else:
            self.sync_replicas()
        return dict(epoch=ck.get("epoch", 0), n_no_improve=ck.get("n_no_improve", 0), best_metric=ck.get("best_metric"), missing=missing,
                    unexpected=unexpected)

    def close(self):
        """Drops the captured graphs (they hold the NCCL kernels of the gradient all-reduces) and drains the device, so that the
        process group can be destroyed afterwards.  Call before dist.destroy_process_group()."""
        if self.on_gpu:
            torch.cuda.synchronize(self.device)
        self.graphs = {}
        self.warm = 0
        if self.on_gpu:
            torch.cuda.synchronize(self.device)
            self.ops.release_tables(id(self))

    def __del__(self):
        try:
            if self.on_gpu and self.graphs:
                self.ops.release_tables(id(self))
        except Exception:
            pass

    # ---------------------------------------------------------------- public API
    def _ensure_static(self, *batch):
        shapes = tuple(tuple(t.shape) for t in batch)
        if self.shapes != shapes:
            dev = self.device
            self.static = [torch.zeros(s, dtype=torch.float32, device=dev) for s in shapes]
            self.pinned = [torch.zeros(s, dtype=torch.float32).pin_memory() if self.on_gpu else torch.zeros(s) for s in shapes]
            # device-side landing buffers of step_async: the H2D of step k + 1 runs on a copy stream while step k computes
            self.stage_dev = [torch.zeros(s, dtype=torch.float32, device=dev) for s in shapes] if self.on_gpu else None
            self.loss_dev = torch.zeros(1, dtype=torch.float32, device=dev)
            self.shapes, self.graphs = shapes, {}
            self.warm = 0
            self._h2d_done = self._landed = self._taken = None

    def step_device(self, *batch):
        """batch = (txt, img, audio, targets) for mmtrvat, (txt, img, audio, poster, targets) for mmtrvapt, already resident on the
        device (fp32).  Returns the device loss tensor (no sync)."""
        self._ensure_static(*batch)
        for s, t in zip(self.static, batch):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self._run()
        return self.loss_dev

    def step(self, *batch):
        """end-to-end step from HOST tensors: pinned staging -> H2D -> step -> D2H of the loss.  Returns a python float (blocks until
        the step has finished, like the `loss.item()` of train.py:393)."""
        return self.step_async(*batch).item()

    def step_async(self, *batch):
        """same as step() but returns at once with a handle; `handle.item()` blocks for THIS step's loss.  Reading the loss one step late
        (enqueue step k+1, then `item()` of step k) keeps the GPU busy while the host prepares the next launch.  Pinned inputs are
        read by DMA straight from the caller's tensors: keep them unchanged until the handle has been read."""
        self._ensure_static(*batch)
        if not self.on_gpu:
            for s, t in zip(self.static, batch):
                s.copy_(t)
        else:
            # copy stream: pinned host -> landing buffers (waits only for the previous step's D2D out of them, which is the first thing
            # that step did); compute stream: landing -> static inputs (D2D, microseconds), then the captured step
            main = torch.cuda.current_stream(self.device)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            cs = self._copy_stream
            if self._h2d_done is not None:
                self._h2d_done.synchronize()                # the previous copies out of this trainer's pinned staging have finished
            if self._taken is not None:
                cs.wait_event(self._taken)
            with torch.cuda.stream(cs):
                for pbuf, sd, t in zip(self.pinned, self.stage_dev, batch):
                    if t.is_pinned() and t.dtype == torch.float32 and t.is_contiguous():
                        sd.copy_(t, non_blocking=True)      # e.g. DataLoader(pin_memory=True): DMA straight from the caller's buffer
                    else:
                        pbuf.copy_(t)                       # pageable input: staged through this trainer's pinned buffers
                        sd.copy_(pbuf, non_blocking=True)
                self._h2d_done = torch.cuda.Event()
                self._h2d_done.record(cs)
            main.wait_event(self._h2d_done)
            for s, sd in zip(self.static, self.stage_dev):
                s.copy_(sd, non_blocking=True)
            self._taken = torch.cuda.Event()
            self._taken.record(main)
        slot = self._loss_slot = (getattr(self, "_loss_slot", 1) + 1) & 1
        self._run()
        self.loss_slots[slot].copy_(self.loss_dev, non_blocking=True)
        ev = None
        if self.on_gpu:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        return _Loss(self.loss_slots[slot], ev)

    def evaluate(self, *batch, output_gates=False):
        """model_forward + the device half of model_eval (train.py:165-186, 283-338) for one batch, on the Trainer's own engine and
        weights: forward in eval mode (no dropout, nothing saved for a backward), BCEWithLogits with this Trainer's pos_weight,
        sigmoid, 0.5 threshold.  batch = (features..., targets), host or device tensors.  Returns a dict of HOST tensors:
        loss (python float), logits, probs, preds (bool), targets [, gates: the final GMU's z, mmtr.py:863-866]."""
        dev = self.device
        *feats, tgt = [t.to(dev, torch.float32, non_blocking=True) for t in batch]
        if self.on_gpu:
            torch.cuda.current_stream(dev).synchronize()        # a captured training step may still be replaying into the same arenas
        self.eng.pack(self.params)
        logits, z = self.eng.forward(*feats, training=False, seed=0)
        loss, _ = self.eng.loss(logits, tgt, self.pos_weight, 1.0)
        B, C, D, Dp = tgt.shape[0], tgt.shape[1], self.model.args.hidden_sz, self.eng.d.Dp          # engine buffers are padded: cut to (B, C) / (B, n*D)
        logits = logits[:, :C]
        probs = torch.sigmoid(logits)
        out = {"logits": logits.detach().cpu(), "probs": probs.cpu(), "preds": (probs > 0.5).cpu(), "targets": tgt.cpu(), "loss": float(loss)}
        if output_gates:
            n_gate = z.numel() // (B * Dp)
            out["gates"] = z.view(B, n_gate, Dp)[:, :, :D].reshape(B, n_gate * D).cpu()
        return out

    def _run(self):
        apply = (self.micro + 1) % self.grad_accum == 0     # train.py:395-398: optimizer step every grad_accum micro-batches
        self.micro = (self.micro + 1) % self.grad_accum
        if not self.use_graph:
            self.loss_dev.copy_(self._enqueue(*self.static, apply=apply))
            self.steps_done += 1
            return
        if apply not in self.graphs:
            if self.warm < 2 * self.grad_accum:             # eager warm-up steps (allocate every arena buffer, NCCL init)
                self.loss_dev.copy_(self._enqueue(*self.static, apply=apply))
                self.warm += 1
                self.steps_done += 1
                return
            torch.cuda.synchronize(self.device)
            self.ops.launches = 0
            self.ops.capture_owner = id(self)               # descriptor tables used by this capture stay alive until close()
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self.loss_dev.copy_(self._enqueue(*self.static, apply=apply))
                self.graphs[apply] = g
                self.launches_per_step = self.ops.launches
            except Exception as e:                         # e.g. a collective that cannot be captured on this stack
                if self.rank == 0:
                    print("bpmult_b200.Trainer: CUDA-graph capture failed (%s); running eagerly" % str(e).splitlines()[0])
                self.use_graph = False
                torch.cuda.synchronize(self.device)
                self.loss_dev.copy_(self._enqueue(*self.static, apply=apply))
                self.steps_done += 1
                return
        self.graphs[apply].replay()
        self.steps_done += 1

    @property
    def graph(self):
        return self.graphs.get(True)

    def bytes_in(self):
        return sum(t.numel() * 4 for t in self.static)
