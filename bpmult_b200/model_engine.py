"""mmtrvat (MultiprojectionMMTransformer3DGMUClf, models/mmtr.py:587-866) as an explicit forward / backward schedule over
EncoderEngine / SeqGmuEngine / HeadEngine.  Owns the per-model buffers; parameters arrive as a dict of reference-named
fp32 tensors and gradients leave the same way."""
import contextlib
import os

import torch

from .engine import Arena, Dims, EncoderEngine, HeadEngine, SeqGmuEngine, round_up
from .ops import Drop

ENC_NAMES = ["l_with_a", "l_with_v", "l_with_v2a", "l_with_a2v",
             "v_with_l", "v_with_a", "v_with_l2a", "v_with_a2l",
             "a_with_l", "a_with_v", "a_with_v2l", "a_with_l2v"]            # ctor order, mmtr.py:639-653
# wave 1: name -> (query modality, source modality)                          mmtr.py:779-786
WAVE1 = {"v_with_a": ("v", "a"), "a_with_v": ("a", "v"), "v_with_l": ("v", "l"), "l_with_v": ("l", "v"),
         "a_with_l": ("a", "l"), "l_with_a": ("l", "a")}
# per target modality: (u, w, p, q) with mid = gmu_m_m([u, w]); top = gmu_m([p + u, q + w]) + mid; p / q read K,V from u / w
TARGETS = {"l": ("v_with_a", "a_with_v", "l_with_a2v", "l_with_v2a"),      # mmtr.py:788-808
           "a": ("l_with_v", "v_with_l", "a_with_v2l", "a_with_l2v"),      # mmtr.py:810-830
           "v": ("l_with_a", "a_with_l", "v_with_a2l", "v_with_l2a")}      # mmtr.py:832-852
HEAD_ORDER = ["l", "v", "a"]                                               # gmu([last_h_l, last_h_v, last_h_a]) :857
# hybrid = True (mmtr.py:631, 662, 680-689, 765-775, 854-855): per modality a bias-free Linear over the time axis (n_vec -> LOW_DIM
# steps), a self-attention encoder of max(layers, 3) layers, first + last step pooling; gmu_early fuses (l, v, a); the result is the
# fourth input of the final TextShiftingNLayer.  The two gate call sites are read as what their callees accept (list vs varargs).
LOW_DIM = 32
EARLY = ["l_early", "v_early", "a_early"]                                   # ctor order :681-683; gmu_early(l, v, a) :775


class Lanes:
    """Independent encoders of a wave are issued round-robin on `n` side streams (fork / join with events; inside the captured
    step they become parallel branches of the CUDA graph).  The GEMM and attention kernels are persistent with one CTA per SM and
    a static tile walk, so a kernel whose tile count is not a multiple of the SM count leaves part of the machine idle in its last
    round (512 tiles on 148 SMs: 4 rounds for 3.46 rounds of work; attention backward: 768 (batch, head) items = 6 rounds for 5.19);
    with two lanes the CTAs of the other lane's kernel start on the SMs that fall idle.  Every lane owns its scratch arena."""

    def __init__(self, device, n):
        self.on_gpu = torch.device(device).type == "cuda"
        self.n = n                                            # (on the CPU the lanes are logical only: same buffers, no streams)
        self.device = device
        self.streams = [torch.cuda.Stream(device=device) for _ in range(self.n)] if (self.n > 1 and self.on_gpu) else []

    def on(self, k):
        return torch.cuda.stream(self.streams[k % self.n]) if self.streams else contextlib.nullcontext()

    def fork(self):
        """the lanes wait for everything enqueued so far on the current stream"""
        if self.streams:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            for s in self.streams:
                s.wait_event(ev)

    def join(self):
        """the current stream waits for everything enqueued so far on the lanes"""
        if self.streams:
            cur = torch.cuda.current_stream(self.device)
            for s in self.streams:
                ev = torch.cuda.Event()
                ev.record(s)
                cur.wait_event(ev)

    def barrier(self):
        self.join()
        self.fork()


def attn_dropout_for(name, args):
    """get_network (mmtr.py:691-697): attention dropout is chosen by the SOURCE modality of the stream."""
    src = name.split("_with_")[1][-1]
    return {"l": args.attn_dropout, "a": args.attn_dropout_a, "v": args.attn_dropout_v}[src]


class MMTrVatEngine:
    def __init__(self, ops, args, dtype=torch.bfloat16, n_vec=512, prune=None):
        """prune (default: environment BPM_PRUNE=1, else off): the wave-2 stacks are crossmodal only -- a query row attends to the K/V
        stream, never to other query rows -- the gated units are row-wise, and only time steps 0 and n_vec-1 of their output reach the
        head (mmtr.py:808,830,852).  With prune the query side of the six wave-2 stacks and the gated units run on those two rows per
        sample: identical logits and gradients (every other row has a zero gradient and no reader), about a third of the step gone.
        The reference computes all rows; so does the default, and so does every headline number."""
        self.ops, self.args, self.T_, self.n_vec = ops, args, dtype, n_vec
        self.prune = bool(int(os.environ.get("BPM_PRUNE", "0"))) if prune is None else bool(prune)
        D, H, L = args.hidden_sz, args.num_heads, args.layers
        self.d = Dims(D, H)
        self.orig = {"l": args.orig_d_l, "a": args.orig_d_a, "v": args.orig_d_v}
        self.Kp = {m: (self.d.Dp if self.orig[m] == D else round_up(self.orig[m], 64)) for m in "lav"}
        self.lanes = Lanes(ops.device, int(os.environ.get("BPM_LANES", "6")))
        self.lane_shared = [Arena(ops) for _ in range(self.lanes.n)]
        self.shared = self.lane_shared[0]
        self.arena = Arena(ops)
        # lane of each encoder: wave 1 alternates, the two encoders of a target modality (p, q) sit on different lanes
        self.lane_of = {n: i % self.lanes.n for i, n in enumerate(WAVE1)}
        for i, m in enumerate(HEAD_ORDER):
            u, w, pn, qn = TARGETS[m]
            self.lane_of[pn], self.lane_of[qn] = (2 * i) % self.lanes.n, (2 * i + 1) % self.lanes.n
        self.enc = {}
        for i, n in enumerate(ENC_NAMES):
            self.enc[n] = EncoderEngine(ops, D, H, L, attn_dropout=attn_dropout_for(n, args), relu_dropout=args.relu_dropout,
                                        res_dropout=args.res_dropout, embed_dropout=args.embed_dropout, attn_mask=args.attn_mask,
                                        biprojection=False, dtype=dtype, uid=i + 1, shared=self.lane_shared[self.lane_of[n]])
        if self.prune:
            for m in HEAD_ORDER:
                for n in TARGETS[m][2:]:
                    # rows b * 4 + {0, 1, 2, 3}: time step 0, then three copies of step n_vec - 1 (the attention backward runs on tensor
                    # cores for T % 4 == 0 only; with 2 rows it fell back to the fp32-math kernels: 20 ms of 98 ms summed device time).
                    # pool_fwd / pool_bwd read and feed the FIRST and LAST row of a sample: rows 1 and 2 never receive a gradient.
                    self.enc[n].prune_pos = (0, n_vec - 1, n_vec - 1, n_vec - 1)
                    self.enc[n].prune_row0 = bool(args.attn_mask)     # future mask of the full sequence: step 0 sees key 0 only
        self.gmu = {}
        self.mod_lane = {m: i % self.lanes.n for i, m in enumerate(HEAD_ORDER)}      # lane of a modality's staging / gated units
        self.hybrid = bool(getattr(args, "hybrid", False))
        self.enc_names = ENC_NAMES + (EARLY if self.hybrid else [])
        if self.hybrid:
            for i, n in enumerate(EARLY):                     # get_network('*_mem', layers=3): every early stack uses attn_dropout (:699-707)
                self.enc[n] = EncoderEngine(ops, D, H, max(L, 3), attn_dropout=args.attn_dropout, relu_dropout=args.relu_dropout,
                                            res_dropout=args.res_dropout, embed_dropout=args.embed_dropout, attn_mask=args.attn_mask,
                                            biprojection=False, dtype=dtype, uid=len(ENC_NAMES) + 1 + i,
                                            shared=self.lane_shared[self.mod_lane[n[0]]])
            self.We = {m: ops.zeros((LOW_DIM, n_vec), torch.float32) for m in "lva"}          # proj_{m}_e.weight, reference layout
            self.Ge = {m: ops.zeros((LOW_DIM, n_vec), torch.float32) for m in "lva"}
            self.gate_early = HeadEngine(ops, D, 3, args.n_classes, prefix="gmu_early.", gate_only=True)
        for m in "lav":
            sh = self.lane_shared[self.mod_lane[m]]
            self.gmu[m + "_m"] = SeqGmuEngine(ops, D, dtype, True, sh, "gmu_%s_m" % m)
            self.gmu[m] = SeqGmuEngine(ops, D, dtype, True, sh, "gmu_%s" % m)
        self.head = HeadEngine(ops, D, 4 if self.hybrid else 3, args.n_classes, out_dropout=args.out_dropout, list_names=self.hybrid)
        z = ops.zeros
        self.Wproj = {m: (z((self.d.Dp, self.Kp[m]), dtype) if self.orig[m] != D else None) for m in "lav"}
        self.Gproj = {m: (z((self.d.Dp, self.Kp[m]), torch.float32) if self.orig[m] != D else None) for m in "lav"}

    # ---------------------------------------------------------------- parameters
    def param_shapes(self):
        s = {}
        D = self.d.D
        for m in ("l_m", "v_m", "a_m", "l", "v", "a"):
            for k, v in self.gmu[m].param_shapes().items():
                s["gmu_%s.%s" % (m, k)] = v
        for m in "lva":
            s["proj_%s.weight" % m] = (D, self.orig[m], 1)
        for n in self.enc_names:
            for k, v in self.enc[n].param_shapes().items():
                s["trans_%s.%s" % (n, k)] = v
        s.update(self.head.param_shapes())
        if self.hybrid:
            s.update(self.gate_early.param_shapes())
            for m in "lva":
                s["proj_%s_e.weight" % m] = (LOW_DIM, self.n_vec)
        return s

    def unused_params(self):
        """parameters the forward never touches (no gradient in the reference either)"""
        u = ["transfm_%s.%s" % (n, k) for n in ("a2l", "v2l", "l2a", "l2v") for k in ("weight", "bias")]
        u += ["proj_%s.weight" % m for m in "lav" if self.orig[m] == self.d.D]
        return u

    def pack(self, params):
        o = self.ops
        o.batch_begin("pack", "model")                        # ONE launch for all ~1400 parameter tensors
        for n in self.enc_names:
            self.enc[n].pack(params, "trans_%s." % n)
        for m, g in self.gmu.items():
            g.pack(params, "gmu_%s." % m)
        self.head.pack(params)
        for m in "lav":
            if self.Wproj[m] is not None:
                w = params["proj_%s.weight" % m]
                o.pack_matrix(w.view(w.shape[0], w.shape[1]), self.Wproj[m])
        if self.hybrid:
            self.gate_early.pack(params)
            for m in "lva":
                o.pack_matrix(params["proj_%s_e.weight" % m], self.We[m])
        o.batch_end()

    def zero_grads(self):
        self.ops.zero_begin()
        try:
            self._zero_grads()
        finally:
            self.ops.zero_end()

    def _zero_grads(self):
        for e in self.enc.values():
            e.zero_grads()
        for g in self.gmu.values():
            g.zero_grads()
        self.head.zero_grads()
        for m in "lav":
            if self.Gproj[m] is not None:
                self.ops.zero_(self.Gproj[m])
        if self.hybrid:
            self.gate_early.zero_grads()
            for m in "lva":
                self.ops.zero_(self.Ge[m])

    def _unpack_rest(self, grads, accumulate):
        """everything but the encoders, inside an open unpack batch"""
        o = self.ops
        for m, g in self.gmu.items():
            g.unpack_grads(grads, "gmu_%s." % m, accumulate=accumulate)
        self.head.unpack_grads(grads, accumulate=accumulate)
        for m in "lav":
            if self.Gproj[m] is not None:
                gw = grads["proj_%s.weight" % m]
                o.unpack_matrix(self.Gproj[m], gw.view(gw.shape[0], gw.shape[1]), accumulate=accumulate)
        if self.hybrid:
            self.gate_early.unpack_grads(grads, accumulate=accumulate)
            for m in "lva":
                o.unpack_matrix(self.Ge[m], grads["proj_%s_e.weight" % m], accumulate=accumulate)

    def unpack_grads(self, grads, accumulate=False):
        self.ops.batch_begin("unpack", "model")
        for n in self.enc_names:
            self.enc[n].unpack_grads(grads, "trans_%s." % n, accumulate)
        self._unpack_rest(grads, accumulate)
        self.ops.batch_end()

    def unpack_misc(self, grads, accumulate=False):
        """gradients of everything but the encoders (the encoders' leave per bucket during backward, see Trainer)"""
        o = self.ops
        o.batch_begin("unpack", "misc")
        self._unpack_rest(grads, accumulate)
        o.batch_end()

    # ---------------------------------------------------------------- forward
    def forward(self, txt, img, audio, training=True, seed=0, seed_ptr=None):
        """txt (B, T_l, orig_d_l), img (B, T_v, orig_d_v), audio (B, T_a, orig_d_a): fp32 device tensors (any strides).
        Returns (logits [B, C], z [B, 3*D]) views of internal buffers."""
        o, d, A, nv = self.ops, self.d, self.arena, self.n_vec
        B = txt.shape[0]
        self.B, self.training, self.seed, self.seed_ptr = B, training, seed, seed_ptr
        M = B * nv
        feats = {"l": txt, "a": audio, "v": img}
        self.in_shapes = {m: tuple(feats[m].shape) for m in "lav"}
        P = {}
        self.X = {}
        ln = self.lanes
        ln.fork()
        for m in "lav":
            # transpose / text embed-dropout / zero-pad to n_vec (mmtr.py:741-761), then Conv1d(k=1) as a row GEMM (:748-750)
            drop = Drop(self.args.embed_dropout, seed, seed_ptr, 7) if (m == "l" and training and self.args.embed_dropout > 0) else None
            X = A.get("X_" + m, (M, self.Kp[m]), self.T_)
            with ln.on(self.mod_lane[m]):
                o.stage_rows(feats[m], X, nv, drop)
                self.X[m] = X
                if self.Wproj[m] is not None:
                    P[m] = A.get("P_" + m, (M, d.Dp), self.T_)
                    o.gemm(X, self.Wproj[m], P[m], M, d.Dp, self.Kp[m])
                else:
                    P[m] = X
        self.P = P
        h = {}
        ln.barrier()
        if self.hybrid:                                                          # "parallel fusion" (:765-775), one lane per modality
            cat_e = self.gate_early.cat_buf(B)
            self.Pe = {}
            for ci, n in enumerate(EARLY):
                m = n[0]
                with ln.on(self.mod_lane[m]):
                    pe = A.get("Pe_" + m, (B * LOW_DIM, d.Dp), self.T_)
                    o.timelin_fwd(P[m], self.We[m], None, pe, B, nv, LOW_DIM, d.D)       # proj_m_e over the time axis (:767-769)
                    self.Pe[m] = pe
                    he = self.enc[n].forward(pe, B, LOW_DIM, training=training, seed=seed, seed_ptr=seed_ptr)
                    o.pool_fwd(he, B, LOW_DIM, cat_e, ci * d.Dp)                          # h[0] + h[-1] (:772-774)
        for n, (qm, km) in WAVE1.items():
            with ln.on(self.lane_of[n]):
                h[n] = self.enc[n].forward(P[qm], B, nv, src_k=P[km], S=nv, training=training, seed=seed, seed_ptr=seed_ptr)
        ln.barrier()                                                            # wave 2 reads wave-1 outputs of either lane
        # rows of the wave-2 query side and of the gated units: all n_vec, or (prune) time step 0 and three copies of step n_vec - 1
        Tq = 4 if self.prune else nv
        Mq = B * Tq

        def two_rows(key, x):
            y = A.get(key, (Mq, d.Dp), x.dtype)
            xv_, yv_ = x.view(B, nv, d.Dp), y.view(B, 4, d.Dp)
            yv_[:, 0].copy_(xv_[:, 0])
            yv_[:, 1:].copy_(xv_[:, nv - 1:nv].expand(B, 3, d.Dp))
            return y
        for m in HEAD_ORDER:
            u, w, pn, qn = TARGETS[m]
            Pq = P[m]
            if self.prune:
                with ln.on(self.lane_of[pn]):
                    Pq = two_rows("Pq_" + m, P[m])
                if ln.streams:                                                    # the other lane reads it too
                    ev = torch.cuda.Event()
                    ev.record(ln.streams[self.lane_of[pn] % ln.n])
                    ln.streams[self.lane_of[qn] % ln.n].wait_event(ev)
            with ln.on(self.lane_of[pn]):
                h[pn] = self.enc[pn].forward(Pq, B, Tq, src_k=h[u], S=nv, training=training, seed=seed, seed_ptr=seed_ptr)
            with ln.on(self.lane_of[qn]):
                h[qn] = self.enc[qn].forward(Pq, B, Tq, src_k=h[w], S=nv, training=training, seed=seed, seed_ptr=seed_ptr)
        ln.barrier()
        cat = self.head.cat_buf(B)
        self.tops = {}
        for ci, m in enumerate(HEAD_ORDER):                                      # the three targets' gated units: one lane each
            u, w, pn, qn = TARGETS[m]
            hp, hq = h[pn], h[qn]
            with ln.on(self.mod_lane[m]):
                hu, hw = (two_rows("u2_" + m, h[u]), two_rows("w2_" + m, h[w])) if self.prune else (h[u], h[w])
                mid = self.gmu[m + "_m"].forward(hu, hw, Mq)                       # "GMU middle"
                a1 = A.get("a1_" + m, (Mq, d.Dp), self.T_)
                a2 = A.get("a2_" + m, (Mq, d.Dp), self.T_)
                o.add(hp, hu, a1)                                                    # residual level 1 -> 2 (:799-800)
                o.add(hq, hw, a2)
                top = self.gmu[m].forward(a1, a2, Mq, addend=mid)                    # "GMU top" + residual level 1 -> 3 (:803-806)
                self.tops[m] = top
                o.pool_fwd(top, B, Tq, cat, ci * d.Dp)                               # h[0] + h[-1] (:808)
        ln.join()
        self.h = h
        if self.hybrid:
            fused_e, _ = self.gate_early.gate_forward(B)                          # gmu_early (:775); its gates are not returned
            cat[:, 3 * d.Dp:4 * d.Dp].copy_(fused_e)                              # fourth input of the final gate (:855)
        logits, z = self.head.forward(B, training, seed, seed_ptr)
        return logits, z

    def loss(self, logits, targets, pos_weight=None, grad_scale=1.0):
        return self.head.loss(logits, targets, pos_weight, grad_scale)

    # ---------------------------------------------------------------- backward
    def backward(self, dlogits, d_inputs=None, on_done=None):
        """dlogits fp32 [B, Cp].  Parameter gradients accumulate in the padded buffers (see unpack_grads).
        d_inputs: optional dict m -> fp32 tensor shaped like the input features, overwritten with input gradients."""
        o, d, A, nv, B = self.ops, self.d, self.arena, self.n_vec, self.B
        M = B * nv
        f32 = torch.float32
        dcat = self.head.backward(dlogits)
        ln = self.lanes
        # every lane accumulates the projection-output gradients in its own buffers (two encoders of different lanes share a query
        # stream); they are summed once at the end
        dPl = [{m: A.get("dP%d_%s" % (k, m), (M, d.Dp), f32) for m in "lav"} for k in range(ln.n)]
        dP = dPl[0]
        dh = {n: A.get("dh_" + n, (M, d.Dp), f32) for n in WAVE1}
        for t in [t for k in range(ln.n) for t in dPl[k].values()] + list(dh.values()):
            o.zero_(t)
        Tq = 4 if self.prune else nv
        Mq = B * Tq
        dtop = {m: A.get("dtop_" + m, (Mq, d.Dp), f32) for m in HEAD_ORDER}
        da1 = {m: A.get("da1_" + m, (Mq, d.Dp), f32) for m in HEAD_ORDER}
        da2 = {m: A.get("da2_" + m, (Mq, d.Dp), f32) for m in HEAD_ORDER}

        def add_two_rows(src4, dst):                               # dst[time steps 0, n_vec - 1] += rows 0 and 3 (rows 1, 2 carry zeros)
            sv_, dv_ = src4.view(B, 4, d.Dp), dst.view(B, nv, d.Dp)
            dv_[:, 0] += sv_[:, 0]
            dv_[:, nv - 1] += sv_[:, 3]
        dcat_e = None
        if self.hybrid:
            dfe = A.get("dfused_e", (B, d.Dp), f32)
            dfe.copy_(dcat[:, 3 * d.Dp:4 * d.Dp])
            dcat_e = self.gate_early.gate_backward(dfe)                           # [B, 3 * Dp]: d(last_h*_early)
        ln.fork()
        if self.hybrid:
            for ci, n in reversed(list(enumerate(EARLY))):
                m = n[0]
                k = self.mod_lane[m]
                with ln.on(k):
                    dhe = A.get("dhe_" + m, (B * LOW_DIM, d.Dp), f32)
                    dpe = A.get("dpe_" + m, (B * LOW_DIM, d.Dp), f32)
                    o.zero_(dhe)
                    o.zero_(dpe)
                    o.pool_bwd(dcat_e, ci * d.Dp, B, LOW_DIM, dhe)
                    self.enc[n].backward(dhe, dpe)
                    if on_done:
                        on_done(n)
                    # back through the time-axis Linear: d proj_x_m += W^T d, d W += d x^T
                    o.timelin_bwd(dpe, self.P[m], self.We[m], dPl[k][m], True, self.Ge[m], None, B, nv, LOW_DIM, d.D)
        for ci, m in reversed(list(enumerate(HEAD_ORDER))):                      # gated fusion units of the three targets: one lane each
            u, w, pn, qn = TARGETS[m]
            with ln.on(self.mod_lane[m]):
                o.zero_(dtop[m])
                o.pool_bwd(dcat, ci * d.Dp, B, Tq, dtop[m])
                o.zero_(da1[m])
                o.zero_(da2[m])
                self.gmu[m].backward(dtop[m], da1[m], da2[m])                    # d(p+u), d(q+w)
                if self.prune:
                    du2 = A.get("du2_" + m, (Mq, d.Dp), f32)
                    dw2 = A.get("dw2_" + m, (Mq, d.Dp), f32)
                    o.zero_(du2)
                    o.zero_(dw2)
                    self.gmu[m + "_m"].backward(dtop[m], du2, dw2)
                    o.axpy_f32(da1[m], du2, True)
                    o.axpy_f32(da2[m], dw2, True)
                    add_two_rows(du2, dh[u])
                    add_two_rows(dw2, dh[w])
                else:
                    self.gmu[m + "_m"].backward(dtop[m], dh[u], dh[w])           # mid consumes u, w directly
                    o.axpy_f32(da1[m], dh[u], True)
                    o.axpy_f32(da2[m], dh[w], True)
        ln.barrier()
        for m in reversed(HEAD_ORDER):
            u, w, pn, qn = TARGETS[m]
            for n, da, dsrc in ((qn, da2[m], dh[w]), (pn, da1[m], dh[u])):
                k = self.lane_of[n]
                with ln.on(k):
                    if self.prune:
                        dq2 = A.get("dq2_" + n, (Mq, d.Dp), f32)
                        o.zero_(dq2)
                        self.enc[n].backward(da, dq2, dsrc)
                        add_two_rows(dq2, dPl[k][m])
                    else:
                        self.enc[n].backward(da, dPl[k][m], dsrc)
                    if on_done:
                        on_done(n)
        ln.barrier()                                                            # wave 1 consumes dh written on either lane
        for n, (qm, km) in reversed(list(WAVE1.items())):
            k = self.lane_of[n]
            with ln.on(k):
                self.enc[n].backward(dh[n], dPl[k][qm], dPl[k][km])
                if on_done:
                    on_done(n)
        ln.barrier()
        for m in "lav":                                                          # input projections: one lane per modality
            with ln.on(self.mod_lane[m]):
                for k in range(1, ln.n):
                    o.axpy_f32(dPl[k][m], dP[m], True)
                if self.Wproj[m] is not None:
                    sh = self.lane_shared[self.mod_lane[m]]
                    g = sh.get("dPc", (M, d.Dp), self.T_)
                    o.cast_drop(dP[m], g, None)
                    o.gemm(g, self.X[m], self.Gproj[m], d.Dp, self.Kp[m], M, ta=1, tb=1, accumulate=True)     # dW = dP^T X
                    if d_inputs is not None and m in d_inputs:
                        dX = sh.get("dX_" + m, (M, self.Kp[m]), f32)
                        o.gemm(g, self.Wproj[m], dX, M, self.Kp[m], d.Dp, tb=1)
                        self._unstage(m, dX, d_inputs[m])
                elif d_inputs is not None and m in d_inputs:
                    self._unstage(m, dP[m], d_inputs[m])
        ln.join()

    def backward_order(self):
        """encoder names in the order their gradients complete during backward (bucket launch order)"""
        order = list(reversed(EARLY)) if self.hybrid else []
        for m in reversed(HEAD_ORDER):
            u, w, pn, qn = TARGETS[m]
            order += [qn, pn]
        return order + list(reversed(list(WAVE1.keys())))

    def _unstage(self, m, g, dst):
        drop = Drop(self.args.embed_dropout, self.seed, self.seed_ptr, 7) if (m == "l" and self.training and self.args.embed_dropout > 0) else None
        self.ops.unstage_rows(g, dst, self.n_vec, False, drop)
