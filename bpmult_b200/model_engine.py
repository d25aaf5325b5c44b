"""mmtrvat (MultiprojectionMMTransformer3DGMUClf, models/mmtr.py:587-866) as an explicit forward / backward schedule over
EncoderEngine / SeqGmuEngine / HeadEngine.  Owns the per-model buffers; parameters arrive as a dict of reference-named
fp32 tensors and gradients leave the same way."""
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


def attn_dropout_for(name, args):
    """get_network (mmtr.py:691-697): attention dropout is chosen by the SOURCE modality of the stream."""
    src = name.split("_with_")[1][-1]
    return {"l": args.attn_dropout, "a": args.attn_dropout_a, "v": args.attn_dropout_v}[src]


class MMTrVatEngine:
    def __init__(self, ops, args, dtype=torch.bfloat16, n_vec=512):
        self.ops, self.args, self.T_, self.n_vec = ops, args, dtype, n_vec
        D, H, L = args.hidden_sz, args.num_heads, args.layers
        self.d = Dims(D, H)
        self.orig = {"l": args.orig_d_l, "a": args.orig_d_a, "v": args.orig_d_v}
        self.Kp = {m: (self.d.Dp if self.orig[m] == D else round_up(self.orig[m], 64)) for m in "lav"}
        self.shared = Arena(ops)
        self.arena = Arena(ops)
        self.enc = {}
        for i, n in enumerate(ENC_NAMES):
            self.enc[n] = EncoderEngine(ops, D, H, L, attn_dropout=attn_dropout_for(n, args), relu_dropout=args.relu_dropout,
                                        res_dropout=args.res_dropout, embed_dropout=args.embed_dropout, attn_mask=args.attn_mask,
                                        biprojection=False, dtype=dtype, uid=i + 1, shared=self.shared)
        self.gmu = {}
        for m in "lav":
            self.gmu[m + "_m"] = SeqGmuEngine(ops, D, dtype, True, self.shared, "gmu_%s_m" % m)
            self.gmu[m] = SeqGmuEngine(ops, D, dtype, True, self.shared, "gmu_%s" % m)
        self.head = HeadEngine(ops, D, 3, args.n_classes, out_dropout=args.out_dropout)
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
        for n in ENC_NAMES:
            for k, v in self.enc[n].param_shapes().items():
                s["trans_%s.%s" % (n, k)] = v
        s.update(self.head.param_shapes())
        return s

    def unused_params(self):
        """parameters the forward never touches (no gradient in the reference either)"""
        u = ["transfm_%s.%s" % (n, k) for n in ("a2l", "v2l", "l2a", "l2v") for k in ("weight", "bias")]
        u += ["proj_%s.weight" % m for m in "lav" if self.orig[m] == self.d.D]
        return u

    def pack(self, params):
        o = self.ops
        o.batch_begin("pack", "model")                        # ONE launch for all ~1400 parameter tensors
        for n in ENC_NAMES:
            self.enc[n].pack(params, "trans_%s." % n)
        for m, g in self.gmu.items():
            g.pack(params, "gmu_%s." % m)
        self.head.pack(params)
        for m in "lav":
            if self.Wproj[m] is not None:
                w = params["proj_%s.weight" % m]
                o.pack_matrix(w.view(w.shape[0], w.shape[1]), self.Wproj[m])
        o.batch_end()

    def zero_grads(self):
        for e in self.enc.values():
            e.zero_grads()
        for g in self.gmu.values():
            g.zero_grads()
        self.head.zero_grads()
        for m in "lav":
            if self.Gproj[m] is not None:
                self.ops.zero_(self.Gproj[m])

    def unpack_grads(self, grads, accumulate=False):
        self.ops.batch_begin("unpack", "model")
        for n in ENC_NAMES:
            self.enc[n].unpack_grads(grads, "trans_%s." % n, accumulate)
        for m, g in self.gmu.items():
            g.unpack_grads(grads, "gmu_%s." % m, accumulate)
        self.head.unpack_grads(grads, accumulate)
        for m in "lav":
            if self.Gproj[m] is not None:
                gw = grads["proj_%s.weight" % m]
                self.ops.unpack_matrix(self.Gproj[m], gw.view(gw.shape[0], gw.shape[1]), accumulate=accumulate)
        self.ops.batch_end()

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
        for m in "lav":
            # transpose / text embed-dropout / zero-pad to n_vec (mmtr.py:741-761), then Conv1d(k=1) as a row GEMM (:748-750)
            drop = Drop(self.args.embed_dropout, seed, seed_ptr, 7) if (m == "l" and training and self.args.embed_dropout > 0) else None
            X = A.get("X_" + m, (M, self.Kp[m]), self.T_)
            o.stage_rows(feats[m], X, nv, drop)
            self.X[m] = X
            if self.Wproj[m] is not None:
                P[m] = A.get("P_" + m, (M, d.Dp), self.T_)
                o.gemm(X, self.Wproj[m], P[m], M, d.Dp, self.Kp[m])
            else:
                P[m] = X
        self.P = P
        h = {}
        for n, (qm, km) in WAVE1.items():
            h[n] = self.enc[n].forward(P[qm], B, nv, src_k=P[km], S=nv, training=training, seed=seed, seed_ptr=seed_ptr)
        cat = self.head.cat_buf(B)
        self.tops = {}
        for ci, m in enumerate(HEAD_ORDER):
            u, w, pn, qn = TARGETS[m]
            hp = self.enc[pn].forward(P[m], B, nv, src_k=h[u], S=nv, training=training, seed=seed, seed_ptr=seed_ptr)
            hq = self.enc[qn].forward(P[m], B, nv, src_k=h[w], S=nv, training=training, seed=seed, seed_ptr=seed_ptr)
            h[pn], h[qn] = hp, hq
            mid = self.gmu[m + "_m"].forward(h[u], h[w], M)                    # "GMU middle"
            a1 = A.get("a1_" + m, (M, d.Dp), self.T_)
            a2 = A.get("a2_" + m, (M, d.Dp), self.T_)
            o.add(hp, h[u], a1)                                                  # residual level 1 -> 2 (:799-800)
            o.add(hq, h[w], a2)
            top = self.gmu[m].forward(a1, a2, M, addend=mid)                     # "GMU top" + residual level 1 -> 3 (:803-806)
            self.tops[m] = top
            o.pool_fwd(top, B, nv, cat, ci * d.Dp)                               # h[0] + h[-1] (:808)
        self.h = h
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
        dP = {m: A.get("dP_" + m, (M, d.Dp), f32) for m in "lav"}
        dh = {n: A.get("dh_" + n, (M, d.Dp), f32) for n in WAVE1}
        for t in list(dP.values()) + list(dh.values()):
            o.zero_(t)
        dtop = A.get("dtop", (M, d.Dp), f32)
        da1 = A.get("da1", (M, d.Dp), f32)
        da2 = A.get("da2", (M, d.Dp), f32)
        for ci, m in reversed(list(enumerate(HEAD_ORDER))):
            u, w, pn, qn = TARGETS[m]
            o.zero_(dtop)
            o.pool_bwd(dcat, ci * d.Dp, B, nv, dtop)
            o.zero_(da1)
            o.zero_(da2)
            self.gmu[m].backward(dtop, da1, da2)                                 # d(p+u), d(q+w)
            self.gmu[m + "_m"].backward(dtop, dh[u], dh[w])                      # mid consumes u, w directly
            o.axpy_f32(da1, dh[u], True)
            o.axpy_f32(da2, dh[w], True)
            self.enc[qn].backward(da2, dP[m], dh[w])
            if on_done:
                on_done(qn)
            self.enc[pn].backward(da1, dP[m], dh[u])
            if on_done:
                on_done(pn)
        for n, (qm, km) in reversed(list(WAVE1.items())):
            self.enc[n].backward(dh[n], dP[qm], dP[km])
            if on_done:
                on_done(n)
        for m in "lav":
            if self.Wproj[m] is not None:
                g = self.shared.get("dPc", (M, d.Dp), self.T_)
                o.cast_drop(dP[m], g, None)
                o.gemm(g, self.X[m], self.Gproj[m], d.Dp, self.Kp[m], M, ta=1, tb=1, accumulate=True)     # dW = dP^T X
                if d_inputs is not None and m in d_inputs:
                    dX = self.shared.get("dX_" + m, (M, self.Kp[m]), f32)
                    o.gemm(g, self.Wproj[m], dX, M, self.Kp[m], d.Dp, tb=1)
                    self._unstage(m, dX, d_inputs[m])
            elif d_inputs is not None and m in d_inputs:
                self._unstage(m, dP[m], d_inputs[m])

    def backward_order(self):
        """encoder names in the order their gradients complete during backward (bucket launch order)"""
        order = []
        for m in reversed(HEAD_ORDER):
            u, w, pn, qn = TARGETS[m]
            order += [qn, pn]
        return order + list(reversed(list(WAVE1.keys())))

    def _unstage(self, m, g, dst):
        drop = Drop(self.args.embed_dropout, self.seed, self.seed_ptr, 7) if (m == "l" and self.training and self.args.embed_dropout > 0) else None
        self.ops.unstage_rows(g, dst, self.n_vec, False, drop)
