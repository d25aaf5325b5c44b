"""mmtrvapt (MultiprojectionMMTransformerGMUClf, models/mmtr.py:278-583): the 4-modality model (text, video, audio, poster) as an
explicit forward / backward schedule over the same engines as mmtrvat.  Differences to model_engine.MMTrVatEngine:
  * per-modality sequence lengths 512 / 200 / 200 (mmtr.py:371-373), so the crossmodal encoders run with T != S;
  * the wave-2 encoders are biprojection stacks (self-attention, then cross-attention with the same weights, 3 LayerNorms; :342-353);
  * wave-1 outputs reach the text-length (or audio / video-length) gated units through nn.Linear layers over the TIME axis
    (transfm_a2l, transfm_v2l : 200 -> 512;  transfm_l2a, transfm_l2v : 512 -> 200; :375-378, 507-508, 530, 553);
  * the poster vector joins at the head: proj_poster (Linear 4096 -> D, no bias, :310,486) and a 4-input TextShifting4Layer (:369,574).
BERT and the AudioEncoder sit upstream of this path (out of scope): text and audio arrive as feature sequences.
hybrid = True is not supported (broken in the reference).  Independent encoders of a wave run on side streams ("lanes",
see model_engine.Lanes): at B = 8 per GPU (README.md:30) a kernel has 1600 .. 4096 rows and fills a fraction of the 148 SMs."""
import os

import torch

from .engine import Arena, AudioEncoderEngine, Dims, EncoderEngine, HeadEngine, SeqGmuEngine, round_up
from .model_engine import ENC_NAMES, HEAD_ORDER, TARGETS, WAVE1, Lanes, attn_dropout_for
from .ops import Drop

BIPROJ = ("l_with_v2a", "l_with_a2v", "v_with_l2a", "v_with_a2l", "a_with_v2l", "a_with_l2v")
NV = {"l": 512, "a": 200, "v": 200}                                          # mmtr.py:371-373
# time-axis linears: name -> (source modality length, target modality length)
TRANSFM = {"a2l": ("a", "l"), "v2l": ("v", "l"), "l2a": ("l", "a"), "l2v": ("l", "v")}


def _mod_of(enc_name):
    """query modality of an encoder = its sequence length (trans_<q>_with_<src>)"""
    return enc_name[0]


class MMTrVaptEngine:
    def __init__(self, ops, args, dtype=torch.bfloat16):
        assert not getattr(args, "hybrid", False), "hybrid=True is broken in the reference (mmtr.py:572) and not supported"
        self.ops, self.args, self.T_ = ops, args, dtype
        D, H, L = args.hidden_sz, args.num_heads, args.layers
        self.d = Dims(D, H)
        self.orig = {"l": args.orig_d_l, "a": args.orig_d_a, "v": args.orig_d_v}
        self.Kp = {m: (self.d.Dp if self.orig[m] == D else round_up(self.orig[m], 64)) for m in "lav"}
        self.Kpp = round_up(args.orig_d_p, 64)
        self.lanes = Lanes(ops.device, int(os.environ.get("BPM_LANES", "6")))
        self.lane_shared = [Arena(ops) for _ in range(self.lanes.n)]
        self.shared = self.lane_shared[0]
        self.arena = Arena(ops)
        # lane of each encoder: wave 1 alternates, the two wave-2 encoders of a target modality sit on different lanes
        self.lane_of = {n: i % self.lanes.n for i, n in enumerate(WAVE1)}
        for i, m in enumerate(HEAD_ORDER):
            u, w, pn, qn = TARGETS[m]
            self.lane_of[pn], self.lane_of[qn] = (2 * i) % self.lanes.n, (2 * i + 1) % self.lanes.n
        self.mod_lane = {m: i % self.lanes.n for i, m in enumerate(HEAD_ORDER)}      # lane of a modality's staging / gated units
        self.enc = {}
        for i, n in enumerate(ENC_NAMES):
            self.enc[n] = EncoderEngine(ops, D, H, L, attn_dropout=attn_dropout_for(n, args), relu_dropout=args.relu_dropout,
                                        res_dropout=args.res_dropout, embed_dropout=args.embed_dropout, attn_mask=args.attn_mask,
                                        biprojection=n in BIPROJ, dtype=dtype, uid=i + 1, shared=self.lane_shared[self.lane_of[n]])
        self.gmu = {}
        for m in "lav":
            sh = self.lane_shared[self.mod_lane[m]]
            self.gmu[m + "_m"] = SeqGmuEngine(ops, D, dtype, True, sh, "gmu_%s_m" % m)
            self.gmu[m] = SeqGmuEngine(ops, D, dtype, True, sh, "gmu_%s" % m)
        self.head = HeadEngine(ops, D, 4, args.n_classes, out_dropout=args.out_dropout)
        # mmtr.py:307,452: the AudioEncoder (raw spectrogram (B, orig_d_a, T_raw) -> (B, orig_d_a, 200)); off = audio arrives as features
        self.audio = AudioEncoderEngine(ops, C=args.orig_d_a, Tp=NV["a"], dtype=dtype) if getattr(args, "audio_encoder", False) else None
        z = ops.zeros
        self.Wproj = {m: (z((self.d.Dp, self.Kp[m]), dtype) if self.orig[m] != D else None) for m in "lav"}
        self.Gproj = {m: (z((self.d.Dp, self.Kp[m]), torch.float32) if self.orig[m] != D else None) for m in "lav"}
        self.Wpost = z((self.d.Dp, self.Kpp), dtype)
        self.Gpost = z((self.d.Dp, self.Kpp), torch.float32)
        # time-axis linears: fp32 in the reference layout [T_out, T_in] (exact SIMT kernels = precision mode).  With bf16 storage the three
        # products run on tensor cores as per-sample GEMMs (y_b = W x_b + bias tile, dW += dy_b x_b^T, dx_b += W^T dy_b): W also as bf16,
        # the bias as a [T_out, Dp] bf16 tile the GEMM epilogue adds as its residual input (zero in the pad columns).
        self.Wt = {n: (z((NV[to], NV[ti]), torch.float32), z((NV[to],), torch.float32)) for n, (ti, to) in TRANSFM.items()}
        self.Gt = {n: (z((NV[to], NV[ti]), torch.float32), z((NV[to],), torch.float32)) for n, (ti, to) in TRANSFM.items()}
        self.tc_time = dtype == torch.bfloat16
        if self.tc_time:
            self.Wt_bf = {n: z((NV[to], NV[ti]), dtype) for n, (ti, to) in TRANSFM.items()}
            self.Bt = {n: z((NV[to], self.d.Dp), dtype) for n, (ti, to) in TRANSFM.items()}

    # ---------------------------------------------------------------- parameters
    def param_shapes(self):
        s = {"proj_poster.weight": (self.d.D, self.args.orig_d_p)}
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
        for n, (ti, to) in TRANSFM.items():
            s["transfm_%s.weight" % n] = (NV[to], NV[ti])
            s["transfm_%s.bias" % n] = (NV[to],)
        if self.audio is not None:
            for k, v in self.audio.param_shapes().items():
                s["audio_enc." + k] = v
        return s

    def unused_params(self):
        return ["proj_%s.weight" % m for m in "lav" if self.orig[m] == self.d.D]

    def pack(self, params):
        o = self.ops
        o.batch_begin("pack", "model4")
        for n in ENC_NAMES:
            self.enc[n].pack(params, "trans_%s." % n)
        for m, g in self.gmu.items():
            g.pack(params, "gmu_%s." % m)
        self.head.pack(params)
        for m in "lav":
            if self.Wproj[m] is not None:
                w = params["proj_%s.weight" % m]
                o.pack_matrix(w.view(w.shape[0], w.shape[1]), self.Wproj[m])
        o.pack_matrix(params["proj_poster.weight"], self.Wpost)
        if self.tc_time:
            for n in TRANSFM:
                o.pack_matrix(params["transfm_%s.weight" % n], self.Wt_bf[n])
        o.batch_end()
        if self.audio is not None:
            self.audio.pack(params, "audio_enc.")
        for n in TRANSFM:
            self.Wt[n][0].copy_(params["transfm_%s.weight" % n])
            self.Wt[n][1].copy_(params["transfm_%s.bias" % n])
            if self.tc_time:
                self.Bt[n][:, :self.d.D].copy_(params["transfm_%s.bias" % n].view(-1, 1).expand(-1, self.d.D))

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
        self.ops.zero_(self.Gpost)
        if self.audio is not None:
            self.audio.zero_grads()
        for gw, gb in self.Gt.values():
            self.ops.zero_(gw)
            self.ops.zero_(gb)

    def unpack_grads(self, grads, accumulate=False):
        for n in ENC_NAMES:
            self.enc[n].unpack_grads(grads, "trans_%s." % n, accumulate)
        self.unpack_misc(grads, accumulate)

    def unpack_misc(self, grads, accumulate=False):
        """gradients of everything but the encoders"""
        o = self.ops
        o.batch_begin("unpack", "misc4")
        for m, g in self.gmu.items():
            g.unpack_grads(grads, "gmu_%s." % m, accumulate)
        self.head.unpack_grads(grads, accumulate)
        for m in "lav":
            if self.Gproj[m] is not None:
                gw = grads["proj_%s.weight" % m]
                o.unpack_matrix(self.Gproj[m], gw.view(gw.shape[0], gw.shape[1]), accumulate=accumulate)
        o.unpack_matrix(self.Gpost, grads["proj_poster.weight"], accumulate=accumulate)
        o.batch_end()
        if self.audio is not None:
            self.audio.unpack_grads(grads, "audio_enc.", accumulate)
        for n in TRANSFM:
            for src, key in ((self.Gt[n][0], "weight"), (self.Gt[n][1], "bias")):
                dst = grads["transfm_%s.%s" % (n, key)]
                dst.copy_(dst + src if accumulate else src)

    # ---------------------------------------------------------------- pieces
    def _time_linear(self, name, h, B):
        ti, to = TRANSFM[name]
        Tin, Tout, Dp = NV[ti], NV[to], self.d.Dp
        y = self.arena.get("t_" + name, (B * Tout, Dp), self.T_)
        if self.tc_time:
            for b in range(B):                                   # A = W [T_out, T_in]; B operand = x_b stored [K = T_in, N = Dp]
                self.ops.gemm(self.Wt_bf[name], h[b * Tin:(b + 1) * Tin], y[b * Tout:(b + 1) * Tout], Tout, Dp, Tin, tb=1, residual=self.Bt[name])
        else:
            self.ops.timelin_fwd(h, self.Wt[name][0], self.Wt[name][1], y, B, Tin, Tout, self.d.D)
        return y

    def _time_linear_bwd(self, name, dy, h, dh, B):
        """dy fp32 [B*T_out, Dp] -> dh fp32 [B*T_in, Dp] += ; weight / bias gradients accumulate"""
        ti, to = TRANSFM[name]
        Tin, Tout, D, Dp, o = NV[ti], NV[to], self.d.D, self.d.Dp, self.ops
        if not self.tc_time:
            o.timelin_bwd(dy, h, self.Wt[name][0], dh, True, self.Gt[name][0], self.Gt[name][1], B, Tin, Tout, D)
            return
        dyb = self.arena.get("t_dy_" + name, (B * Tout, Dp), self.T_)
        o.cast_drop(dy, dyb)
        o.timelin_bwd(dy, h, self.Wt[name][0], None, False, None, self.Gt[name][1], B, Tin, Tout, D)      # bias gradient only
        for b in range(B):
            g, x = dyb[b * Tout:(b + 1) * Tout], h[b * Tin:(b + 1) * Tin]
            o.gemm(g, x, self.Gt[name][0], Tout, Tin, Dp, accumulate=True)                               # dW += dy_b x_b^T   (K = Dp)
            # dx_b += W^T dy_b: A = W stored [K = T_out, M = T_in], B operand = dy_b stored [K, N]; N = D keeps the pad columns untouched
            o.gemm(self.Wt_bf[name], g, dh[b * Tin:(b + 1) * Tin], Tin, D, Tout, ta=1, tb=1, accumulate=True)

    # ---------------------------------------------------------------- forward
    def forward(self, txt, img, audio, poster, training=True, seed=0, seed_ptr=None):
        """txt (B, T_l<=512, orig_d_l), img (B, T_v<=200, orig_d_v), audio (B, T_a<=200, orig_d_a), poster (B, orig_d_p): fp32 device tensors.
        Returns (logits [B, Cp], z [B, 4*Dp]) views of internal buffers."""
        o, d, A = self.ops, self.d, self.arena
        B = txt.shape[0]
        self.B, self.training, self.seed, self.seed_ptr = B, training, seed, seed_ptr
        feats = {"l": txt, "a": audio, "v": img}
        self.in_shapes = {m: tuple(feats[m].shape) for m in "lav"}
        P = {}
        self.X = {}
        ln = self.lanes
        ln.fork()
        for m in "lav":
            drop = Drop(self.args.embed_dropout, seed, seed_ptr, 7) if (m == "l" and training and self.args.embed_dropout > 0) else None
            X = A.get("X_" + m, (B * NV[m], self.Kp[m]), self.T_, zero=True)
            with ln.on(self.mod_lane[m]):
                if m == "a" and self.audio is not None:
                    self.audio.forward(feats[m], X)                             # raw spectrogram (B, C, T_raw) -> 200 feature rows (mmtr.py:452)
                else:
                    o.stage_rows(feats[m], X, NV[m], drop)                      # transpose / embed-dropout / zero-pad to num_vectors_m
                self.X[m] = X
                if self.Wproj[m] is not None:
                    P[m] = A.get("P_" + m, (B * NV[m], d.Dp), self.T_)
                    o.gemm(X, self.Wproj[m], P[m], B * NV[m], d.Dp, self.Kp[m])
                else:
                    P[m] = X
        self.P = P
        h = {}
        ln.barrier()
        for n, (qm, km) in WAVE1.items():
            with ln.on(self.lane_of[n]):
                h[n] = self.enc[n].forward(P[qm], B, NV[qm], src_k=P[km], S=NV[km], training=training, seed=seed, seed_ptr=seed_ptr)
        ln.barrier()                                                            # wave 2 reads wave-1 outputs of either lane
        for m in HEAD_ORDER:
            u, w, pn, qn = TARGETS[m]
            su, sw = NV[_mod_of(u)], NV[_mod_of(w)]
            with ln.on(self.lane_of[pn]):
                h[pn] = self.enc[pn].forward(P[m], B, NV[m], src_k=h[u], S=su, training=training, seed=seed, seed_ptr=seed_ptr)
            with ln.on(self.lane_of[qn]):
                h[qn] = self.enc[qn].forward(P[m], B, NV[m], src_k=h[w], S=sw, training=training, seed=seed, seed_ptr=seed_ptr)
        ln.barrier()
        cat = self.head.cat_buf(B)
        self.tsrc = {}
        for ci, m in enumerate(HEAD_ORDER):                                      # the three targets' gated units: one lane each
            u, w, pn, qn = TARGETS[m]
            Mm = B * NV[m]
            su, sw = NV[_mod_of(u)], NV[_mod_of(w)]
            hp, hq = h[pn], h[qn]
            with ln.on(self.mod_lane[m]):
                # wave-1 outputs at this target's length: through the time-axis linear when the lengths differ (:507-508,530,553)
                tname_u = "%s2%s" % (_mod_of(u), m) if su != NV[m] else None
                tname_w = "%s2%s" % (_mod_of(w), m) if sw != NV[m] else None
                tu = self._time_linear(tname_u, h[u], B) if tname_u else h[u]
                tw = self._time_linear(tname_w, h[w], B) if tname_w else h[w]
                self.tsrc[m] = (tname_u, tname_w, tu, tw)
                mid = self.gmu[m + "_m"].forward(tu, tw, Mm)                      # "GMU middle"
                a1 = A.get("a1_" + m, (Mm, d.Dp), self.T_)
                a2 = A.get("a2_" + m, (Mm, d.Dp), self.T_)
                o.add(hp, tu, a1)                                                # residual level 1 -> 2
                o.add(hq, tw, a2)
                top = self.gmu[m].forward(a1, a2, Mm, addend=mid)                # "GMU top" + residual level 1 -> 3
                o.pool_fwd(top, B, NV[m], cat, ci * d.Dp)                        # h[0] + h[-1]
        ln.join()
        # poster: Linear(orig_d_p -> D, no bias) straight into the 4th block of the head's input
        Xp = A.get("X_p", (B, self.Kpp), self.T_)
        o.stage_rows(poster.view(B, 1, -1), Xp, 1, None)
        self.Xp = Xp
        o.gemm(Xp, self.Wpost, cat[:, 3 * d.Dp:], B, d.Dp, self.Kpp)
        self.h = h
        return self.head.forward(B, training, seed, seed_ptr)

    def loss(self, logits, targets, pos_weight=None, grad_scale=1.0):
        return self.head.loss(logits, targets, pos_weight, grad_scale)

    # ---------------------------------------------------------------- backward
    def backward_order(self):
        """encoder names in the order their gradients complete during backward (gradient-bucket order of the Trainer)"""
        order = []
        for m in reversed(HEAD_ORDER):
            u, w, pn, qn = TARGETS[m]
            order += [qn, pn]
        return order + list(reversed(list(WAVE1.keys())))

    def backward(self, dlogits, d_inputs=None, on_done=None):
        """dlogits fp32 [B, Cp].  Parameter gradients accumulate in the padded buffers (see unpack_grads).
        d_inputs: optional dict m -> fp32 tensor shaped like the input features ("l", "a", "v"), overwritten with input gradients.
        on_done(name): called when an encoder's parameter gradients are complete."""
        o, d, A, B, h = self.ops, self.d, self.arena, self.B, self.h
        f32 = torch.float32
        dcat = self.head.backward(dlogits)
        # poster projection: dW = dpost^T Xp
        gpo = A.get("dpost", (B, d.Dp), self.T_)
        gpo.copy_(dcat[:, 3 * d.Dp:])                                            # (strided slice + cast: a torch copy, off the hot path)
        o.gemm(gpo, self.Xp, self.Gpost, d.Dp, self.Kpp, B, ta=1, tb=1, accumulate=True)
        ln = self.lanes
        # every lane accumulates the projection-output gradients in its own buffers (two encoders of different lanes share a query
        # stream); they are summed once at the end
        dPl = [{m: A.get("dP%d_%s" % (k, m), (B * NV[m], d.Dp), f32) for m in "lav"} for k in range(ln.n)]
        dP = dPl[0]
        dh = {n: A.get("dh_" + n, (B * NV[_mod_of(n)], d.Dp), f32) for n in WAVE1}
        for t in [t for k in range(ln.n) for t in dPl[k].values()] + list(dh.values()):
            o.zero_(t)
        da = {}
        ln.fork()
        for ci, m in reversed(list(enumerate(HEAD_ORDER))):                      # gated fusion units of the three targets: one lane each
            u, w, pn, qn = TARGETS[m]
            Mm = B * NV[m]
            tname_u, tname_w, tu, tw = self.tsrc[m]
            with ln.on(self.mod_lane[m]):
                dtop = A.get("dtop_" + m, (Mm, d.Dp), f32)
                da1 = A.get("da1_" + m, (Mm, d.Dp), f32)
                da2 = A.get("da2_" + m, (Mm, d.Dp), f32)
                dtu = A.get("dtu_" + m, (Mm, d.Dp), f32) if tname_u else dh[u]   # gradient wrt the (time-mapped) wave-1 outputs
                dtw = A.get("dtw_" + m, (Mm, d.Dp), f32) if tname_w else dh[w]
                for t in [dtop, da1, da2] + ([dtu] if tname_u else []) + ([dtw] if tname_w else []):
                    o.zero_(t)
                o.pool_bwd(dcat, ci * d.Dp, B, NV[m], dtop)
                self.gmu[m].backward(dtop, da1, da2)                             # d(p + tu), d(q + tw)
                self.gmu[m + "_m"].backward(dtop, dtu, dtw)
                o.axpy_f32(da1, dtu, True)
                o.axpy_f32(da2, dtw, True)
                if tname_u:
                    self._time_linear_bwd(tname_u, dtu, h[u], dh[u], B)
                if tname_w:
                    self._time_linear_bwd(tname_w, dtw, h[w], dh[w], B)
                da[m] = (da1, da2)
        ln.barrier()
        for m in reversed(HEAD_ORDER):
            u, w, pn, qn = TARGETS[m]
            for n, g_, dsrc in ((qn, da[m][1], dh[w]), (pn, da[m][0], dh[u])):
                with ln.on(self.lane_of[n]):
                    self.enc[n].backward(g_, dPl[self.lane_of[n]][m], dsrc)
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
                sh = self.lane_shared[self.mod_lane[m]]
                if self.Wproj[m] is not None:
                    g = sh.get("dPc_" + m, (B * NV[m], d.Dp), self.T_)
                    o.cast_drop(dP[m], g, None)
                    o.gemm(g, self.X[m], self.Gproj[m], d.Dp, self.Kp[m], B * NV[m], ta=1, tb=1, accumulate=True)     # dW = dP^T X
                    enc_a = m == "a" and self.audio is not None               # the audio encoder sits upstream of this projection
                    if enc_a or (d_inputs is not None and m in d_inputs):
                        dX = sh.get("dX_" + m, (B * NV[m], self.Kp[m]), f32)
                        o.gemm(g, self.Wproj[m], dX, B * NV[m], self.Kp[m], d.Dp, tb=1)
                        if enc_a:
                            self.audio.backward(dX)
                        else:
                            self._unstage(m, dX, d_inputs[m])
                elif m == "a" and self.audio is not None:
                    self.audio.backward(dP[m])
                elif d_inputs is not None and m in d_inputs:
                    self._unstage(m, dP[m], d_inputs[m])
        ln.join()

    def _unstage(self, m, g, dst):
        drop = Drop(self.args.embed_dropout, self.seed, self.seed_ptr, 7) if (m == "l" and self.training and self.args.embed_dropout > 0) else None
        self.ops.unstage_rows(g, dst, NV[m], False, drop)
