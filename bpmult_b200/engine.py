"""Host-side executor of the BPMulT fusion trunk: explicit forward / backward kernel sequences over the padded HBM
layout, with hand-derived gradients (no autograd inside).  Every device op goes through an `ops` object
(bpmult_b200.ops.CudaOps -> our CUDA kernels via the C ABI).

Layout: activations are batch-major rows r = b*T + t with pitch Dp = round_up(D, 64);
per-head tensors use pitch HP = H * dhp with dhp = round_up(D/H, 32); the FFN hidden uses FP = round_up(4D, 64).
Pad columns are zero by construction (zero-padded weights / biases / LayerNorm affine), so no kernel special-cases them.
Storage type T is bf16 (tensor-core mode) or fp32 (precision mode); the residual stream, all gradients that are
accumulated from several consumers, LayerNorm statistics, the softmax LSE and the whole [B, D] head stay fp32.

Reference semantics implemented here (paths under /root/reference/bpmult):
  TransformerEncoder.forward        models/transformer.py:52-93
  TransformerEncoderLayer.forward   models/transformer.py:141-195 (cross / self-only / biprojection variants)
  MultiheadAttention.forward        models/multihead_attention.py:52-135
  GatedMultimodalLayerFeatures      models/mmtr.py:179-195;  TextShifting3/4Layer models/mmtr.py:197-247
  mmtrvat forward                   models/mmtr.py:735-866;  BCEWithLogitsLoss train.py:99-106,333
"""
import math

import torch

from .ops import Drop


def round_up(x, m):
    return (x + m - 1) // m * m


class Dims:
    def __init__(self, D, H):
        assert D % H == 0, "embed_dim must be divisible by num_heads"            # multihead_attention.py:22
        self.D, self.H = D, H
        self.dh = D // H
        self.dhp = round_up(self.dh, 32)
        self.Dp = round_up(D, 64)
        self.HP = H * self.dhp
        self.F = 4 * D
        self.FP = round_up(4 * D, 64)
        self.scaling = self.dh ** -0.5
        assert self.Dp <= 1024, "hidden sizes above 1024 are not supported by the LayerNorm kernel"


def sinusoid_table(num_pos, D, Dp, device):
    """models/position_embedding.py:44-60, evaluated with the same fp32 torch ops as the reference (so the table is
    bit-identical), zero-padded to Dp columns; row 0 (padding position) is zero."""
    half = D // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(num_pos, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(num_pos, -1)
    out = torch.zeros(num_pos, Dp)
    out[:, :e.shape[1]] = e
    out[0, :] = 0
    return out.to(device)


def carve(ops, dtype, shapes):
    """one flat zeroed buffer + a view per shape (128-element aligned starts): lets a whole group be cleared with ONE memset"""
    offs, total = [], 0
    for shp in shapes:
        n = 1
        for x in shp:
            n *= x
        offs.append(total)
        total += round_up(n, 128)
    flat = ops.zeros((total,), dtype)
    views = []
    for shp, o in zip(shapes, offs):
        n = 1
        for x in shp:
            n *= x
        views.append(flat[o:o + n].view(shp))
    return flat, views


class Arena:
    """Named, persistent device buffers (static addresses => the step can be captured in a CUDA graph)."""

    def __init__(self, ops):
        self.ops = ops
        self.bufs = {}

    def get(self, key, shape, dtype, zero=False):
        shape = tuple(shape)
        k = (key, shape, dtype)                 # one buffer per (name, shape): addresses never change once created
        t = self.bufs.get(k)
        if t is None:
            t = self.ops.zeros(shape, dtype) if zero else self.ops.empty(shape, dtype)
            self.bufs[k] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


# ============================================================================================== encoder
class EncoderEngine:
    """One TransformerEncoder (L layers + final LayerNorm).  `uid` separates its dropout streams."""
    _go_valid = False       # the shared "go" buffer already holds the next block's gradient operand (fused into a LayerNorm backward)
    fuse_cast = True        # LayerNorm backward also writes the next block's (dropped, storage-type) gradient operand
    fold_kv = True          # hoist the K / V LayerNorm out of the layer loop (x_hat once per encoder, affine folded into in_proj)

    # parameter table: name-suffix -> (kind, packer)
    def __init__(self, ops, D, H, L, attn_dropout=0.0, relu_dropout=0.0, res_dropout=0.0, embed_dropout=0.0, attn_mask=False,
                 biprojection=False, dtype=torch.bfloat16, uid=0, shared=None, with_embed=True, with_final_ln=True):
        self.ops, self.d, self.L = ops, Dims(D, H), L
        self.with_embed, self.with_final_ln = with_embed, with_final_ln        # False: run bare layers (standalone layer module)
        self._mask_override = None
        self.p_attn, self.p_relu, self.p_res, self.p_embed = attn_dropout, relu_dropout, res_dropout, embed_dropout
        self.attn_mask, self.biproj, self.T_ = attn_mask, biprojection, dtype
        self.uid = uid
        self.n_ln = 3 if biprojection else 2
        self.kv_ln = 1 if biprojection else 0                     # which LayerNorm the crossmodal K / V inputs go through
        self.arena = Arena(ops)
        self.shared = shared if shared is not None else Arena(ops)     # scratch shared between encoders (backward temporaries)
        self.pe = None
        # pruned query rows (crossmodal stacks only; see MMTrVatEngine): the T rows handed to forward() are rows `prune_pos` (0-based time
        # steps) of a longer sequence -- they take those positions' embeddings, attend without the index mask, and, under the future
        # mask of the full sequence (|S - T_full| = 0), row 0 sees key 0 alone: its attention output is the (dropout-scaled) value row 0.
        self.prune_pos, self.prune_row0 = None, False
        self._alloc_weights()

    # ---------------------------------------------------------------- weights
    def param_shapes(self):
        D = self.d.D
        s = {}
        for i in range(self.L):
            p = "layers.%d." % i
            s[p + "self_attn.in_proj_weight"] = (3 * D, D)
            s[p + "self_attn.in_proj_bias"] = (3 * D,)
            s[p + "self_attn.out_proj.weight"] = (D, D)
            s[p + "self_attn.out_proj.bias"] = (D,)
            s[p + "fc1.weight"] = (4 * D, D)
            s[p + "fc1.bias"] = (4 * D,)
            s[p + "fc2.weight"] = (D, 4 * D)
            s[p + "fc2.bias"] = (D,)
            for j in range(self.n_ln):
                s[p + "layer_norms.%d.weight" % j] = (D,)
                s[p + "layer_norms.%d.bias" % j] = (D,)
        if self.with_final_ln:
            s["layer_norm.weight"] = (D,)
            s["layer_norm.bias"] = (D,)
        return s

    def _alloc_weights(self):
        d, T_, f32 = self.d, self.T_, torch.float32
        z = self.ops.zeros
        self.W, self.G = [], []                     # packed weights / padded fp32 gradient accumulators per layer
        gshapes = []
        for _ in range(self.L):
            gshapes += [(3 * d.HP, d.Dp), (3 * d.HP,), (d.Dp, d.HP), (d.Dp,), (d.FP, d.Dp), (d.FP,), (d.Dp, d.FP), (d.Dp,)] + [(d.Dp,)] * (2 * self.n_ln)
        gshapes += [(d.Dp,), (d.Dp,)]
        gshapes += [(self.L * d.HP, d.Dp), (self.L * d.HP, d.Dp), (self.L * d.HP,), (self.L * d.HP,)]   # folded K / V projections of ALL layers
        self.G_flat, gv = carve(self.ops, f32, gshapes)      # every gradient accumulator of this encoder: zeroed with one memset
        it = iter(gv)
        for _ in range(self.L):
            w = dict(Wqkv=z((3 * d.HP, d.Dp), T_), bqkv=z((3 * d.HP,), f32), Wo=z((d.Dp, d.HP), T_), bo=z((d.Dp,), f32),
                     W1=z((d.FP, d.Dp), T_), b1=z((d.FP,), f32), W2=z((d.Dp, d.FP), T_), b2=z((d.Dp,), f32),
                     ln_g=[z((d.Dp,), f32) for _ in range(self.n_ln)], ln_b=[z((d.Dp,), f32) for _ in range(self.n_ln)])
            g = dict(Wqkv=next(it), bqkv=next(it), Wo=next(it), bo=next(it), W1=next(it), b1=next(it), W2=next(it), b2=next(it))
            g["ln_g"] = [next(it) for _ in range(self.n_ln)]
            g["ln_b"] = [next(it) for _ in range(self.n_ln)]
            self.W.append(w)
            self.G.append(g)
        self.Wf = dict(g=z((d.Dp,), f32), b=z((d.Dp,), f32))
        self.Gf = dict(g=next(it), b=next(it))
        # folded K / V projections (x_hat operands) of all L layers, stacked: one GEMM per encoder produces every layer's K (V)
        self.Gkv = dict(Wk=next(it), Wv=next(it), bk=next(it), bv=next(it))
        self.Wkv = dict(Wk=z((self.L * d.HP, d.Dp), T_), Wv=z((self.L * d.HP, d.Dp), T_), bk=z((self.L * d.HP,), f32), bv=z((self.L * d.HP,), f32))
        # unit LayerNorm affine (x_hat of the K / V inputs) and a sink for its unused affine gradients
        self.unit_g, self.unit_b, self.ln_sink = z((d.Dp,), f32), z((d.Dp,), f32), z((2, d.Dp), f32)
        self.unit_g[:d.D] = 1.0
        self.P = [None] * self.L                    # per layer: reference-layout (in_proj_weight, in_proj_bias, K/V LayerNorm gamma, beta)

    def pack(self, params, pfx=""):
        """reference-layout fp32 parameters -> zero-padded kernel operands (run whenever the parameters changed)."""
        o, d = self.ops, self.d
        hm = (d.dh, d.dhp)
        o.batch_begin("pack", ("enc", self.uid, pfx))
        for l in range(self.L):
            p = "%slayers.%d." % (pfx, l)
            w = self.W[l]
            self.pack_attention(l, params[p + "self_attn.in_proj_weight"], params[p + "self_attn.in_proj_bias"],
                                params[p + "self_attn.out_proj.weight"], params[p + "self_attn.out_proj.bias"])
            o.pack_matrix(params[p + "fc1.weight"], w["W1"])
            o.pack_matrix(params[p + "fc1.bias"].view(1, -1), w["b1"].view(1, -1))
            o.pack_matrix(params[p + "fc2.weight"], w["W2"])
            o.pack_matrix(params[p + "fc2.bias"].view(1, -1), w["b2"].view(1, -1))
            for j in range(self.n_ln):
                o.pack_matrix(params[p + "layer_norms.%d.weight" % j].view(1, -1), w["ln_g"][j].view(1, -1))
                o.pack_matrix(params[p + "layer_norms.%d.bias" % j].view(1, -1), w["ln_b"][j].view(1, -1))
            self.P[l] = (params[p + "self_attn.in_proj_weight"], params[p + "self_attn.in_proj_bias"],
                         params[p + "layer_norms.%d.weight" % self.kv_ln], params[p + "layer_norms.%d.bias" % self.kv_ln])
        if self.with_final_ln:
            o.pack_matrix(params[pfx + "layer_norm.weight"].view(1, -1), self.Wf["g"].view(1, -1))
            o.pack_matrix(params[pfx + "layer_norm.bias"].view(1, -1), self.Wf["b"].view(1, -1))
        o.batch_end()
        if self.fold_kv:
            o.fold_batch_begin("fwd")
            for l in range(self.L):                 # K' = W_k diag(gamma), b_k' = b_k + W_k beta  (and the same for V)
                ipw, ipb, g_, b_ = self.P[l]
                r = slice(l * d.HP, (l + 1) * d.HP)
                o.ln_fold_fwd(ipw[d.D:2 * d.D], ipb[d.D:2 * d.D], g_, b_, self.Wkv["Wk"][r], self.Wkv["bk"][r], row_map=hm)
                o.ln_fold_fwd(ipw[2 * d.D:], ipb[2 * d.D:], g_, b_, self.Wkv["Wv"][r], self.Wkv["bv"][r], row_map=hm)
            o.fold_batch_end()

    def pack_attention(self, l, ipw, ipb, ow, ob):
        o, d, w = self.ops, self.d, self.W[l]
        hm = (d.dh, d.dhp)
        for j in range(3):          # rows [0:D] = Q, [D:2D] = K, [2D:3D] = V  (multihead_attention.py:137-158)
            o.pack_matrix(ipw[j * d.D:(j + 1) * d.D], w["Wqkv"][j * d.HP:(j + 1) * d.HP], row_map=hm)
            o.pack_matrix(ipb[j * d.D:(j + 1) * d.D].view(1, -1), w["bqkv"][j * d.HP:(j + 1) * d.HP].view(1, -1), col_map=hm)
        o.pack_matrix(ow, w["Wo"], col_map=hm)
        o.pack_matrix(ob.view(1, -1), w["bo"].view(1, -1))

    def unpack_attention_grads(self, l):
        """fresh reference-layout gradient tensors (in_proj_weight, in_proj_bias, out_proj.weight, out_proj.bias) of layer l"""
        o, d, g = self.ops, self.d, self.G[l]
        hm = (d.dh, d.dhp)
        ipw, ipb = o.zeros((3 * d.D, d.D), torch.float32), o.zeros((3 * d.D,), torch.float32)
        ow, ob = o.zeros((d.D, d.D), torch.float32), o.zeros((d.D,), torch.float32)
        for j in range(3):
            o.unpack_matrix(g["Wqkv"][j * d.HP:(j + 1) * d.HP], ipw[j * d.D:(j + 1) * d.D], row_map=hm)
            o.unpack_matrix(g["bqkv"][j * d.HP:(j + 1) * d.HP].view(1, -1), ipb[j * d.D:(j + 1) * d.D].view(1, -1), col_map=hm)
        o.unpack_matrix(g["Wo"], ow, col_map=hm)
        o.unpack_matrix(g["bo"].view(1, -1), ob.view(1, -1))
        return ipw, ipb, ow, ob

    def zero_grads(self):
        self.ops.zero_(self.G_flat)

    def unpack_grads(self, grads, pfx="", accumulate=False):
        """padded fp32 gradient accumulators -> reference-layout gradient tensors `grads[name]`."""
        o, d = self.ops, self.d
        hm = (d.dh, d.dhp)
        o.batch_begin("unpack", ("enc", self.uid, pfx))
        for l in range(self.L):
            p = "%slayers.%d." % (pfx, l)
            g = self.G[l]
            ipw, ipb = grads[p + "self_attn.in_proj_weight"], grads[p + "self_attn.in_proj_bias"]
            for j in range(3):
                o.unpack_matrix(g["Wqkv"][j * d.HP:(j + 1) * d.HP], ipw[j * d.D:(j + 1) * d.D], row_map=hm, accumulate=accumulate)
                o.unpack_matrix(g["bqkv"][j * d.HP:(j + 1) * d.HP].view(1, -1), ipb[j * d.D:(j + 1) * d.D].view(1, -1), col_map=hm,
                                accumulate=accumulate)
            o.unpack_matrix(g["Wo"], grads[p + "self_attn.out_proj.weight"], col_map=hm, accumulate=accumulate)
            o.unpack_matrix(g["bo"].view(1, -1), grads[p + "self_attn.out_proj.bias"].view(1, -1), accumulate=accumulate)
            o.unpack_matrix(g["W1"], grads[p + "fc1.weight"], accumulate=accumulate)
            o.unpack_matrix(g["b1"].view(1, -1), grads[p + "fc1.bias"].view(1, -1), accumulate=accumulate)
            o.unpack_matrix(g["W2"], grads[p + "fc2.weight"], accumulate=accumulate)
            o.unpack_matrix(g["b2"].view(1, -1), grads[p + "fc2.bias"].view(1, -1), accumulate=accumulate)
            for j in range(self.n_ln):
                o.unpack_matrix(g["ln_g"][j].view(1, -1), grads[p + "layer_norms.%d.weight" % j].view(1, -1), accumulate=accumulate)
                o.unpack_matrix(g["ln_b"][j].view(1, -1), grads[p + "layer_norms.%d.bias" % j].view(1, -1), accumulate=accumulate)
        if self.with_final_ln:
            o.unpack_matrix(self.Gf["g"].view(1, -1), grads[pfx + "layer_norm.weight"].view(1, -1), accumulate=accumulate)
            o.unpack_matrix(self.Gf["b"].view(1, -1), grads[pfx + "layer_norm.bias"].view(1, -1), accumulate=accumulate)
        o.batch_end()

    # ---------------------------------------------------------------- helpers
    def _site(self, layer, idx):
        return (self.uid << 20) | ((layer + 1) << 8) | idx

    def _drop(self, p, layer, idx):
        if not self.training or p <= 0.0:
            return None
        return Drop(p, self.seed, self.seed_ptr, self._site(layer, idx))

    def _mask_off(self, T, S):
        # models/transformer.py:209-216: masked iff j - i >= 1 + |S - T|  <=>  visible iff j <= i + |S - T|
        if self._mask_override is not None:
            return self._mask_override
        if self.prune_pos is not None:
            return -1                                       # the visible set of the pruned rows is handled in _attn_fwd / _attn_bwd
        return abs(S - T) if self.attn_mask else -1

    # ---------------------------------------------------------------- attention block (in-proj, attention, out-proj + residual)
    def _attn_fwd(self, l, blk, q_in, k_in, v_in, B, T, S, x_res, x_out, res_drop=True, folded=False):
        """folded: k_in / v_in ARE this layer's K / V (column slices of the all-layers projections of x_hat, made once per encoder)"""
        o, d, A, w = self.ops, self.d, self.arena, self.W[l]
        M, Ms = B * T, B * S
        key = "L%d.%s." % (l, blk)
        q = A.get(key + "q", (M, d.HP), self.T_)
        a = A.get(key + "a", (M, d.HP), self.T_)
        lse = A.get(key + "lse", (B * d.H * T,), torch.float32)
        Wq, Wk, Wv = w["Wqkv"][:d.HP], w["Wqkv"][d.HP:2 * d.HP], w["Wqkv"][2 * d.HP:]
        bq, bk, bv = w["bqkv"][:d.HP], w["bqkv"][d.HP:2 * d.HP], w["bqkv"][2 * d.HP:]
        o.gemm(q_in, Wq, q, M, d.HP, d.Dp, bias=bq, alpha=d.scaling)          # q = (x Wq^T + bq) * dh^-0.5  (:86)
        if folded:                                                              # this layer's columns of the all-layers K / V projections
            k, v = k_in, v_in
        else:
            k = A.get(key + "k", (Ms, d.HP), self.T_)
            o.gemm(k_in, Wk, k, Ms, d.HP, d.Dp, bias=bk)
            v = A.get(key + "v", (Ms, d.HP), self.T_)
            o.gemm(v_in, Wv, v, Ms, d.HP, d.Dp, bias=bv)
        adrop = self._drop(self.p_attn, l, 10 + (blk == "x"))
        bits = A.get(key + "bits", (B * d.H * T * ((S + 31) // 32),), torch.int32) if adrop is not None else None
        o.xattn_fwd(q, k, v, a, lse, B, T, S, d.H, d.dh, d.dhp, mask_off=self._mask_off(T, S), drop=adrop, drop_bits=bits)
        keep0 = None
        if self.prune_row0 and blk == "x":
            # time step 0 under the future mask sees key 0 only: softmax = 1, output = dropout(1) * v[key 0] per head
            keep0 = A.get(key + "keep0", (B, d.H, 1), torch.float32)
            if adrop is not None:
                keep0.copy_((torch.rand(B, d.H, 1, device=keep0.device) >= self.p_attn).float() / (1.0 - self.p_attn))
            else:
                keep0.fill_(1.0)
            v0 = v.view(B, S, d.H, d.dhp)[:, 0].float() * keep0
            a.view(B, T, d.HP)[:, 0].copy_(v0.view(B, d.HP))
        # x_out = x_res + dropout(a Wo^T + bo)                                   (transformer.py:174-175)
        o.gemm(a, w["Wo"], x_out, M, d.Dp, d.HP, bias=w["bo"], drop=self._drop(self.p_res if res_drop else 0.0, l, 20 + (blk == "x")),
               residual=x_res)
        return dict(q=q, k=k, v=v, a=a, lse=lse, q_in=q_in, k_in=k_in, v_in=v_in, S=S, bits=bits, folded=folded, keep0=keep0)

    def _attn_bwd(self, l, blk, sv, B, T, gx, res_drop=True, dkv=None):
        """gx: fp32 [M, Dp] gradient wrt the block output x_out (= also flows to x_res unchanged).
        Returns (dq_in, dk_in, dv_in) in storage type (gradients wrt the projection inputs).  Folded K / V projections: dK / dV go
        into dkv = (this layer's column slices of the all-layers dK / dV buffers) and (dq_in, None, None) is returned; their weight
        and input gradients are taken for all layers at once at the end of the encoder backward."""
        o, d, w, g, Sh = self.ops, self.d, self.W[l], self.G[l], self.shared
        S = sv["S"]
        M, Ms = B * T, B * S
        Wq, Wk, Wv = w["Wqkv"][:d.HP], w["Wqkv"][d.HP:2 * d.HP], w["Wqkv"][2 * d.HP:]
        gWq, gWk, gWv = g["Wqkv"][:d.HP], g["Wqkv"][d.HP:2 * d.HP], g["Wqkv"][2 * d.HP:]
        gbq, gbk, gbv = g["bqkv"][:d.HP], g["bqkv"][d.HP:2 * d.HP], g["bqkv"][2 * d.HP:]
        go = Sh.get("go", (M, d.Dp), self.T_)
        if not self._go_valid:                                                  # (else: fused into the LayerNorm backward that last wrote gx)
            o.cast_drop(gx, go, self._drop(self.p_res if res_drop else 0.0, l, 20 + (blk == "x")))   # grad wrt out_proj output
        self._go_valid = False
        o.gemm(go, sv["a"], g["Wo"], d.Dp, d.HP, M, ta=1, tb=1, accumulate=True, colsum=g["bo"])   # dWo = go^T a, dbo = colsum(go)
        da = Sh.get("da", (M, d.HP), self.T_)
        o.gemm(go, w["Wo"], da, M, d.HP, d.Dp, tb=1)                                           # da = go Wo
        dq = Sh.get("dq", (M, d.HP), self.T_)
        if sv["folded"]:
            dk, dv = dkv
        else:
            dk = Sh.get("dk", (Ms, d.HP), self.T_)
            dv = Sh.get("dv", (Ms, d.HP), self.T_)
        # [0] rowsum(dO*O), [1] lse*log2e (+ the fp32 dQ accumulator of the head-dim-128 tensor-core backward)
        delta = Sh.get("delta", (o.xattn_bwd_workspace(self.T_, B, T, S, d.H, d.dh, d.dhp),), torch.float32)
        g0 = None
        if sv.get("keep0") is not None:
            # row 0 (single visible key): no gradient reaches the scores; its output gradient goes to value row 0 alone.  With that row
            # of dO cleared the kernel adds nothing for it (dP = 0, delta = 0 => dS = 0).
            da0 = da.view(B, T, d.HP)[:, 0]
            g0 = da0.float().view(B, d.H, d.dhp) * sv["keep0"]
            da0.zero_()
        o.xattn_bwd(sv["q"], sv["k"], sv["v"], sv["a"], da, sv["lse"], delta, dq, d.scaling, dk, dv, B, T, S, d.H, d.dh, d.dhp,
                    mask_off=self._mask_off(T, S), drop=self._drop(self.p_attn, l, 10 + (blk == "x")), drop_bits=sv["bits"])
        if g0 is not None:
            dv0 = dv.view(B, S, d.HP)[:, 0]
            dv0.copy_((dv0.float() + g0.view(B, d.HP)).to(dv.dtype))
        # dq already carries the dh^-0.5 factor => it is the gradient wrt (x Wq^T + bq)
        o.gemm(dq, sv["q_in"], gWq, d.HP, d.Dp, M, ta=1, tb=1, accumulate=True, colsum=gbq)
        dq_in = Sh.get("dq_in", (M, d.Dp), self.T_)
        o.gemm(dq, Wq, dq_in, M, d.Dp, d.HP, tb=1)
        if sv["folded"]:
            return dq_in, None, None
        o.gemm(dk, sv["k_in"], gWk, d.HP, d.Dp, Ms, ta=1, tb=1, accumulate=True, colsum=gbk)
        o.gemm(dv, sv["v_in"], gWv, d.HP, d.Dp, Ms, ta=1, tb=1, accumulate=True, colsum=gbv)
        dk_in = Sh.get("dk_in", (Ms, d.Dp), self.T_)
        dv_in = Sh.get("dv_in", (Ms, d.Dp), self.T_)
        o.gemm(dk, Wk, dk_in, Ms, d.Dp, d.HP, tb=1)
        o.gemm(dv, Wv, dv_in, Ms, d.Dp, d.HP, tb=1)
        return dq_in, dk_in, dv_in

    # ---------------------------------------------------------------- FFN block
    def _ffn_fwd(self, l, ln_idx, x_in, x_out, M):
        o, d, A, w = self.ops, self.d, self.arena, self.W[l]
        key = "L%d.ffn." % l
        hn = A.get(key + "hn", (M, d.Dp), self.T_)
        mean = A.get(key + "mean", (M,), torch.float32)
        rstd = A.get(key + "rstd", (M,), torch.float32)
        h = A.get(key + "h", (M, d.FP), self.T_)
        o.layernorm_fwd(x_in, w["ln_g"][ln_idx], w["ln_b"][ln_idx], d.D, hn, mean, rstd)
        o.gemm(hn, w["W1"], h, M, d.FP, d.Dp, bias=w["b1"], act=1, drop=self._drop(self.p_relu, l, 30))       # relu + dropout (:186-187)
        o.gemm(h, w["W2"], x_out, M, d.Dp, d.FP, bias=w["b2"], drop=self._drop(self.p_res, l, 31), residual=x_in)  # (:188-190)
        return dict(hn=hn, mean=mean, rstd=rstd, h=h, x_in=x_in, ln=ln_idx)

    def _ffn_bwd(self, l, sv, M, gx):
        """gx (fp32, in/out): on entry d/d x_out, on exit d/d x_in."""
        o, d, w, g, Sh = self.ops, self.d, self.W[l], self.G[l], self.shared
        g2 = Sh.get("go", (M, d.Dp), self.T_)
        if not self._go_valid:
            o.cast_drop(gx, g2, self._drop(self.p_res, l, 31))
        self._go_valid = False
        o.gemm(g2, sv["h"], g["W2"], d.Dp, d.FP, M, ta=1, tb=1, accumulate=True, colsum=g["b2"])   # dW2 = g2^T h, db2 = colsum(g2)
        dh = Sh.get("dh", (M, d.FP), self.T_)
        keep = 1.0 / (1.0 - self.p_relu) if (self.training and self.p_relu > 0) else 1.0
        o.gemm(g2, w["W2"], dh, M, d.FP, d.Dp, tb=1, gate=sv["h"], gate_scale=keep)            # through dropout + relu
        o.gemm(dh, sv["hn"], g["W1"], d.FP, d.Dp, M, ta=1, tb=1, accumulate=True, colsum=g["b1"])  # dW1 = dh^T hn, db1 = colsum(dh)
        dhn = Sh.get("dq_in", (M, d.Dp), self.T_)
        o.gemm(dh, w["W1"], dhn, M, d.Dp, d.FP, tb=1)
        # the attention block of this layer is next: its out_proj gradient operand (dropout site 20 / 21) leaves this kernel too
        blk_x = getattr(self, "cross", False)
        fuse = self.fuse_cast
        o.layernorm_bwd(dhn, sv["x_in"], sv["mean"], sv["rstd"], w["ln_g"][sv["ln"]], d.D, gx, True, g["ln_g"][sv["ln"]], g["ln_b"][sv["ln"]],
                        cast_out=g2 if fuse else None, cast_drop=self._drop(self.p_res, l, 20 + int(blk_x)) if fuse else None)
        self._go_valid = fuse

    # ---------------------------------------------------------------- LN helper with saved stats
    def _ln_fwd(self, key, x, l, idx, rows):
        o, d, A, w = self.ops, self.d, self.arena, self.W[l]
        y = A.get(key + "y", (rows, d.Dp), self.T_)
        mean = A.get(key + "mean", (rows,), torch.float32)
        rstd = A.get(key + "rstd", (rows,), torch.float32)
        o.layernorm_fwd(x, w["ln_g"][idx], w["ln_b"][idx], d.D, y, mean, rstd)
        return dict(y=y, mean=mean, rstd=rstd, x=x, idx=idx)

    def _ln_bwd(self, l, sv, dy, dx, accumulate=True, cast_next=False):
        """cast_next: dx is now final for this layer -> also emit the FFN gradient operand of layer l-1 (dropout site 31)"""
        w, g = self.W[l], self.G[l]
        cast_out = cast_drop = None
        if cast_next and l > 0 and self.fuse_cast:
            cast_out = self.shared.get("go", tuple(dx.shape), self.T_)
            cast_drop = self._drop(self.p_res, l - 1, 31)
            self._go_valid = True
        self.ops.layernorm_bwd(dy, sv["x"], sv["mean"], sv["rstd"], w["ln_g"][sv["idx"]], self.d.D, dx, accumulate, g["ln_g"][sv["idx"]],
                               g["ln_b"][sv["idx"]], cast_out=cast_out, cast_drop=cast_drop)

    # ---------------------------------------------------------------- forward
    def forward(self, src_q, B, T, src_k=None, S=None, src_v=None, training=True, seed=0, seed_ptr=None):
        """src_q: storage-type rows [B*T, Dp] (the un-embedded x_in); src_k / src_v: [B*S, Dp] or None (self-attention stack).
        Returns the encoder output rows [B*T, Dp] in storage type."""
        o, d, A = self.ops, self.d, self.arena
        self.training, self.seed, self.seed_ptr = training, seed, seed_ptr
        self.B, self.T, self.S = B, T, S
        self.cross = src_k is not None
        M = B * T
        Ms = B * S if self.cross else 0
        need = max(T, S or 0, (max(self.prune_pos) + 1) if self.prune_pos is not None else 0) + 1
        if self.pe is None or self.pe.shape[0] < need:
            self.pe = sinusoid_table(need, d.D, d.Dp, src_q.device)
        scale = math.sqrt(d.D)
        # x = dropout(sqrt(D) x_in + PE)                                         (transformer.py:66-69)
        xs = [A.get("x%d" % i, (M, d.Dp), torch.float32) for i in range(2 * self.L + 1 + (self.L if self.biproj and self.cross else 0))]
        xk = xv = None
        self.kv_shared = False
        if not self.with_embed:                                                  # bare layers: inputs are used as they are
            o.axpy_f32(src_q, xs[0], False)
            if self.cross:
                xk = src_k
                self.kv_shared = src_v is None or src_v is src_k
                xv = xk if self.kv_shared else src_v
        else:
            pe_q = self.pe
            if self.prune_pos is not None:                                       # row t of a sample sits at time step prune_pos[t]
                assert self.cross and not self.biproj and len(self.prune_pos) == T
                if getattr(self, "_pe_q_key", None) != (tuple(self.prune_pos), self.pe.shape[0]):
                    self.pe_q = torch.cat([self.pe[:1]] + [self.pe[p + 1:p + 2] for p in self.prune_pos], 0).contiguous()
                    self._pe_q_key = (tuple(self.prune_pos), self.pe.shape[0])
                pe_q = self.pe_q
            o.embed_fwd(src_q, pe_q, B, T, d.D, scale, xs[0], self._drop(self.p_embed, -1, 1))
        if self.cross and self.with_embed:
            xk = A.get("xk", (Ms, d.Dp), self.T_)
            o.embed_fwd(src_k, self.pe, B, S, d.D, scale, xk, self._drop(self.p_embed, -1, 2))
            # x_k and x_v are the same tensor unless embed dropout draws two masks (transformer.py:78-79)
            self.kv_shared = (src_v is None or src_v is src_k) and not (training and self.p_embed > 0)
            if self.kv_shared:
                xv = xk
            else:
                xv = A.get("xv", (Ms, d.Dp), self.T_)
                o.embed_fwd(src_k if src_v is None else src_v, self.pe, B, S, d.D, scale, xv, self._drop(self.p_embed, -1, 3))
        self.saved = []
        self.folded = self.cross and self.fold_kv
        nk = nv = None
        if self.folded:                                                          # x_hat of the K / V inputs, shared by all L layers
            def xhat(key, src):
                y = A.get(key + "y", (Ms, d.Dp), self.T_)
                mean = A.get(key + "mean", (Ms,), torch.float32)
                rstd = A.get(key + "rstd", (Ms,), torch.float32)
                o.layernorm_fwd(src, self.unit_g, self.unit_b, d.D, y, mean, rstd)
                return dict(y=y, mean=mean, rstd=rstd, x=src)
            nk = xhat("nk.", xk)
            nv = nk if self.kv_shared else xhat("nv.", xv)
            self.nkv = (nk, nv)
            LH = self.L * d.HP                                                    # K / V of every layer with one GEMM each
            K_all = A.get("K_all", (Ms, LH), self.T_)
            V_all = A.get("V_all", (Ms, LH), self.T_)
            o.gemm(nk["y"], self.Wkv["Wk"], K_all, Ms, LH, d.Dp, bias=self.Wkv["bk"])
            o.gemm(nv["y"], self.Wkv["Wv"], V_all, Ms, LH, d.Dp, bias=self.Wkv["bv"])
            kv_of = lambda l_: (K_all[:, l_ * d.HP:(l_ + 1) * d.HP], V_all[:, l_ * d.HP:(l_ + 1) * d.HP])
        xi = 0
        x = xs[0]
        for l in range(self.L):
            sv = {}
            ln_q = self._ln_fwd("L%d.lnq." % l, x, l, 0, M)                      # pre-norm (normalize_before, :132,153)
            sv["ln_q"] = ln_q
            if not self.cross:                                                   # :158-159
                x1 = xs[xi + 1]
                sv["self"] = self._attn_fwd(l, "s", ln_q["y"], ln_q["y"], ln_q["y"], B, T, T, x, x1)
                xi += 1
                ffn_ln = 1
            elif self.biproj:                                                    # :160-169
                x1 = xs[xi + 1]
                sv["self"] = self._attn_fwd(l, "s", ln_q["y"], ln_q["y"], ln_q["y"], B, T, T, x, x1)
                if self.folded:
                    k_op, v_op = kv_of(l)
                else:
                    ln_k = self._ln_fwd("L%d.lnk." % l, xk, l, 1, Ms)
                    ln_v = ln_k if self.kv_shared else self._ln_fwd("L%d.lnv." % l, xv, l, 1, Ms)
                    sv["ln_k"], sv["ln_v"] = ln_k, ln_v
                    k_op, v_op = ln_k["y"], ln_v["y"]
                # the cross-attention query is the residual stream itself, NOT re-normalised (:169)
                qc = A.get("L%d.qcast" % l, (M, d.Dp), self.T_)
                o.cast_drop(x1, qc, None)
                x2 = xs[xi + 2]
                sv["cross"] = self._attn_fwd(l, "x", qc, k_op, v_op, B, T, S, x1, x2, folded=self.folded)
                x1 = x2
                xi += 2
                ffn_ln = 2
            else:                                                                # :170-173
                if self.folded:
                    k_op, v_op = kv_of(l)
                else:
                    ln_k = self._ln_fwd("L%d.lnk." % l, xk, l, 0, Ms)
                    ln_v = ln_k if self.kv_shared else self._ln_fwd("L%d.lnv." % l, xv, l, 0, Ms)
                    sv["ln_k"], sv["ln_v"] = ln_k, ln_v
                    k_op, v_op = ln_k["y"], ln_v["y"]
                x1 = xs[xi + 1]
                sv["cross"] = self._attn_fwd(l, "x", ln_q["y"], k_op, v_op, B, T, S, x, x1, folded=self.folded)
                xi += 1
                ffn_ln = 1
            x2 = xs[xi + 1]
            sv["ffn"] = self._ffn_fwd(l, ffn_ln, x1, x2, M)
            xi += 1
            x = x2
            self.saved.append(sv)
        out = A.get("out", (M, d.Dp), self.T_)
        if not self.with_final_ln:
            o.cast_drop(x, out, None)
            return out
        mean = A.get("f.mean", (M,), torch.float32)
        rstd = A.get("f.rstd", (M,), torch.float32)
        o.layernorm_fwd(x, self.Wf["g"], self.Wf["b"], d.D, out, mean, rstd)    # :90-91
        self.final = dict(x=x, mean=mean, rstd=rstd)
        return out

    # ---------------------------------------------------------------- backward
    def backward(self, dout, d_src_q, d_src_k=None, d_src_v=None):
        """dout: fp32 [B*T, Dp] gradient of the encoder output.  Accumulates (+=) into the fp32 gradient buffers of the
        un-embedded inputs (d_src_*; d_src_v defaults to d_src_k) and into the padded parameter-gradient buffers."""
        o, d, Sh = self.ops, self.d, self.shared
        B, T, S = self.B, self.T, self.S
        M = B * T
        Ms = B * S if self.cross else 0
        gx = Sh.get("gx", (M, d.Dp), torch.float32)
        if self.with_final_ln:
            o.layernorm_bwd(dout, self.final["x"], self.final["mean"], self.final["rstd"], self.Wf["g"], d.D, gx, False, self.Gf["g"], self.Gf["b"])
        else:
            o.axpy_f32(dout, gx, False)
        gxk = gxv = gnk = gnv = None
        if self.cross:
            gxk = Sh.get("gxk", (Ms, d.Dp), torch.float32)
            o.zero_(gxk)
            gxv = gxk
            if not self.kv_shared:
                gxv = Sh.get("gxv", (Ms, d.Dp), torch.float32)
                o.zero_(gxv)
            if self.folded:                         # every layer's dK / dV side by side: one wgrad and one dgrad GEMM per encoder and side
                LH = self.L * d.HP
                dK_all = Sh.get("dK_all", (Ms, LH), self.T_)
                dV_all = Sh.get("dV_all", (Ms, LH), self.T_)
                dkv_of = lambda l_: (dK_all[:, l_ * d.HP:(l_ + 1) * d.HP], dV_all[:, l_ * d.HP:(l_ + 1) * d.HP])
        self._go_valid = False
        for l in reversed(range(self.L)):
            sv = self.saved[l]
            self._ffn_bwd(l, sv["ffn"], M, gx)
            if not self.cross:
                dq_in, dk_in, dv_in = self._attn_bwd(l, "s", sv["self"], B, T, gx)
                for t in (dq_in, dk_in, dv_in):
                    self._ln_bwd(l, sv["ln_q"], t, gx, cast_next=t is dv_in)
            elif self.biproj:
                dq_in, dk_in, dv_in = self._attn_bwd(l, "x", sv["cross"], B, T, gx, dkv=dkv_of(l) if self.folded else None)
                o.axpy_f32(dq_in, gx, True)                                      # query path is the raw residual stream
                if not self.folded:
                    self._ln_bwd(l, sv["ln_k"], dk_in, gxk)
                    self._ln_bwd(l, sv["ln_v"], dv_in, gxv)
                dq_in, dk_in, dv_in = self._attn_bwd(l, "s", sv["self"], B, T, gx)
                for t in (dq_in, dk_in, dv_in):
                    self._ln_bwd(l, sv["ln_q"], t, gx, cast_next=t is dv_in)
            else:
                dq_in, dk_in, dv_in = self._attn_bwd(l, "x", sv["cross"], B, T, gx, dkv=dkv_of(l) if self.folded else None)
                self._ln_bwd(l, sv["ln_q"], dq_in, gx, cast_next=True)
                if not self.folded:
                    self._ln_bwd(l, sv["ln_k"], dk_in, gxk)
                    self._ln_bwd(l, sv["ln_v"], dv_in, gxv)
        if self.cross and self.folded:
            nk, nv = self.nkv
            hm = (d.dh, d.dhp)
            Wkv, Gkv = self.Wkv, self.Gkv
            # weight / bias gradients of all layers' folded projections: dW_k'(all) = dK(all)^T x_hat, db_k'(all) = colsum dK(all)
            o.gemm(dK_all, nk["y"], Gkv["Wk"], LH, d.Dp, Ms, ta=1, tb=1, accumulate=True, colsum=Gkv["bk"])
            o.gemm(dV_all, nv["y"], Gkv["Wv"], LH, d.Dp, Ms, ta=1, tb=1, accumulate=True, colsum=Gkv["bv"])
            # d x_hat = dK(all) W_k'(all) (+ dV(all) W_v'(all) when K and V share their input), fp32; then ONE LayerNorm backward
            gnk = Sh.get("gnk", (Ms, d.Dp), torch.float32)
            o.gemm(dK_all, Wkv["Wk"], gnk, Ms, d.Dp, LH, tb=1)
            if self.kv_shared:
                o.gemm(dV_all, Wkv["Wv"], gnk, Ms, d.Dp, LH, tb=1, residual=gnk)
            o.layernorm_bwd(gnk, nk["x"], nk["mean"], nk["rstd"], self.unit_g, d.D, gxk, True, self.ln_sink[0], self.ln_sink[1])
            if not self.kv_shared:
                gnv = Sh.get("gnv", (Ms, d.Dp), torch.float32)
                o.gemm(dV_all, Wkv["Wv"], gnv, Ms, d.Dp, LH, tb=1)
                o.layernorm_bwd(gnv, nv["x"], nv["mean"], nv["rstd"], self.unit_g, d.D, gxv, True, self.ln_sink[0], self.ln_sink[1])
            o.fold_batch_begin("bwd")
            for l in range(self.L):                 # gradients of (W_k', b_k') -> W_k, b_k and the layer's LayerNorm affine
                ipw, ipb, g_, b_ = self.P[l]
                G = self.G[l]
                r = slice(l * d.HP, (l + 1) * d.HP)
                o.ln_fold_bwd(ipw[d.D:2 * d.D], g_, b_, Gkv["Wk"][r], Gkv["bk"][r], G["Wqkv"][d.HP:2 * d.HP], G["bqkv"][d.HP:2 * d.HP],
                              G["ln_g"][self.kv_ln], G["ln_b"][self.kv_ln], row_map=hm)
                o.ln_fold_bwd(ipw[2 * d.D:], g_, b_, Gkv["Wv"][r], Gkv["bv"][r], G["Wqkv"][2 * d.HP:], G["bqkv"][2 * d.HP:],
                              G["ln_g"][self.kv_ln], G["ln_b"][self.kv_ln], row_map=hm)
            o.fold_batch_end()
        scale = math.sqrt(d.D)
        if not self.with_embed:
            if d_src_q is not None:
                o.axpy_f32(gx, d_src_q, True)
            if self.cross:
                if d_src_k is not None:
                    o.axpy_f32(gxk, d_src_k, True)
                if not self.kv_shared:
                    o.axpy_f32(gxv, d_src_v if d_src_v is not None else d_src_k, True)
            return
        if d_src_q is not None:
            o.embed_bwd(gx, d.D, scale, d_src_q, True, self._drop(self.p_embed, -1, 1))
        if self.cross:
            if d_src_v is None:
                d_src_v = d_src_k
            if d_src_k is not None:
                o.embed_bwd(gxk, d.D, scale, d_src_k, True, self._drop(self.p_embed, -1, 2))
            if not self.kv_shared and d_src_v is not None:
                o.embed_bwd(gxv, d.D, scale, d_src_v, True, self._drop(self.p_embed, -1, 3))


# ============================================================================================== sequence GMU
class SeqGmuEngine:
    """GatedMultimodalLayerFeatures / GatedMultimodalLayer on [rows, Dp] (models/mmtr.py:161-195); size_in == size_out == D."""

    def __init__(self, ops, D, dtype, features=True, shared=None, name="gmu"):
        self.ops, self.D, self.Dp, self.T_, self.features = ops, D, round_up(D, 64), dtype, features
        z = ops.zeros
        Dp = self.Dp
        self.W = dict(h1=z((Dp, Dp), dtype), h2=z((Dp, Dp), dtype), zg=z((Dp, 2 * Dp), dtype))
        self.G = dict(h1=z((Dp, Dp), torch.float32), h2=z((Dp, Dp), torch.float32), zg=z((Dp, 2 * Dp), torch.float32))
        self.arena = Arena(ops)
        self.shared = shared if shared is not None else Arena(ops)
        self.name = name

    def param_shapes(self):
        D = self.D
        return {"hidden1.weight": (D, D), "hidden2.weight": (D, D), "x_gate.weight": (D, 2 * D)}

    def pack(self, params, pfx=""):
        o = self.ops
        o.pack_matrix(params[pfx + "hidden1.weight"], self.W["h1"])
        o.pack_matrix(params[pfx + "hidden2.weight"], self.W["h2"])
        o.pack_matrix(params[pfx + "x_gate.weight"], self.W["zg"], col_map=(self.D, self.Dp))   # [x1 | x2] column blocks

    def zero_grads(self):
        for t in self.G.values():
            self.ops.zero_(t)

    def unpack_grads(self, grads, pfx="", accumulate=False):
        o = self.ops
        o.unpack_matrix(self.G["h1"], grads[pfx + "hidden1.weight"], accumulate=accumulate)
        o.unpack_matrix(self.G["h2"], grads[pfx + "hidden2.weight"], accumulate=accumulate)
        o.unpack_matrix(self.G["zg"], grads[pfx + "x_gate.weight"], col_map=(self.D, self.Dp), accumulate=accumulate)

    def forward(self, a1, a2, rows, addend=None, want_gate=False):
        o, A, Dp, W = self.ops, self.arena, self.Dp, self.W
        h1p = A.get("h1p", (rows, Dp), self.T_)
        h2p = A.get("h2p", (rows, Dp), self.T_)
        zp = A.get("zp", (rows, Dp), self.T_)
        y = A.get("y", (rows, Dp), self.T_)
        o.gemm(a1, W["h1"], h1p, rows, Dp, Dp)
        o.gemm(a2, W["h2"], h2p, rows, Dp, Dp)
        o.gemm(a1, W["zg"][:, :Dp], zp, rows, Dp, Dp)                       # z_pre = [a1 | a2] Wz^T, as two accumulating GEMMs
        o.gemm(a2, W["zg"][:, Dp:], zp, rows, Dp, Dp, residual=zp)
        zo = A.get("z", (rows, Dp), self.T_) if want_gate else None
        o.gmu_fwd(self.features, a1, a2, h1p, h2p, zp, addend, y, zo)
        self.sv = dict(a1=a1, a2=a2, h1p=h1p, h2p=h2p, zp=zp, rows=rows, z=zo)
        return y

    def backward(self, dy, da1, da2):
        """dy fp32 [rows, Dp]; accumulates into fp32 da1 / da2 (and the parameter-gradient buffers)."""
        o, Sh, Dp, W, G, sv = self.ops, self.shared, self.Dp, self.W, self.G, self.sv
        rows = sv["rows"]
        dh1 = Sh.get("gmu.dh1", (rows, Dp), self.T_)
        dh2 = Sh.get("gmu.dh2", (rows, Dp), self.T_)
        dz = Sh.get("gmu.dz", (rows, Dp), self.T_)
        o.gmu_bwd(self.features, sv["a1"], sv["a2"], sv["h1p"], sv["h2p"], sv["zp"], dy, dh1, dh2, dz, da1, da2)
        o.gemm(dh1, sv["a1"], G["h1"], Dp, Dp, rows, ta=1, tb=1, accumulate=True)
        o.gemm(dh2, sv["a2"], G["h2"], Dp, Dp, rows, ta=1, tb=1, accumulate=True)
        o.gemm(dz, sv["a1"], G["zg"][:, :Dp], Dp, Dp, rows, ta=1, tb=1, accumulate=True)
        o.gemm(dz, sv["a2"], G["zg"][:, Dp:], Dp, Dp, rows, ta=1, tb=1, accumulate=True)
        # input gradients through the three linears (fp32 accumulation via the residual epilogue)
        o.gemm(dh1, W["h1"], da1, rows, Dp, Dp, tb=1, residual=da1)
        o.gemm(dz, W["zg"][:, :Dp], da1, rows, Dp, Dp, tb=1, residual=da1)
        o.gemm(dh2, W["h2"], da2, rows, Dp, Dp, tb=1, residual=da2)
        o.gemm(dz, W["zg"][:, Dp:], da2, rows, Dp, Dp, tb=1, residual=da2)


# ============================================================================================== head (fp32, [B, D] rows)
class HeadEngine:
    """pooling (mmtr.py:808,830,852) -> TextShiftingN (mmtr.py:197-247) -> residual MLP (:860-861) -> out_layer (:866)
    -> BCEWithLogits (train.py:104,333).  B rows only: runs entirely in fp32 on the FFMA GEMM + small fused kernels."""

    def __init__(self, ops, D, n_in, n_classes, out_dropout=0.0, uid=4095, prefix="gmu.", list_names=False, gate_only=False):
        """prefix / list_names: parameter names of the gate -- `<prefix>hidden{i+1}.weight`, `<prefix>x{i+1}_gate.weight` (TextShifting3/4Layer)
        or `<prefix>hiddens.{i}.weight`, `<prefix>x_gates.{i}.weight` (TextShiftingNLayer, an nn.ModuleList).  gate_only: the gate alone
        (gate_forward / gate_backward), without the residual MLP and the classifier -- the hybrid branch's gmu_early (mmtr.py:631,775)."""
        self.ops, self.D, self.Dp, self.n_in, self.C = ops, D, round_up(D, 64), n_in, n_classes
        self.Cp = round_up(n_classes, 8)
        self.p_out, self.uid = out_dropout, uid
        self.prefix, self.list_names, self.gate_only = prefix, list_names, gate_only
        z, f32, Dp, n = ops.zeros, torch.float32, self.Dp, n_in
        mk = lambda: dict(h=[z((Dp, Dp), f32) for _ in range(n)], zg=[z((Dp, n * Dp), f32) for _ in range(n)])
        self.W, self.G = mk(), mk()
        if not gate_only:
            for d in (self.W, self.G):
                d.update(p1=z((Dp, Dp), f32), p1b=z((Dp,), f32), p2=z((Dp, Dp), f32), p2b=z((Dp,), f32),
                         out=z((self.Cp, Dp), f32), outb=z((self.Cp,), f32))
        self.arena = Arena(ops)

    def _hn(self, i):
        return "%shiddens.%d.weight" % (self.prefix, i) if self.list_names else "%shidden%d.weight" % (self.prefix, i + 1)

    def _zn(self, i):
        return "%sx_gates.%d.weight" % (self.prefix, i) if self.list_names else "%sx%d_gate.weight" % (self.prefix, i + 1)

    def param_shapes(self):
        D, n, s = self.D, self.n_in, {}
        for i in range(n):
            s[self._hn(i)] = (D, D)
        for i in range(n):
            s[self._zn(i)] = (D, n * D)
        if not self.gate_only:
            s.update({"proj1.weight": (D, D), "proj1.bias": (D,), "proj2.weight": (D, D), "proj2.bias": (D,),
                      "out_layer.weight": (self.C, D), "out_layer.bias": (self.C,)})
        return s

    def pack(self, params):
        o, W = self.ops, self.W
        for i in range(self.n_in):
            o.pack_matrix(params[self._hn(i)], W["h"][i])
            o.pack_matrix(params[self._zn(i)], W["zg"][i], col_map=(self.D, self.Dp))
        if self.gate_only:
            return
        o.pack_matrix(params["proj1.weight"], W["p1"]); o.pack_matrix(params["proj1.bias"].view(1, -1), W["p1b"].view(1, -1))
        o.pack_matrix(params["proj2.weight"], W["p2"]); o.pack_matrix(params["proj2.bias"].view(1, -1), W["p2b"].view(1, -1))
        o.pack_matrix(params["out_layer.weight"], W["out"]); o.pack_matrix(params["out_layer.bias"].view(1, -1), W["outb"].view(1, -1))

    def zero_grads(self):
        for v in self.G.values():
            for t in (v if isinstance(v, list) else [v]):
                self.ops.zero_(t)

    def unpack_grads(self, grads, accumulate=False):
        o, G = self.ops, self.G
        for i in range(self.n_in):
            o.unpack_matrix(G["h"][i], grads[self._hn(i)], accumulate=accumulate)
            o.unpack_matrix(G["zg"][i], grads[self._zn(i)], col_map=(self.D, self.Dp), accumulate=accumulate)
        if self.gate_only:
            return
        o.unpack_matrix(G["p1"], grads["proj1.weight"], accumulate=accumulate)
        o.unpack_matrix(G["p1b"].view(1, -1), grads["proj1.bias"].view(1, -1), accumulate=accumulate)
        o.unpack_matrix(G["p2"], grads["proj2.weight"], accumulate=accumulate)
        o.unpack_matrix(G["p2b"].view(1, -1), grads["proj2.bias"].view(1, -1), accumulate=accumulate)
        o.unpack_matrix(G["out"], grads["out_layer.weight"], accumulate=accumulate)
        o.unpack_matrix(G["outb"].view(1, -1), grads["out_layer.bias"].view(1, -1), accumulate=accumulate)

    def forward(self, B, training=True, seed=0, seed_ptr=None):
        """expects self.cat (fp32 [B, n_in*Dp]) filled by pool_fwd / direct inputs.  Returns logits [B, Cp] (first C valid)."""
        o, A, Dp, n, W = self.ops, self.arena, self.Dp, self.n_in, self.W
        self.B, self.training, self.seed, self.seed_ptr = B, training, seed, seed_ptr
        fused, z = self.gate_forward(B)
        h1 = A.get("h1", (B, Dp), torch.float32)
        y = A.get("y", (B, Dp), torch.float32)
        logits = A.get("logits", (B, self.Cp), torch.float32)
        drop = Drop(self.p_out, seed, seed_ptr, (self.uid << 20) | 1) if (training and self.p_out > 0) else None
        o.gemm(fused, W["p1"], h1, B, Dp, Dp, bias=W["p1b"], act=1, drop=drop)
        o.gemm(h1, W["p2"], y, B, Dp, Dp, bias=W["p2b"], residual=fused)
        o.gemm(y, W["out"], logits, B, self.Cp, Dp, bias=W["outb"])
        self.sv.update(fused=fused, h1=h1, y=y, z=z)
        return logits, z

    def gate_forward(self, B):
        """TextShiftingN on self.cat_buf(B): returns (fused [B, Dp], z [B, n*Dp])"""
        o, A, Dp, n, W = self.ops, self.arena, self.Dp, self.n_in, self.W
        self.B = B
        cat = self.cat_buf(B)
        hpre = A.get("hpre", (n, B, Dp), torch.float32)
        zpre = A.get("zpre", (n, B, Dp), torch.float32)
        for i in range(n):
            o.gemm(cat[:, i * Dp:(i + 1) * Dp], W["h"][i], hpre[i], B, Dp, Dp)
            o.gemm(cat, W["zg"][i], zpre[i], B, Dp, n * Dp)
        fused = A.get("fused", (B, Dp), torch.float32)
        z = A.get("z", (B, n * Dp), torch.float32)
        o.tsgate_fwd(hpre, zpre, n, B, Dp, fused, z)
        self.sv = dict(cat=cat, hpre=hpre, zpre=zpre, fused=fused, z=z)
        return fused, z

    def cat_buf(self, B):
        return self.arena.get("cat", (B, self.n_in * self.Dp), torch.float32)

    def loss(self, logits, targets, pos_weight, grad_scale=1.0):
        o, A = self.ops, self.arena
        loss = A.get("loss", (1,), torch.float32)
        dlogits = A.get("dlogits", (self.B, self.Cp), torch.float32, zero=True)
        o.bce_fwd_bwd(logits, targets, pos_weight, self.B, self.C, grad_scale, loss, dlogits)
        return loss, dlogits

    def backward(self, dlogits):
        """dlogits fp32 [B, Cp] -> returns dcat fp32 [B, n_in*Dp] (gradient wrt the pooled inputs)."""
        o, A, Dp, n, W, G, sv, B = self.ops, self.arena, self.Dp, self.n_in, self.W, self.G, self.sv, self.B
        o.colsum(dlogits, self.Cp, G["outb"])
        o.gemm(dlogits, sv["y"], G["out"], self.Cp, Dp, B, ta=1, tb=1, accumulate=True)
        dy = A.get("dy", (B, Dp), torch.float32)
        o.gemm(dlogits, W["out"], dy, B, Dp, self.Cp, tb=1)
        o.colsum(dy, Dp, G["p2b"])
        o.gemm(dy, sv["h1"], G["p2"], Dp, Dp, B, ta=1, tb=1, accumulate=True)
        dh1 = A.get("dh1", (B, Dp), torch.float32)
        keep = 1.0 / (1.0 - self.p_out) if (self.training and self.p_out > 0) else 1.0
        o.gemm(dy, W["p2"], dh1, B, Dp, Dp, tb=1, gate=sv["h1"], gate_scale=keep)
        o.colsum(dh1, Dp, G["p1b"])
        o.gemm(dh1, sv["fused"], G["p1"], Dp, Dp, B, ta=1, tb=1, accumulate=True)
        dfused = A.get("dfused", (B, Dp), torch.float32)
        o.gemm(dh1, W["p1"], dfused, B, Dp, Dp, tb=1, residual=dy)            # + residual branch (last_hs_proj += last_hs)
        return self.gate_backward(dfused)

    def gate_backward(self, dfused):
        o, A, Dp, n, W, G, sv, B = self.ops, self.arena, self.Dp, self.n_in, self.W, self.G, self.sv, self.B
        dhpre = A.get("dhpre", (n, B, Dp), torch.float32)
        dzpre = A.get("dzpre", (n, B, Dp), torch.float32)
        o.tsgate_bwd(sv["hpre"], sv["zpre"], dfused, n, B, Dp, dhpre, dzpre)
        dcat = A.get("dcat", (B, n * Dp), torch.float32)
        o.zero_(dcat)
        cat = sv["cat"]
        for i in range(n):
            o.gemm(dhpre[i], cat[:, i * Dp:(i + 1) * Dp], G["h"][i], Dp, Dp, B, ta=1, tb=1, accumulate=True)
            o.gemm(dzpre[i], cat, G["zg"][i], Dp, n * Dp, B, ta=1, tb=1, accumulate=True)
            o.gemm(dhpre[i], W["h"][i], dcat[:, i * Dp:(i + 1) * Dp], B, Dp, Dp, tb=1, residual=dcat[:, i * Dp:(i + 1) * Dp])
            o.gemm(dzpre[i], W["zg"][i], dcat, B, n * Dp, Dp, tb=1, residual=dcat)
        return dcat


# ============================================================================================== audio encoder (mmtr.py:93-108)
class AudioEncoderEngine:
    """AudioEncoder of the 4-modality model: Conv1d(C, C, k=128, stride=2) x 2 + AdaptiveAvgPool1d(Tp) on a raw spectrogram (B, C, T_raw).
    Each strided convolution is an implicit GEMM with K = 128 * C (im2col rows -> the tensor-core GEMM); activations are time-major rows
    [B*T, C].  forward() writes the pooled features straight into the caller's staged input rows (pitch = that buffer's)."""
    KW, STRIDE = 128, 2

    def __init__(self, ops, C=96, Tp=200, dtype=torch.bfloat16):
        self.ops, self.C, self.Tp, self.T_ = ops, C, Tp, dtype
        z = ops.zeros
        K = self.KW * C
        self.W = [dict(w=z((C, K), dtype), b=z((C,), torch.float32)) for _ in range(2)]
        self.G = [dict(w=z((C, K), torch.float32), b=z((C,), torch.float32)) for _ in range(2)]
        self.arena = Arena(ops)

    def param_shapes(self):
        C = self.C
        return {"conv_layers.0.weight": (C, C, self.KW), "conv_layers.0.bias": (C,), "conv_layers.1.weight": (C, C, self.KW), "conv_layers.1.bias": (C,)}

    def pack(self, params, pfx=""):
        for i in range(2):
            self.ops.conv1d_pack_weight(params["%sconv_layers.%d.weight" % (pfx, i)], self.W[i]["w"])
            self.W[i]["b"].copy_(params["%sconv_layers.%d.bias" % (pfx, i)])

    def zero_grads(self):
        for g in self.G:
            self.ops.zero_(g["w"])
            self.ops.zero_(g["b"])

    def unpack_grads(self, grads, pfx="", accumulate=False):
        for i in range(2):
            self.ops.conv1d_unpack_wgrad(self.G[i]["w"], grads["%sconv_layers.%d.weight" % (pfx, i)], accumulate)
            gb = grads["%sconv_layers.%d.bias" % (pfx, i)]
            gb.copy_(gb + self.G[i]["b"] if accumulate else self.G[i]["b"])

    def out_len(self, T):
        return (T - self.KW) // self.STRIDE + 1

    def forward(self, audio, out_rows):
        """audio fp32 (B, C, T_raw) (channel-major, as the reference's Conv1d takes it); out_rows [B*Tp, ld >= C] receives the features"""
        o, A, C = self.ops, self.arena, self.C
        B, Cc, T0 = audio.shape
        assert Cc == C and T0 >= 3 * self.KW + 2, "AudioEncoder: (B, %d, T >= %d) expected" % (C, 3 * self.KW + 2)
        T1 = self.out_len(T0)
        T2 = self.out_len(T1)
        x0 = A.get("x0", (B * T0, C), self.T_)
        o.stage_rows(audio.permute(0, 2, 1), x0, T0)                            # (B, C, T) -> time-major rows [B*T, C]
        col1 = A.get("col1", (B * T1, self.KW * C), self.T_)
        o.conv1d_im2col(x0, B, T0, C, self.KW, self.STRIDE, col1, T1)
        y1 = A.get("y1", (B * T1, C), self.T_)
        o.gemm(col1, self.W[0]["w"], y1, B * T1, C, self.KW * C, bias=self.W[0]["b"])
        col2 = A.get("col2", (B * T2, self.KW * C), self.T_)
        o.conv1d_im2col(y1, B, T1, C, self.KW, self.STRIDE, col2, T2)
        y2 = A.get("y2", (B * T2, C), self.T_)
        o.gemm(col2, self.W[1]["w"], y2, B * T2, C, self.KW * C, bias=self.W[1]["b"])
        o.adaptive_pool_fwd(y2, B, T2, C, self.Tp, out_rows)
        self.sv = dict(B=B, T1=T1, T2=T2, col1=col1, col2=col2)

    def backward(self, dout):
        """dout fp32 [B*Tp, ld >= C]: gradient of the pooled features.  Accumulates the weight / bias gradients (the spectrogram gets none)."""
        o, A, C, sv = self.ops, self.arena, self.C, self.sv
        B, T1, T2, K = sv["B"], sv["T1"], sv["T2"], self.KW * C
        dy2 = A.get("dy2", (B * T2, C), torch.float32)
        o.adaptive_pool_bwd(dout, B, T2, C, self.Tp, dy2)
        g2 = A.get("g2", (B * T2, C), self.T_)
        o.cast_drop(dy2, g2, None)
        o.gemm(g2, sv["col2"], self.G[1]["w"], C, K, B * T2, ta=1, tb=1, accumulate=True, colsum=self.G[1]["b"])     # dW2 = dY2^T col2
        dcol2 = A.get("dcol2", (B * T2, K), self.T_)
        o.gemm(g2, self.W[1]["w"], dcol2, B * T2, K, C, tb=1)                                                       # dcol2 = dY2 W2
        dy1 = A.get("dy1", (B * T1, C), torch.float32)
        o.conv1d_col2im(dcol2, B, T1, C, self.KW, self.STRIDE, T2, dy1)
        g1 = A.get("g1", (B * T1, C), self.T_)
        o.cast_drop(dy1, g1, None)
        o.gemm(g1, sv["col1"], self.G[0]["w"], C, K, B * T1, ta=1, tb=1, accumulate=True, colsum=self.G[0]["b"])     # dW1 = dY1^T col1
