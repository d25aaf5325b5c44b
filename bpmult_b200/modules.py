"""Drop-in replacements for the reference's modules on the hot path (same constructor / forward signatures, same
state_dict keys, same parameter initialisation order => same weights under the same seed), executing on our CUDA kernels.

  reference class (file:line under /root/reference/bpmult/models)            here
  TransformerEncoder            transformer.py:9-100                          TransformerEncoder
  TransformerEncoderLayer       transformer.py:102-202                        TransformerEncoderLayer
  MultiheadAttention            multihead_attention.py:10-158                 MultiheadAttention
  SinusoidalPositionalEmbedding position_embedding.py:30-80                   SinusoidalPositionalEmbedding
  GatedMultimodalLayer[Features] mmtr.py:161-195                              same names
  TextShifting3Layer / 4Layer   mmtr.py:197-247                               same names
  MultiprojectionMMTransformer3DGMUClf ("mmtrvat") mmtr.py:587-866            same name
  get_model / MODELS            models/__init__.py:6-14                       same names

Each forward is ONE torch.autograd.Function whose forward/backward run the explicit kernel schedules of engine.py /
model_engine.py; torch only owns the tensors.  `precision="bf16"` (default: bf16 storage, fp32 accumulation, tcgen05
tensor cores) or `"fp32"` (exact-fp32 kernels).  The reference's feature extractors (BERT, AudioEncoder) are bypassed:
`txt` is a float feature sequence (B, L, orig_d_l)."""
import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Parameter

from . import engine as E
from .model_engine import ENC_NAMES, MMTrVatEngine, attn_dropout_for

_DT = {"bf16": torch.bfloat16, "fp32": torch.float32}
_OPS = {}


def _ops_for(device):
    from .ops import CudaOps
    if device.type != "cuda":
        raise RuntimeError("bpmult_b200 runs on CUDA (sm_100a) only: move the module and its inputs to the GPU "
                           "(there is no CPU fallback)")
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _OPS:
        _OPS[key] = CudaOps(torch.device("cuda", key))
    return _OPS[key]


_SEED = [0x5EED]


def _next_seed():
    _SEED[0] = (_SEED[0] * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
    return _SEED[0] >> 1


def manual_seed(seed):
    """seeds the dropout streams of the eager module path (the Trainer keeps its own device-side seed)"""
    _SEED[0] = int(seed) & 0xFFFFFFFFFFFFFFFF


# ============================================================================================== parameter containers
def Linear(in_features, out_features, bias=True):                      # transformer.py:219-224
    m = nn.Linear(in_features, out_features, bias)
    nn.init.xavier_uniform_(m.weight)
    if bias:
        nn.init.constant_(m.bias, 0.)
    return m


def LayerNorm(embedding_dim):                                          # transformer.py:227-229
    return nn.LayerNorm(embedding_dim)


def buffered_future_mask(tensor, tensor2=None):
    """transformer.py:209-216.  Returns the additive (T, S) mask for API compatibility; our attention kernel evaluates the
    predicate j - i >= 1 + |S - T| from indices and never reads this tensor (it is tagged with the offset)."""
    dim1 = dim2 = tensor.size(0)
    if tensor2 is not None:
        dim2 = tensor2.size(0)
    m = torch.triu(torch.full((dim1, dim2), float("-inf"), device=tensor.device), 1 + abs(dim2 - dim1))
    m._bpm_mask_off = abs(dim2 - dim1)
    return m


class SinusoidalPositionalEmbedding(nn.Module):
    """position_embedding.py:30-80.  forward(input (bsz, seqlen) float) -> (bsz, seqlen, D); position = t + 1 where
    input != padding_idx (0), else 0 (zero row)."""

    def __init__(self, embedding_dim, padding_idx=0, left_pad=0, init_size=128):
        super().__init__()
        assert padding_idx == 0 and not left_pad, "the trunk only uses padding_idx=0, left_pad=0 (transformer.py:28)"
        self.embedding_dim, self.padding_idx, self.left_pad = embedding_dim, padding_idx, left_pad
        self.register_buffer("_float_tensor", torch.zeros(1))
        self._pe = None

    def forward(self, input):
        bsz, seq_len = input.shape
        ops = _ops_for(input.device)
        D, Dp = self.embedding_dim, E.round_up(self.embedding_dim, 64)
        if self._pe is None or self._pe.shape[0] < seq_len + 1 or self._pe.device != input.device:
            self._pe = E.sinusoid_table(seq_len + 1, D, Dp, input.device)
        x = ops.zeros((bsz * seq_len, Dp), torch.float32)
        x[:, 0] = input.reshape(-1).float()
        y = ops.empty((bsz * seq_len, Dp), torch.float32)
        ops.embed_fwd(x, self._pe, bsz, seq_len, D, 0.0, y, None)
        return y.view(bsz, seq_len, Dp)[:, :, :D].detach()

    def max_positions(self):
        return int(1e5)


class MultiheadAttention(nn.Module):
    """multihead_attention.py:10-158: packed in_proj_weight (3D, D) rows [Q; K; V], out_proj Linear."""

    def __init__(self, embed_dim, num_heads, attn_dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False):
        super().__init__()
        assert bias and not add_bias_kv and not add_zero_attn, "the trunk never enables add_bias_kv / add_zero_attn / bias=False"
        self.embed_dim, self.num_heads, self.attn_dropout = embed_dim, num_heads, attn_dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        self.scaling = self.head_dim ** -0.5
        self.in_proj_weight = Parameter(torch.Tensor(3 * embed_dim, embed_dim))
        self.in_proj_bias = Parameter(torch.Tensor(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.bias_k = self.bias_v = None
        self.add_zero_attn = add_zero_attn
        self.precision = "bf16"
        self._eng = None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.constant_(self.in_proj_bias, 0.)
        nn.init.constant_(self.out_proj.bias, 0.)

    def forward(self, query, key, value, attn_mask=None, need_weights=True):
        """Time x Batch x Channel in, returns (attn (T, B, D), head-averaged weights (B, T, S) or None)."""
        T, B, D = query.shape
        assert D == self.embed_dim and key.shape == value.shape
        S = key.shape[0]
        mask_off = -1
        if attn_mask is not None:
            mask_off = getattr(attn_mask, "_bpm_mask_off", None)
            if mask_off is None:
                mask_off = abs(S - T)
                if not torch.equal(attn_mask, buffered_future_mask(query, key).to(attn_mask.dtype)):
                    raise NotImplementedError("bpmult_b200 attention supports the reference's future mask "
                                              "(transformer.py:209-216) or no mask; arbitrary additive masks are not implemented")
        return _MHAFn.apply(self, query, key, value, mask_off, need_weights, self.in_proj_weight, self.in_proj_bias,
                            self.out_proj.weight, self.out_proj.bias)


class _MHAFn(torch.autograd.Function):
    """standalone attention block: one-layer EncoderEngine pieces without LayerNorm / residual / FFN"""

    @staticmethod
    def forward(ctx, mod, query, key, value, mask_off, need_weights, ipw, ipb, ow, ob):
        T, B, D = query.shape
        S = key.shape[0]
        ops = _ops_for(query.device)
        dt = _DT[mod.precision]
        if mod._eng is None or mod._eng.T_ != dt or mod._eng.ops is not ops:
            mod._eng = E.EncoderEngine(ops, D, mod.num_heads, 1, attn_dropout=mod.attn_dropout, attn_mask=True, dtype=dt, uid=900)
        eng = mod._eng
        d = eng.d
        eng.pack_attention(0, ipw.detach(), ipb.detach(), ow.detach(), ob.detach())
        eng.training, eng.seed, eng.seed_ptr = mod.training, _next_seed(), None
        eng._mask_override = mask_off
        xs = []
        for t, n in ((query, T), (key, S), (value, S)):
            r = ops.empty((B * n, d.Dp), dt)
            ops.stage_rows(t.detach().float().permute(1, 0, 2), r, n)
            xs.append(r)
        zero = eng.arena.get("mha.zero", (B * T, d.Dp), torch.float32, zero=True)
        out = eng.arena.get("mha.out", (B * T, d.Dp), torch.float32)
        sv = eng._attn_fwd(0, "x", xs[0], xs[1], xs[2], B, T, S, zero, out, res_drop=False)
        ctx.mod, ctx.sv, ctx.dims = mod, sv, (T, B, S, D)
        w = None
        if need_weights:
            w = ops.empty((B, T, S), torch.float32)
            ops.xattn_weights(sv["q"], sv["k"], sv["lse"], w, B, T, S, d.H, d.dh, d.dhp, mask_off=mask_off,
                              drop=eng._drop(eng.p_attn, 0, 11))
        ctx.mark_non_differentiable(*([w] if w is not None else []))
        return out.view(B, T, d.Dp)[:, :, :D].permute(1, 0, 2).clone(), w

    @staticmethod
    def backward(ctx, g, _gw):
        mod, sv = ctx.mod, ctx.sv
        T, B, S, D = ctx.dims
        eng = mod._eng
        ops, d = eng.ops, eng.d
        eng.zero_grads()
        gx = ops.empty((B * T, d.Dp), torch.float32)
        ops.stage_rows(g.float().permute(1, 0, 2), gx, T)
        dq_in, dk_in, dv_in = eng._attn_bwd(0, "x", sv, B, T, gx, res_drop=False)
        outs = []
        for t, n in ((dq_in, T), (dk_in, S), (dv_in, S)):
            outs.append(t.float().view(B, n, d.Dp)[:, :, :D].permute(1, 0, 2).clone())
        gr = eng.unpack_attention_grads(0)
        return (None, outs[0], outs[1], outs[2], None, None) + gr


class TransformerEncoderLayer(nn.Module):
    """transformer.py:102-202 (parameter container + standalone forward)."""

    def __init__(self, embed_dim, num_heads=4, attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.1, attn_mask=False, biprojection=False):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.self_attn = MultiheadAttention(embed_dim=embed_dim, num_heads=num_heads, attn_dropout=attn_dropout)
        self.attn_mask, self.biprojection = attn_mask, biprojection
        self.attn_dropout, self.relu_dropout, self.res_dropout = attn_dropout, relu_dropout, res_dropout
        self.normalize_before = True
        self.fc1 = Linear(embed_dim, 4 * embed_dim)
        self.fc2 = Linear(4 * embed_dim, embed_dim)
        self.layer_norms = nn.ModuleList([LayerNorm(embed_dim) for _ in range(3 if biprojection else 2)])
        self.precision = "bf16"
        self._enc = None

    def forward(self, x, x_k=None, x_v=None):
        """(T, B, D) -> (T, B, D): one layer without input embedding / final LayerNorm."""
        if self._enc is None:
            self._enc = _LayerRunner(self)
        return self._enc(x, x_k, x_v)


class TransformerEncoder(nn.Module):
    """transformer.py:9-100.  forward(x_in (T,B,D), x_in_k=None, x_in_v=None) -> (T,B,D), time-major like the reference."""

    def __init__(self, embed_dim, num_heads, layers, attn_dropout=0.0, relu_dropout=0.0, res_dropout=0.0, embed_dropout=0.0,
                 attn_mask=False, biprojection=False, precision="bf16"):
        super().__init__()
        self.dropout = embed_dropout
        self.attn_dropout, self.relu_dropout, self.res_dropout = attn_dropout, relu_dropout, res_dropout
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.embed_scale = math.sqrt(embed_dim)
        self.embed_positions = SinusoidalPositionalEmbedding(embed_dim)
        self.attn_mask, self.biprojection = attn_mask, biprojection
        self.layers = nn.ModuleList([TransformerEncoderLayer(embed_dim, num_heads=num_heads, attn_dropout=attn_dropout,
                                                             relu_dropout=relu_dropout, res_dropout=res_dropout, attn_mask=attn_mask,
                                                             biprojection=biprojection) for _ in range(layers)])
        self.register_buffer("version", torch.Tensor([2]))
        self.normalize = True
        self.layer_norm = LayerNorm(embed_dim)
        self.precision = precision
        self._eng = None
        self.with_embed, self.with_final_ln = True, True

    def _engine(self, device):
        ops = _ops_for(device)
        dt = _DT[self.precision]
        if self._eng is None or self._eng.T_ != dt or self._eng.ops is not ops:
            self._eng = E.EncoderEngine(ops, self.embed_dim, self.num_heads, len(self.layers), attn_dropout=self.attn_dropout,
                                        relu_dropout=self.relu_dropout, res_dropout=self.res_dropout, embed_dropout=self.dropout,
                                        attn_mask=self.attn_mask, biprojection=self.biprojection, dtype=dt, uid=800,
                                        with_embed=self.with_embed, with_final_ln=self.with_final_ln)
        return self._eng

    def forward(self, x_in, x_in_k=None, x_in_v=None):
        names = [n for n, _ in self.named_parameters()]
        params = [p for _, p in self.named_parameters()]
        return _EncoderFn.apply(self, names, x_in, x_in_k, x_in_v, *params)

    def max_positions(self):
        return self.embed_positions.max_positions()


class _LayerRunner:
    """runs a single TransformerEncoderLayer through a 1-layer EncoderEngine (no embedding, no final LayerNorm)"""

    def __init__(self, layer):
        self.layer = layer
        enc = TransformerEncoder.__new__(TransformerEncoder)
        nn.Module.__init__(enc)
        enc.dropout, enc.attn_dropout, enc.relu_dropout, enc.res_dropout = 0.0, layer.attn_dropout, layer.relu_dropout, layer.res_dropout
        enc.embed_dim, enc.num_heads = layer.embed_dim, layer.num_heads
        enc.attn_mask, enc.biprojection = layer.attn_mask, layer.biprojection
        enc.layers = nn.ModuleList([layer])
        enc.precision = layer.precision
        enc._eng = None
        enc.with_embed, enc.with_final_ln = False, False
        self.enc = enc

    def __call__(self, x, x_k, x_v):
        self.enc.precision = self.layer.precision
        self.enc.train(self.layer.training)
        return self.enc(x, x_k, x_v)


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, names, x_in, x_in_k, x_in_v, *params):
        eng = mod._engine(x_in.device)
        ops, d, dt = eng.ops, eng.d, eng.T_
        T, B, D = x_in.shape
        eng.pack({n: p.detach() for n, p in zip(names, params)})
        xq = ops.empty((B * T, d.Dp), dt)
        ops.stage_rows(x_in.detach().float().permute(1, 0, 2), xq, T)
        xk = xv = None
        S = None
        if x_in_k is not None and x_in_v is not None:
            S = x_in_k.shape[0]
            xk = ops.empty((B * S, d.Dp), dt)
            ops.stage_rows(x_in_k.detach().float().permute(1, 0, 2), xk, S)
            if x_in_v is not x_in_k:
                xv = ops.empty((B * S, d.Dp), dt)
                ops.stage_rows(x_in_v.detach().float().permute(1, 0, 2), xv, S)
        out = eng.forward(xq, B, T, src_k=xk, S=S, src_v=xv, training=mod.training, seed=_next_seed())
        ctx.mod, ctx.names, ctx.dims, ctx.kv = mod, names, (T, B, S, D), (x_in_k is not None, x_in_v is not x_in_k)
        return out.float().view(B, T, d.Dp)[:, :, :D].permute(1, 0, 2).clone()

    @staticmethod
    def backward(ctx, g):
        mod, names = ctx.mod, ctx.names
        T, B, S, D = ctx.dims
        has_kv, v_distinct = ctx.kv
        eng = mod._eng
        ops, d = eng.ops, eng.d
        eng.zero_grads()
        dout = ops.empty((B * T, d.Dp), torch.float32)
        ops.stage_rows(g.float().permute(1, 0, 2), dout, T)
        dq = ops.zeros((B * T, d.Dp), torch.float32)
        dk = ops.zeros((B * S, d.Dp), torch.float32) if has_kv else None
        dv = ops.zeros((B * S, d.Dp), torch.float32) if (has_kv and v_distinct) else None
        eng.backward(dout, dq, dk, dv)
        grads = {n: torch.zeros_like(p) for n, p in zip(names, [p for _, p in mod.named_parameters()])}
        eng.unpack_grads(grads)
        un = lambda t, n: t.view(B, n, d.Dp)[:, :, :D].permute(1, 0, 2).clone()
        gk = un(dk, S) if has_kv else None
        gv = (un(dv, S) if v_distinct else None) if has_kv else None
        return (None, None, un(dq, T), gk, gv) + tuple(grads[n] for n in names)


# ============================================================================================== GMUs
class _SeqGmuBase(nn.Module):
    FEATURES = True

    def __init__(self, size_in1, size_in2, size_out):
        super().__init__()
        self.size_in1, self.size_in2, self.size_out = size_in1, size_in2, size_out
        self.hidden1 = nn.Linear(size_in1, size_out, bias=False)
        self.hidden2 = nn.Linear(size_in2, size_out, bias=False)
        self.x_gate = nn.Linear(size_in1 + size_in2, size_out, bias=False)
        self.precision = "bf16"
        self._eng = None

    def forward(self, xs):
        assert self.size_in1 == self.size_in2 == self.size_out, "the fused GMU kernel needs size_in1 == size_in2 == size_out"
        return _SeqGmuFn.apply(self, xs[0], xs[1], self.hidden1.weight, self.hidden2.weight, self.x_gate.weight)


class GatedMultimodalLayer(_SeqGmuBase):                               # mmtr.py:161-177
    FEATURES = False


class GatedMultimodalLayerFeatures(_SeqGmuBase):                       # mmtr.py:179-195
    FEATURES = True


class _SeqGmuFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, x1, x2, w1, w2, wz):
        ops = _ops_for(x1.device)
        dt = _DT[mod.precision]
        D = mod.size_out
        if mod._eng is None or mod._eng.T_ != dt or mod._eng.ops is not ops:
            mod._eng = E.SeqGmuEngine(ops, D, dt, mod.FEATURES)
        eng = mod._eng
        eng.pack({"hidden1.weight": w1.detach(), "hidden2.weight": w2.detach(), "x_gate.weight": wz.detach()})
        shape = x1.shape
        rows = x1.numel() // D
        a = []
        for x in (x1, x2):
            r = ops.empty((rows, eng.Dp), dt)
            ops.stage_rows(x.detach().float().reshape(1, rows, D), r, rows)
            a.append(r)
        y = eng.forward(a[0], a[1], rows, want_gate=True)
        ctx.mod, ctx.shape, ctx.rows = mod, shape, rows
        z = eng.sv["z"].float()[:, :D].reshape(shape)
        ctx.mark_non_differentiable(z)
        out = y.float()[:, :D].reshape(shape)
        return out, torch.cat((z, 1 - z), dim=-1)

    @staticmethod
    def backward(ctx, g, _gz):
        mod, shape, rows = ctx.mod, ctx.shape, ctx.rows
        eng = mod._eng
        ops, D = eng.ops, eng.D
        eng.zero_grads()
        dy = ops.empty((rows, eng.Dp), torch.float32)
        ops.stage_rows(g.float().reshape(1, rows, D), dy, rows)
        da1, da2 = ops.zeros((rows, eng.Dp), torch.float32), ops.zeros((rows, eng.Dp), torch.float32)
        eng.backward(dy, da1, da2)
        grads = {"hidden1.weight": torch.zeros_like(mod.hidden1.weight), "hidden2.weight": torch.zeros_like(mod.hidden2.weight),
                 "x_gate.weight": torch.zeros_like(mod.x_gate.weight)}
        eng.unpack_grads(grads)
        return (None, da1[:, :D].reshape(shape), da2[:, :D].reshape(shape), grads["hidden1.weight"], grads["hidden2.weight"],
                grads["x_gate.weight"])


class _TextShiftingBase(nn.Module):
    N_IN = 3

    def _build(self, sizes_in, size_out):
        n = len(sizes_in)
        for i, s in enumerate(sizes_in):
            setattr(self, "hidden%d" % (i + 1), nn.Linear(s, size_out, bias=False))
        tot = self._gate_in
        for i in range(n):
            setattr(self, "x%d_gate" % (i + 1), nn.Linear(tot, size_out, bias=False))
        self.size_out = size_out
        self._eng = None

    def forward(self, xs):
        n = self.N_IN
        ws = [getattr(self, "hidden%d" % (i + 1)).weight for i in range(n)] + [getattr(self, "x%d_gate" % (i + 1)).weight for i in range(n)]
        return _TextShiftingFn.apply(self, n, *list(xs[:n]), *ws)


class TextShifting3Layer(_TextShiftingBase):
    """mmtr.py:197-219.  The reference class takes 5 sizes but its only call site passes 4 (mmtr.py:663 -> TypeError as
    shipped); both arities are accepted here, with the intended semantics: 3 inputs, gates over their concatenation."""
    N_IN = 3

    def __init__(self, size_in1, size_in2, size_in3, size_in4, size_out=None):
        super().__init__()
        if size_out is None:
            size_in4, size_out = 0, size_in4
        self.size_in1, self.size_in2, self.size_in3 = size_in1, size_in2, size_in3
        self._gate_in = size_in1 + size_in2 + size_in3 + size_in4
        assert size_in4 == 0 and size_in1 == size_in2 == size_in3 == size_out, "fused head kernel: equal sizes, no 4th gate input"
        self._build([size_in1, size_in2, size_in3], size_out)


class TextShifting4Layer(_TextShiftingBase):                          # mmtr.py:221-247
    N_IN = 4

    def __init__(self, size_in1, size_in2, size_in3, size_in4, size_out):
        super().__init__()
        self._gate_in = size_in1 + size_in2 + size_in3 + size_in4
        assert size_in1 == size_in2 == size_in3 == size_in4 == size_out
        self._build([size_in1, size_in2, size_in3, size_in4], size_out)


class TextShiftingNLayer(nn.Module):                                  # mmtr.py:249-273
    """N-input gate (`hybrid=True` head): parameters `hiddens.i.weight (out, in_i)`, `x_gates.i.weight (out, sum in)`, forward(*xs)."""

    def __init__(self, sizes_in, size_out):
        super().__init__()
        self.sizes_in, self.size_out = list(sizes_in), size_out
        assert all(s == size_out for s in self.sizes_in), "fused head kernel: equal sizes"
        self.hiddens = nn.ModuleList([nn.Linear(s, size_out, bias=False) for s in self.sizes_in])
        self.x_gates = nn.ModuleList([nn.Linear(sum(self.sizes_in), size_out, bias=False) for _ in self.sizes_in])
        self._eng = None

    def forward(self, *xs):
        n = len(self.sizes_in)
        assert len(xs) == n, "TextShiftingNLayer: expected %d inputs" % n
        ws = [h.weight for h in self.hiddens] + [g.weight for g in self.x_gates]
        return _TextShiftingFn.apply(self, n, *xs, *ws)


class _TextShiftingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, n, *args):
        xs, ws = args[:n], args[n:]
        ops = _ops_for(xs[0].device)
        D = mod.size_out
        if mod._eng is None or mod._eng.ops is not ops:
            mod._eng = E.HeadEngine(ops, D, n, 8)
        eng = mod._eng
        B = xs[0].shape[0]
        Dp = eng.Dp
        for i in range(n):
            ops.pack_matrix(ws[i].detach(), eng.W["h"][i])
            ops.pack_matrix(ws[n + i].detach(), eng.W["zg"][i], col_map=(D, Dp))
        cat = eng.cat_buf(B)
        cat.zero_()
        for i in range(n):
            cat[:, i * Dp:i * Dp + D] = xs[i].detach().float()
        fused, z = eng.gate_forward(B)
        ctx.mod, ctx.n, ctx.B = mod, n, B
        ctx.wshapes = [tuple(w.shape) for w in ws]
        zz = z.view(B, n, Dp)[:, :, :D].reshape(B, n * D).clone()
        ctx.mark_non_differentiable(zz)
        return fused[:, :D].clone(), zz

    @staticmethod
    def backward(ctx, g, _gz):
        mod, n, B = ctx.mod, ctx.n, ctx.B
        eng = mod._eng
        ops, D, Dp = eng.ops, eng.D, eng.Dp
        eng.zero_grads()
        df = ops.zeros((B, Dp), torch.float32)
        df[:, :D] = g.float()
        dcat = eng.gate_backward(df)
        gx = tuple(dcat[:, i * Dp:i * Dp + D].clone() for i in range(n))
        gw = []
        for i in range(n):
            t = torch.zeros(ctx.wshapes[i], dtype=torch.float32, device=g.device)
            ops.unpack_matrix(eng.G["h"][i], t)
            gw.append(t)
        for i in range(n):
            t = torch.zeros(ctx.wshapes[n + i], dtype=torch.float32, device=g.device)
            ops.unpack_matrix(eng.G["zg"][i], t, col_map=(D, Dp))
            gw.append(t)
        return (None, None) + gx + tuple(gw)


# ============================================================================================== the model
class FeatureEncoder(nn.Module):
    """stands in for mmtr.BertEncoder (mmtr.py:144-158): the text arrives as a float feature sequence (north star:
    feature extractors are bypassed)."""

    def __init__(self, args=None):
        super().__init__()

    def forward(self, txt, mask=None, segment=None):
        return txt


class MultiprojectionMMTransformer3DGMUClf(nn.Module):
    """"mmtrvat" -- mmtr.py:587-866.  forward(txt, mask, segment, img, audio, output_gate=False) -> logits (B, C)
    [, z (B, 3*D)].  txt (B, L, orig_d_l) float, img (B, T_v, orig_d_v), audio (B, T_a, orig_d_a); all are zero-padded
    to 512 time steps inside (mmtr.py:664-670,756-761)."""

    def __init__(self, args, precision="bf16"):
        super().__init__()
        self.args = args
        self.orig_d_l, self.orig_d_v, self.orig_d_a = args.orig_d_l, args.orig_d_v, args.orig_d_a
        self.d_l = self.d_a = self.d_v = D = args.hidden_sz
        self.vonly, self.lonly, self.aonly = args.vonly, args.lonly, args.aonly
        if not (self.vonly and self.lonly and self.aonly):
            raise NotImplementedError("the reference forward itself requires lonly = aonly = vonly (last_h_* undefined otherwise, mmtr.py:857)")
        if getattr(args, "hybrid", False):
            raise NotImplementedError("hybrid=True is broken in the reference (list-vs-varargs call sites, mmtr.py:572,855); not implemented")
        self.num_heads, self.layers_n = args.num_heads, args.layers
        self.precision = precision
        self.enc = FeatureEncoder(args)
        mk = lambda: GatedMultimodalLayerFeatures(D, D, D)
        self.gmu_l_m, self.gmu_v_m, self.gmu_a_m = mk(), mk(), mk()          # construction order = reference (same RNG stream)
        self.gmu_l, self.gmu_v, self.gmu_a = mk(), mk(), mk()
        self.proj_l = nn.Conv1d(self.orig_d_l, D, kernel_size=1, padding=0, bias=False)
        self.proj_v = nn.Conv1d(self.orig_d_v, D, kernel_size=1, padding=0, bias=False)
        self.proj_a = nn.Conv1d(self.orig_d_a, D, kernel_size=1, padding=0, bias=False)
        for n in ENC_NAMES:
            setattr(self, "trans_" + n, self.get_network(n))
        self.proj1 = nn.Linear(D, D)
        self.proj2 = nn.Linear(D, D)
        self.out_layer = nn.Linear(D, args.n_classes)
        self.gmu = TextShifting3Layer(D, D, D, D)
        self.num_vectors_l = self.num_vectors_a = self.num_vectors_v = 512
        self.transfm_a2l = nn.Linear(512, 512)                                  # present (unused) in the reference: kept for
        self.transfm_v2l = nn.Linear(512, 512)                                  # state_dict compatibility
        self.transfm_l2a = nn.Linear(512, 512)
        self.transfm_l2v = nn.Linear(512, 512)
        self._eng = None

    def get_network(self, name):
        a = self.args
        return TransformerEncoder(embed_dim=a.hidden_sz, num_heads=a.num_heads, layers=a.layers, attn_dropout=attn_dropout_for(name, a),
                                  relu_dropout=a.relu_dropout, res_dropout=a.res_dropout, embed_dropout=a.embed_dropout,
                                  attn_mask=a.attn_mask)

    # ---- engine plumbing
    def trunk_named_parameters(self):
        skip = ("_float_tensor", "version")
        return [(n, p) for n, p in self.named_parameters() if not n.endswith(skip)]

    def engine(self, device=None):
        device = device or next(self.parameters()).device
        ops = _ops_for(device)
        dt = _DT[self.precision]
        if self._eng is None or self._eng.T_ != dt or self._eng.ops is not ops:
            self._eng = MMTrVatEngine(ops, self.args, dtype=dt, n_vec=self.num_vectors_l)
        return self._eng

    def forward(self, txt, mask, segment, img, audio, output_gate=False):
        x_l = self.enc(txt, mask, segment)
        named = self.trunk_named_parameters()
        names = [n for n, _ in named]
        logits, z = _MMTrVatFn.apply(self, names, x_l, img, audio, *[p for _, p in named])
        return (logits, z) if output_gate else logits


class _MMTrVatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, names, txt, img, audio, *params):
        eng = mod.engine(txt.device)
        for nm, t, dim in (("text", txt, mod.orig_d_l), ("img", img, mod.orig_d_v), ("audio", audio, mod.orig_d_a)):
            assert t.dim() == 3 and t.shape[2] == dim, "%s features must be (B, T, %d)" % (nm, dim)
        eng.pack({n: p.detach() for n, p in zip(names, params)})
        logits, z = eng.forward(txt.detach().float(), img.detach().float(), audio.detach().float(), training=mod.training, seed=_next_seed())
        B, C, D, Dp = txt.shape[0], mod.args.n_classes, mod.args.hidden_sz, eng.d.Dp
        ctx.mod, ctx.names, ctx.shapes = mod, names, (txt.shape, img.shape, audio.shape)
        ctx.need_in = (txt.requires_grad, img.requires_grad, audio.requires_grad)
        zz = z.view(B, 3, Dp)[:, :, :D].reshape(B, 3 * D).clone()
        ctx.mark_non_differentiable(zz)
        return logits[:, :C].clone(), zz

    @staticmethod
    def backward(ctx, g, _gz):
        mod, names = ctx.mod, ctx.names
        eng = mod._eng
        ops = eng.ops
        B, C = g.shape
        dl = ops.zeros((B, eng.head.Cp), torch.float32)
        dl[:, :C] = g.float()
        eng.zero_grads()
        d_in = {}
        for m, need, shp in zip("lva", ctx.need_in, ctx.shapes):
            if need:
                d_in[m] = ops.zeros(tuple(shp), torch.float32)
        eng.backward(dl, d_in)
        pmap = dict(mod.trunk_named_parameters())
        grads = {n: torch.zeros_like(pmap[n]) for n in eng.param_shapes()}
        eng.unpack_grads(grads)
        return (None, None, d_in.get("l"), d_in.get("v"), d_in.get("a")) + tuple(grads.get(n) for n in names)


class AudioFeatures(nn.Module):
    """stands in for mmtr.AudioEncoder (mmtr.py:452): the audio arrives as post-encoder features (B, T_a, orig_d_a)."""

    def __init__(self, args=None):
        super().__init__()

    def forward(self, audio):
        return audio


class MultiprojectionMMTransformerGMUClf(nn.Module):
    """"mmtrvapt" -- mmtr.py:278-583.  forward(txt, mask, segment, img, audio, poster, output_gate=False) -> logits (B, C) [, z (B, 4*D)].
    txt (B, L <= 512, orig_d_l) float, img (B, T_v <= 200, orig_d_v), audio (B, T_a <= 200, orig_d_a) post-encoder features,
    poster (B, orig_d_p).  Sequences are zero-padded to 512 / 200 / 200 inside (mmtr.py:371-373, 461-466)."""

    def __init__(self, args, precision="bf16"):
        super().__init__()
        from .model_engine4 import BIPROJ, NV, TRANSFM
        self.args = args
        self.orig_d_l, self.orig_d_v, self.orig_d_a, self.orig_d_p = args.orig_d_l, args.orig_d_v, args.orig_d_a, args.orig_d_p
        self.d_l = self.d_a = self.d_v = D = args.hidden_sz
        self.vonly, self.lonly, self.aonly = args.vonly, args.lonly, args.aonly
        if not (self.vonly and self.lonly and self.aonly):
            raise NotImplementedError("the reference forward itself requires lonly = aonly = vonly (last_h_* undefined otherwise, mmtr.py:572)")
        if getattr(args, "hybrid", False):
            raise NotImplementedError("hybrid=True is broken in the reference (list-vs-varargs call sites, mmtr.py:572); not implemented")
        self.precision = precision
        self.enc = FeatureEncoder(args)
        self.audio_enc = AudioFeatures(args)
        self.proj_poster = nn.Linear(self.orig_d_p, D, bias=False)             # construction order = reference (mmtr.py:310-378)
        mk = lambda: GatedMultimodalLayerFeatures(D, D, D)
        self.gmu_l_m, self.gmu_v_m, self.gmu_a_m = mk(), mk(), mk()
        self.gmu_l, self.gmu_v, self.gmu_a = mk(), mk(), mk()
        self.proj_l = nn.Conv1d(self.orig_d_l, D, kernel_size=1, padding=0, bias=False)
        self.proj_v = nn.Conv1d(self.orig_d_v, D, kernel_size=1, padding=0, bias=False)
        self.proj_a = nn.Conv1d(self.orig_d_a, D, kernel_size=1, padding=0, bias=False)
        for n in ("l_with_a", "l_with_v", "l_with_v2a", "l_with_a2v", "v_with_l", "v_with_a", "v_with_l2a", "v_with_a2l",
                  "a_with_l", "a_with_v", "a_with_v2l", "a_with_l2v"):
            setattr(self, "trans_" + n, self.get_network(n, biprojection=n in BIPROJ))
        self.proj1 = nn.Linear(D, D)
        self.proj2 = nn.Linear(D, D)
        self.out_layer = nn.Linear(D, args.n_classes)
        self.gmu = TextShifting4Layer(D, D, D, D, D)
        self.num_vectors_l, self.num_vectors_a, self.num_vectors_v = NV["l"], NV["a"], NV["v"]
        for n in ("a2l", "v2l", "l2a", "l2v"):
            ti, to = TRANSFM[n]
            setattr(self, "transfm_" + n, nn.Linear(NV[ti], NV[to]))
        self._eng = None

    def get_network(self, name, biprojection=False):
        a = self.args
        return TransformerEncoder(embed_dim=a.hidden_sz, num_heads=a.num_heads, layers=a.layers, attn_dropout=attn_dropout_for(name, a),
                                  relu_dropout=a.relu_dropout, res_dropout=a.res_dropout, embed_dropout=a.embed_dropout,
                                  attn_mask=a.attn_mask, biprojection=biprojection)

    def trunk_named_parameters(self):
        skip = ("_float_tensor", "version")
        return [(n, p) for n, p in self.named_parameters() if not n.endswith(skip)]

    def engine(self, device=None):
        from .model_engine4 import MMTrVaptEngine
        device = device or next(self.parameters()).device
        ops = _ops_for(device)
        dt = _DT[self.precision]
        if self._eng is None or self._eng.T_ != dt or self._eng.ops is not ops:
            self._eng = MMTrVaptEngine(ops, self.args, dtype=dt)
        return self._eng

    def forward(self, txt, mask, segment, img, audio, poster, output_gate=False):
        x_l = self.enc(txt, mask, segment)
        x_a = self.audio_enc(audio)
        named = self.trunk_named_parameters()
        names = [n for n, _ in named]
        logits, z = _MMTrVaptFn.apply(self, names, x_l, img, x_a, poster, *[p for _, p in named])
        return (logits, z) if output_gate else logits


class _MMTrVaptFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, names, txt, img, audio, poster, *params):
        eng = mod.engine(txt.device)
        for nm, t, dim in (("text", txt, mod.orig_d_l), ("img", img, mod.orig_d_v), ("audio", audio, mod.orig_d_a)):
            assert t.dim() == 3 and t.shape[2] == dim, "%s features must be (B, T, %d)" % (nm, dim)
        assert poster.dim() == 2 and poster.shape[1] == mod.orig_d_p, "poster must be (B, %d)" % mod.orig_d_p
        assert txt.shape[1] <= mod.num_vectors_l and audio.shape[1] <= mod.num_vectors_a and img.shape[1] <= mod.num_vectors_v, \
            "a sequence exceeds its fixed length (the reference raises on a negative pad size, mmtr.py:431-441)"
        eng.pack({n: p.detach() for n, p in zip(names, params)})
        logits, z = eng.forward(txt.detach().float(), img.detach().float(), audio.detach().float(), poster.detach().float(),
                                training=mod.training, seed=_next_seed())
        B, C, D, Dp = txt.shape[0], mod.args.n_classes, mod.args.hidden_sz, eng.d.Dp
        ctx.mod, ctx.names, ctx.shapes = mod, names, (txt.shape, img.shape, audio.shape)
        ctx.need_in = (txt.requires_grad, img.requires_grad, audio.requires_grad)
        zz = z.view(B, 4, Dp)[:, :, :D].reshape(B, 4 * D).clone()
        ctx.mark_non_differentiable(zz)
        return logits[:, :C].clone(), zz

    @staticmethod
    def backward(ctx, g, _gz):
        mod, names = ctx.mod, ctx.names
        eng = mod._eng
        ops = eng.ops
        B, C = g.shape
        dl = ops.zeros((B, eng.head.Cp), torch.float32)
        dl[:, :C] = g.float()
        eng.zero_grads()
        d_in = {}
        for m, need, shp in zip("lva", ctx.need_in, ctx.shapes):
            if need:
                d_in[m] = ops.zeros(tuple(shp), torch.float32)
        eng.backward(dl, d_in)
        pmap = dict(mod.trunk_named_parameters())
        grads = {n: torch.zeros_like(pmap[n]) for n in eng.param_shapes()}
        eng.unpack_grads(grads)
        return (None, None, d_in.get("l"), d_in.get("v"), d_in.get("a"), None) + tuple(grads.get(n) for n in names)     # (no poster input gradient)


MODELS = {"mmtrvat": MultiprojectionMMTransformer3DGMUClf, "mmtrvapt": MultiprojectionMMTransformerGMUClf}    # models/__init__.py:6-9


def get_model(args):                                                   # models/__init__.py:12-14
    return MODELS[args.model](args)
