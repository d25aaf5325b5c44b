"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU restatement (plain torch, dtype-generic, autograd-differentiable) of the reference algorithm on the
hot path: BPMulT fusion trunk, dropout = 0.  Each function cites the reference lines it follows
(paths relative to /root/reference/bpmult).  It takes a reference-named `state_dict` so that the same
weights drive the reference, this oracle and the CUDA path.

PARITY PIN: the reference ships no tests / golden vectors, so this oracle is pinned
against outputs of the reference itself, run here with the shims in `oracle/ref_shim.py`:
`oracle/make_golden.py` asserts restatement == reference (max-abs diff 0.0 .. 2e-6 in fp32) and writes the
fixtures in `tests/golden/`.  `tests/test_oracle.py` re-checks the restatement against those fixtures
(and against the live reference when `/root/reference` is present).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline/reference legs may import this.
"""
import math

import torch
import torch.nn.functional as F


PROBE = None        # optional dict filled with intermediate activations by the functions below (test diagnostics only)


# ----------------------------------------------------------------------------- positional embedding
def sinusoid_table(num_pos, dim, dtype=torch.float32):
    """models/position_embedding.py:44-60 (get_embedding): halves concatenated [sin | cos], row 0 zeroed,
    an extra zero column when dim is odd.  Computed in fp32 exactly like the reference, then cast."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(num_pos, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(num_pos, -1)
    if dim % 2 == 1:
        e = torch.cat([e, torch.zeros(num_pos, 1)], dim=1)
    e[0, :] = 0
    return e.to(dtype)


def positional_embedding(x_in):
    """models/position_embedding.py:8-27,62-76 as called at models/transformer.py:68:
    position(t, b) = t + 1 if x_in[t, b, 0] != 0 else 0 (padding_idx = 0, left_pad = 0).  x_in is (T, B, D).
    Returns (T, B, D), detached."""
    T, B, D = x_in.shape
    tab = sinusoid_table(T + 1, D, x_in.dtype).to(x_in.device)
    ch0 = x_in[:, :, 0]
    pos = torch.arange(1, T + 1, device=x_in.device).unsqueeze(1).expand(T, B)
    pos = torch.where(ch0 != 0, pos, torch.zeros_like(pos))
    return tab[pos.reshape(-1)].view(T, B, D).detach()


# ----------------------------------------------------------------------------- attention
def future_mask(T, S, dtype, device="cpu"):
    """models/transformer.py:209-216: M[i, j] = -inf if j - i >= 1 + |S - T| else 0."""
    i = torch.arange(T, device=device).unsqueeze(1)
    j = torch.arange(S, device=device).unsqueeze(0)
    m = torch.zeros(T, S, dtype=dtype, device=device)
    m[(j - i) >= 1 + abs(S - T)] = float("-inf")
    return m


def multihead_attention(sd, pfx, query, key, value, num_heads, mask=None, need_weights=False):
    """models/multihead_attention.py:52-135.  in_proj rows [0:D]=Q, [D:2D]=K, [2D:3D]=V (:137-158);
    q scaled by head_dim^-0.5 AFTER the bias (:86); heads = contiguous column blocks (:95-99);
    softmax over keys in fp32-or-wider (:121); out_proj (:130)."""
    T, B, D = query.shape
    S = key.shape[0]
    H = num_heads
    dh = D // H
    W, bias = sd[pfx + "in_proj_weight"], sd[pfx + "in_proj_bias"]
    q = F.linear(query, W[:D], bias[:D]) * dh ** -0.5
    k = F.linear(key, W[D:2 * D], bias[D:2 * D])
    v = F.linear(value, W[2 * D:], bias[2 * D:])
    q = q.contiguous().view(T, B * H, dh).transpose(0, 1)
    k = k.contiguous().view(S, B * H, dh).transpose(0, 1)
    v = v.contiguous().view(S, B * H, dh).transpose(0, 1)
    w = torch.bmm(q, k.transpose(1, 2))
    if mask is not None:
        w = w + mask.unsqueeze(0)
    w = F.softmax(w, dim=-1)
    a = torch.bmm(w, v).transpose(0, 1).contiguous().view(T, B, D)
    a = F.linear(a, sd[pfx + "out_proj.weight"], sd[pfx + "out_proj.bias"])
    if need_weights:
        return a, w.view(B, H, T, S).sum(dim=1) / H
    return a, None


def _ln(sd, pfx, x):
    return F.layer_norm(x, (x.shape[-1],), sd[pfx + "weight"], sd[pfx + "bias"], 1e-5)


def encoder_layer(sd, pfx, x, x_k, x_v, num_heads, attn_mask, biprojection=False):
    """models/transformer.py:141-195 (pre-norm: normalize_before = True, :132)."""
    T = x.shape[0]
    mk = lambda S: future_mask(T, S, x.dtype, x.device) if attn_mask else None
    residual = x
    xn = _ln(sd, pfx + "layer_norms.0.", x)
    if x_k is None and x_v is None:                                        # :158-159 self-attention only
        a, _ = multihead_attention(sd, pfx + "self_attn.", xn, xn, xn, num_heads, mk(T))
        x = residual + a
        ffn_ln = 1
    elif biprojection:                                                     # :160-169
        a, _ = multihead_attention(sd, pfx + "self_attn.", xn, xn, xn, num_heads, mk(T))
        x = residual + a
        residual = x
        kn = _ln(sd, pfx + "layer_norms.1.", x_k)
        vn = _ln(sd, pfx + "layer_norms.1.", x_v)
        a, _ = multihead_attention(sd, pfx + "self_attn.", x, kn, vn, num_heads, mk(x_k.shape[0]))
        x = residual + a
        ffn_ln = 2
    else:                                                                  # :170-173
        kn = _ln(sd, pfx + "layer_norms.0.", x_k)
        vn = _ln(sd, pfx + "layer_norms.0.", x_v)
        a, _ = multihead_attention(sd, pfx + "self_attn.", xn, kn, vn, num_heads, mk(x_k.shape[0]))
        x = residual + a
        ffn_ln = 1
    residual = x                                                           # :181-190
    h = _ln(sd, pfx + "layer_norms.%d." % ffn_ln, x)
    pre = F.linear(h, sd[pfx + "fc1.weight"], sd[pfx + "fc1.bias"])
    if PROBE is not None:                                                  # tests: FFN pre-activations (ReLU ties, see tests/test_fullshape_gpu.py)
        PROBE[pfx + "fc1_pre"] = pre.detach()
    h = F.relu(pre)
    h = F.linear(h, sd[pfx + "fc2.weight"], sd[pfx + "fc2.bias"])
    return residual + h


def transformer_encoder(sd, pfx, x_in, x_in_k, x_in_v, num_heads, layers, attn_mask, biprojection=False):
    """models/transformer.py:52-93.  (T, B, D) time-major in and out."""
    D = x_in.shape[-1]
    x = math.sqrt(D) * x_in + positional_embedding(x_in)
    x_k = x_v = None
    if x_in_k is not None and x_in_v is not None:
        x_k = math.sqrt(D) * x_in_k + positional_embedding(x_in_k)
        x_v = math.sqrt(D) * x_in_v + positional_embedding(x_in_v)
    for l in range(layers):
        x = encoder_layer(sd, "%slayers.%d." % (pfx, l), x, x_k, x_v, num_heads, attn_mask, biprojection)
    return _ln(sd, pfx + "layer_norm.", x)


# ----------------------------------------------------------------------------- GMUs
def gmu_features(sd, pfx, x1, x2):
    """models/mmtr.py:179-195 GatedMultimodalLayerFeatures."""
    h1 = torch.tanh(F.linear(x1, sd[pfx + "hidden1.weight"]))
    h2 = torch.tanh(F.linear(x2, sd[pfx + "hidden2.weight"]))
    z = torch.sigmoid(F.linear(torch.cat([x1, x2], -1), sd[pfx + "x_gate.weight"]))
    return z * h1 * x1 + (1 - z) * h2 * x2, torch.cat((z, 1 - z), -1)


def gmu_plain(sd, pfx, x1, x2):
    """models/mmtr.py:161-177 GatedMultimodalLayer."""
    h1 = torch.tanh(F.linear(x1, sd[pfx + "hidden1.weight"]))
    h2 = torch.tanh(F.linear(x2, sd[pfx + "hidden2.weight"]))
    z = torch.sigmoid(F.linear(torch.cat([x1, x2], -1), sd[pfx + "x_gate.weight"]))
    return z * h1 + (1 - z) * h2, torch.cat((z, 1 - z), -1)


def text_shifting(sd, pfx, xs):
    """models/mmtr.py:197-247 TextShifting3Layer / TextShifting4Layer (N = len(xs))."""
    cat = torch.cat(xs, -1)
    hs = [torch.tanh(F.linear(x, sd["%shidden%d.weight" % (pfx, i + 1)])) for i, x in enumerate(xs)]
    zs = [torch.sigmoid(F.linear(cat, sd["%sx%d_gate.weight" % (pfx, i + 1)])) for i in range(len(xs))]
    out = zs[0] * hs[0]
    for z, h in zip(zs[1:], hs[1:]):
        out = out + z * h
    return out, torch.cat(zs, -1)


def text_shifting_n(sd, pfx, xs):
    """models/mmtr.py:249-273 TextShiftingNLayer (ModuleList parameter names hiddens.i / x_gates.i; forward(*xs))."""
    cat = torch.cat(xs, -1)
    hs = [torch.tanh(F.linear(x, sd["%shiddens.%d.weight" % (pfx, i)])) for i, x in enumerate(xs)]
    zs = [torch.sigmoid(F.linear(cat, sd["%sx_gates.%d.weight" % (pfx, i)])) for i in range(len(xs))]
    out = zs[0] * hs[0]
    for z, h in zip(zs[1:], hs[1:]):
        out = out + z * h
    return out, torch.cat(zs, -1)


# ----------------------------------------------------------------------------- loss
def bce_with_logits(logits, targets, pos_weight=None):
    """train.py:99-106,333: nn.BCEWithLogitsLoss(pos_weight=w), mean over (B, C)."""
    return F.binary_cross_entropy_with_logits(logits, targets, pos_weight=pos_weight)


# ----------------------------------------------------------------------------- mmtrvat
def _pad_time(x, n):
    """models/mmtr.py:722-732 transfm_2dim(dim=0)."""
    if x.shape[0] == n:
        return x
    return torch.cat([x, torch.zeros(n - x.shape[0], x.shape[1], x.shape[2], dtype=x.dtype, device=x.device)], 0)


def mmtrvat_forward(sd, cfg, txt, img, audio, n_vec=512, return_intermediates=False):
    """models/mmtr.py:735-866 MultiprojectionMMTransformer3DGMUClf.forward, dropout = 0, BERT bypassed (txt is float
    (B, L, orig_d_l)).  cfg: hidden_sz, num_heads, layers, attn_mask, orig_d_*, hybrid.
    hybrid (:765-775, :854-855; ctor :631, :662, :680-689), with the two gate call sites read as what their callees accept
    (oracle/ref_shim.py shim 6): every projected stream goes through a bias-free Linear over its TIME axis (512 -> 32 steps), a
    SELF-attention encoder of max(layers, 3) layers, first + last step pooling; gmu_early fuses the three; the result is the fourth
    input of the final TextShiftingNLayer."""
    D, H, L, am = cfg.hidden_sz, cfg.num_heads, cfg.layers, cfg.attn_mask
    hybrid = bool(getattr(cfg, "hybrid", False))

    def proj(x, key, orig_d):                                              # :742-753
        x = x.transpose(1, 2)
        if orig_d != D:
            x = F.conv1d(x, sd[key])
        return _pad_time(x.permute(2, 0, 1), n_vec)                        # :756-761

    p_l = proj(txt, "proj_l.weight", cfg.orig_d_l)
    p_a = proj(audio, "proj_a.weight", cfg.orig_d_a)
    p_v = proj(img, "proj_v.weight", cfg.orig_d_v)
    enc = lambda name, q, kv: transformer_encoder(sd, "trans_%s." % name, q, kv, kv, H, L, am)
    last_early = None
    if hybrid:                                                             # :765-775
        pooled = []
        for m, p in (("l", p_l), ("v", p_v), ("a", p_a)):
            pe = F.linear(p.permute(2, 1, 0), sd["proj_%s_e.weight" % m]).permute(2, 1, 0)          # (32, B, D)
            he = transformer_encoder(sd, "trans_%s_early." % m, pe, None, None, H, max(L, 3), am)
            pooled.append(he[0] + he[-1])
        last_early, _ = text_shifting(sd, "gmu_early.", pooled)
    h = {}
    h["v_with_a"] = enc("v_with_a", p_v, p_a)                              # :779-786
    h["a_with_v"] = enc("a_with_v", p_a, p_v)
    h["v_with_l"] = enc("v_with_l", p_v, p_l)
    h["l_with_v"] = enc("l_with_v", p_l, p_v)
    h["a_with_l"] = enc("a_with_l", p_a, p_l)
    h["l_with_a"] = enc("l_with_a", p_l, p_a)

    # target l (:788-808)
    l_v2a = enc("l_with_v2a", p_l, h["a_with_v"])
    l_a2v = enc("l_with_a2v", p_l, h["v_with_a"])
    mid, _ = gmu_features(sd, "gmu_l_m.", h["v_with_a"], h["a_with_v"])
    top, _ = gmu_features(sd, "gmu_l.", l_a2v + h["v_with_a"], l_v2a + h["a_with_v"])
    top_l = top + mid
    last_l = top_l[0] + top_l[-1]
    # target a (:810-830)
    a_v2l = enc("a_with_v2l", p_a, h["l_with_v"])
    a_l2v = enc("a_with_l2v", p_a, h["v_with_l"])
    mid, _ = gmu_features(sd, "gmu_a_m.", h["l_with_v"], h["v_with_l"])
    top, _ = gmu_features(sd, "gmu_a.", a_v2l + h["l_with_v"], a_l2v + h["v_with_l"])
    top_a = top + mid
    last_a = top_a[0] + top_a[-1]
    # target v (:832-852)
    v_a2l = enc("v_with_a2l", p_v, h["l_with_a"])
    v_l2a = enc("v_with_l2a", p_v, h["a_with_l"])
    mid, _ = gmu_features(sd, "gmu_v_m.", h["l_with_a"], h["a_with_l"])
    top, _ = gmu_features(sd, "gmu_v.", v_a2l + h["l_with_a"], v_l2a + h["a_with_l"])
    top_v = top + mid
    last_v = top_v[0] + top_v[-1]
    # head (:857-866)
    if hybrid:
        fused, z = text_shifting_n(sd, "gmu.", [last_l, last_v, last_a, last_early])                  # :854-855
    else:
        fused, z = text_shifting(sd, "gmu.", [last_l, last_v, last_a])
    y = F.linear(F.relu(F.linear(fused, sd["proj1.weight"], sd["proj1.bias"])),
                 sd["proj2.weight"], sd["proj2.bias"]) + fused
    logits = F.linear(y, sd["out_layer.weight"], sd["out_layer.bias"])
    if return_intermediates:
        return logits, z, dict(h, p_l=p_l, p_a=p_a, p_v=p_v, last_l=last_l, last_a=last_a, last_v=last_v,
                               l_with_v2a=l_v2a, l_with_a2v=l_a2v, fused=fused)
    return logits, z


# ----------------------------------------------------------------------------- mmtrvapt
def time_linear(sd, name, h):
    """models/mmtr.py:507-508,530,553: transfm_x2y(h.permute(2,1,0)).permute(2,1,0) -- nn.Linear over the TIME axis of (T, B, D)."""
    return F.linear(h.permute(2, 1, 0), sd["transfm_%s.weight" % name], sd["transfm_%s.bias" % name]).permute(2, 1, 0)


def audio_encoder(sd, pfx, x, out_len=200):
    """models/mmtr.py:93-108 AudioEncoder: Conv1d(96, 96, 128, stride=2) x 2 + AdaptiveAvgPool1d(200); x (B, 96, T_raw) -> (B, 96, 200)."""
    x = F.conv1d(x, sd[pfx + "conv_layers.0.weight"], sd[pfx + "conv_layers.0.bias"], stride=2)
    x = F.conv1d(x, sd[pfx + "conv_layers.1.weight"], sd[pfx + "conv_layers.1.bias"], stride=2)
    return F.adaptive_avg_pool1d(x, out_len)


def mmtrvapt_forward(sd, cfg, txt, img, audio, poster, nv=(512, 200, 200)):
    """models/mmtr.py:444-583 MultiprojectionMMTransformerGMUClf.forward, hybrid = False, dropout = 0, BERT and AudioEncoder bypassed
    (txt float (B, T_l, orig_d_l); audio float (B, T_a, orig_d_a) post-encoder); wave-2 encoders are biprojection layers (:342-353)."""
    D, H, L, am = cfg.hidden_sz, cfg.num_heads, cfg.layers, cfg.attn_mask
    nl, na, nvv = nv

    def proj(x, key, orig_d, n):                                           # :448-468
        x = x.transpose(1, 2)
        if orig_d != D:
            x = F.conv1d(x, sd[key])
        return _pad_time(x.permute(2, 0, 1), n)

    p_l = proj(txt, "proj_l.weight", cfg.orig_d_l, nl)
    if "audio_enc.conv_layers.0.weight" in sd:                             # :452 x_a = self.audio_enc(audio): raw (B, 96, T_raw) -> (B, 96, 200)
        audio = audio_encoder(sd, "audio_enc.", audio).transpose(1, 2)
    p_a = proj(audio, "proj_a.weight", cfg.orig_d_a, na)
    p_v = proj(img, "proj_v.weight", cfg.orig_d_v, nvv)
    post = F.linear(poster, sd["proj_poster.weight"])                      # :486
    enc = lambda name, q, kv, bi=False: transformer_encoder(sd, "trans_%s." % name, q, kv, kv, H, L, am, biprojection=bi)
    h = {}
    h["v_with_a"] = enc("v_with_a", p_v, p_a)                              # :491-498
    h["a_with_v"] = enc("a_with_v", p_a, p_v)
    h["v_with_l"] = enc("v_with_l", p_v, p_l)
    h["l_with_v"] = enc("l_with_v", p_l, p_v)
    h["a_with_l"] = enc("a_with_l", p_a, p_l)
    h["l_with_a"] = enc("l_with_a", p_l, p_a)
    # target l (:501-523)
    l_v2a = enc("l_with_v2a", p_l, h["a_with_v"], True)
    l_a2v = enc("l_with_a2v", p_l, h["v_with_a"], True)
    t_a = time_linear(sd, "a2l", h["a_with_v"])
    t_v = time_linear(sd, "v2l", h["v_with_a"])
    mid, _ = gmu_features(sd, "gmu_l_m.", t_v, t_a)
    top, _ = gmu_features(sd, "gmu_l.", l_a2v + t_v, l_v2a + t_a)
    top = top + mid
    last_l = top[0] + top[-1]
    # target a (:525-546)
    a_v2l = enc("a_with_v2l", p_a, h["l_with_v"], True)
    a_l2v = enc("a_with_l2v", p_a, h["v_with_l"], True)
    t_l = time_linear(sd, "l2a", h["l_with_v"])
    t_v = h["v_with_l"]
    mid, _ = gmu_features(sd, "gmu_a_m.", t_l, t_v)
    top, _ = gmu_features(sd, "gmu_a.", a_v2l + t_l, a_l2v + t_v)
    top = top + mid
    last_a = top[0] + top[-1]
    # target v (:548-569)
    v_a2l = enc("v_with_a2l", p_v, h["l_with_a"], True)
    v_l2a = enc("v_with_l2a", p_v, h["a_with_l"], True)
    t_l = time_linear(sd, "l2v", h["l_with_a"])
    t_a = h["a_with_l"]
    mid, _ = gmu_features(sd, "gmu_v_m.", t_l, t_a)
    top, _ = gmu_features(sd, "gmu_v.", v_a2l + t_l, v_l2a + t_a)
    top = top + mid
    last_v = top[0] + top[-1]
    # head (:574-583)
    fused, z = text_shifting(sd, "gmu.", [last_l, last_v, last_a, post])
    y = F.linear(F.relu(F.linear(fused, sd["proj1.weight"], sd["proj1.bias"])), sd["proj2.weight"], sd["proj2.bias"]) + fused
    return F.linear(y, sd["out_layer.weight"], sd["out_layer.bias"]), z


# ----------------------------------------------------------------------------- error metrics
def max_rel(a, b):
    """max|a-b| / max|b| per tensor (logits / activations)."""
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    """||a-b|| / ||b|| per tensor (gradients)."""
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
