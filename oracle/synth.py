"""TEST INFRASTRUCTURE ONLY.  Deterministic synthetic weights / inputs shared by the golden generator,
the parity tests and the CPU-baseline leg of bench.py.  Weights are generated from a seed (not stored), so
fixtures stay small: the reference, the oracle and the CUDA path all load the same generated state_dict.
Key names / shapes follow SURVEY.md section 8b ([probe] 233 keys for mmtrvat) and are verified against the
reference's own state_dict by oracle/make_golden.py."""
from argparse import Namespace

import torch

ENC_NAMES = ["l_with_a", "l_with_v", "l_with_v2a", "l_with_a2v",
             "v_with_l", "v_with_a", "v_with_l2a", "v_with_a2l",
             "a_with_l", "a_with_v", "a_with_v2l", "a_with_l2v"]          # ctor order, mmtr.py:639-653


def encoder_shapes(D, L, biprojection=False, pfx=""):
    s = {}
    for i in range(L):
        p = "%slayers.%d." % (pfx, i)
        s[p + "self_attn.in_proj_weight"] = (3 * D, D)
        s[p + "self_attn.in_proj_bias"] = (3 * D,)
        s[p + "self_attn.out_proj.weight"] = (D, D)
        s[p + "self_attn.out_proj.bias"] = (D,)
        s[p + "fc1.weight"] = (4 * D, D)
        s[p + "fc1.bias"] = (4 * D,)
        s[p + "fc2.weight"] = (D, 4 * D)
        s[p + "fc2.bias"] = (D,)
        for j in range(3 if biprojection else 2):
            s[p + "layer_norms.%d.weight" % j] = (D,)
            s[p + "layer_norms.%d.bias" % j] = (D,)
    s[pfx + "layer_norm.weight"] = (D,)
    s[pfx + "layer_norm.bias"] = (D,)
    return s


def mmtrvat_shapes(cfg, n_vec=512):
    D, C = cfg.hidden_sz, cfg.n_classes
    s = {}
    for m in ("l_m", "v_m", "a_m", "l", "v", "a"):
        s["gmu_%s.hidden1.weight" % m] = (D, D)
        s["gmu_%s.hidden2.weight" % m] = (D, D)
        s["gmu_%s.x_gate.weight" % m] = (D, 2 * D)
    s["proj_l.weight"] = (D, cfg.orig_d_l, 1)
    s["proj_v.weight"] = (D, cfg.orig_d_v, 1)
    s["proj_a.weight"] = (D, cfg.orig_d_a, 1)
    for n in ENC_NAMES:
        s.update(encoder_shapes(D, cfg.layers, False, "trans_%s." % n))
    s["proj1.weight"] = (D, D); s["proj1.bias"] = (D,)
    s["proj2.weight"] = (D, D); s["proj2.bias"] = (D,)
    s["out_layer.weight"] = (C, D); s["out_layer.bias"] = (C,)
    hybrid = bool(getattr(cfg, "hybrid", False))
    if hybrid:                                                             # TextShiftingNLayer([D] * 4, D), mmtr.py:662
        for i in range(4):
            s["gmu.hiddens.%d.weight" % i] = (D, D)
        for i in range(4):
            s["gmu.x_gates.%d.weight" % i] = (D, 4 * D)
    else:
        for i in (1, 2, 3):
            s["gmu.hidden%d.weight" % i] = (D, D)
        for i in (1, 2, 3):
            s["gmu.x%d_gate.weight" % i] = (D, 3 * D)
    for n in ("a2l", "v2l", "l2a", "l2v"):                                 # unused by forward (no grad)
        s["transfm_%s.weight" % n] = (n_vec, n_vec)
        s["transfm_%s.bias" % n] = (n_vec,)
    if hybrid:                                                             # mmtr.py:631, 680-689
        for i in (1, 2, 3):
            s["gmu_early.hidden%d.weight" % i] = (D, D)
        for i in (1, 2, 3):
            s["gmu_early.x%d_gate.weight" % i] = (D, 3 * D)
        for m in "lva":
            s.update(encoder_shapes(D, max(cfg.layers, 3), False, "trans_%s_early." % m))
            s["proj_%s_e.weight" % m] = (32, n_vec)
    return s


BIPROJ = ("l_with_v2a", "l_with_a2v", "v_with_l2a", "v_with_a2l", "a_with_v2l", "a_with_l2v")      # mmtr.py:342-353
NV = {"l": 512, "a": 200, "v": 200}                                          # mmtr.py:371-373 num_vectors_{l,a,v}


def mmtrvapt_shapes(cfg):
    """state_dict of MultiprojectionMMTransformerGMUClf (mmtr.py:278-396), hybrid = False, BERT / AudioEncoder bypassed"""
    D, C = cfg.hidden_sz, cfg.n_classes
    s = {"proj_poster.weight": (D, cfg.orig_d_p)}
    for m in ("l_m", "v_m", "a_m", "l", "v", "a"):
        s["gmu_%s.hidden1.weight" % m] = (D, D)
        s["gmu_%s.hidden2.weight" % m] = (D, D)
        s["gmu_%s.x_gate.weight" % m] = (D, 2 * D)
    s["proj_l.weight"] = (D, cfg.orig_d_l, 1)
    s["proj_v.weight"] = (D, cfg.orig_d_v, 1)
    s["proj_a.weight"] = (D, cfg.orig_d_a, 1)
    for n in ENC_NAMES:
        s.update(encoder_shapes(D, cfg.layers, n in BIPROJ, "trans_%s." % n))
    s["proj1.weight"] = (D, D); s["proj1.bias"] = (D,)
    s["proj2.weight"] = (D, D); s["proj2.bias"] = (D,)
    s["out_layer.weight"] = (C, D); s["out_layer.bias"] = (C,)
    for i in (1, 2, 3, 4):
        s["gmu.hidden%d.weight" % i] = (D, D)
    for i in (1, 2, 3, 4):
        s["gmu.x%d_gate.weight" % i] = (D, 4 * D)
    for n, (ti, to) in (("a2l", (NV["a"], NV["l"])), ("v2l", (NV["v"], NV["l"])), ("l2a", (NV["l"], NV["a"])), ("l2v", (NV["l"], NV["v"]))):
        s["transfm_%s.weight" % n] = (to, ti)
        s["transfm_%s.bias" % n] = (to,)
    return s


def audio_encoder_shapes(C=96, pfx=""):
    return {pfx + "conv_layers.0.weight": (C, C, 128), pfx + "conv_layers.0.bias": (C,),
            pfx + "conv_layers.1.weight": (C, C, 128), pfx + "conv_layers.1.bias": (C,)}


def mmtrvapt_inputs(cfg, B, T_l, T_a, T_v, seed=2024):
    """text (B, T_l, orig_d_l), video (B, T_v, orig_d_v), audio (B, T_a, orig_d_a) post-encoder features, poster (B, orig_d_p), targets"""
    g = torch.Generator().manual_seed(seed)
    txt = torch.randn(B, T_l, cfg.orig_d_l, generator=g)
    img = torch.randn(B, T_v, cfg.orig_d_v, generator=g)
    audio = torch.randn(B, T_a, cfg.orig_d_a, generator=g)
    poster = torch.randn(B, cfg.orig_d_p, generator=g)
    tgt = (torch.rand(B, cfg.n_classes, generator=g) < 0.3).float()
    return txt, img, audio, poster, tgt


def make_state_dict(shapes, seed, dtype=torch.float32, gain=1.0):
    """randn / sqrt(fan_in) matrices, LayerNorm weight 1 + 0.1 randn, non-zero biases 0.1 randn --
    biases are zero at reference init (transformer.py:219-224) but a trained checkpoint has them,
    so parity is exercised with them non-zero."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if len(shp) >= 2:
            fan_in = 1
            for x in shp[1:]:                                      # (conv weights: in_channels * kernel_size; k = 1 for the projections)
                fan_in *= x
            t = torch.randn(shp, generator=g) * (gain / fan_in ** 0.5)
        elif "layer_norm" in k and k.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            t = 0.1 * torch.randn(shp, generator=g)
        sd[k] = t.to(dtype)
    return sd


def randn(shape, seed, dtype=torch.float32):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)).to(dtype)


def mmtrvat_inputs(cfg, B, T_l, T_a, T_v, seed=2024, dtype=torch.float32):
    """features N(0,1) fp32 (so channel 0 is never exactly 0), targets Bernoulli(0.3)."""
    g = torch.Generator().manual_seed(seed)
    txt = torch.randn(B, T_l, cfg.orig_d_l, generator=g).to(dtype)
    img = torch.randn(B, T_v, cfg.orig_d_v, generator=g).to(dtype)
    audio = torch.randn(B, T_a, cfg.orig_d_a, generator=g).to(dtype)
    tgt = (torch.rand(B, cfg.n_classes, generator=g) < 0.3).to(dtype)
    return txt, img, audio, tgt


def tiny_cfg(**kw):
    d = dict(orig_d_l=24, orig_d_v=7, orig_d_a=12, orig_d_p=16, hidden_sz=40, num_heads=4, layers=2,
             vonly=True, lonly=True, aonly=True, attn_mask=True, hybrid=False, n_classes=6,
             attn_dropout=0.0, attn_dropout_v=0.0, attn_dropout_a=0.0, relu_dropout=0.0, res_dropout=0.0,
             out_dropout=0.0, embed_dropout=0.0, bert_model="none")
    d.update(kw)
    return Namespace(**d)


def summarize(t, n=64):
    """Compact fingerprint of a (gradient) tensor: L2 norm, abs-sum, and a strided sample of n values."""
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return dict(norm=f.norm().item(), asum=f.abs().sum().item(), idx=idx, val=f[idx].float().clone(),
                shape=tuple(t.shape))
