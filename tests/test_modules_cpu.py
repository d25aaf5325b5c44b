"""Host-logic tests of the drop-in nn.Module layer (no GPU): constructor / state_dict / init-order compatibility with
the reference, and autograd wiring of each module, with the device ops replaced by the test-only emulation."""
import json
import os
from argparse import Namespace

import pytest
import torch

import bpmult_b200.modules as M
from emu_ops import EmuOps
from helpers import GOLD, load_gold
from oracle import functional as Fn
from oracle import synth


@pytest.fixture(autouse=True)
def emu(monkeypatch):
    o = EmuOps()
    monkeypatch.setattr(M, "_ops_for", lambda device: o)
    return o


def test_state_dict_keys_shapes_and_init_match_reference():
    ref = json.load(open(os.path.join(GOLD, "mmtrvat_state_dict.json")))
    torch.manual_seed(1234)
    m = M.MultiprojectionMMTransformer3DGMUClf(synth.tiny_cfg())
    sd = m.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == ref["keys"]
    for k, (s, a) in ref["init_fingerprint_seed1234"].items():        # same RNG consumption order => identical initial weights
        assert abs(float(sd[k].double().sum()) - s) <= 1e-9 + 1e-12 * abs(s), k
        assert abs(float(sd[k].double().abs().sum()) - a) <= 1e-9 + 1e-12 * abs(a), k


def test_live_reference_init_equality_when_available():
    from oracle.ref_shim import load_reference
    ref = load_reference()
    if ref is None:
        pytest.skip("reference tree not present")
    cfg = synth.tiny_cfg()
    torch.manual_seed(7)
    a = ref.mmtr.MultiprojectionMMTransformer3DGMUClf(cfg).state_dict()
    torch.manual_seed(7)
    b = M.MultiprojectionMMTransformer3DGMUClf(cfg).state_dict()
    assert a.keys() == b.keys()
    for k in a:
        if not k.endswith("_float_tensor"):
            assert torch.equal(a[k], b[k]), k


ENC = load_gold("encoder.pt")


@pytest.mark.parametrize("rec", [r for r in ENC if r["case"][4] <= 48], ids=[r["case"][0] for r in ENC if r["case"][4] <= 48])
def test_transformer_encoder_module_autograd(rec):
    name, T, S, B, D, H, L, bi, mask, self_only, zt = rec["case"]
    sd = synth.make_state_dict(synth.encoder_shapes(D, L, bi), rec["seed"])
    m = M.TransformerEncoder(D, H, L, attn_mask=mask, biprojection=bi, precision="fp32")
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert set(missing) <= {"version", "embed_positions._float_tensor"} and not unexpected
    m.train()
    x = synth.randn((T, B, D), rec["seed"] + 100).requires_grad_()
    k = synth.randn((S, B, D), rec["seed"] + 101).requires_grad_()
    g = synth.randn((T, B, D), rec["seed"] + 102)
    with torch.no_grad():
        if zt:
            x[T - zt:] = 0
            k[S - zt:] = 0
    out = m(x) if self_only else m(x, k, k)
    (out * g).sum().backward()
    assert Fn.max_rel(out, rec["out"]) < 2e-5
    assert Fn.max_rel(x.grad, rec["dx"]) < 5e-5
    if not self_only:
        assert Fn.max_rel(k.grad, rec["dk"]) < 5e-5
    for n, p in m.named_parameters():
        assert Fn.rel_l2(p.grad, rec["pgrads"][n]) < 5e-5, n


def test_layer_module_equals_one_layer_of_oracle():
    D, H, T, S, B = 40, 4, 7, 9, 2
    sd = synth.make_state_dict({k[len("layers.0."):]: v for k, v in synth.encoder_shapes(D, 1).items() if k.startswith("layers.0.")}, 5)
    layer = M.TransformerEncoderLayer(D, H, attn_dropout=0, relu_dropout=0, res_dropout=0, attn_mask=True)
    layer.precision = "fp32"
    layer.load_state_dict(sd)
    x, k = synth.randn((T, B, D), 1).requires_grad_(), synth.randn((S, B, D), 2).requires_grad_()
    out = layer(x, k, k)
    xo, ko = x.detach().clone().requires_grad_(), k.detach().clone().requires_grad_()
    sdo = {"l." + n: v.clone().requires_grad_() for n, v in sd.items()}
    ref = Fn.encoder_layer(sdo, "l.", xo, ko, ko, H, True)
    g = synth.randn((T, B, D), 3)
    (out * g).sum().backward()
    (ref * g).sum().backward()
    assert Fn.max_rel(out, ref) < 2e-5 and Fn.max_rel(x.grad, xo.grad) < 5e-5 and Fn.max_rel(k.grad, ko.grad) < 5e-5
    for n, p in layer.named_parameters():
        assert Fn.rel_l2(p.grad, sdo["l." + n].grad) < 5e-5, n


def test_multihead_attention_module_matches_golden():
    g = load_gold("modules.pt")["mha"]
    D, H, T, S, B = g["dims"]
    m = M.MultiheadAttention(D, H)
    m.precision = "fp32"
    m.load_state_dict(synth.make_state_dict({"in_proj_weight": (3 * D, D), "in_proj_bias": (3 * D,), "out_proj.weight": (D, D), "out_proj.bias": (D,)}, g["seed"]))
    q, k, v = [synth.randn((n, B, D), g["seed"] + i).requires_grad_() for i, n in ((1, T), (2, S), (3, S))]
    a, w = m(q, k, v, attn_mask=M.buffered_future_mask(q, k))
    assert Fn.max_rel(a, g["out"]) < 2e-5 and Fn.max_rel(w, g["weights"]) < 2e-5
    # gradients vs the oracle
    sd = {n: p.detach().clone().requires_grad_() for n, p in m.named_parameters()}
    qo, ko, vo = [t.detach().clone().requires_grad_() for t in (q, k, v)]
    ao, _ = Fn.multihead_attention(sd, "", qo, ko, vo, H, Fn.future_mask(T, S, torch.float32))
    gg = synth.randn(a.shape, 9)
    (a * gg).sum().backward()
    (ao * gg).sum().backward()
    for x, y in ((q, qo), (k, ko), (v, vo)):
        assert Fn.max_rel(x.grad, y.grad) < 5e-5
    for n, p in m.named_parameters():
        assert Fn.rel_l2(p.grad, sd[n].grad) < 5e-5, n
    with pytest.raises(NotImplementedError):
        m(q, k, v, attn_mask=torch.zeros(T, S))


def test_positional_embedding_module_matches_golden():
    for rec in load_gold("modules.pt")["pe"]:
        T, B, D = rec["dims"]
        x = synth.randn((T, B, D), rec["seed"])
        x[T - 2:, :, 0] = 0
        x[1, 0, 0] = 0
        out = M.SinusoidalPositionalEmbedding(D)(x.transpose(0, 1)[:, :, 0]).transpose(0, 1)
        assert torch.equal(out[:, :, ::7], rec["out"])


def test_gmu_modules_match_golden():
    g = load_gold("modules.pt")["gmu"]
    D, rows = g["dims"]
    x = [synth.randn((rows, D), g["seed"] + i) for i in range(4)]
    for cls, nm in ((M.GatedMultimodalLayerFeatures, "features"), (M.GatedMultimodalLayer, "plain")):
        m = cls(D, D, D)
        m.precision = "fp32"
        m.load_state_dict(synth.make_state_dict({"hidden1.weight": (D, D), "hidden2.weight": (D, D), "x_gate.weight": (D, 2 * D)}, g["seed"]))
        xs = [t.clone().requires_grad_() for t in x[:2]]
        o, z = m(xs)
        (o * synth.randn(o.shape, g["seed"] + 9)).sum().backward()
        r = g[nm]
        assert Fn.max_rel(o, r["out"]) < 2e-5 and Fn.max_rel(z, r["z"]) < 2e-5
        for a, b in zip(xs, r["dx"]):
            assert Fn.max_rel(a.grad, b) < 5e-5
        for n, p in m.named_parameters():
            assert Fn.rel_l2(p.grad, r["pgrads"][n]) < 5e-5, (nm, n)
    for n_in, cls in ((3, M.TextShifting3Layer), (4, M.TextShifting4Layer)):
        m = cls(D, D, D, D) if n_in == 3 else cls(D, D, D, D, D)
        shp = {}
        for i in range(n_in):
            shp["hidden%d.weight" % (i + 1)] = (D, D)
        for i in range(n_in):
            shp["x%d_gate.weight" % (i + 1)] = (D, n_in * D)
        m.load_state_dict(synth.make_state_dict(shp, g["seed"] + n_in))
        xs = [t.clone().requires_grad_() for t in x[:n_in]]
        o, z = m(xs)
        (o * synth.randn(o.shape, g["seed"] + 9)).sum().backward()
        r = g["ts%d" % n_in]
        assert Fn.max_rel(o, r["out"]) < 2e-5 and Fn.max_rel(z, r["z"]) < 2e-5
        for a, b in zip(xs, r["dx"]):
            assert Fn.max_rel(a.grad, b) < 5e-5
        for n, p in m.named_parameters():
            assert Fn.rel_l2(p.grad, r["pgrads"][n]) < 5e-5, n


def test_text_shifting_n_layer_matches_reference_golden():
    """mmtr.py:249-273 (the `hybrid=True` head): ModuleList parameter names, varargs forward, N = 5 inputs"""
    g = load_gold("modules.pt")["gmu"]
    D, rows = g["dims"]
    n_in = 5
    x = [synth.randn((rows, D), g["seed"] + i) for i in range(n_in)]
    m = M.TextShiftingNLayer([D] * n_in, D)
    shp = {"hiddens.%d.weight" % i: (D, D) for i in range(n_in)}
    shp.update({"x_gates.%d.weight" % i: (D, n_in * D) for i in range(n_in)})
    assert set(m.state_dict().keys()) == set(shp.keys())
    m.load_state_dict(synth.make_state_dict(shp, g["seed"] + 20))
    xs = [t.clone().requires_grad_() for t in x]
    o, z = m(*xs)
    (o * synth.randn(o.shape, g["seed"] + 9)).sum().backward()
    r = g["tsN"]
    assert Fn.max_rel(o, r["out"]) < 2e-5 and Fn.max_rel(z, r["z"]) < 2e-5
    for a, b in zip(xs, r["dx"]):
        assert Fn.max_rel(a.grad, b) < 5e-5
    for n, p in m.named_parameters():
        assert Fn.rel_l2(p.grad, r["pgrads"][n]) < 5e-5, n
    import pytest
    with pytest.raises(AssertionError):
        m(*xs[:3])


def test_mmtrvapt_module_autograd_matches_reference_golden():
    """the 4-modality model behind the reference's module API (mmtr.py:278-583): state_dict names, forward signature with poster"""
    from helpers import check_fingerprints
    rec = load_gold("mmtrvapt_tiny.pt")
    cfg = Namespace(**rec["cfg"])
    m = M.MultiprojectionMMTransformerGMUClf(cfg, precision="fp32")
    shapes = synth.mmtrvapt_shapes(cfg)
    keys = {k for k in m.state_dict().keys() if not (k.endswith(".version") or k.endswith("_float_tensor"))}
    assert keys == set(shapes.keys()), sorted(keys ^ set(shapes.keys()))[:8]
    m.load_state_dict(synth.make_state_dict(shapes, rec["seed"]), strict=False)
    m.train()
    B, T_l, T_a, T_v = rec["dims"]
    txt, img, audio, poster, tgt = synth.mmtrvapt_inputs(cfg, B, T_l, T_a, T_v)
    txt.requires_grad_()
    logits, z = m(txt, None, None, img, audio, poster, output_gate=True)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=rec["pos_weight"])(logits, tgt)
    loss.backward()
    assert Fn.max_rel(logits, rec["logits"]) < 2e-5 and Fn.max_rel(z, rec["z"]) < 2e-5
    assert Fn.max_rel(txt.grad, rec["dtxt"]) < 1e-4
    check_fingerprints({n: p.grad for n, p in m.named_parameters()}, rec["pgrad_fp"], 2e-4)
    assert M.get_model(Namespace(model="mmtrvapt", **rec["cfg"])).__class__ is M.MultiprojectionMMTransformerGMUClf


def test_mmtrvat_module_autograd_matches_reference_golden():
    rec = load_gold("mmtrvat_tiny.pt")
    cfg = Namespace(**rec["cfg"])
    m = M.MultiprojectionMMTransformer3DGMUClf(cfg, precision="fp32")
    m.load_state_dict(synth.make_state_dict(synth.mmtrvat_shapes(cfg), rec["seed"]), strict=False)
    m.train()
    B, T_l, T_a, T_v = rec["dims"]
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, B, T_l, T_a, T_v)
    txt.requires_grad_()
    logits, z = m(txt, None, None, img, audio, output_gate=True)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=rec["pos_weight"])(logits, tgt)      # the reference's own criterion (train.py:104)
    loss.backward()
    assert Fn.max_rel(logits, rec["logits"]) < 2e-5 and Fn.max_rel(z, rec["z"]) < 2e-5
    assert Fn.max_rel(txt.grad, rec["dtxt"]) < 1e-4
    pm = dict(m.named_parameters())
    for n, ref in rec["pgrads"].items():
        assert Fn.rel_l2(pm[n].grad, ref) < 2e-4, n
    assert sorted(n for n, p in m.named_parameters() if p.grad is None) == sorted(rec["nograd"])
    with pytest.raises(Exception):
        m(torch.randn(1, 600, cfg.orig_d_l), None, None, img[:1], audio[:1])        # longer than the fixed 512 steps (mmtr.py:722-732)


def test_audio_encoder_module_matches_reference_golden():
    """mmtr.py:93-108: parameter names, forward and parameter gradients of the standalone module (host logic on the emulation)"""
    from helpers import check_fingerprints
    g = load_gold("audio_encoder.pt")["audio_encoder"]
    m = M.AudioEncoder()
    m.precision = "fp32"
    sd = synth.make_state_dict(synth.audio_encoder_shapes(96), g["seed"])
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd)
    x = synth.randn((2, 96, 900), g["seed"] + 1)
    y = m(x)
    (y * synth.randn((2, 96, 200), g["seed"] + 2)).sum().backward()
    assert Fn.max_rel(y, g["out"]) < 2e-5
    check_fingerprints({n: p.grad for n, p in m.named_parameters()}, g["pgrad_fp"], 1e-4)


def test_mmtrvapt_with_audio_encoder_matches_reference_golden():
    """the 4-modality model as the reference ships it: raw spectrogram through audio_enc (mmtr.py:307,452) into the trunk"""
    from helpers import check_fingerprints
    rec = load_gold("audio_encoder.pt")["mmtrvapt_audio"]
    cfg = Namespace(**rec["cfg"])
    m = M.MultiprojectionMMTransformerGMUClf(cfg, precision="fp32")
    assert m.with_audio_enc
    shapes = synth.mmtrvapt_shapes(cfg)
    shapes.update(synth.audio_encoder_shapes(96, "audio_enc."))
    keys = {k for k in m.state_dict().keys() if not (k.endswith(".version") or k.endswith("_float_tensor"))}
    assert keys == set(shapes.keys()), sorted(keys ^ set(shapes.keys()))[:8]
    m.load_state_dict(synth.make_state_dict(shapes, rec["seed"]), strict=False)
    m.train()
    B, T_l, T_raw, T_v = rec["dims"]
    txt, img, _, poster, tgt = synth.mmtrvapt_inputs(cfg, B, T_l, 30, T_v)
    audio = synth.randn((B, 96, T_raw), rec["seed"] + 5)
    txt.requires_grad_()
    logits, z = m(txt, None, None, img, audio, poster, output_gate=True)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=rec["pos_weight"])(logits, tgt)
    loss.backward()
    assert Fn.max_rel(logits, rec["logits"]) < 2e-5 and Fn.max_rel(z, rec["z"]) < 2e-5
    assert Fn.max_rel(txt.grad, rec["dtxt"]) < 1e-4
    check_fingerprints({n: p.grad for n, p in m.named_parameters()}, rec["pgrad_fp"], 2e-4)
