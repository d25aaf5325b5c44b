"""GPU tests of the smaller drop-in modules through the nn.Module API (CUDA kernels behind the C ABI, no emulation)."""
import pytest
import torch

from helpers import load_gold
from oracle import functional as Fn
from oracle import synth

pytestmark = pytest.mark.gpu


def test_text_shifting_n_layer_on_gpu_matches_reference_golden():
    """mmtr.py:249-273 TextShiftingNLayer (the `hybrid=True` head): N = 5 inputs, golden from the reference module"""
    import bpmult_b200.modules as M
    g = load_gold("modules.pt")["gmu"]
    D, rows = g["dims"]
    n_in = 5
    m = M.TextShiftingNLayer([D] * n_in, D)
    shp = {"hiddens.%d.weight" % i: (D, D) for i in range(n_in)}
    shp.update({"x_gates.%d.weight" % i: (D, n_in * D) for i in range(n_in)})
    m.load_state_dict(synth.make_state_dict(shp, g["seed"] + 20))
    m.cuda()
    xs = [synth.randn((rows, D), g["seed"] + i).cuda().requires_grad_() for i in range(n_in)]
    o, z = m(*xs)
    (o * synth.randn(o.shape, g["seed"] + 9).cuda()).sum().backward()
    torch.cuda.synchronize()
    r = g["tsN"]
    assert Fn.max_rel(o.detach().cpu(), r["out"]) < 2e-5 and Fn.max_rel(z.cpu(), r["z"]) < 2e-5
    for a, b in zip(xs, r["dx"]):
        assert Fn.max_rel(a.grad.cpu(), b) < 5e-5
    for n, p in m.named_parameters():
        assert Fn.rel_l2(p.grad.cpu(), r["pgrads"][n]) < 5e-5, n
    with pytest.raises(AssertionError):
        m(*xs[:3])


@pytest.mark.parametrize("n_in", [3, 4])
def test_text_shifting_3_4_on_gpu_match_reference_golden(n_in):
    import bpmult_b200.modules as M
    g = load_gold("modules.pt")["gmu"]
    D, rows = g["dims"]
    cls = M.TextShifting3Layer if n_in == 3 else M.TextShifting4Layer
    m = cls(D, D, D, D) if n_in == 3 else cls(D, D, D, D, D)
    shp = {}
    for i in range(n_in):
        shp["hidden%d.weight" % (i + 1)] = (D, D)
    for i in range(n_in):
        shp["x%d_gate.weight" % (i + 1)] = (D, n_in * D)
    m.load_state_dict(synth.make_state_dict(shp, g["seed"] + n_in))
    m.cuda()
    xs = [synth.randn((rows, D), g["seed"] + i).cuda().requires_grad_() for i in range(n_in)]
    o, z = m(xs)
    (o * synth.randn(o.shape, g["seed"] + 9).cuda()).sum().backward()
    r = g["ts%d" % n_in]
    assert Fn.max_rel(o.detach().cpu(), r["out"]) < 2e-5 and Fn.max_rel(z.cpu(), r["z"]) < 2e-5
    for a, b in zip(xs, r["dx"]):
        assert Fn.max_rel(a.grad.cpu(), b) < 5e-5
    for n, p in m.named_parameters():
        assert Fn.rel_l2(p.grad.cpu(), r["pgrads"][n]) < 5e-5, n


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_audio_encoder_on_gpu_matches_reference_golden(precision):
    """mmtr.py:93-108: Conv1d(96, 96, 128, stride 2) x 2 + AdaptiveAvgPool1d(200) as im2col rows + tensor-core GEMMs"""
    import bpmult_b200.modules as M
    from helpers import check_fingerprints
    g = load_gold("audio_encoder.pt")["audio_encoder"]
    m = M.AudioEncoder()
    m.precision = precision
    m.load_state_dict(synth.make_state_dict(synth.audio_encoder_shapes(96), g["seed"]))
    m.cuda()
    x = synth.randn((2, 96, 900), g["seed"] + 1).cuda()
    y = m(x)
    (y * synth.randn((2, 96, 200), g["seed"] + 2).cuda()).sum().backward()
    torch.cuda.synchronize()
    fp32 = precision == "fp32"
    e = Fn.max_rel(y.detach().cpu(), g["out"])
    print("AudioEncoder %s: out max-rel %.3e" % (precision, e))
    assert e < (2e-5 if fp32 else 1e-2)
    check_fingerprints({n: p.grad.cpu() for n, p in m.named_parameters()}, g["pgrad_fp"], 1e-4 if fp32 else 2e-2)


def test_mmtrvapt_with_audio_encoder_on_gpu_matches_reference_golden():
    """the 4-modality model with its AudioEncoder upstream of the trunk (mmtr.py:307,452), raw spectrogram input"""
    from argparse import Namespace
    import bpmult_b200.modules as M
    from helpers import check_fingerprints
    rec = load_gold("audio_encoder.pt")["mmtrvapt_audio"]
    cfg = Namespace(**rec["cfg"])
    m = M.MultiprojectionMMTransformerGMUClf(cfg, precision="fp32")
    shapes = synth.mmtrvapt_shapes(cfg)
    shapes.update(synth.audio_encoder_shapes(96, "audio_enc."))
    m.load_state_dict(synth.make_state_dict(shapes, rec["seed"]), strict=False)
    m.cuda().train()
    B, T_l, T_raw, T_v = rec["dims"]
    txt, img, _, poster, tgt = [t.cuda() for t in synth.mmtrvapt_inputs(cfg, B, T_l, 30, T_v)]
    audio = synth.randn((B, 96, T_raw), rec["seed"] + 5).cuda()
    txt.requires_grad_()
    logits, z = m(txt, None, None, img, audio, poster, output_gate=True)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=rec["pos_weight"].cuda())(logits, tgt)
    loss.backward()
    torch.cuda.synchronize()
    assert Fn.max_rel(logits.detach().cpu(), rec["logits"]) < 1e-4 and Fn.max_rel(z.cpu(), rec["z"]) < 1e-4
    assert Fn.rel_l2(txt.grad.cpu(), rec["dtxt"]) < 2e-4
    from helpers import fingerprint_errors
    errs = fingerprint_errors({n: p.grad.cpu() for n, p in m.named_parameters()}, rec["pgrad_fp"])
    bad = {n: e for n, e in errs.items() if e >= 5e-4}
    print("mmtrvapt + AudioEncoder fp32: %d gradient tensors, worst audio_enc %.2e, above 5e-4: %s" % (
        len(errs), max(e for n, e in errs.items() if n.startswith("audio_enc.")), sorted(bad.items(), key=lambda kv: -kv[1])[:4]))
    # A ReLU tie (see tests/test_fullshape_gpu.py) in this 40-wide, 1-layer toy moves the gradient of ONE hidden unit, i.e. the FFN-side
    # gradients of that encoder by ~1/sqrt(rows * units) and, more weakly, everything upstream of it (its K/V source stream, here the audio
    # path).  Accepted only with that signature: the worst tensor is an fc1 tensor, the rest of the excess is small and a minority.
    if bad:
        worst = max(bad, key=bad.get)
        assert ".fc1." in worst and all(e < 3e-2 for e in bad.values()) and len(bad) <= len(errs) // 4, bad
        clean = [e for n, e in errs.items() if n not in bad]
        assert max(clean) < 5e-4


def test_modules_trace_under_torch_compile():
    """every module forward is ONE registered torch op with a fake kernel and a registered autograd formula, so the drop-in
    model traces with fullgraph=True (no graph break at the C-ABI boundary) and gives the eager result"""
    from argparse import Namespace
    import bpmult_b200.modules as M
    cfg = synth.tiny_cfg(layers=1)
    m = M.MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg)), precision="fp32")
    m.load_state_dict(synth.make_state_dict(synth.mmtrvat_shapes(cfg), 5), strict=False)
    m.cuda().train()
    txt, img, audio, tgt = [t.cuda() for t in synth.mmtrvat_inputs(cfg, 2, 10, 30, 25)]
    ref = m(txt, None, None, img, audio)
    fn = torch.compile(lambda a, b, c: m(a, None, None, b, c), fullgraph=True, backend="aot_eager")
    out = fn(txt, img, audio)
    assert Fn.max_rel(out.detach().cpu(), ref.detach().cpu()) < 1e-6
    torch.nn.functional.binary_cross_entropy_with_logits(out, tgt).backward()
    assert m.proj1.weight.grad is not None and float(m.proj1.weight.grad.abs().sum()) > 0
    assert "bpmult_b200::mmtrvat" in str(torch.ops.bpmult_b200.mmtrvat.default._schema)
    torch.library.opcheck(torch.ops.bpmult_b200.seq_gmu.default,
                          (M._handle(m.gmu_l), torch.randn(6, cfg.hidden_sz, device="cuda"), torch.randn(6, cfg.hidden_sz, device="cuda"),
                           m.gmu_l.hidden1.weight, m.gmu_l.hidden2.weight, m.gmu_l.x_gate.weight), test_utils=("test_schema", "test_faketensor"))
