"""Host-logic tests (no GPU): the product engines (kernel sequencing, padded layout, hand-derived backward) driven by
the pure-torch op emulation in tests/emu_ops.py must reproduce the golden vectors generated from the reference."""
import pytest
import torch

from emu_ops import EmuOps
from helpers import load_gold, run_encoder_engine, run_model_engine
from oracle import functional as Fn
from oracle import synth

ENC = load_gold("encoder.pt")


@pytest.mark.parametrize("rec", ENC, ids=[r["case"][0] for r in ENC])
def test_encoder_engine_fp32_matches_reference_golden(rec):
    name, T, S, B, D, H, L, bi, mask, self_only, zt = rec["case"]
    sd = synth.make_state_dict(synth.encoder_shapes(D, L, bi), rec["seed"])
    x = synth.randn((T, B, D), rec["seed"] + 100)
    k = synth.randn((S, B, D), rec["seed"] + 101)
    g = synth.randn((T, B, D), rec["seed"] + 102)
    if zt:
        x[T - zt:] = 0
        k[S - zt:] = 0
    out, dx, dk, grads, _ = run_encoder_engine(EmuOps(), sd, x, None if self_only else k, g, H, L, mask, bi, self_only)
    assert Fn.max_rel(out, rec["out"]) < 2e-5
    assert Fn.max_rel(dx, rec["dx"]) < 5e-5
    if not self_only:
        assert Fn.max_rel(dk, rec["dk"]) < 5e-5
    if "pgrads" in rec:
        for n, ref in rec["pgrads"].items():
            assert Fn.rel_l2(grads[n], ref) < 5e-5, n
    else:
        for n, s in rec["pgrad_summ"].items():
            f = grads[n].reshape(-1).double()
            assert abs(f.norm().item() - s["norm"]) <= 5e-5 * s["norm"] + 1e-9, n
            assert torch.allclose(f[s["idx"]].float(), s["val"], rtol=2e-3, atol=2e-5 * max(1e-6, s["norm"])), n


def test_oracle_mmtrvapt_restatement_matches_reference_golden():
    """the 4-modality model (mmtr.py:278-583): oracle restatement against the fixture generated from the shimmed reference"""
    from argparse import Namespace
    from helpers import check_fingerprints
    rec = load_gold("mmtrvapt_tiny.pt")
    cfg = Namespace(**rec["cfg"])
    B, T_l, T_a, T_v = rec["dims"]
    sd = {k: v.requires_grad_() for k, v in synth.make_state_dict(synth.mmtrvapt_shapes(cfg), rec["seed"]).items()}
    txt, img, audio, poster, tgt = synth.mmtrvapt_inputs(cfg, B, T_l, T_a, T_v)
    txt.requires_grad_()
    logits, z = Fn.mmtrvapt_forward(sd, cfg, txt, img, audio, poster)
    loss = Fn.bce_with_logits(logits, tgt, rec["pos_weight"])
    loss.backward()
    assert Fn.max_rel(logits, rec["logits"]) < 2e-5 and Fn.max_rel(z, rec["z"]) < 2e-5
    assert abs(loss.item() - rec["loss"].item()) < 1e-6 and Fn.rel_l2(txt.grad, rec["dtxt"]) < 1e-4
    check_fingerprints({n: v.grad for n, v in sd.items()}, rec["pgrad_fp"], 1e-4)


def test_mmtrvapt_engine_fp32_matches_reference_golden():
    """host schedule of the 4-modality model (device ops emulated): T != S encoders, biprojection wave 2, time-axis linears, poster"""
    from helpers import check_fingerprints, run_model4_engine
    rec = load_gold("mmtrvapt_tiny.pt")
    logits, z, loss, dtxt, grads, eng = run_model4_engine(EmuOps(), rec)
    assert Fn.max_rel(logits, rec["logits"]) < 2e-5 and Fn.max_rel(z, rec["z"]) < 2e-5
    assert abs(loss.item() - rec["loss"].item()) < 1e-6
    assert Fn.rel_l2(dtxt, rec["dtxt"]) < 1e-4
    check_fingerprints(grads, rec["pgrad_fp"], 2e-4)


def test_mmtrvat_engine_fp32_matches_reference_golden():
    rec = load_gold("mmtrvat_tiny.pt")
    logits, z, loss, dtxt, grads, eng = run_model_engine(EmuOps(), rec)
    assert Fn.max_rel(logits, rec["logits"]) < 2e-5
    assert Fn.max_rel(z, rec["z"]) < 2e-5
    assert abs(loss.item() - rec["loss"].item()) < 1e-5
    assert Fn.max_rel(dtxt, rec["dtxt"]) < 1e-4
    assert sorted(eng.unused_params()) == sorted(rec["nograd"])
    worst = 0.0
    for n, ref in rec["pgrads"].items():
        e = Fn.rel_l2(grads[n], ref)
        worst = max(worst, e)
        assert e < 2e-4, (n, e)
    print("worst param-grad rel-l2", worst)


@pytest.mark.parametrize("bi", [False, True])
def test_fused_schedules_match_the_plain_ones_with_dropout_on(bi):
    """The K/V LayerNorm hoist (fold_kv) and the cast fused into the LayerNorm backward (fuse_cast) are schedule changes only:
    with every dropout site live they must reproduce the plain per-layer schedule (same masks, same gradients)."""
    from emu_ops import EmuOps
    from bpmult_b200.engine import EncoderEngine
    from oracle import synth
    torch.manual_seed(0)
    T, S, B, D, H, L = 12, 9, 2, 40, 4, 2
    shapes = EncoderEngine(EmuOps(), D, H, L, biprojection=bi).param_shapes()
    sd = synth.make_state_dict(shapes, 7)
    x, k, g = torch.randn(T, B, D), torch.randn(S, B, D), torch.randn(T, B, D)
    p = dict(attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.2, embed_dropout=0.25)
    res = {}
    try:
        for fold, fuse in ((False, False), (True, True)):
            EncoderEngine.fold_kv, EncoderEngine.fuse_cast = fold, fuse
            res[fold] = run_encoder_engine(EmuOps(), sd, x, k, g, H, L, True, bi, False, p=p, training=True, seed=11)
    finally:
        EncoderEngine.fold_kv, EncoderEngine.fuse_cast = True, True
    (o0, dx0, dk0, g0, _), (o1, dx1, dk1, g1, _) = res[False], res[True]
    assert Fn.max_rel(o1, o0) < 1e-5 and Fn.rel_l2(dx1, dx0) < 1e-5 and Fn.rel_l2(dk1, dk0) < 1e-5
    for n in g0:
        assert Fn.rel_l2(g1[n], g0[n]) < 1e-5, n
