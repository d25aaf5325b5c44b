"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, shimmed by oracle/ref_shim.py) on seeded synthetic inputs and weights, and asserts that the
restatement in oracle/functional.py reproduces it.  Run here (the reference cannot travel to the GPU box):

    python oracle/make_golden.py

Fixtures hold only seeds, configs, outputs and gradients (weights are re-generated from the seed by
oracle/synth.py), so they stay small."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import functional as Fn                      # noqa: E402
from oracle import synth                                 # noqa: E402
from oracle.ref_shim import load_reference               # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
TOL = 5e-6

ENC_CASES = [
    # name, T, S, B, D, H, L, biproj, mask, self_only, zero_tail
    ("cross_causal", 12, 12, 2, 40, 4, 2, False, True, False, 0),
    ("cross_6x4", 6, 4, 2, 40, 4, 1, False, True, False, 0),
    ("cross_4x6", 4, 6, 3, 40, 4, 1, False, True, False, 0),
    ("cross_nomask", 9, 17, 2, 40, 4, 2, False, False, False, 0),
    ("biproj", 8, 13, 2, 40, 4, 2, True, True, False, 0),
    ("self_only", 10, 10, 2, 40, 4, 2, False, True, True, 0),
    ("zero_tail", 12, 12, 2, 40, 4, 2, False, True, False, 3),
    ("d300", 24, 40, 2, 300, 12, 1, False, True, False, 5),
    ("d96", 16, 16, 2, 96, 4, 1, False, True, False, 0),
    ("odd_d", 7, 9, 2, 45, 5, 1, False, True, False, 0),
]


def enc_inputs(T, S, B, D, zero_tail, seed):
    x = synth.randn((T, B, D), seed)
    k = synth.randn((S, B, D), seed + 1)
    g = synth.randn((T, B, D), seed + 2)
    if zero_tail:
        x[T - zero_tail:] = 0
        k[S - zero_tail:] = 0
    return x, k, g


def run_encoder_case(ref, case, seed):
    name, T, S, B, D, H, L, bi, mask, self_only, zt = case
    shapes = synth.encoder_shapes(D, L, bi)
    sd = synth.make_state_dict(shapes, seed)
    x, k, g = enc_inputs(T, S, B, D, zt, seed + 100)
    # ---- reference
    m = ref.tr.TransformerEncoder(D, H, L, attn_mask=mask, biprojection=bi)
    ref_keys = {kk for kk in m.state_dict().keys() if kk not in ("version", "embed_positions._float_tensor")}
    assert ref_keys == set(shapes.keys()), (ref_keys ^ set(shapes.keys()))
    m.load_state_dict(sd, strict=False)
    m.train()
    xr, kr = x.clone().requires_grad_(), k.clone().requires_grad_()
    out = m(xr) if self_only else m(xr, kr, kr)
    (out * g).sum().backward()
    rgr = {n: p.grad.clone() for n, p in m.named_parameters()}
    # ---- restatement
    sdo = {kk: v.clone().requires_grad_() for kk, v in sd.items()}
    xo, ko = x.clone().requires_grad_(), k.clone().requires_grad_()
    o2 = Fn.transformer_encoder(sdo, "", xo, None if self_only else ko, None if self_only else ko, H, L, mask, bi)
    (o2 * g).sum().backward()
    err = Fn.max_rel(o2, out)
    assert err < TOL, (name, "out", err)
    assert Fn.max_rel(xo.grad, xr.grad) < TOL, (name, "dx")
    if not self_only:
        assert Fn.max_rel(ko.grad, kr.grad) < TOL, (name, "dk")
    for n in rgr:
        e = Fn.max_rel(sdo[n].grad, rgr[n])
        assert e < 2e-5, (name, n, e)
    rec = dict(case=case, seed=seed, out=out.detach(), dx=xr.grad, dk=None if self_only else kr.grad)
    if D <= 48:
        rec["pgrads"] = rgr
    else:
        rec["pgrad_summ"] = {n: synth.summarize(v) for n, v in rgr.items()}
    print("encoder %-14s ok  out err %.2e" % (name, err))
    return rec


def run_mha(ref, seed):
    D, H, T, S, B = 40, 4, 5, 7, 2
    m = ref.mha.MultiheadAttention(D, H)
    shapes = {"in_proj_weight": (3 * D, D), "in_proj_bias": (3 * D,), "out_proj.weight": (D, D), "out_proj.bias": (D,)}
    sd = synth.make_state_dict(shapes, seed)
    m.load_state_dict(sd)
    q, k, v = synth.randn((T, B, D), seed + 1), synth.randn((S, B, D), seed + 2), synth.randn((S, B, D), seed + 3)
    mask = Fn.future_mask(T, S, torch.float32)
    a, w = m(q, k, v, attn_mask=ref.tr.buffered_future_mask(q, k))
    a2, w2 = Fn.multihead_attention(sd, "", q, k, v, H, mask, need_weights=True)
    assert Fn.max_rel(a2, a) < TOL and Fn.max_rel(w2, w) < TOL
    print("mha ok")
    return dict(seed=seed, dims=(D, H, T, S, B), out=a.detach(), weights=w.detach())


def run_pe(ref):
    recs = []
    for (T, B, D) in [(12, 3, 40), (9, 2, 45), (600, 1, 300)]:
        x = synth.randn((T, B, D), 7)
        x[T - 2:, :, 0] = 0            # channel-0 zero => padding position
        x[1, 0, 0] = 0
        m = ref.pe.SinusoidalPositionalEmbedding(D)
        r = m(x.transpose(0, 1)[:, :, 0]).transpose(0, 1)
        o = Fn.positional_embedding(x)
        assert torch.equal(r, o), (T, B, D)
        recs.append(dict(dims=(T, B, D), seed=7, out=r[:, :, ::7].clone()))
    print("pe ok")
    return recs


def run_gmus(ref, seed):
    D, M = 40, 11
    rec = {}
    x = [synth.randn((M, D), seed + i) for i in range(4)]
    for cls, fn, nm in [(ref.mmtr.GatedMultimodalLayerFeatures, Fn.gmu_features, "features"),
                        (ref.mmtr.GatedMultimodalLayer, Fn.gmu_plain, "plain")]:
        m = cls(D, D, D)
        sd = synth.make_state_dict({"hidden1.weight": (D, D), "hidden2.weight": (D, D), "x_gate.weight": (D, 2 * D)}, seed)
        m.load_state_dict(sd)
        xs = [t.clone().requires_grad_() for t in x[:2]]
        o, z = m(xs)
        g = synth.randn(o.shape, seed + 9)
        (o * g).sum().backward()
        o2, z2 = fn(sd, "", x[0], x[1])
        assert Fn.max_rel(o2, o) < TOL and Fn.max_rel(z2, z) < TOL
        rec[nm] = dict(out=o.detach(), z=z.detach(), dx=[t.grad for t in xs],
                       pgrads={n: p.grad.clone() for n, p in m.named_parameters()})
    for n_in, cls in [(3, ref.mmtr.TextShifting3Layer), (4, ref.mmtr.TextShifting4Layer)]:
        m = cls(D, D, D, D) if n_in == 3 else cls(D, D, D, D, D)
        shp = {}
        for i in range(n_in):
            shp["hidden%d.weight" % (i + 1)] = (D, D)
        for i in range(n_in):
            shp["x%d_gate.weight" % (i + 1)] = (D, n_in * D)
        sd = synth.make_state_dict(shp, seed + n_in)
        assert set(m.state_dict().keys()) == set(shp.keys()), set(m.state_dict().keys()) ^ set(shp.keys())
        m.load_state_dict(sd)
        xs = [t.clone().requires_grad_() for t in x[:n_in]]
        o, z = m(xs)
        g = synth.randn(o.shape, seed + 9)
        (o * g).sum().backward()
        o2, z2 = Fn.text_shifting(sd, "", x[:n_in])
        assert Fn.max_rel(o2, o) < TOL and Fn.max_rel(z2, z) < TOL
        rec["ts%d" % n_in] = dict(out=o.detach(), z=z.detach(), dx=[t.grad for t in xs],
                                  pgrads={n: p.grad.clone() for n, p in m.named_parameters()})
    # TextShiftingNLayer (mmtr.py:249-273), N = 5 (one more than any fixed-arity class), varargs forward
    n_in = 5
    x5 = x + [synth.randn((M, D), seed + 4)]
    m = ref.mmtr.TextShiftingNLayer([D] * n_in, D)
    shp = {"hiddens.%d.weight" % i: (D, D) for i in range(n_in)}
    shp.update({"x_gates.%d.weight" % i: (D, n_in * D) for i in range(n_in)})
    sd = synth.make_state_dict(shp, seed + 20)
    assert set(m.state_dict().keys()) == set(shp.keys())
    m.load_state_dict(sd)
    xs = [t.clone().requires_grad_() for t in x5]
    o, z = m(*xs)
    g = synth.randn(o.shape, seed + 9)
    (o * g).sum().backward()
    o2, z2 = Fn.text_shifting_n(sd, "", x5)
    assert Fn.max_rel(o2, o) < TOL and Fn.max_rel(z2, z) < TOL
    rec["tsN"] = dict(out=o.detach(), z=z.detach(), dx=[t.grad for t in xs], pgrads={n: p.grad.clone() for n, p in m.named_parameters()})
    rec["seed"], rec["dims"] = seed, (D, M)
    print("gmus ok")
    return rec


def run_bce(seed):
    B, C = 5, 6
    x = synth.randn((B, C), seed).requires_grad_()
    y = (synth.randn((B, C), seed + 1) > 0.5).float()
    w = torch.rand(C, generator=torch.Generator().manual_seed(seed)) * 3 + 0.5
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=w)(x, y)            # train.py:104
    loss.backward()
    assert abs(Fn.bce_with_logits(x.detach(), y, w).item() - loss.item()) < 1e-7
    print("bce ok")
    return dict(seed=seed, x=x.detach(), y=y, w=w, loss=loss.detach(), dx=x.grad)


def run_mmtrvapt(ref, cfg, B, T_l, T_a, T_v, seed):
    """4-modality model (mmtr.py:278-583): lengths 512 / 200 / 200, biprojection wave-2 encoders, time-axis transfm_* linears, poster,
    TextShifting4Layer head.  Parameter gradients are stored as fingerprints (norm + strided sample) to keep the fixture small."""
    shapes = synth.mmtrvapt_shapes(cfg)
    sd = synth.make_state_dict(shapes, seed)
    m = ref.mmtr.MultiprojectionMMTransformerGMUClf(cfg)
    ref_keys = {k for k in m.state_dict().keys() if not (k.endswith(".version") or k.endswith("_float_tensor"))}
    assert ref_keys == set(shapes.keys()), sorted(ref_keys ^ set(shapes.keys()))[:10]
    m.load_state_dict(sd, strict=False)
    m.train()
    txt, img, audio, poster, tgt = synth.mmtrvapt_inputs(cfg, B, T_l, T_a, T_v)
    pw = torch.linspace(0.5, 2.0, cfg.n_classes)
    txt_r = txt.clone().requires_grad_()
    logits, z = m(txt_r, None, None, img, audio, poster, output_gate=True)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=pw)(logits, tgt)
    loss.backward()
    rgr = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    nograd = sorted(n for n, p in m.named_parameters() if p.grad is None)
    sdo = {k: v.clone().requires_grad_() for k, v in sd.items()}
    l2, z2 = Fn.mmtrvapt_forward(sdo, cfg, txt, img, audio, poster)
    loss2 = Fn.bce_with_logits(l2, tgt, pw)
    loss2.backward()
    e = Fn.max_rel(l2, logits)
    assert e < 2e-5, ("logits", e)
    assert Fn.max_rel(z2, z) < 2e-5
    worst = 0.0
    for n in rgr:
        ee = Fn.rel_l2(sdo[n].grad, rgr[n])
        worst = max(worst, ee)
        assert ee < 1e-4, (n, ee)
    print("mmtrvapt D=%d L=%d ok  logits err %.2e  worst grad rel-l2 %.2e  (no-grad params: %d)" % (cfg.hidden_sz, cfg.layers, e, worst, len(nograd)))
    return dict(cfg=vars(cfg), dims=(B, T_l, T_a, T_v), seed=seed, pos_weight=pw, logits=logits.detach(), z=z.detach(), loss=loss.detach(),
                dtxt=txt_r.grad, nograd=nograd, pgrad_fp={n: synth.summarize(g) for n, g in rgr.items()})


def run_mmtrvat(ref, cfg, B, T_l, T_a, T_v, seed, full_pgrads):
    shapes = synth.mmtrvat_shapes(cfg)
    sd = synth.make_state_dict(shapes, seed)
    m = ref.mmtr.MultiprojectionMMTransformer3DGMUClf(cfg)
    ref_keys = {k for k in m.state_dict().keys() if not (k.endswith(".version") or k.endswith("_float_tensor"))}
    assert ref_keys == set(shapes.keys()), sorted(ref_keys ^ set(shapes.keys()))[:10]
    m.load_state_dict(sd, strict=False)
    m.train()
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, B, T_l, T_a, T_v)
    pw = torch.linspace(0.5, 2.0, cfg.n_classes)
    txt_r = txt.clone().requires_grad_()
    logits, z = m(txt_r, None, None, img, audio, output_gate=True)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=pw)(logits, tgt)
    loss.backward()
    rgr = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    nograd = sorted(n for n, p in m.named_parameters() if p.grad is None)
    # restatement
    sdo = {k: v.clone().requires_grad_() for k, v in sd.items()}
    l2, z2 = Fn.mmtrvat_forward(sdo, cfg, txt, img, audio)
    loss2 = Fn.bce_with_logits(l2, tgt, pw)
    loss2.backward()
    e = Fn.max_rel(l2, logits)
    assert e < 2e-5, ("logits", e)
    assert Fn.max_rel(z2, z) < 2e-5
    worst = 0.0
    for n in rgr:
        ee = Fn.rel_l2(sdo[n].grad, rgr[n])
        worst = max(worst, ee)
        assert ee < 1e-4, (n, ee)
    rec = dict(cfg=vars(cfg), dims=(B, T_l, T_a, T_v), seed=seed, pos_weight=pw, logits=logits.detach(),
               z=z.detach(), loss=loss.detach(), dtxt=txt_r.grad, nograd=nograd)
    if full_pgrads:
        rec["pgrads"] = rgr
    else:
        rec["pgrad_summ"] = {n: synth.summarize(v) for n, v in rgr.items()}
    print("mmtrvat D=%d L=%d ok  logits err %.2e  worst grad rel-l2 %.2e  (no-grad params: %d)"
          % (cfg.hidden_sz, cfg.layers, e, worst, len(nograd)))
    return rec


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = load_reference()
    assert ref is not None, "reference tree not found"
    os.makedirs(OUT, exist_ok=True)
    torch.save([run_encoder_case(ref, c, 1000 + 17 * i) for i, c in enumerate(ENC_CASES)],
               os.path.join(OUT, "encoder.pt"))
    torch.save(dict(mha=run_mha(ref, 31), pe=run_pe(ref), gmu=run_gmus(ref, 41), bce=run_bce(51)),
               os.path.join(OUT, "modules.pt"))
    torch.save(run_mmtrvat(ref, synth.tiny_cfg(), 2, 10, 30, 25, 77, True), os.path.join(OUT, "mmtrvat_tiny.pt"))
    cfg = synth.tiny_cfg(orig_d_l=96, orig_d_v=35, orig_d_a=74, hidden_sz=96, num_heads=4, layers=1)
    torch.save(run_mmtrvat(ref, cfg, 2, 50, 60, 40, 78, False), os.path.join(OUT, "mmtrvat_d96.pt"))
    cfg = synth.tiny_cfg(layers=1, n_classes=13, orig_d_p=48)
    torch.save(run_mmtrvapt(ref, cfg, 2, 20, 30, 25, 79), os.path.join(OUT, "mmtrvapt_tiny.pt"))
    # hybrid = True (mmtrvat): the reference's own branch with the two gate call sites accepted in either convention (ref_shim shim 6)
    torch.save(run_mmtrvat(ref, synth.tiny_cfg(layers=1, hybrid=True), 2, 10, 30, 25, 81, True), os.path.join(OUT, "mmtrvat_tiny_hybrid.pt"))
    # state_dict contract: key names + shapes of the reference model, and its init under the default seed
    import json
    from oracle.ref_shim import mmtrvat_args
    torch.manual_seed(1234)
    m = ref.mmtr.MultiprojectionMMTransformer3DGMUClf(synth.tiny_cfg())
    sdm = m.state_dict()
    keys = {k: list(v.shape) for k, v in sdm.items()}
    fp = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sdm.items() if not k.endswith("_float_tensor")}
    json.dump(dict(keys=keys, init_fingerprint_seed1234=fp), open(os.path.join(OUT, "mmtrvat_state_dict.json"), "w"), indent=0)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
