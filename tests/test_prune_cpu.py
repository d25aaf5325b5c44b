"""Pruned mode (MMTrVatEngine(prune=True)): the query side of the six wave-2 stacks and the gated units on time steps 0 and n_vec-1 only.
It must reproduce the reference's logits, gates, loss and EVERY gradient (goldens of the plain and of the hybrid model) -- host logic on
the ops emulation."""
import os
import sys
from argparse import Namespace

import pytest
import torch

from oracle import functional as Fn
from oracle import synth

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import load_gold  # noqa: E402


def _run(rec, prune, sd=None, cfg_over=None, training=True):
    from emu_ops import EmuOps
    from bpmult_b200.model_engine import MMTrVatEngine
    cfg = Namespace(**rec["cfg"])
    for k, v in (cfg_over or {}).items():
        setattr(cfg, k, v)
    B, T_l, T_a, T_v = rec["dims"]
    sd = sd or synth.make_state_dict(synth.mmtrvat_shapes(cfg), rec["seed"])
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, B, T_l, T_a, T_v)
    eng = MMTrVatEngine(EmuOps(), cfg, dtype=torch.float32, prune=prune)
    eng.pack({k: v for k, v in sd.items()})
    logits, z = eng.forward(txt, img, audio, training=training)
    loss, dlogits = eng.loss(logits, tgt, rec["pos_weight"])
    eng.zero_grads()
    dtxt, dimg = torch.zeros_like(txt), torch.zeros_like(img)
    eng.backward(dlogits, {"l": dtxt, "v": dimg})
    grads = {n: torch.zeros(s) for n, s in eng.param_shapes().items()}
    eng.unpack_grads(grads)
    D, Dp, C = cfg.hidden_sz, eng.d.Dp, cfg.n_classes
    ng = 4 if getattr(cfg, "hybrid", False) else 3
    return logits[:, :C].clone(), z.view(B, ng, Dp)[:, :, :D].reshape(B, ng * D).clone(), float(loss), dtxt, dimg, grads, eng


@pytest.mark.parametrize("gold", ["mmtrvat_tiny.pt", "mmtrvat_tiny_hybrid.pt"])
def test_pruned_engine_reproduces_the_reference_golden(gold):
    rec = load_gold(gold)
    logits, z, loss, dtxt, dimg, grads, eng = _run(rec, True)
    assert eng.prune and eng.enc["l_with_a2v"].prune_pos == (0, 511, 511, 511) and eng.enc["l_with_a2v"].prune_row0
    assert eng.enc["l_with_a2v"].T == 4 and eng.enc["l_with_a"].T == 512
    assert Fn.max_rel(logits, rec["logits"]) < 2e-5 and Fn.max_rel(z, rec["z"]) < 2e-5
    assert abs(loss - rec["loss"].item()) < 1e-5
    assert Fn.max_rel(dtxt, rec["dtxt"]) < 1e-4
    for n, ref in rec["pgrads"].items():
        assert Fn.rel_l2(grads[n], ref) < 2e-4, (n, Fn.rel_l2(grads[n], ref))


def test_pruned_equals_full_without_the_future_mask_and_for_other_inputs():
    """attn_mask = False (no single-key row), input gradients of a second stream, longer real sequences"""
    rec = load_gold("mmtrvat_tiny.pt")
    rec = dict(rec, dims=(3, 40, 70, 512))
    a = _run(rec, False, cfg_over=dict(attn_mask=False))
    b = _run(rec, True, cfg_over=dict(attn_mask=False))
    assert not b[6].enc["v_with_l2a"].prune_row0
    assert Fn.max_rel(b[0], a[0]) < 1e-5 and Fn.max_rel(b[1], a[1]) < 1e-5 and abs(a[2] - b[2]) < 1e-6
    assert Fn.max_rel(b[3], a[3]) < 1e-4 and Fn.max_rel(b[4], a[4]) < 1e-4
    for n in a[5]:
        assert Fn.rel_l2(b[5][n], a[5][n]) < 2e-4, n
    # with the mask on, and the last time step of the video stream real (T_v = 512): the row T-1 attends to everything
    a = _run(rec, False)
    b = _run(rec, True)
    assert Fn.max_rel(b[0], a[0]) < 1e-5 and Fn.max_rel(b[4], a[4]) < 1e-4
    for n in a[5]:
        assert Fn.rel_l2(b[5][n], a[5][n]) < 2e-4, n


def test_pruned_mode_trains_with_every_dropout_site_live():
    import bpmult_b200.modules as M
    from emu_ops import EmuOps
    from bpmult_b200.trainer import Trainer
    o = EmuOps()
    M._ops_for = lambda device: o
    os.environ["BPM_PRUNE"] = "1"
    try:
        cfg = synth.tiny_cfg(layers=1, attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.1, embed_dropout=0.25, out_dropout=0.1)
        m = M.MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg)), precision="fp32")
        m.load_state_dict(synth.make_state_dict(synth.mmtrvat_shapes(cfg), 5), strict=False)
        tr = Trainer(m.train(), lr=1e-2, use_graph=False)
        assert tr.eng.prune
        batch = synth.mmtrvat_inputs(cfg, 4, 8, 12, 10)
        losses = [tr.step(*batch) for _ in range(6)]
        assert all(l == l for l in losses) and min(losses[3:]) < losses[0]
    finally:
        del os.environ["BPM_PRUNE"]
