"""hybrid = True of mmtrvat (mmtr.py:631, 662, 680-689, 765-775, 854-855): the reference's own branch, run with
its two gate call sites accepted in either calling convention (oracle/ref_shim.py shim 6), is the golden; the restatement, the
engine (on the ops emulation), the drop-in module and the Trainer are checked against it on CPU."""
import os
import sys
from argparse import Namespace

import pytest
import torch

from oracle import functional as Fn
from oracle import synth

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import load_gold, run_model_engine  # noqa: E402


def _modules():
    import bpmult_b200.modules as M
    from emu_ops import EmuOps
    o = EmuOps()
    M._ops_for = lambda device: o
    return M


def test_restatement_matches_the_reference_hybrid_golden():
    rec = load_gold("mmtrvat_tiny_hybrid.pt")
    cfg = Namespace(**rec["cfg"])
    assert cfg.hybrid
    B, T_l, T_a, T_v = rec["dims"]
    sd = {k: v.clone().requires_grad_() for k, v in synth.make_state_dict(synth.mmtrvat_shapes(cfg), rec["seed"]).items()}
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, B, T_l, T_a, T_v)
    logits, z = Fn.mmtrvat_forward(sd, cfg, txt, img, audio)
    Fn.bce_with_logits(logits, tgt, rec["pos_weight"]).backward()
    assert Fn.max_rel(logits, rec["logits"]) < 1e-6 and Fn.max_rel(z, rec["z"]) < 1e-6 and z.shape[1] == 4 * cfg.hidden_sz
    for n, ref in rec["pgrads"].items():
        assert Fn.rel_l2(sd[n].grad, ref) < 1e-5, n
    early = [n for n in rec["pgrads"] if "early" in n or n.startswith(("proj_l_e", "proj_v_e", "proj_a_e"))]
    assert len(early) == 3 * (3 * 12 + 2) + 6 + 3, len(early)                 # 3 encoders of max(1, 3) = 3 layers, gmu_early, proj_*_e


def test_engine_on_the_emulation_matches_the_hybrid_golden():
    from emu_ops import EmuOps
    rec = load_gold("mmtrvat_tiny_hybrid.pt")
    logits, z, loss, dtxt, grads, eng = run_model_engine(EmuOps(), rec)
    assert eng.hybrid and eng.head.n_in == 4
    assert Fn.max_rel(logits, rec["logits"]) < 2e-5 and Fn.max_rel(z, rec["z"]) < 2e-5
    assert abs(loss.item() - rec["loss"].item()) < 1e-5
    assert Fn.max_rel(dtxt, rec["dtxt"]) < 1e-4
    assert sorted(eng.unused_params()) == sorted(rec["nograd"])
    for n, ref in rec["pgrads"].items():
        assert Fn.rel_l2(grads[n], ref) < 2e-4, (n, Fn.rel_l2(grads[n], ref))
    assert eng.backward_order()[:3] == ["a_early", "v_early", "l_early"] and len(eng.backward_order()) == 15


def test_module_keys_order_autograd_and_trainer():
    M = _modules()
    rec = load_gold("mmtrvat_tiny_hybrid.pt")
    cfg = Namespace(**rec["cfg"])
    m = M.MultiprojectionMMTransformer3DGMUClf(cfg, precision="fp32")
    sd = synth.make_state_dict(synth.mmtrvat_shapes(cfg), rec["seed"])
    keys = [k for k in m.state_dict().keys() if not k.endswith(("_float_tensor", "version"))]
    assert set(keys) == set(sd.keys())
    tops = []
    for k in keys:
        if k.split(".")[0] not in tops:
            tops.append(k.split(".")[0])
    # registration order of the reference constructor (mmtr.py:588-689): it fixes the RNG stream of the initialisation
    assert tops == ["gmu_l_m", "gmu_v_m", "gmu_a_m", "gmu_l", "gmu_v", "gmu_a", "gmu_early", "proj_l", "proj_v", "proj_a"] + \
        ["trans_" + n for n in M.ENC_NAMES] + ["proj1", "proj2", "out_layer", "gmu", "transfm_a2l", "transfm_v2l", "transfm_l2a", "transfm_l2v",
                                              "trans_l_early", "trans_v_early", "trans_a_early", "proj_l_e", "proj_v_e", "proj_a_e"]
    m.load_state_dict(sd, strict=False)
    m.train()
    B, T_l, T_a, T_v = rec["dims"]
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, B, T_l, T_a, T_v)
    txt.requires_grad_()
    logits, z = m(txt, None, None, img, audio, output_gate=True)
    torch.nn.BCEWithLogitsLoss(pos_weight=rec["pos_weight"])(logits, tgt).backward()
    assert Fn.max_rel(logits, rec["logits"]) < 2e-5 and Fn.max_rel(z, rec["z"]) < 2e-5
    assert Fn.max_rel(txt.grad, rec["dtxt"]) < 1e-4
    pm = dict(m.named_parameters())
    for n, ref in rec["pgrads"].items():
        assert Fn.rel_l2(pm[n].grad, ref) < 2e-4, n
    assert sorted(n for n, p in m.named_parameters() if p.grad is None) == sorted(rec["nograd"])
    # one optimizer step through the Trainer: flat buffers and buckets include the early encoders, the loss matches the golden
    from bpmult_b200.trainer import Trainer
    m2 = M.MultiprojectionMMTransformer3DGMUClf(cfg, precision="fp32")
    m2.load_state_dict(sd, strict=False)
    tr = Trainer(m2.train(), lr=1e-3, use_graph=False, pos_weight=rec["pos_weight"])
    assert [b[0] for b in tr.buckets][:3] == ["a_early", "v_early", "l_early"] and tr.buckets[-1][0] == "misc"
    l0 = tr.step(txt.detach(), img, audio, tgt)
    assert abs(l0 - rec["loss"].item()) < 1e-5
    l1 = tr.step(txt.detach(), img, audio, tgt)
    assert l1 < l0


def test_reference_init_is_reproduced_with_hybrid_on():
    """same constructor order => same initial weights as the reference under the same seed (when the reference tree is present)"""
    from oracle.ref_shim import load_reference
    ref = load_reference()
    if ref is None:
        pytest.skip("no reference tree")
    M = _modules()
    cfg = synth.tiny_cfg(layers=1, hybrid=True)
    torch.manual_seed(1234)
    a = ref.mmtr.MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg)))
    torch.manual_seed(1234)
    b = M.MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg)), precision="fp32")
    sa, sb = a.state_dict(), b.state_dict()
    for k, v in sa.items():
        if k.endswith(("_float_tensor", "version")):
            continue
        assert torch.equal(v, sb[k]), k
