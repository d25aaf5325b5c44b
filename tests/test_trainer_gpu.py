"""GPU tests of the public module API and the graph-captured Trainer (flat buffers, fused Adam, CUDA-graph replay)."""
from argparse import Namespace

import pytest
import torch

from oracle import functional as Fn
from oracle import synth

pytestmark = pytest.mark.gpu


def _model(cfg, precision, seed=5):
    from bpmult_b200 import MultiprojectionMMTransformer3DGMUClf
    m = MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg)), precision=precision)
    m.load_state_dict(synth.make_state_dict(synth.mmtrvat_shapes(cfg), seed), strict=False)
    return m.cuda().train()


def test_trainer_graph_matches_autograd_plus_torch_adam():
    cfg = synth.tiny_cfg()
    txt, img, audio, tgt = [t.cuda() for t in synth.mmtrvat_inputs(cfg, 2, 10, 30, 25)]
    from bpmult_b200 import Trainer
    a, b = _model(cfg, "fp32"), _model(cfg, "fp32")
    opt = torch.optim.Adam([p for p in a.parameters()], lr=1e-3)
    tr = Trainer(b, lr=1e-3)
    la, lb = [], []
    for _ in range(5):                                         # 2 eager warm-up steps, capture, 2 replays
        opt.zero_grad()
        loss = torch.nn.BCEWithLogitsLoss()(a(txt, None, None, img, audio), tgt)      # reference criterion (train.py:104)
        loss.backward()
        opt.step()
        la.append(float(loss))
        lb.append(float(tr.step_device(txt, img, audio, tgt)[0]))
    assert tr.graph is not None and tr.use_graph
    assert max(abs(x - y) for x, y in zip(la, lb)) < 2e-5, (la, lb)
    assert la[-1] < la[0]
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    # Adam normalises gradients, so tensors whose true gradient is zero (e.g. the K bias: softmax is shift-invariant) move by
    # +-lr on rounding noise alone; compare in relative L2 over all parameters instead of per element
    num = sum(float((pb[n].detach() - pa[n].detach()).double().pow(2).sum()) for n in pa if pa[n].grad is not None)
    den = sum(float(pa[n].detach().double().pow(2).sum()) for n in pa if pa[n].grad is not None)
    assert (num / den) ** 0.5 < 1e-3, (num / den) ** 0.5


def test_trainer_e2e_host_api_and_dropout_determinism():
    cfg = synth.tiny_cfg(embed_dropout=0.25, attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.1)
    host = synth.mmtrvat_inputs(cfg, 2, 10, 30, 25)
    from bpmult_b200 import Trainer
    runs = []
    for graph in (True, False):
        tr = Trainer(_model(cfg, "bf16"), lr=1e-3, seed=99, use_graph=graph)
        runs.append([tr.step(*host) for _ in range(5)])
    assert all(abs(x - y) < 2e-3 for x, y in zip(*runs)), runs     # same device-side seed stream with and without the graph
    assert len(set(round(x, 6) for x in runs[0])) > 1


def test_step_async_reads_each_loss_one_step_late_and_matches_step():
    cfg = synth.tiny_cfg()
    host = synth.mmtrvat_inputs(cfg, 2, 10, 30, 25)
    pinned = [t.pin_memory() for t in host]
    from bpmult_b200 import Trainer
    a, b = Trainer(_model(cfg, "fp32"), lr=1e-3, seed=5), Trainer(_model(cfg, "fp32"), lr=1e-3, seed=5)
    la = [a.step(*host) for _ in range(5)]                    # pageable inputs, blocking
    lb, pend = [], []
    for _ in range(5):                                         # pinned inputs, two steps in flight
        pend.append(b.step_async(*pinned))
        if len(pend) > 1:
            lb.append(pend.pop(0).item())
    lb.append(pend.pop(0).item())
    assert max(abs(x - y) for x, y in zip(la, lb)) < 1e-6, (la, lb)
    assert len(set(lb)) == 5


def test_gradient_accumulation_and_lr_change_under_graph_replay():
    """train.py:390-398 with gradient_accumulation_steps=2: the accumulate-only and the apply micro-step are two captured graphs;
    the learning rate is read from device memory, so set_lr() acts on a replayed graph"""
    cfg = synth.tiny_cfg()
    txt, img, audio, tgt = [t.cuda() for t in synth.mmtrvat_inputs(cfg, 4, 10, 30, 25)]
    from bpmult_b200 import Trainer
    a = _model(cfg, "fp32")
    opt = torch.optim.Adam([p for p in a.parameters()], lr=1e-3)
    tr = Trainer(_model(cfg, "fp32"), lr=1e-3, grad_accum=2)
    la, lb = [], []
    for it in range(5):                                        # 4 eager micro-steps, then both graphs captured and replayed
        if it == 3:
            tr.set_lr(5e-4)
            opt.param_groups[0]["lr"] = 5e-4
        opt.zero_grad()
        loss = torch.nn.BCEWithLogitsLoss()(a(txt, None, None, img, audio), tgt)
        loss.backward()
        opt.step()
        la.append(float(loss))
        l0 = float(tr.step_device(txt[:2], img[:2], audio[:2], tgt[:2])[0])
        l1 = float(tr.step_device(txt[2:], img[2:], audio[2:], tgt[2:])[0])
        lb.append(0.5 * (l0 + l1))
    assert set(tr.graphs) == {True, False}
    assert max(abs(x - y) for x, y in zip(la, lb)) < 2e-5, (la, lb)
    pa, pb = dict(a.named_parameters()), dict(tr.model.named_parameters())
    num = sum(float((pb[n].detach() - pa[n].detach()).double().pow(2).sum()) for n in pa if pa[n].grad is not None)
    den = sum(float(pa[n].detach().double().pow(2).sum()) for n in pa if pa[n].grad is not None)
    assert (num / den) ** 0.5 < 1e-3, (num / den) ** 0.5


def test_encoder_module_on_gpu_bf16_and_fp32():
    from bpmult_b200 import TransformerEncoder
    D, H, L, T, S, B = 300, 12, 2, 50, 70, 3
    sd = synth.make_state_dict(synth.encoder_shapes(D, L), 3)
    x, k = synth.randn((T, B, D), 1), synth.randn((S, B, D), 2)
    sdo = {n: v.clone().requires_grad_() for n, v in sd.items()}
    xo = x.clone().requires_grad_()
    ref = Fn.transformer_encoder(sdo, "", xo, k, k, H, L, True)
    ref.sum().backward()
    for prec, tol in (("fp32", 1e-4), ("bf16", 1e-2)):
        m = TransformerEncoder(D, H, L, attn_mask=True, precision=prec)
        m.load_state_dict(sd, strict=False)
        m.cuda().train()
        xg = x.cuda().requires_grad_()
        out = m(xg, k.cuda(), k.cuda())
        out.sum().backward()
        assert Fn.max_rel(out.detach().cpu(), ref.detach()) < tol
        assert Fn.rel_l2(xg.grad.cpu(), xo.grad) < (2e-4 if prec == "fp32" else 5e-2)


def test_mmtrvapt_module_on_gpu_matches_reference_golden():
    """the 4-modality model through the nn.Module API on the GPU (fp32 precision mode)"""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import check_fingerprints, load_gold
    from bpmult_b200 import MultiprojectionMMTransformerGMUClf
    rec = load_gold("mmtrvapt_tiny.pt")
    cfg = Namespace(**rec["cfg"])
    m = MultiprojectionMMTransformerGMUClf(cfg, precision="fp32")
    m.load_state_dict(synth.make_state_dict(synth.mmtrvapt_shapes(cfg), rec["seed"]), strict=False)
    m.cuda().train()
    B, T_l, T_a, T_v = rec["dims"]
    txt, img, audio, poster, tgt = [t.cuda() for t in synth.mmtrvapt_inputs(cfg, B, T_l, T_a, T_v)]
    txt.requires_grad_()
    logits, z = m(txt, None, None, img, audio, poster, output_gate=True)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=rec["pos_weight"].cuda())(logits, tgt)
    loss.backward()
    assert Fn.max_rel(logits.detach().cpu(), rec["logits"]) < 1e-4 and Fn.max_rel(z.cpu(), rec["z"]) < 1e-4
    assert Fn.rel_l2(txt.grad.cpu(), rec["dtxt"]) < 2e-4
    check_fingerprints({n: p.grad for n, p in m.named_parameters()}, rec["pgrad_fp"], 5e-4)


def test_trainer_runs_the_4_modality_model_under_graph_replay():
    """mmtrvapt through the same Trainer (flat buffers, buckets in backward order, fused Adam, CUDA graph): 5-tensor batch"""
    from bpmult_b200 import MultiprojectionMMTransformerGMUClf, Trainer
    cfg = synth.tiny_cfg(layers=1, n_classes=13, orig_d_p=48)

    def mk():
        m = MultiprojectionMMTransformerGMUClf(Namespace(**vars(cfg)), precision="fp32")
        m.load_state_dict(synth.make_state_dict(synth.mmtrvapt_shapes(cfg), 5), strict=False)
        return m.cuda().train()
    txt, img, audio, poster, tgt = [t.cuda() for t in synth.mmtrvapt_inputs(cfg, 2, 20, 30, 25)]
    a, b = mk(), mk()
    opt = torch.optim.Adam([p for p in a.parameters()], lr=1e-3)
    tr = Trainer(b, lr=1e-3)
    la, lb = [], []
    for _ in range(5):
        opt.zero_grad()
        loss = torch.nn.BCEWithLogitsLoss()(a(txt, None, None, img, audio, poster), tgt)
        loss.backward()
        opt.step()
        la.append(float(loss))
        lb.append(float(tr.step_device(txt, img, audio, poster, tgt)[0]))
    assert tr.graph is not None and tr.use_graph
    assert max(abs(x - y) for x, y in zip(la, lb)) < 2e-5, (la, lb)
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    num = sum(float((pb[n].detach() - pa[n].detach()).double().pow(2).sum()) for n in pa if pa[n].grad is not None)
    den = sum(float(pa[n].detach().double().pow(2).sum()) for n in pa if pa[n].grad is not None)
    assert (num / den) ** 0.5 < 1e-3, (num / den) ** 0.5


def test_no_cpu_fallback():
    from bpmult_b200 import TransformerEncoder
    m = TransformerEncoder(40, 4, 1)
    with pytest.raises(RuntimeError):
        m(torch.randn(5, 2, 40))


def test_reference_format_checkpoint_round_trip_on_gpu(tmp_path):
    """SURVEY 8 f3 on the CUDA path (utils/utils.py:21-30, train.py:372-379,419-430): a `checkpoint.pt` written the way the reference
    writes it (nn.DataParallel `module.` prefix, torch Adam state) resumes in model + Trainer under CUDA-graph replay exactly as torch's
    Adam would continue; the Trainer's own checkpoint loads back into torch Adam and, when the reference tree is on the box
    (baseline/_ref), into the UNMODIFIED reference model with identical logits."""
    from bpmult_b200 import Trainer
    cfg = synth.tiny_cfg()
    txt, img, audio, tgt = [t.cuda() for t in synth.mmtrvat_inputs(cfg, 2, 10, 30, 25)]
    crit = torch.nn.BCEWithLogitsLoss()

    def torch_steps(model, opt, n):
        out = []
        for _ in range(n):
            opt.zero_grad()
            loss = crit(model(txt, None, None, img, audio), tgt)
            loss.backward()
            opt.step()
            out.append(float(loss.detach()))
        return out
    a = _model(cfg, "fp32")
    opt_a = torch.optim.Adam(a.parameters(), lr=1e-3)
    torch_steps(a, opt_a, 2)
    path = str(tmp_path / "checkpoint.pt")
    torch.save({"epoch": 3, "state_dict": {"module." + k: v for k, v in a.state_dict().items()}, "optimizer": opt_a.state_dict(),
                "scheduler": {}, "n_no_improve": 1, "best_metric": 0.5}, path)
    la = torch_steps(a, opt_a, 4)
    b = _model(cfg, "fp32", seed=99)                               # other weights: everything must come from the checkpoint
    tr = Trainer(b, lr=0.5)
    info = tr.load_checkpoint(path)
    assert info["epoch"] == 3 and not info["missing"] and not info["unexpected"] and int(tr.step_t) == 2
    lb = [float(tr.step_device(txt, img, audio, tgt)[0]) for _ in range(4)]          # 2 eager steps, capture, replay
    assert tr.graph is not None
    assert max(abs(x - y) for x, y in zip(la, lb)) < 2e-5, (la, lb)
    ck2 = tr.checkpoint(epoch=4)
    c = _model(cfg, "fp32", seed=7)
    c.load_state_dict(ck2["state_dict"], strict=False)
    opt_c = torch.optim.Adam(c.parameters(), lr=1.0)
    opt_c.load_state_dict(ck2["optimizer"])
    lc = torch_steps(c, opt_c, 1)
    ld = float(tr.step_device(txt, img, audio, tgt)[0])
    assert abs(lc[0] - ld) < 2e-5
    tr.close()
    from oracle.ref_shim import load_reference, zero_dropout
    ref = load_reference()
    if ref is not None:
        rm = ref.mmtr.MultiprojectionMMTransformer3DGMUClf(zero_dropout(Namespace(**vars(cfg))))
        missing, unexpected = rm.load_state_dict({k: v.cpu() for k, v in ck2["state_dict"].items()}, strict=False)
        assert not unexpected
        rm.eval()
        pad = lambda t: torch.cat([t.cpu(), torch.zeros(t.shape[0], 512 - t.shape[1], t.shape[2])], 1)     # (the reference's own padding calls .cuda())
        with torch.no_grad():
            lr_ = rm(pad(txt), None, None, pad(img), pad(audio))
        b2 = _model(cfg, "fp32", seed=3)
        b2.load_state_dict(ck2["state_dict"], strict=False)
        b2.eval()
        with torch.no_grad():
            lo = b2(txt, None, None, img, audio).cpu()
        assert float((lr_ - lo).abs().max()) < 2e-5


def test_evaluate_between_graph_replays_matches_eval_mode_forward_and_the_oracle():
    """train.py:165-186 / 283-338: the eval pass runs on the Trainer's engine between captured training steps (dropout sites off),
    returns what the module forward returns in eval mode, and leaves the training trajectory untouched"""
    from bpmult_b200 import Trainer, model_eval
    cfg = synth.tiny_cfg(attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.1, embed_dropout=0.25, out_dropout=0.1)
    batch = synth.mmtrvat_inputs(cfg, 2, 10, 30, 25)
    pw = torch.tensor([1.0, 2.0, 0.5, 3.0, 1.0, 1.5])
    a, b = Trainer(_model(cfg, "fp32"), lr=1e-3, pos_weight=pw), Trainer(_model(cfg, "fp32"), lr=1e-3, pos_weight=pw)
    for i in range(5):
        la = a.step(*batch)
        if i == 3:
            metrics, arrays = model_eval([batch, batch], b, "moviescope", output_gates=True)
        lb = b.step(*batch)
        # same seeds, same dropout masks: evaluate() changed nothing (two runs differ by the order of atomic / split-K additions only)
        assert abs(la - lb) < 2e-6 * max(1.0, abs(la)), (i, la, lb)
    rel = ((a.flat_p - b.flat_p).double().norm() / a.flat_p.double().norm()).item()
    assert rel < 2e-4 and b.graph is not None, rel      # (Adam turns rounding noise on near-zero gradients into +-lr moves)
    m = b.model.eval()
    with torch.no_grad():
        logits, z = m(batch[0].cuda(), None, None, batch[1].cuda(), batch[2].cuda(), True)
    r = b.evaluate(*batch, output_gates=True)
    assert torch.equal(r["logits"], logits.cpu()) and torch.equal(r["gates"], z.cpu())
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    cfg0 = synth.tiny_cfg()
    lo, zo = Fn.mmtrvat_forward(sd, cfg0, *batch[:3])
    assert Fn.max_rel(r["logits"], lo) < 1e-4 and Fn.max_rel(r["gates"], zo) < 1e-4
    assert abs(r["loss"] - float(Fn.bce_with_logits(lo, batch[3], pw))) < 1e-5
    assert arrays["preds"].shape == (4, 6) and arrays["gates"].shape == (4, 3 * cfg.hidden_sz) and "auc_pr_samples" in metrics
    a.close()
    b.close()


def test_hybrid_model_trains_under_graph_replay_and_matches_autograd():
    """hybrid = True through the drop-in module's autograd and through the graph-captured Trainer: same losses"""
    from bpmult_b200 import MultiprojectionMMTransformer3DGMUClf, Trainer
    cfg = synth.tiny_cfg(layers=1, hybrid=True)
    sd = synth.make_state_dict(synth.mmtrvat_shapes(cfg), 5)

    def mk():
        m = MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg)), precision="fp32")
        m.load_state_dict(sd, strict=False)
        return m.cuda().train()
    txt, img, audio, tgt = [t.cuda() for t in synth.mmtrvat_inputs(cfg, 2, 10, 30, 25)]
    a, b = mk(), mk()
    opt = torch.optim.Adam([p for p in a.parameters()], lr=1e-3)
    tr = Trainer(b, lr=1e-3)
    la, lb = [], []
    for _ in range(5):
        opt.zero_grad()
        out, z = a(txt, None, None, img, audio, True)
        loss = torch.nn.BCEWithLogitsLoss()(out, tgt)
        loss.backward()
        opt.step()
        la.append(float(loss))
        lb.append(float(tr.step_device(txt, img, audio, tgt)[0]))
    assert z.shape == (2, 4 * cfg.hidden_sz)
    assert tr.graph is not None and [x[0] for x in tr.buckets][:3] == ["a_early", "v_early", "l_early"]
    assert max(abs(x - y) for x, y in zip(la, lb)) < 2e-5, (la, lb)
    assert la[-1] < la[0]
    tr.close()


def test_pruned_trainer_under_graph_replay_follows_the_full_one(monkeypatch):
    """BPM_PRUNE=1 through the captured Trainer (dropout off so that both draw nothing): the same losses and parameters as the full model"""
    from bpmult_b200 import Trainer
    cfg = synth.tiny_cfg()
    batch = [t.cuda() for t in synth.mmtrvat_inputs(cfg, 2, 10, 30, 25)]
    full = Trainer(_model(cfg, "fp32"), lr=1e-3)
    monkeypatch.setenv("BPM_PRUNE", "1")
    pr = Trainer(_model(cfg, "fp32"), lr=1e-3)
    assert pr.eng.prune and not full.eng.prune
    lf = [float(full.step_device(*batch)[0]) for _ in range(5)]
    lp = [float(pr.step_device(*batch)[0]) for _ in range(5)]
    assert pr.graph is not None and max(abs(a - b) for a, b in zip(lf, lp)) < 2e-5, (lf, lp)
    rel = ((full.flat_p - pr.flat_p).double().norm() / full.flat_p.double().norm()).item()
    assert rel < 2e-4, rel
    assert pr.launches_per_step < full.launches_per_step + 200
    full.close()
    pr.close()
