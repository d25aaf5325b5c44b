"""Host logic of the Trainer on CPU (device ops = the test-only emulation): gradient accumulation as train.py:390-398, the
device-resident learning rate behind a ReduceLROnPlateau scheduler (train.py:128-136, 408) and optimizer checkpoint / resume
(train.py:374-379, 417-425)."""
import os
import sys
from argparse import Namespace

import torch

from oracle import synth

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _build(cfg):
    import bpmult_b200.modules as M
    from emu_ops import EmuOps
    o = EmuOps()
    M._ops_for = lambda device: o
    m = M.MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg)), precision="fp32")
    m.load_state_dict(synth.make_state_dict(synth.mmtrvat_shapes(cfg), 5), strict=False)
    return m.train()


def test_two_micro_batches_equal_one_full_batch_step():
    from bpmult_b200.trainer import Trainer
    cfg = synth.tiny_cfg(layers=1)
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, 4, 8, 12, 10)
    full = Trainer(_build(cfg), lr=1e-2, use_graph=False)
    acc = Trainer(_build(cfg), lr=1e-2, use_graph=False, grad_accum=2)
    for _ in range(2):
        lf = full.step(txt, img, audio, tgt)
        p_before = acc.flat_p.clone()
        l0 = acc.step(txt[:2], img[:2], audio[:2], tgt[:2])
        assert torch.equal(acc.flat_p, p_before)              # no optimizer step on the first micro-batch
        l1 = acc.step(txt[2:], img[2:], audio[2:], tgt[2:])
        assert abs(0.5 * (l0 + l1) - lf) < 1e-6
    assert int(acc.step_t) == 2 and int(full.step_t) == 2
    assert float(acc.flat_g.abs().max()) == 0.0               # cleared after the optimizer step
    rel = ((full.flat_p - acc.flat_p).double().norm() / full.flat_p.double().norm()).item()
    assert rel < 2e-4, rel


def test_plateau_scheduler_drives_the_device_learning_rate():
    from bpmult_b200.trainer import Trainer
    cfg = synth.tiny_cfg(layers=1)
    batch = synth.mmtrvat_inputs(cfg, 2, 8, 12, 10)
    tr = Trainer(_build(cfg), lr=1e-2, use_graph=False)
    tr.step(*batch)
    p0 = tr.flat_p.clone()
    tr.set_lr(0.0)                                            # a zero rate must freeze the parameters: proves the kernel reads lr_t
    tr.step(*batch)
    assert torch.equal(tr.flat_p, p0)
    # the reference's scheduler object works through a one-group shim
    shim = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-2)
    sched = torch.optim.lr_scheduler.ReduceLROnPlateau(shim, "min", patience=0, factor=0.5)
    for metric in (1.0, 2.0):
        sched.step(metric)
    tr.set_lr(shim.param_groups[0]["lr"])
    assert tr.get_lr() == 5e-3 and abs(float(tr.lr_t) - 5e-3) < 1e-9


def test_optimizer_state_round_trip_resumes_bit_exactly():
    from bpmult_b200.trainer import Trainer
    cfg = synth.tiny_cfg(layers=1, attn_dropout=0.1, res_dropout=0.1)
    batch = synth.mmtrvat_inputs(cfg, 2, 8, 12, 10)
    a = Trainer(_build(cfg), lr=1e-2, use_graph=False, seed=3)
    for _ in range(2):
        a.step(*batch)
    model_sd = {k: v.clone() for k, v in a.model.state_dict().items()}
    opt_sd = a.optimizer_state_dict()
    la = [a.step(*batch) for _ in range(2)]
    m = _build(cfg)
    b = Trainer(m, lr=1.0, use_graph=False, seed=77)
    m.load_state_dict(model_sd)                               # loads in place, i.e. into the flat parameter buffer
    b.load_optimizer_state_dict(opt_sd)
    lb = [b.step(*batch) for _ in range(2)]
    assert la == lb
    assert torch.equal(a.flat_p, b.flat_p)


def test_trainer_drives_the_4_modality_model():
    """mmtrvapt through the Trainer's host logic (5-tensor batch, flat buffers incl. the time-axis linears and the poster projection)"""
    import bpmult_b200.modules as M
    from emu_ops import EmuOps
    from bpmult_b200.trainer import Trainer
    o = EmuOps()
    M._ops_for = lambda device: o
    cfg = synth.tiny_cfg(layers=1, n_classes=13, orig_d_p=48)

    def mk():
        m = M.MultiprojectionMMTransformerGMUClf(Namespace(**vars(cfg)), precision="fp32")
        m.load_state_dict(synth.make_state_dict(synth.mmtrvapt_shapes(cfg), 5), strict=False)
        return m.train()
    txt, img, audio, poster, tgt = synth.mmtrvapt_inputs(cfg, 2, 20, 30, 25)
    a, b = mk(), mk()
    opt = torch.optim.Adam([p for p in a.parameters()], lr=1e-3)
    tr = Trainer(b, lr=1e-3, use_graph=False)
    for _ in range(2):
        opt.zero_grad()
        loss = torch.nn.BCEWithLogitsLoss()(a(txt, None, None, img, audio, poster), tgt)
        loss.backward()
        opt.step()
        lb = tr.step(txt, img, audio, poster, tgt)
        assert abs(float(loss) - lb) < 1e-5
    names = [bk[0] for bk in tr.buckets]
    assert names[-1] == "misc" and len(names) == 13
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    num = sum(float((pb[n].detach() - pa[n].detach()).double().pow(2).sum()) for n in pa)
    den = sum(float(pa[n].detach().double().pow(2).sum()) for n in pa)
    assert (num / den) ** 0.5 < 1e-3, (num / den) ** 0.5


def test_reference_format_checkpoint_round_trip_with_torch_adam():
    """SURVEY 8 f3 (utils/utils.py:21-30, train.py:372-379,419-430): a checkpoint written the way the reference writes it -- model
    state_dict with the nn.DataParallel `module.` prefix, `optim.Adam(model.parameters()).state_dict()` -- loads into model + Trainer
    and training continues exactly as torch's Adam would; the Trainer's own checkpoint loads back into torch's Adam."""
    from bpmult_b200.trainer import Trainer
    cfg = synth.tiny_cfg(layers=1)
    batch = synth.mmtrvat_inputs(cfg, 2, 8, 12, 10)
    txt, img, audio, tgt = batch
    crit = torch.nn.BCEWithLogitsLoss()

    def torch_steps(model, opt, n):
        out = []
        for _ in range(n):
            opt.zero_grad()
            loss = crit(model(txt, None, None, img, audio), tgt)
            loss.backward()
            opt.step()
            out.append(float(loss))
        return out
    # "reference side": the drop-in module driven by torch's own Adam, checkpointed like train.py:419-430 under DataParallel
    a = _build(cfg)
    opt_a = torch.optim.Adam(a.parameters(), lr=1e-2)
    torch_steps(a, opt_a, 2)
    ck = {"epoch": 3, "state_dict": {"module." + k: v.clone() for k, v in a.state_dict().items()}, "optimizer": opt_a.state_dict(),
          "scheduler": {}, "n_no_improve": 1, "best_metric": 0.5}
    ck = {k: (v if k != "optimizer" else torch.load(_roundtrip(v), weights_only=False)) for k, v in ck.items()}
    la = torch_steps(a, opt_a, 2)
    # Trainer side: fresh model with other weights, resume from the checkpoint
    b = _build(cfg)
    with torch.no_grad():
        for p in b.parameters():
            p.add_(0.25)
    tr = Trainer(b, lr=123.0, use_graph=False)
    info = tr.load_checkpoint(ck)
    assert info["epoch"] == 3 and info["n_no_improve"] == 1 and not info["unexpected"] and not info["missing"]
    assert abs(tr.get_lr() - 1e-2) < 1e-12 and int(tr.step_t) == 2
    lb = [tr.step(*batch) for _ in range(2)]
    assert max(abs(x - y) for x, y in zip(la, lb)) < 1e-5, (la, lb)
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    num = sum(float((pb[n].detach() - pa[n].detach()).double().pow(2).sum()) for n in pa)
    den = sum(float(pa[n].detach().double().pow(2).sum()) for n in pa)
    assert (num / den) ** 0.5 < 1e-3
    # and back: the Trainer's checkpoint resumes under torch's Adam on a reference-layout model
    ck2 = tr.checkpoint(epoch=4)
    c = _build(cfg)
    c.load_state_dict(ck2["state_dict"], strict=False)
    opt_c = torch.optim.Adam(c.parameters(), lr=1.0)
    opt_c.load_state_dict(ck2["optimizer"])
    assert opt_c.param_groups[0]["lr"] == 1e-2
    lc = torch_steps(c, opt_c, 1)
    ld = [tr.step(*batch)]
    assert abs(lc[0] - ld[0]) < 1e-5
    # ReduceLROnPlateau-style code can drive the rate through param_groups
    tr.param_groups[0]["lr"] = 5e-3
    assert tr.get_lr() == 5e-3


def _roundtrip(obj):
    import io
    buf = io.BytesIO()
    torch.save(obj, buf)
    buf.seek(0)
    return buf


def test_reference_checkpoint_loads_into_the_live_reference_when_available():
    """a Trainer checkpoint's state_dict loads into the UNMODIFIED reference model (shimmed import) and gives the same logits"""
    import pytest
    from oracle.ref_shim import load_reference, zero_dropout
    ref = load_reference()
    if ref is None:
        pytest.skip("reference tree not present")
    from bpmult_b200.trainer import Trainer
    cfg = synth.tiny_cfg(layers=1)
    batch = synth.mmtrvat_inputs(cfg, 2, 8, 12, 10)
    tr = Trainer(_build(cfg), lr=1e-2, use_graph=False)
    tr.step(*batch)
    ck = tr.checkpoint(epoch=1)
    rm = ref.mmtr.MultiprojectionMMTransformer3DGMUClf(zero_dropout(Namespace(**vars(cfg))))
    missing, unexpected = rm.load_state_dict(ck["state_dict"], strict=False)
    assert not unexpected and all(k.endswith("_float_tensor") or k.endswith("version") for k in missing), (missing, unexpected)
    rm.eval()
    tr.model.eval()
    with torch.no_grad():
        lr_ = rm(batch[0], None, None, batch[1], batch[2])
        lo = tr.model(batch[0], None, None, batch[1], batch[2])
    assert float((lr_ - lo).abs().max()) < 1e-5
    opt = torch.optim.Adam(rm.parameters(), lr=1.0)
    opt.load_state_dict(ck["optimizer"])                       # same parameter order as the reference's model.parameters()
    assert opt.param_groups[0]["lr"] == 1e-2
