"""Shared test helpers: drive the product engines on reference-layout tensors."""
import os

import torch

from bpmult_b200.engine import EncoderEngine, round_up
from oracle import functional as Fn
from oracle import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def to_dev(sd, dev):
    return {k: v.to(dev) for k, v in sd.items()}


def run_encoder_engine(ops, sd, x, k, g, H, L, mask, biproj, self_only, dtype=torch.float32, p=None, training=True, seed=0):
    """x (T,B,D), k (S,B,D) time-major fp32 CPU tensors; g upstream gradient (T,B,D).
    Returns out (T,B,D), dx, dk, param grads (reference layout) as CPU fp32."""
    dev = ops.device
    T, B, D = x.shape
    S = k.shape[0] if k is not None else None
    p = p or {}
    eng = EncoderEngine(ops, D, H, L, attn_mask=mask, biprojection=biproj, dtype=dtype, uid=3, **p)
    params = to_dev(sd, dev)
    eng.pack(params)
    Dp = eng.d.Dp
    xq = ops.empty((B * T, Dp), dtype)
    ops.stage_rows(x.to(dev).permute(1, 0, 2), xq, T)
    xk = None
    if not self_only:
        xk = ops.empty((B * S, Dp), dtype)
        ops.stage_rows(k.to(dev).permute(1, 0, 2), xk, S)
    out = eng.forward(xq, B, T, src_k=xk, S=S, training=training, seed=seed)
    out_tbd = out.float().view(B, T, Dp)[:, :, :D].permute(1, 0, 2).cpu()
    # backward
    eng.zero_grads()
    dout = ops.empty((B * T, Dp), torch.float32)
    ops.stage_rows(g.to(dev).permute(1, 0, 2), dout, T)
    dq = ops.zeros((B * T, Dp), torch.float32)
    dk = ops.zeros((B * S, Dp), torch.float32) if not self_only else None
    eng.backward(dout, dq, dk)
    grads = {n: torch.zeros(s, device=dev) for n, s in eng.param_shapes().items()}
    eng.unpack_grads(grads)
    dx = dq.view(B, T, Dp)[:, :, :D].permute(1, 0, 2).cpu()
    dkk = dk.view(B, S, Dp)[:, :, :D].permute(1, 0, 2).cpu() if dk is not None else None
    return out_tbd, dx, dkk, {n: v.cpu() for n, v in grads.items()}, eng


def run_model_engine(ops, rec, dtype=torch.float32, full=True, sd=None):
    from argparse import Namespace
    from bpmult_b200.model_engine import MMTrVatEngine
    cfg = Namespace(**rec["cfg"])
    B, T_l, T_a, T_v = rec["dims"]
    if sd is None:
        sd = synth.make_state_dict(synth.mmtrvat_shapes(cfg), rec["seed"])
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, B, T_l, T_a, T_v)
    eng = MMTrVatEngine(ops, cfg, dtype=dtype)
    dev = ops.device
    params = {k: v.to(dev) for k, v in sd.items()}
    assert set(eng.param_shapes().keys()) | set(eng.unused_params()) >= set(sd.keys())
    eng.pack(params)
    logits, z = eng.forward(txt.to(dev), img.to(dev), audio.to(dev), training=True)
    loss, dlogits = eng.loss(logits, tgt.to(dev), rec["pos_weight"].to(dev))
    eng.zero_grads()
    dtxt = torch.zeros_like(txt, device=dev)
    eng.backward(dlogits, {"l": dtxt})
    grads = {n: torch.zeros(s, device=dev) for n, s in eng.param_shapes().items()}
    eng.unpack_grads(grads)
    D, Dp, C = cfg.hidden_sz, eng.d.Dp, cfg.n_classes
    ng = 4 if getattr(cfg, "hybrid", False) else 3
    zz = z.view(B, ng, Dp)[:, :, :D].reshape(B, ng * D)
    return logits[:, :C].cpu(), zz.cpu(), loss.cpu(), dtxt.cpu(), {n: v.cpu() for n, v in grads.items()}, eng




def fingerprint_errors(grads, fps):
    """per-tensor error of gradients against stored fingerprints (max of the strided-sample rel-L2 and the norm error)"""
    out = {}
    for n, fp in fps.items():
        g = grads[n].detach().reshape(-1).double().cpu()
        ref = fp["val"].double()
        e = float((g[fp["idx"]] - ref).norm() / ref.norm().clamp_min(1e-30))
        en = abs(float(g.norm()) - fp["norm"]) / max(fp["norm"], 1e-30)
        out[n] = max(e, en)
    return out


def check_fingerprints(grads, fps, tol):
    """gradients against the stored fingerprints (norm + strided sample, oracle/synth.py:summarize) of the reference's"""
    worst = (0.0, "")
    for n, fp in fps.items():
        g = grads[n].detach().reshape(-1).double().cpu()
        ref = fp["val"].double()
        e = float((g[fp["idx"]] - ref).norm() / ref.norm().clamp_min(1e-30))
        en = abs(float(g.norm()) - fp["norm"]) / max(fp["norm"], 1e-30)
        worst = max(worst, (max(e, en), n))
        assert tuple(grads[n].shape) == tuple(fp["shape"]), n
    assert worst[0] < tol, worst
    return worst


def run_model4_engine(ops, rec, dtype=torch.float32, sd=None):
    """MMTrVaptEngine on a golden record: returns logits, z, loss, dtxt, reference-layout parameter gradients"""
    from argparse import Namespace
    from bpmult_b200.model_engine4 import MMTrVaptEngine
    cfg = Namespace(**rec["cfg"])
    B, T_l, T_a, T_v = rec["dims"]
    if sd is None:
        sd = synth.make_state_dict(synth.mmtrvapt_shapes(cfg), rec["seed"])
    txt, img, audio, poster, tgt = synth.mmtrvapt_inputs(cfg, B, T_l, T_a, T_v)
    eng = MMTrVaptEngine(ops, cfg, dtype=dtype)
    dev = ops.device
    params = {k: v.to(dev) for k, v in sd.items()}
    assert set(eng.param_shapes().keys()) == set(sd.keys()), set(eng.param_shapes().keys()) ^ set(sd.keys())
    for k, shp in eng.param_shapes().items():
        assert tuple(sd[k].shape) == tuple(shp), k
    eng.pack(params)
    logits, z = eng.forward(txt.to(dev), img.to(dev), audio.to(dev), poster.to(dev), training=True)
    loss, dlogits = eng.loss(logits, tgt.to(dev), rec["pos_weight"].to(dev))
    eng.zero_grads()
    dtxt = torch.zeros_like(txt, device=dev)
    eng.backward(dlogits, {"l": dtxt})
    grads = {n: torch.zeros(s, device=dev) for n, s in eng.param_shapes().items()}
    eng.unpack_grads(grads)
    D, Dp, C = cfg.hidden_sz, eng.d.Dp, cfg.n_classes
    zz = z.view(B, 4, Dp)[:, :, :D].reshape(B, 4 * D)
    return logits[:, :C].cpu(), zz.cpu(), loss.cpu(), dtxt.cpu(), {n: v.cpu() for n, v in grads.items()}, eng
