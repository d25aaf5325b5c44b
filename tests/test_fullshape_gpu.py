"""GPU parity at the BENCHMARKED shapes (BASELINE.json configs[1] and configs[2]), product path vs the oracle restatement
(oracle/functional.py, pinned to the reference by oracle/make_golden.py) evaluated in true fp32 on the same GPU:

  mmtrvat   D=300, H=12 (head dim 25 -> 32), L=8, all streams padded to T=S=512, cfg-2 input widths 768 / 35 / 74, B=2
  mmtrvapt  D=768, H=6 (head dim 128), L=5, lengths 512 / 200 / 200, video 4096-d, poster 4096-d, B=1

dropout 0, attn_mask=True, train() mode.  Bars (north star): fp32 mode <= 1e-4, bf16 mode logits <= 1e-2 max-rel; bf16 parameter
gradients in relative L2 against max(2e-2, 2 x the error of torch's own bf16 autocast of the oracle), both numbers printed.
These are the shapes `bench.py` times: folded K/V GEMMs over 8 layers, 6 lanes, CTA-pair GEMMs, the lse/delta-folded attention
backward (mmtrvat) and the head-dim-128 tensor-core attention (mmtrvapt) all run here exactly as in the benchmark."""
from argparse import Namespace

import pytest
import torch

from helpers import run_model4_engine, run_model_engine
from oracle import functional as Fn
from oracle import synth

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def ops():
    from bpmult_b200.ops import CudaOps
    return CudaOps()


def _rec_vat():
    cfg = synth.tiny_cfg(hidden_sz=300, num_heads=12, layers=8, orig_d_l=768, orig_d_v=35, orig_d_a=74, n_classes=6)
    return dict(cfg=vars(cfg), dims=(2, 50, 500, 500), seed=4242, pos_weight=torch.ones(6))


def _rec_vapt():
    cfg = synth.tiny_cfg(hidden_sz=768, num_heads=6, layers=5, orig_d_l=768, orig_d_v=4096, orig_d_a=96, orig_d_p=4096, n_classes=13)
    return dict(cfg=vars(cfg), dims=(1, 512, 200, 200), seed=777, pos_weight=torch.ones(13))


def _oracle(rec, autocast, four):
    """the oracle restatement on the GPU (fp32, TF32 off; or under torch's bf16 autocast): logits, z, dtxt, parameter gradients"""
    cfg = Namespace(**rec["cfg"])
    B, T_l, T_a, T_v = rec["dims"]
    shapes = synth.mmtrvapt_shapes(cfg) if four else synth.mmtrvat_shapes(cfg)
    sd = synth.make_state_dict(shapes, rec["seed"])
    ins = synth.mmtrvapt_inputs(cfg, B, T_l, T_a, T_v) if four else synth.mmtrvat_inputs(cfg, B, T_l, T_a, T_v)
    ins = [t.cuda() for t in ins]
    sdo = {k: v.cuda().requires_grad_() for k, v in sd.items()}
    ins[0].requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits, z = (Fn.mmtrvapt_forward if four else Fn.mmtrvat_forward)(sdo, cfg, *ins[:-1])
    loss = Fn.bce_with_logits(logits.float(), ins[-1], rec["pos_weight"].cuda())
    loss.backward()
    out = (logits.float().detach().cpu(), z.float().detach().cpu(), float(loss), ins[0].grad.cpu(),
           {n: v.grad.cpu() for n, v in sdo.items() if v.grad is not None})
    del sdo, ins, logits, z, loss
    torch.cuda.empty_cache()
    return out


def _check(tag, dtype, ours, ref32, refac):
    logits, z, loss, dtxt, grads = ours
    l32, z32, loss32, dtxt32, pg32 = ref32
    fp32 = dtype == torch.float32
    e_log, e_z, e_dtxt = Fn.max_rel(logits, l32), Fn.max_rel(z, z32), Fn.rel_l2(dtxt, dtxt32)
    report = sorted(((Fn.rel_l2(grads[n], pg32[n]), n) for n in pg32), reverse=True)
    print("%s %s: logits max-rel %.3e, gates max-rel %.3e, loss %.6f vs %.6f, dtxt rel-l2 %.3e, worst param grads %s"
          % (tag, "fp32" if fp32 else "bf16", e_log, e_z, loss, loss32, e_dtxt, ["%.2e %s" % r for r in report[:3]]))
    assert set(pg32) <= set(grads)
    if fp32:
        assert e_log < 1e-4 and e_z < 1e-4 and abs(loss - loss32) < 1e-5
        assert e_dtxt < 1e-4
        assert report[0][0] < 1e-4, report[0]
        return
    lac, zac, lossac, dtxtac, pgac = refac
    e_log_ac = Fn.max_rel(lac, l32)
    ac = {n: Fn.rel_l2(pgac[n], pg32[n]) for n in pg32}
    ac_worst = max(ac.values())
    print("%s bf16: torch-autocast of the oracle: logits max-rel %.3e, dtxt rel-l2 %.3e, worst param grad rel-l2 %.3e; ours on that tensor %.3e"
          % (tag, e_log_ac, Fn.rel_l2(dtxtac, dtxt32), ac_worst, Fn.rel_l2(grads[max(ac, key=ac.get)], pg32[max(ac, key=ac.get)])))
    assert e_log < 1e-2, "bf16 logits miss the 1e-2 bar: %.3e (torch autocast: %.3e)" % (e_log, e_log_ac)
    assert e_z < 1e-2
    assert abs(loss - loss32) < 1e-2 * max(1.0, abs(loss32))
    assert e_dtxt < max(2e-2, 2.0 * Fn.rel_l2(dtxtac, dtxt32))
    for e, n in report:
        assert e < max(2e-2, 2.0 * ac_worst), (n, e, ac[n], ac_worst)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_mmtrvat_benchmark_shape(ops, dtype):
    """mmtr.py:735-866 + train.py:99-106 at the cfg-2 shape"""
    rec = _rec_vat()
    logits, z, loss, dtxt, grads, eng = run_model_engine(ops, rec, dtype=dtype)
    torch.cuda.synchronize()
    assert eng.lanes.n == 6 and eng.enc["l_with_a"].fold_kv
    ours = (logits, z, float(loss), dtxt, grads)
    del eng
    torch.cuda.empty_cache()
    ref32 = _oracle(rec, False, False)
    refac = _oracle(rec, True, False) if dtype == torch.bfloat16 else None
    _check("mmtrvat D=300 H=12 L=8 T=512 B=2", dtype, ours, ref32, refac)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_mmtrvapt_benchmark_shape(ops, dtype):
    """mmtr.py:444-583 at the cfg-3 shape (head dim 128: the tensor-core attention for dh = 128 in bf16 mode)"""
    rec = _rec_vapt()
    logits, z, loss, dtxt, grads, eng = run_model4_engine(ops, rec, dtype=dtype)
    torch.cuda.synchronize()
    ours = (logits, z, float(loss), dtxt, grads)
    del eng
    torch.cuda.empty_cache()
    ref32 = _oracle(rec, False, True)
    refac = _oracle(rec, True, True) if dtype == torch.bfloat16 else None
    _check("mmtrvapt D=768 H=6 L=5 512/200/200 B=1", dtype, ours, ref32, refac)
