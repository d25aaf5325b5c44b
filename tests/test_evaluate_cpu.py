"""Evaluation path (train.py:165-280): the host half (metrics dictionary per task) against independent statements of the same
definitions, and the device half (Trainer.evaluate) against the oracle's eval-mode forward -- on CPU with the ops emulation."""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

from oracle import functional as Fn
from oracle import synth

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _loop_weighted_acc(preds, truths):
    """the definition, element by element (train.py:140-163)"""
    tp = tn = p = n = 0
    for pr, tr in zip(preds, truths):
        if tr == 0:
            n += 1
            tn += int(pr == 0)
        elif tr == 1:
            p += 1
            tp += int(pr == 1)
    fp, fn = n - tn, p - tp
    rec, prec = tp / (tp + fn + 1e-8), tp / (tp + fp + 1e-8)
    return (tp * n / p + tn) / (2 * n), 2 * rec * prec / (rec + prec + 1e-8)


def _fake(n=64, c=6, seed=0):
    g = np.random.default_rng(seed)
    tgts = (g.random((n, c)) < 0.35).astype(np.float32)
    raw = np.clip(0.5 * tgts + 0.6 * g.random((n, c)), 0, 1).astype(np.float32)
    return tgts, raw > 0.5, raw


def test_weighted_acc_matches_its_definition():
    from bpmult_b200.evaluate import weighted_acc
    tgts, preds, _ = _fake()
    for c in range(tgts.shape[1]):
        a, f = weighted_acc(preds[:, c], tgts[:, c])
        a0, f0 = _loop_weighted_acc(preds[:, c].astype(int), tgts[:, c].astype(int))
        assert abs(a - a0) < 1e-12 and abs(f - f0) < 1e-12


def test_metric_keys_and_values_per_task():
    from sklearn.metrics import average_precision_score, f1_score
    from bpmult_b200.evaluate import multilabel_metrics
    tgts, preds, raw = _fake(seed=3)
    m = multilabel_metrics("moviescope", tgts, preds, raw, [0.5, 0.7])
    assert list(m) == ["loss", "macro_f1", "micro_f1", "auc_pr_macro", "auc_pr_micro", "auc_pr_samples"] and abs(m["loss"] - 0.6) < 1e-12
    assert m["macro_f1"] == f1_score(tgts, preds, average="macro") and m["auc_pr_samples"] == average_precision_score(tgts, raw, average="samples")
    m = multilabel_metrics("mmimdb", tgts, preds, raw, [1.0])
    # the reference's own key assignment for this task (train.py:202-207): names and contents differ, checkpoints are selected on them
    assert m["micro_f1"] == average_precision_score(tgts, raw, average="micro")
    assert m["auc_pr_macro"] == f1_score(tgts, preds, average="weighted")
    assert m["auc_pr_micro"] == f1_score(tgts, preds, average="micro")
    assert m["auc_pr_samples"] == f1_score(tgts, preds, average="samples")
    m = multilabel_metrics("cmu-mosei", tgts, preds, raw, [1.0])
    per = [_loop_weighted_acc(preds[:, c].astype(int), tgts[:, c].astype(int)) for c in range(6)]
    assert [k for k in m if k.startswith("f1_emo") and k != "f1_emos"] == ["f1_emo%d" % i for i in range(1, 7)]
    assert abs(m["wacc_emo3"] - per[2][0]) < 1e-12 and abs(m["f1_emo5"] - per[4][1]) < 1e-12
    assert abs(m["f1_emos"] - np.mean([f for _, f in per])) < 1e-12
    assert abs(m["auc_pr_micro"] - np.mean([a for a, _ in per])) < 1e-12          # train.py:261: the mean weighted accuracy lives under this key
    assert m["wacc_emos"] == average_precision_score(tgts, raw, average="micro")  # train.py:260
    from sklearn.metrics import accuracy_score
    m = multilabel_metrics("counseling", tgts[:, :2], preds[:, :2], raw[:, :2], [1.0])
    assert list(m) == ["loss", "f1_low", "f1_high", "acc", "auc_pr_micro"]
    assert abs(m["f1_low"] - per[1][1]) < 1e-12 and abs(m["f1_high"] - per[0][1]) < 1e-12          # train.py:229-230 (overwrite the sklearn values)
    assert m["acc"] == accuracy_score(tgts[:, :2], preds[:, :2]) and m["auc_pr_micro"] == average_precision_score(tgts[:, :2], raw[:, :2], average="micro")
    try:
        multilabel_metrics("cmu-mosi", tgts, preds, raw, [1.0])
        assert False
    except ValueError:
        pass


def test_model_eval_equals_the_oracle_in_eval_mode():
    import bpmult_b200.modules as M
    from bpmult_b200.evaluate import model_eval
    from bpmult_b200.trainer import Trainer
    from emu_ops import EmuOps
    o = EmuOps()
    M._ops_for = lambda device: o
    cfg = synth.tiny_cfg(layers=1)
    cfg.attn_dropout, cfg.relu_dropout, cfg.res_dropout, cfg.embed_dropout, cfg.out_dropout = 0.1, 0.1, 0.1, 0.25, 0.1   # must be OFF in eval
    sd = synth.make_state_dict(synth.mmtrvat_shapes(cfg), 5)
    m = M.MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg)), precision="fp32")
    m.load_state_dict(sd, strict=False)
    pw = torch.tensor([1.0, 2.0, 0.5, 3.0][:cfg.n_classes] + [1.0] * max(0, cfg.n_classes - 4))
    tr = Trainer(m.train(), lr=1e-3, use_graph=False, pos_weight=pw)
    batches = [synth.mmtrvat_inputs(cfg, 3, 8, 12, 10, seed=s) for s in (1, 2)]
    metrics, arrays = model_eval(batches, tr, "moviescope", output_gates=True)
    cfg0 = Namespace(**vars(cfg))
    cfg0.attn_dropout = cfg0.relu_dropout = cfg0.res_dropout = cfg0.embed_dropout = cfg0.out_dropout = 0.0
    cfg0.attn_dropout_a = cfg0.attn_dropout_v = 0.0
    losses, probs, zs = [], [], []
    for txt, img, audio, tgt in batches:
        lo, z = Fn.mmtrvat_forward(sd, cfg0, txt, img, audio)
        losses.append(float(Fn.bce_with_logits(lo, tgt, pw)))
        probs.append(torch.sigmoid(lo))
        zs.append(z)
    assert abs(metrics["loss"] - np.mean(losses)) < 1e-5
    assert np.abs(arrays["preds_raw"] - torch.cat(probs).numpy()).max() < 1e-5
    assert np.abs(arrays["gates"] - torch.cat(zs).numpy()).max() < 1e-5
    assert arrays["tgts"].shape == (6, cfg.n_classes) and arrays["preds"].dtype == bool
    # evaluating must not disturb training: a step after evaluate equals a step without it
    tr2 = Trainer(M.MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg0)), precision="fp32").train(), lr=1e-3, use_graph=False, pos_weight=pw)
    tr3 = Trainer(M.MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg0)), precision="fp32").train(), lr=1e-3, use_graph=False, pos_weight=pw)
    for t in (tr2, tr3):
        t.model.load_state_dict(sd, strict=False)
    tr2.evaluate(*batches[0])
    l2, l3 = tr2.step(*batches[1]), tr3.step(*batches[1])
    assert l2 == l3 and torch.equal(tr2.flat_p, tr3.flat_p)
