"""Evaluation loop of the reference (train.py:165-280 `model_eval`, :140-163 `weighted_acc`) over the drop-in: the device half is
`Trainer.evaluate` (eval-mode forward, loss, sigmoid, 0.5 threshold); this module is the host half -- stacking the batches and
computing the metrics dictionary the reference logs and selects checkpoints by (`metrics[...]` keys per task, train.py:191-262,
tuning metric train.py:404-407).  Multilabel tasks of the two fusion models only (cmu-mosei, counseling for mmtrvat; moviescope, mmimdb
for mmtrvapt); the regression / single-label tasks of the reference's other models are outside SURVEY section 8.

The key -> metric assignment deliberately follows the reference, including where a key's name and its content disagree (mmimdb's
"micro_f1" holds the micro average precision, cmu-mosei's "auc_pr_micro" the mean weighted accuracy, ...): checkpoints are selected on
these keys, so a drop-in has to produce the same numbers under the same names."""
import numpy as np


def weighted_acc(preds, truths):
    """train.py:140-163 for one label column: weighted accuracy (tp * n / p + tn) / (2 n) and the eps-smoothed F1.  Like the reference it
    raises ZeroDivisionError for a column without positives (or without negatives): evaluate on a set that has both."""
    preds, truths = np.asarray(preds).reshape(-1).astype(bool), np.asarray(truths).reshape(-1)
    pos, neg = truths == 1, truths == 0
    p, n = int(pos.sum()), int(neg.sum())
    tp, tn = int((preds & pos).sum()), int((~preds & neg).sum())
    w_acc = (tp * n / p + tn) / (2 * n)
    fp, fn = n - tn, p - tp
    recall, precision = tp / (tp + fn + 1e-8), tp / (tp + fp + 1e-8)
    return w_acc, 2 * recall * precision / (recall + precision + 1e-8)


def _f1(avg):
    def fn(tgts, preds, raw):
        from sklearn.metrics import f1_score
        return f1_score(tgts, preds, average=avg)
    return fn


def _ap(avg):
    def fn(tgts, preds, raw):
        from sklearn.metrics import average_precision_score
        return average_precision_score(tgts, raw, average=avg)
    return fn


# task -> ordered (key, metric) pairs; train.py:195-207
_TABLE = {
    "moviescope": (("macro_f1", _f1("macro")), ("micro_f1", _f1("micro")), ("auc_pr_macro", _ap("macro")), ("auc_pr_micro", _ap("micro")),
                   ("auc_pr_samples", _ap("samples"))),
    "mmimdb": (("macro_f1", _f1("macro")), ("micro_f1", _ap("micro")), ("auc_pr_macro", _f1("weighted")), ("auc_pr_micro", _f1("micro")),
               ("auc_pr_samples", _f1("samples"))),
}


def multilabel_metrics(task, tgts, preds, raw_preds, losses):
    """the metrics dict of train.py:191-262 from stacked targets (N, C), thresholded predictions (N, C) and sigmoid outputs (N, C)"""
    tgts, preds, raw = np.asarray(tgts), np.asarray(preds), np.asarray(raw_preds)
    metrics = {"loss": float(np.mean(losses))}
    if task in _TABLE:
        for key, fn in _TABLE[task]:
            metrics[key] = fn(tgts, preds, raw)
    elif task == "cmu-mosei":
        # train.py:233-262: per-emotion weighted accuracy and F1, their means, and the micro average precision (under "wacc_emos")
        per = [weighted_acc(preds[:, c], tgts[:, c]) for c in range(tgts.shape[1])]
        accs, f1s = [a for a, _ in per], [f for _, f in per]
        for c in range(len(per)):
            metrics["f1_emo%d" % (c + 1)] = f1s[c]
        for c in range(len(per)):
            metrics["wacc_emo%d" % (c + 1)] = accs[c]
        metrics["f1_emos"] = float(np.average(f1s))
        metrics["wacc_emos"] = _ap("micro")(tgts, preds, raw)
        metrics["auc_pr_micro"] = float(np.average(accs))
    elif task == "counseling":
        # train.py:208-231: two labels; f1_low / f1_high end up as the eps-smoothed F1 of label 1 / label 0 (the sklearn values assigned
        # first are overwritten), acc = subset accuracy, auc_pr_micro = micro average precision
        from sklearn.metrics import accuracy_score
        per = [weighted_acc(preds[:, c], tgts[:, c]) for c in range(2)]
        metrics["f1_low"] = per[1][1]
        metrics["f1_high"] = per[0][1]
        metrics["acc"] = accuracy_score(tgts, preds)
        metrics["auc_pr_micro"] = _ap("micro")(tgts, preds, raw)
    else:
        raise ValueError("multilabel_metrics: task %r is not one of moviescope, mmimdb, cmu-mosei, counseling" % (task,))
    return metrics


def model_eval(data, trainer, task, output_gates=False):
    """train.py:165-280.  `data` yields batches (features..., targets) in the Trainer's order ((txt, img, audio, tgt) for mmtrvat,
    (txt, img, audio, poster, tgt) for mmtrvapt).  Returns (metrics, arrays) with arrays = {"tgts", "preds", "preds_raw"[, "gates"]}:
    what store_preds_to_disk (utils.py:45-84) writes out."""
    losses, preds, raws, tgts, gates = [], [], [], [], []
    for batch in data:
        r = trainer.evaluate(*batch, output_gates=output_gates)
        losses.append(r["loss"])
        preds.append(r["preds"].numpy())
        raws.append(r["probs"].numpy())
        tgts.append(r["targets"].numpy())
        if output_gates:
            gates.append(r["gates"].numpy())
    arrays = {"tgts": np.vstack(tgts), "preds": np.vstack(preds), "preds_raw": np.vstack(raws)}
    if output_gates:
        arrays["gates"] = np.vstack(gates)
    return multilabel_metrics(task, arrays["tgts"], arrays["preds"], arrays["preds_raw"], losses), arrays
