"""Diagnostic (GPU): where do the largest parameter-gradient errors of the benchmark-shape parity test sit?  For every tensor whose
relative-L2 error exceeds `thr` prints how concentrated the error is (share of the squared error carried by the k largest elements)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def main():
    import test_fullshape_gpu as Tf
    from helpers import run_model4_engine, run_model_engine
    from bpmult_b200.ops import CudaOps
    from oracle import functional as Fn
    ops = CudaOps()
    which = sys.argv[1] if len(sys.argv) > 1 else "vat"
    dtype = torch.float32 if (len(sys.argv) < 3 or sys.argv[2] == "fp32") else torch.bfloat16
    rec = Tf._rec_vat() if which == "vat" else Tf._rec_vapt()
    run = run_model_engine if which == "vat" else run_model4_engine
    logits, z, loss, dtxt, grads, eng = run(ops, rec, dtype=dtype)
    del eng
    torch.cuda.empty_cache()
    l32, z32, loss32, dtxt32, pg32, _ = Tf._oracle(rec, False, which != "vat")
    rep = sorted(((Fn.rel_l2(grads[n], pg32[n]), n) for n in pg32), reverse=True)
    thr = 5e-5 if dtype == torch.float32 else 5e-2
    print("tensors above %.0e: %d of %d" % (thr, sum(1 for e, _ in rep if e > thr), len(rep)))
    for e, n in rep[:24]:
        d = (grads[n].double() - pg32[n].double()).reshape(-1)
        sq = d.pow(2)
        tot = float(sq.sum())
        top = torch.topk(sq, min(8, sq.numel()))
        shares = [float(top.values[:k].sum()) / max(tot, 1e-300) for k in (1, 2, 4, 8)]
        idx = [int(i) for i in top.indices[:4]]
        cols = pg32[n].shape[-1] if pg32[n].dim() > 1 else 0
        where = [(i // cols, i % cols) if cols else i for i in idx]
        print("%.3e %-58s shape %-14s share of err^2 in top 1/2/4/8 elements: %.2f %.2f %.2f %.2f  at %s" %
              (e, n, tuple(pg32[n].shape), shares[0], shares[1], shares[2], shares[3], where))


if __name__ == "__main__":
    main()
