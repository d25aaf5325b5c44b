import sys
sys.path.insert(0, ".")
import torch
exec(open("scripts/bench_gemm.py").read().split("cases = [")[0])
for name, (Mm, N) in (("W2 320x1216", (d.Dp, d.FP)), ("W1 1216x320", (d.FP, d.Dp)), ("Wq 384x320", (d.HP, d.Dp)), ("Wo 320x384", (d.Dp, d.HP))):
    A = mk((M, Mm)); B = mk((M, N))
    Cs = [torch.zeros((Mm, N), device=dev) for _ in range(NSET)]
    cs = torch.zeros(Mm, device=dev)
    row = "%-14s" % name
    for sk in (0, 6, 8, 9, 10, 12, 16, 19, 24, 32):
        def fn(i, sk=sk):
            ops.gemm(A[i], B[i], Cs[i], Mm, N, M, ta=1, tb=1, accumulate=True, colsum=cs, split_k=sk)
        try:
            row += "  sk%-2d %5.1f" % (sk, timeit(fn))
        except Exception as e:
            row += "  sk%-2d  err" % sk
    print(row, flush=True)
