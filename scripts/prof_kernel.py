"""Launches one hot kernel at its cfg-2 shape a few times (for `ncu --set full -k regex:<name>`)."""
import sys

import torch

sys.path.insert(0, ".")
from bpmult_b200.engine import Dims
from bpmult_b200.ops import CudaOps, Drop

which = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ops = CudaOps()
dev = ops.device
WIDE = which.startswith("attn128")                     # head dim 128 kernels (attn_tc128.cu): hidden 768 / 6 heads
d = Dims(768, 6) if WIDE else Dims(300, 12)
B, T = int(sys.argv[3]) if len(sys.argv) > 3 else 64, 512
M = B * T
bf = torch.bfloat16
if which in ("attn_fwd", "attn_bwd", "attn_fwd_drop", "attn_bwd_drop", "attn128_fwd", "attn128_bwd", "attn128_fwd_drop", "attn128_bwd_drop"):
    q = torch.randn(M, d.HP, device=dev).to(bf) * 0.3
    k = torch.randn(M, d.HP, device=dev).to(bf) * 0.3
    v = torch.randn(M, d.HP, device=dev).to(bf)
    o = torch.empty(M, d.HP, device=dev, dtype=bf)
    lse = torch.empty(B * d.H * T, device=dev)
    drop = Drop(0.1, 1, None, 3) if which.endswith("drop") else None
    bits = torch.zeros(B * d.H * T * (T // 32), dtype=torch.int32, device=dev) if drop else None
    for _ in range(iters):
        ops.xattn_fwd(q, k, v, o, lse, B, T, T, d.H, d.dh, d.dhp, mask_off=0, drop=drop, drop_bits=bits)
    if "bwd" in which:
        do = torch.randn(M, d.HP, device=dev).to(bf)
        dq, dk, dv = [torch.empty(M, d.HP, device=dev, dtype=bf) for _ in range(3)]
        delta = torch.empty(ops.xattn_bwd_workspace(bf, B, T, T, d.H, d.dh, d.dhp), device=dev)
        for _ in range(iters):
            ops.xattn_bwd(q, k, v, o, do, lse, delta, dq, d.scaling, dk, dv, B, T, T, d.H, d.dh, d.dhp, mask_off=0, drop=drop, drop_bits=bits)
elif which.startswith("gemm"):
    shapes = {"gemm_fc1": (M, d.FP, d.Dp), "gemm_q": (M, d.HP, d.Dp), "gemm_fc2": (M, d.Dp, d.FP)}
    Mm, N, K = shapes[which]
    A = torch.randn(Mm, K, device=dev).to(bf)
    W = torch.randn(N, K, device=dev).to(bf)
    bias = torch.zeros(N, device=dev)
    if which == "gemm_fc2":
        x = torch.randn(Mm, N, device=dev)
        C = torch.empty(Mm, N, device=dev)
        for _ in range(iters):
            ops.gemm(A, W, C, Mm, N, K, bias=bias, drop=Drop(0.1, 1, None, 3), residual=x)
    else:
        C = torch.empty(Mm, N, device=dev, dtype=bf)
        for _ in range(iters):
            ops.gemm(A, W, C, Mm, N, K, bias=bias, act=1 if which == "gemm_fc1" else 0, drop=Drop(0.1, 1, None, 3) if which == "gemm_fc1" else None)
elif which.startswith("wgrad"):
    # weight gradients dW = dY^T X (+ bias gradient): [N_out, M rows] x [M rows, K_in], split-K over the M = 32768 rows
    shapes = {"wgrad_q": (d.HP, d.Dp), "wgrad_o": (d.Dp, d.HP), "wgrad_fc1": (d.FP, d.Dp), "wgrad_fc2": (d.Dp, d.FP), "wgrad_kv": (8 * d.HP, d.Dp)}
    No, Ki = shapes[which]
    dY = torch.randn(M, No, device=dev).to(bf)
    X = torch.randn(M, Ki, device=dev).to(bf)
    G = torch.zeros(No, Ki, device=dev)
    gb = torch.zeros(No, device=dev)
    for _ in range(iters):
        ops.gemm(dY, X, G, No, Ki, M, ta=1, tb=1, accumulate=True, colsum=gb)
elif which == "ln_fwd":
    x = torch.randn(M, d.Dp, device=dev)
    y = torch.empty(M, d.Dp, device=dev, dtype=bf)
    g, b = torch.ones(d.Dp, device=dev), torch.zeros(d.Dp, device=dev)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    for _ in range(iters):
        ops.layernorm_fwd(x, g, b, d.D, y, mean, rstd)
elif which == "ln_bwd":
    x = torch.randn(M, d.Dp, device=dev)
    dy = torch.randn(M, d.Dp, device=dev).to(bf)
    dx = torch.randn(M, d.Dp, device=dev)
    co = torch.empty(M, d.Dp, device=dev, dtype=bf)
    g = torch.ones(d.Dp, device=dev)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    dg, db = torch.zeros(d.Dp, device=dev), torch.zeros(d.Dp, device=dev)
    for _ in range(iters):
        ops.layernorm_bwd(dy, x, mean, rstd, g, d.D, dx, True, dg, db, co, None)
torch.cuda.synchronize()
print("ok", which)
