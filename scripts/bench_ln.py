"""Times the LayerNorm kernels at the cfg-2 shape (rows = 64 x 512, D = 300 stored 320; CUDA-graph captured iterations, rotating buffers)."""
import sys

import torch

sys.path.insert(0, ".")
from bpmult_b200.engine import Dims
from bpmult_b200.ops import CudaOps

import os
ops = CudaOps()
dev = ops.device
d = Dims(int(os.environ.get("LN_D", "300")), 12 if os.environ.get("LN_D", "300") == "300" else 6)
M = int(os.environ.get("LN_ROWS", str(64 * 512)))
if os.environ.get("LN_RPW"):
    ops.lib.bpm_debug_set(2, int(os.environ["LN_RPW"]))
NS = 4
bf = torch.bfloat16
x = [torch.randn(M, d.Dp, device=dev) for _ in range(NS)]
y = [torch.empty(M, d.Dp, device=dev, dtype=bf) for _ in range(NS)]
dy = [torch.randn(M, d.Dp, device=dev).to(bf) for _ in range(NS)]
dx = [torch.randn(M, d.Dp, device=dev) for _ in range(NS)]
co = [torch.empty(M, d.Dp, device=dev, dtype=bf) for _ in range(NS)]
g, b = torch.ones(d.Dp, device=dev), torch.zeros(d.Dp, device=dev)
mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
dg, db = torch.zeros(d.Dp, device=dev), torch.zeros(d.Dp, device=dev)


def timeit(fn, iters=16):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3):
            fn(i % NS)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for i in range(iters):
                fn(i % NS)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        gr.replay()
        e1.record(st)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / iters


t = timeit(lambda i: ops.layernorm_fwd(x[i], g, b, d.D, y[i], mean, rstd))
print("ln fwd (fp32 -> bf16):                         %6.1f us  %5.2f TB/s" % (t, M * d.Dp * 6 / t / 1e6))
t = timeit(lambda i: ops.layernorm_bwd(dy[i], x[i], mean, rstd, g, d.D, dx[i], True, dg, db, co[i], None))
print("ln bwd (bf16 dy, fp32 x, dx += , bf16 cast):   %6.1f us  %5.2f TB/s" % (t, M * d.Dp * (2 + 4 + 8 + 2) / t / 1e6))
t = timeit(lambda i: ops.layernorm_bwd(dy[i], x[i], mean, rstd, g, d.D, dx[i], True, dg, db))
print("ln bwd (bf16 dy, fp32 x, dx += ):              %6.1f us  %5.2f TB/s" % (t, M * d.Dp * (2 + 4 + 8) / t / 1e6))
# calibration: plain device copies of the same size / larger (what the memory system gives a trivially parallel stream)
xs = [torch.randn(M, d.Dp, device=dev) for _ in range(NS)]
t = timeit(lambda i: xs[i].copy_(x[i]))
print("torch copy fp32 [rows, 320] (42 MB -> 42 MB):  %6.1f us  %5.2f TB/s" % (t, M * d.Dp * 8 / t / 1e6))
big_a = [torch.empty(64 << 20, device=dev) for _ in range(2)]
big_b = [torch.empty(64 << 20, device=dev) for _ in range(2)]
t = timeit(lambda i: big_b[i % 2].copy_(big_a[i % 2]), iters=8)
print("torch copy fp32 256 MB -> 256 MB:              %6.1f us  %5.2f TB/s" % (t, (64 << 20) * 8 / t / 1e6))
