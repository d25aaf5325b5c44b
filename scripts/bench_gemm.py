"""Times bpm_gemm at the cfg-2 layer shapes (CUDA events, rotating buffer sets so operands are not L2-resident),
optionally with the diagnostic stage-bypass knobs (bpm_debug_set slot 0) to see which stage bounds the kernel."""
import sys

import torch

sys.path.insert(0, ".")
from bpmult_b200.engine import Dims
from bpmult_b200.ops import CudaOps, Drop

ops = CudaOps()
dev = ops.device
d = Dims(300, 12)
M = 64 * 512
bf = torch.bfloat16
import os
NSET = int(os.environ.get("NSET", "3"))


def mk(shape, dtype=bf):
    return [torch.randn(shape, device=dev, dtype=torch.float32).to(dtype) for _ in range(NSET)]


def timeit(fn, iters=12):
    """the iterations are captured in one CUDA graph: a python/ctypes launch costs ~20 us of CPU, more than most of these kernels"""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3):
            fn(i % NSET)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(iters):
                fn(i % NSET)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        g.replay()
        g.replay()
        e1.record(st)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / (2 * iters)


def case(name, Mm, N, K, ta=0, tb=0, out_dtype=bf, bias=False, act=0, drop=False, res=False, gate=False, acc=False, colsum=False):
    if ta:
        A = mk((K, Mm))
    else:
        A = mk((Mm, K))
    if tb:
        B = mk((K, N))
    else:
        B = mk((N, K))
    Cs = [torch.zeros((Mm, N), device=dev, dtype=out_dtype) for _ in range(NSET)]
    bias_t = torch.zeros(N, device=dev) if bias else None
    R = mk((Mm, N), out_dtype) if res else None
    G = mk((Mm, N), out_dtype) if gate else None
    cs = torch.zeros(Mm, device=dev) if colsum else None

    def fn(i):
        ops.gemm(A[i], B[i], Cs[i], Mm, N, K, ta=ta, tb=tb, bias=bias_t, act=act, drop=Drop(0.1, 1, None, 3) if drop else None,
                 residual=R[i] if res else None, gate=G[i] if gate else None, accumulate=acc, colsum=cs)
    return name, fn, 2.0 * Mm * N * K


cases = [
    case("q      M x384 x320 bias", M, d.HP, d.Dp, bias=True),
    case("out    M x320 x384 bias drop res f32", M, d.Dp, d.HP, out_dtype=torch.float32, bias=True, drop=True, res=True),
    case("out    M x320 x384 plain bf16", M, d.Dp, d.HP),
    case("out    M x320 x384 plain f32", M, d.Dp, d.HP, out_dtype=torch.float32),
    case("out    M x320 x384 bias res f32", M, d.Dp, d.HP, out_dtype=torch.float32, bias=True, res=True),
    case("out    M x320 x384 bias drop f32", M, d.Dp, d.HP, out_dtype=torch.float32, bias=True, drop=True),
    case("fc1    M x1216x320 bias relu drop", M, d.FP, d.Dp, bias=True, act=1, drop=True),
    case("fc1    M x1216x320 bias relu", M, d.FP, d.Dp, bias=True, act=1),
    case("fc1    M x1216x320 plain", M, d.FP, d.Dp),
    case("fc2    M x320 x1216 bias drop res f32", M, d.Dp, d.FP, out_dtype=torch.float32, bias=True, drop=True, res=True),
    case("fc2    M x320 x1216 plain bf16", M, d.Dp, d.FP),
    case("dgrad dh  M x1216x320 tb gate", M, d.FP, d.Dp, tb=1, gate=True),
    case("dgrad dhn M x320 x1216 tb", M, d.Dp, d.FP, tb=1),
    case("dgrad da  M x384 x320 tb", M, d.HP, d.Dp, tb=1),
    case("wgrad W2 320 x1216x M colsum", d.Dp, d.FP, M, ta=1, tb=1, out_dtype=torch.float32, acc=True, colsum=True),
    case("wgrad W1 1216x320 x M colsum", d.FP, d.Dp, M, ta=1, tb=1, out_dtype=torch.float32, acc=True, colsum=True),
    case("wgrad Wq 384 x320 x M colsum", d.HP, d.Dp, M, ta=1, tb=1, out_dtype=torch.float32, acc=True, colsum=True),
]
knobs = [0] if len(sys.argv) < 2 else [int(x) for x in sys.argv[1].split(",")]
if len(sys.argv) > 2:
    ops.lib.bpm_debug_set(2, int(sys.argv[2]))      # max BN override
if len(sys.argv) > 3:
    ops.lib.bpm_debug_set(3, int(sys.argv[3]))      # staging buffers per epilogue warp
print("%-40s" % "case" + "".join("  dbg=%-3d us (TF/s)" % k for k in knobs))
import os
for name, fn, fl in cases:
    if os.environ.get("BG_FILTER", "") not in name:
        continue
    row = "%-40s" % name
    for k in knobs:
        has_in = ("res" in name) or ("gate" in name)
        if (k & 8) and has_in:
            row += "  %18s" % "-"
            continue
        ops.lib.bpm_debug_set(0, k)
        us = timeit(fn)
        ops.lib.bpm_debug_set(0, 0)
        row += "  %8.1f (%6.0f)" % (us, fl / us * 1e-6)
    print(row, flush=True)
