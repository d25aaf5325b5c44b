"""Times the crossmodal attention kernels (CUDA-graph captured iterations, CUDA events).
    python scripts/bench_attn.py [more]           cfg-2 shape: hidden 300 / 12 heads (head dim 25 -> 32)
    python scripts/bench_attn.py wide [B]         cfg-3 shape: hidden 768 / 6 heads (head dim 128), the lengths of the 4-modality model"""
import sys

import torch

sys.path.insert(0, ".")
from bpmult_b200.engine import Dims
from bpmult_b200.ops import CudaOps, Drop

ops = CudaOps()
dev = ops.device
WIDE = len(sys.argv) > 1 and sys.argv[1] == "wide"
d = Dims(768, 6) if WIDE else Dims(300, 12)
bf = torch.bfloat16


def timeit(fn, iters=6):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        g.replay()
        e1.record(st)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / iters


def run(B, T, S, mask_off, p):
    M, Ms = B * T, B * S
    q = torch.randn(M, d.HP, device=dev).to(bf) * 0.3
    k = torch.randn(Ms, d.HP, device=dev).to(bf) * 0.3
    v = torch.randn(Ms, d.HP, device=dev).to(bf)
    o = torch.empty(M, d.HP, device=dev, dtype=bf)
    lse = torch.empty(B * d.H * T, device=dev)
    drop = Drop(p, 1, None, 3) if p > 0 else None
    bits = torch.zeros(B * d.H * T * ((S + 31) // 32), dtype=torch.int32, device=dev) if drop else None
    do = torch.randn(M, d.HP, device=dev).to(bf)
    dq = torch.empty(M, d.HP, device=dev, dtype=bf)
    dk, dv = [torch.empty(Ms, d.HP, device=dev, dtype=bf) for _ in range(2)]
    delta = torch.empty(ops.xattn_bwd_workspace(bf, B, T, S, d.H, d.dh, d.dhp), device=dev)
    tf = timeit(lambda: ops.xattn_fwd(q, k, v, o, lse, B, T, S, d.H, d.dh, d.dhp, mask_off=mask_off, drop=drop, drop_bits=bits))
    tb = timeit(lambda: ops.xattn_bwd(q, k, v, o, do, lse, delta, dq, d.scaling, dk, dv, B, T, S, d.H, d.dh, d.dhp, mask_off=mask_off, drop=drop,
                                      drop_bits=bits))
    # algorithmic flops: 4*B*H*T*S*dh*rho fwd, 2.5x bwd
    if mask_off >= 0:
        vis = sum(min(S, i + mask_off + 1) for i in range(T))
    else:
        vis = T * S
    fl = 4.0 * B * d.H * vis * d.dh
    print("B=%d T=%d S=%d mask_off=%d p=%.1f: fwd %7.1f us (%6.1f TF/s)   bwd(+delta) %7.1f us (%6.1f TF/s)" % (
        B, T, S, mask_off, p, tf, fl / tf * 1e-6, tb, 2.5 * fl / tb * 1e-6), flush=True)


if WIDE:
    Bw = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    for (T, S, off, p) in ((512, 512, 0, 0.0), (512, 512, 0, 0.1), (512, 200, 312, 0.0), (200, 512, 312, 0.0), (200, 200, 0, 0.0), (512, 512, -1, 0.0)):
        run(Bw, T, S, off, p)
    run(64, 512, 512, 0, 0.0)
    run(16, 2048, 2048, 0, 0.0)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1].startswith("dbg="):
    for kv in sys.argv[1][4:].split(","):
        ops.lib.bpm_debug_set(1, int(kv))
        print("dbg", kv, end=": ")
        run(64, 512, 512, -1, 0.0)
    sys.exit(0)
run(64, 512, 512, 0, 0.0)
run(64, 512, 512, 0, 0.1)
run(64, 512, 512, -1, 0.0)
if len(sys.argv) > 1:
    run(8, 512, 512, 0, 0.0)
    run(64, 512, 200, 312, 0.0)
    run(64, 200, 512, 312, 0.0)
