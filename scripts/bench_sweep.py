"""BASELINE.json configs[4]: crossmodal attention + GMU microbenchmark sweep against the roofline.

    python scripts/bench_sweep.py [quick] [out.txt]                       one GPU
    python -m torch.distributed.run --nproc-per-node N scripts/bench_sweep.py ...   N independent replicas (no communication in the timed
                                                                          region; the slowest rank's time is reported, throughput x N)

Attention: source / target lengths 50 .. 4096, batch 8 .. 512, head dim 25 stored as 32 (12 heads, hidden 300) and head dim 128 (6 heads,
hidden 768), mask off / offset-causal, dropout 0 / 0.1; forward and forward+backward.  Algorithmic FLOPs: 4*B*H*T*S*dh*rho
forward, 2.5x that backward, rho = visible fraction of the score matrix.  Sequence GMU: rows x D for D in {300, 768} (3 GEMMs + the
combine kernel), against the HBM and tensor rooflines.  Times: CUDA events around a captured graph of launches (scripts/bench_attn.py)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bpmult_b200.engine import Dims, SeqGmuEngine  # noqa: E402
from bpmult_b200.ops import CudaOps, Drop  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ops = CudaOps(torch.device("cuda", local))
dev = ops.device
bf = torch.bfloat16
try:
    PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PK = {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
quick = "quick" in sys.argv
out_path = [a for a in sys.argv[1:] if a.endswith(".txt")]
lines = []


def emit(s):
    if rank == 0:
        print(s, flush=True)
        lines.append(s)


def timeit(fn, iters):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                fn()
        g.replay()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        g.replay()
        e1.record(st)
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000.0 / iters
    if world > 1:
        t = torch.tensor([us], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        us = float(t[0])
    return us


def attn_case(d, B, T, S, causal, p):
    M, Ms = B * T, B * S
    off = abs(S - T) if causal else -1
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
    vis = sum(min(S, i + off + 1) for i in range(T)) if causal else T * S
    fl = 4.0 * B * d.H * vis * d.dh
    iters = 3 if fl > 2e12 else 6
    tf = timeit(lambda: ops.xattn_fwd(q, k, v, o, lse, B, T, S, d.H, d.dh, d.dhp, mask_off=off, drop=drop, drop_bits=bits), iters)
    tc_bwd = T % 4 == 0                                      # (lse / delta rows travel as 16-byte bulk copies; other lengths: exact-fp32 kernel)
    if tc_bwd or fl < 3e11:
        tb = timeit(lambda: ops.xattn_bwd(q, k, v, o, do, lse, delta, dq, d.scaling, dk, dv, B, T, S, d.H, d.dh, d.dhp, mask_off=off, drop=drop,
                                          drop_bits=bits), iters)
        bwd = "%9.1f us %7.1f TF/s %5.1f%%%s" % (tb, 2.5 * fl / tb * 1e-6 * world, 100 * 2.5 * fl / tb * 1e-6 / PK["bf16_tflops"], "" if tc_bwd else " (fp32 kernel)")
    else:
        bwd = "      (fp32-math kernel for head dim 32 with T > 512: not timed at this size)"
    emit("dh=%3d B=%3d T=%4d S=%4d %-6s p=%.1f | fwd %9.1f us %7.1f TF/s %5.1f%% | bwd %s" % (
        d.dh, B, T, S, "causal" if causal else "full", p, tf, fl / tf * 1e-6 * world, 100 * fl / tf * 1e-6 / PK["bf16_tflops"], bwd))


def gmu_case(D, rows):
    eng = SeqGmuEngine(ops, D, bf, True)
    Dp = eng.Dp
    a1 = torch.randn(rows, Dp, device=dev).to(bf)
    a2 = torch.randn(rows, Dp, device=dev).to(bf)
    for w in eng.W.values():
        w.copy_(torch.randn_like(w.float()).to(bf) * 0.05)
    t = timeit(lambda: eng.forward(a1, a2, rows), 6)
    fl = 2.0 * rows * D * D * 4                               # hidden1, hidden2 and the gate over the concatenation (2D -> D)
    by = rows * D * 3 * 2 + 4 * D * D * 2                     # gates not written
    emit("seq-GMU D=%3d rows=%7d | fwd (4 GEMMs + combine) %8.1f us  %6.1f TF/s %4.1f%% of tensor peak  %6.0f GB/s algorithmic = %4.1f%% of HBM peak" % (
        D, rows, t, fl / t * 1e-6 * world, 100 * fl / t * 1e-6 / PK["bf16_tflops"], by / t * 1e-3 * world, 100 * by / t * 1e-3 / PK["hbm_gbs"]))


emit("# crossmodal attention + GMU sweep, %d GPU(s) (independent replicas), peaks: %.0f TFLOP/s bf16, %.0f GB/s HBM (MEASURED_PEAKS.json)" % (
    world, PK["bf16_tflops"], PK["hbm_gbs"]))
emit("# %% columns are per GPU against the burst bf16 peak; TF/s columns are the aggregate over the replicas")
d32, d128 = Dims(300, 12), Dims(768, 6)
lens = [50, 200, 512, 1024, 2048, 4096]
for d in (d32, d128):
    for T in lens:
        for B in ([8, 64] if quick else [8, 64, 512]):
            if B * T * T * d.H > 512 * 2048 * 2048 * 6:       # keep the largest points within seconds
                continue
            attn_case(d, B, T, T, True, 0.0)
            if T in (512, 2048) and B == 64:
                attn_case(d, B, T, T, False, 0.0)
                attn_case(d, B, T, T, True, 0.1)
    for (T, S) in ((512, 200), (200, 512), (4096, 512), (512, 4096)):
        attn_case(d, 64 if max(T, S) <= 512 else 16, T, S, True, 0.0)
for D in (300, 768):
    for rows in ([4096, 32768] if quick else [1600, 4096, 32768, 262144]):
        gmu_case(D, rows)
if rank == 0 and out_path:
    open(out_path[0], "w").write("\n".join(lines) + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
