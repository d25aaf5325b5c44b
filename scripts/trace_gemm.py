"""Per-k-block event trace of CTA 0 of gemm_tc_kernel (diagnostics only).  Needs a library built with the trace points:
    make -C bpmult_b200/csrc clean && make -C bpmult_b200/csrc EXTRA=-DBPM_GEMM_TRACE
Prints, for one launch: load latency (slot acquired by the producer -> full barrier seen by the MMA issuer), issue time of a
k-block's MMAs, slot recycle time (MMAs issued -> the producer sees the slot empty again) and the k-block period."""
import sys

import torch

sys.path.insert(0, ".")
from bpmult_b200.engine import Dims
from bpmult_b200.ops import CudaOps

ops = CudaOps()
dev = ops.device
d = Dims(300, 12)
bf = torch.bfloat16
M = 64 * 512
case = sys.argv[1] if len(sys.argv) > 1 else "w1"
stages = int(sys.argv[2]) if len(sys.argv) > 2 else 7
if len(sys.argv) > 3:
    ops.lib.bpm_debug_set(0, int(sys.argv[3]))
from bpmult_b200.ops import Drop
shapes = {"w1": (d.FP, d.Dp, M, 1, 1, True), "wq": (d.HP, d.Dp, M, 1, 1, True), "w2": (d.Dp, d.FP, M, 1, 1, True),
          "fc1": (M, d.FP, d.Dp, 0, 0, False), "fc2": (M, d.Dp, d.FP, 0, 0, False), "fc1d": (M, d.FP, d.Dp, 0, 0, False)}
Mm, N, K, ta, tb, acc = shapes[case]
kw = dict(bias=torch.zeros(N, device=dev), act=1, drop=Drop(0.1, 1, None, 3)) if case == "fc1d" else {}
A = torch.randn((K, Mm) if ta else (Mm, K), device=dev).to(bf)
B = torch.randn((K, N) if tb else (N, K), device=dev).to(bf)
C = torch.zeros((Mm, N), device=dev, dtype=torch.float32 if acc else bf)
cs = torch.zeros(Mm, device=dev) if acc else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    ops.gemm(A, B, C, Mm, N, K, ta=ta, tb=tb, accumulate=acc, colsum=cs, **kw)
flush.zero_()
torch.cuda.synchronize()
tr = torch.zeros(3 * 1024, dtype=torch.int64, device=dev)
ops.lib.bpm_debug_set_ptr(tr.data_ptr())
ops.gemm(A, B, C, Mm, N, K, ta=ta, tb=tb, accumulate=acc, colsum=cs, **kw)
torch.cuda.synchronize()
ops.lib.bpm_debug_set_ptr(0)
tr = tr.cpu().view(3, 1024)
ev = [[(int(x) >> 8, int(x) & 255) for x in tr[r].tolist() if x != 0] for r in range(3)]
acq = [t for t, e in ev[0] if e == 1]
full = [t for t, e in ev[1] if e == 2]
done = [t for t, e in ev[1] if e == 3]
t0 = acq[0]
n = min(len(acq), len(full), len(done))
print("case %s: %d k-blocks traced, stages %d" % (case, n, stages))
names = {4: "acc full", 6: "tmem ld done", 7: "math done", 8: "staging free", 9: "stored", 5: "end"}
prev = None
for t, e in ev[2][:60]:
    print("   epilogue warp 0: %-13s @%7d (+%d)" % (names.get(e, e), t - t0, 0 if prev is None else t - prev))
    prev = t
print("%4s %9s %9s %9s %9s %9s" % ("kb", "acquired", "load lat", "mma issue", "recycle", "period"))
for i in range(n):
    rec = acq[i + stages] - done[i] if i + stages < len(acq) else -1
    print("%4d %9d %9d %9d %9d %9d" % (i, acq[i] - t0, full[i] - acq[i], done[i] - full[i], rec, full[i] - full[i - 1] if i else 0))
lat = [full[i] - acq[i] for i in range(stages, n)]
per = [full[i] - full[i - 1] for i in range(stages + 1, n)]
print("steady state: load latency mean %.0f clk, k-block period mean %.0f clk" % (sum(lat) / max(1, len(lat)), sum(per) / max(1, len(per))))
