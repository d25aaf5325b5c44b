"""Dumps the per-role event trace of CTA 0 of the attention backward kernel (diagnostics only).
Needs a library built with the trace points compiled in:  make -C bpmult_b200/csrc clean && make -C bpmult_b200/csrc EXTRA=-DBPM_ATTN_TRACE"""
import sys

import torch

sys.path.insert(0, ".")
from bpmult_b200.engine import Dims
from bpmult_b200.ops import CudaOps

ops = CudaOps()
dev = ops.device
d = Dims(300, 12)
bf = torch.bfloat16
B, T, S = 64, 512, 512
mask_off = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dbg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
M = B * T
q = torch.randn(M, d.HP, device=dev).to(bf) * 0.3
k = torch.randn(M, d.HP, device=dev).to(bf) * 0.3
v = torch.randn(M, d.HP, device=dev).to(bf)
o = torch.empty(M, d.HP, device=dev, dtype=bf)
lse = torch.empty(B * d.H * T, device=dev)
do = torch.randn(M, d.HP, device=dev).to(bf)
dq, dk, dv = [torch.empty(M, d.HP, device=dev, dtype=bf) for _ in range(3)]
delta = torch.empty(2 * B * d.H * T, device=dev)
ops.xattn_fwd(q, k, v, o, lse, B, T, S, d.H, d.dh, d.dhp, mask_off=mask_off)
ops.lib.bpm_debug_set(1, dbg)
for _ in range(2):
    ops.xattn_bwd(q, k, v, o, do, lse, delta, dq, d.scaling, dk, dv, B, T, S, d.H, d.dh, d.dhp, mask_off=mask_off)
torch.cuda.synchronize()
tr = torch.zeros(5 * 4096, dtype=torch.int64, device=dev)
ops.lib.bpm_debug_set_ptr(tr.data_ptr())
ops.xattn_bwd(q, k, v, o, do, lse, delta, dq, d.scaling, dk, dv, B, T, S, d.H, d.dh, d.dhp, mask_off=mask_off)
torch.cuda.synchronize()
ops.lib.bpm_debug_set_ptr(0)
tr = tr.cpu().view(5, 4096)
t0 = min(int(tr[r][0]) >> 8 for r in range(5) if int(tr[r][0]) != 0)
names = ["producer", "mma st", "mma acc", "compute h0", "compute h1"]
for r in range(5):
    ev = [(int(x) >> 8, int(x) & 255) for x in tr[r].tolist() if x != 0]
    print("== role", names[r], len(ev), "events")
    prev = None
    line = []
    for t, e in ev[:int(sys.argv[3]) if len(sys.argv) > 3 else 400]:
        line.append("%d@%d(+%d)" % (e, t - t0, 0 if prev is None else t - prev))
        prev = t
        if len(line) == 8:
            print("  " + "  ".join(line))
            line = []
    if line:
        print("  " + "  ".join(line))
