import sys, os, math
sys.path.insert(0, ".")
import torch
import bench
from bpmult_b200 import MultiprojectionMMTransformer3DGMUClf, Trainer
from bpmult_b200 import _lib
def run(dbg, steps):
    lib = _lib.load() if hasattr(_lib, "load") else None
    from bpmult_b200.ops import CudaOps
    ops = CudaOps()
    ops.lib.bpm_debug_set(1, dbg)
    torch.manual_seed(1234)
    cfg = bench.make_config("cfg2")
    args = cfg["args"]
    dev = torch.device("cuda", 0)
    model = MultiprojectionMMTransformer3DGMUClf(args, precision="bf16").to(dev).train()
    tr = Trainer(model, lr=1e-3, seed=1234)
    host = [t.to(dev) for t in bench.synth_batch(cfg, 64, 2024)]
    out = []
    for i in range(steps):
        out.append(float(tr.step_device(*host)[0]))
    ops.lib.bpm_debug_set(1, 0)
    return out
a = run(0, 40)
print("default  :", " ".join("%.4f" % v for v in a[:8]), "...", " ".join("%.4f" % v for v in a[-4:]))
assert all(math.isfinite(v) for v in a) and a[-1] < a[0] * 0.5, "loss did not fall"
b = run(2048 + 8192 + 32768 + 4096, 8)     # no lse/delta fold, no ones column, 4 softmax warps, 8 compute warps
print("no-fold  :", " ".join("%.4f" % v for v in b))
print("max |diff| over 8 steps: %.2e" % max(abs(x - y) for x, y in zip(a, b)))
