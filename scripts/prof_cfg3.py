import sys
sys.path.insert(0, ".")
import runpy, torch
from argparse import Namespace
from bpmult_b200.model_engine4 import MMTrVaptEngine
from bpmult_b200.ops import CudaOps
B=8
cfg = Namespace(orig_d_l=768, orig_d_v=4096, orig_d_a=96, orig_d_p=4096, hidden_sz=768, num_heads=6, layers=5, vonly=True, lonly=True,
                aonly=True, attn_mask=True, hybrid=False, n_classes=13, attn_dropout=0.1, attn_dropout_v=0.0, attn_dropout_a=0.0,
                relu_dropout=0.1, res_dropout=0.1, out_dropout=0.0, embed_dropout=0.25)
ops = CudaOps(); dev = ops.device
eng = MMTrVaptEngine(ops, cfg, dtype=torch.bfloat16)
g = torch.Generator().manual_seed(1)
params = {k: (torch.randn(s, generator=g) * 0.02).to(dev) for k, s in eng.param_shapes().items()}
eng.pack(params)
txt = torch.randn(B, 512, 768, generator=g).to(dev); img = torch.randn(B, 200, 4096, generator=g).to(dev)
audio = torch.randn(B, 200, 96, generator=g).to(dev); poster = torch.randn(B, 4096, generator=g).to(dev)
tgt = (torch.rand(B, 13, generator=g) < 0.3).float().to(dev)
def step():
    logits, _ = eng.forward(txt, img, audio, poster, training=True, seed=3)
    loss, dl = eng.loss(logits, tgt, None); eng.zero_grads(); eng.backward(dl)
for _ in range(2): step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in ev:
    n = e.name.split("(")[0][:60]; a = agg.setdefault(n, [0.0, 0]); a[0] += e.device_time; a[1] += 1
tot = sum(v[0] for v in agg.values())
print("kernels %d total %.1f ms" % (len(ev), tot/1000))
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print("%-60s %8.2f ms %5.1f%% %5d x %8.1f us" % (n, t/1000, 100*t/tot, c, t/c))
