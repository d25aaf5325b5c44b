"""Times forward + BCE + backward of the 4-modality model (mmtrvapt) at the cfg-3 shape (Moviescope, B = 8 per GPU,
D = 768, H = 6 (head dim 128 -> exact-fp32-math attention kernels, no tensor-core attention yet), L = 5, lengths 512 / 200 / 200)."""
import sys
from argparse import Namespace

import torch

sys.path.insert(0, ".")
from bpmult_b200.model_engine4 import MMTrVaptEngine
from bpmult_b200.ops import CudaOps

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = Namespace(orig_d_l=768, orig_d_v=4096, orig_d_a=96, orig_d_p=4096, hidden_sz=768, num_heads=6, layers=5, vonly=True, lonly=True,
                aonly=True, attn_mask=True, hybrid=False, n_classes=13, attn_dropout=0.1, attn_dropout_v=0.0, attn_dropout_a=0.0,
                relu_dropout=0.1, res_dropout=0.1, out_dropout=0.0, embed_dropout=0.25)
ops = CudaOps()
dev = ops.device
eng = MMTrVaptEngine(ops, cfg, dtype=torch.bfloat16)
g = torch.Generator().manual_seed(1)
params = {k: (torch.randn(s, generator=g) * 0.02).to(dev) for k, s in eng.param_shapes().items()}
eng.pack(params)
txt = torch.randn(B, 512, 768, generator=g).to(dev)
img = torch.randn(B, 200, 4096, generator=g).to(dev)
audio = torch.randn(B, 200, 96, generator=g).to(dev)
poster = torch.randn(B, 4096, generator=g).to(dev)
tgt = (torch.rand(B, 13, generator=g) < 0.3).float().to(dev)


def step():
    logits, _ = eng.forward(txt, img, audio, poster, training=True, seed=3)
    loss, dl = eng.loss(logits, tgt, None)
    eng.zero_grads()
    eng.backward(dl)
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("mmtrvapt cfg 3  B=%d  fwd+bwd %.1f ms  (%.1f samples/s, eager launches, no optimizer)  loss %.4f" % (B, ms, B / ms * 1e3, float(loss)))

# full training step (Adam, CUDA graph) through the module API + Trainer
del eng
torch.cuda.empty_cache()
from bpmult_b200 import MultiprojectionMMTransformerGMUClf, Trainer  # noqa: E402
torch.manual_seed(1234)
model = MultiprojectionMMTransformerGMUClf(cfg, precision="bf16").to(dev).train()
tr = Trainer(model, lr=1e-4)
for _ in range(4):
    tr.step_device(txt, img, audio, poster, tgt)
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    tr.step_device(txt, img, audio, poster, tgt)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("mmtrvapt cfg 3  B=%d  train step (fwd + BCE + bwd + Adam, CUDA graph %s) %.1f ms  (%.1f samples/s)  params %.1f M  loss %.4f" % (
    B, tr.graph is not None, ms, B / ms * 1e3, tr.n_params / 1e6, float(tr.loss_dev[0])))
