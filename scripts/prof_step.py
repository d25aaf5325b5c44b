"""Per-kernel time breakdown of one training step with torch.profiler (CUPTI kernel durations, no replay / serialisation):
a cheap complement to the ncu launch list (scripts/gpu_profile.sh)."""
import os
import sys

import torch

sys.path.insert(0, ".")
if os.environ.get("PROF_GRAPH", "1") != "1":            # default: profile the replayed CUDA graph (what bench.py times)
    os.environ["BPM_NO_GRAPH"] = "1"
import bench  # noqa: E402
from bpmult_b200 import MultiprojectionMMTransformer3DGMUClf  # noqa: E402
from bpmult_b200 import Trainer  # noqa: E402

torch.manual_seed(0)
cfg = bench.make_config(os.environ.get("PROF_CFG", "cfg2"))
args = cfg["args"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
model = MultiprojectionMMTransformer3DGMUClf(args, precision="bf16").to(dev)
model.train()
tr = Trainer(model, lr=1e-3)
g = torch.Generator().manual_seed(2024)
host = list(bench.synth_batch(cfg, B, 2024))
devb = [t.to(dev) for t in host]
for _ in range(5):
    tr.step_device(*devb)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step_device(*devb)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in ev:
    n = e.name.split("(")[0][:60]
    a = agg.setdefault(n, [0.0, 0])
    a[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    a[1] += 1
tot = sum(v[0] for v in agg.values())
print("kernels: %d   summed device time: %.2f ms" % (len(ev), tot / 1000))
for n, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print("%-62s %9.3f ms %5.1f%% %6d launches %8.2f us" % (n, t / 1000, 100 * t / tot, c, t / c))

# ---- concurrency: how much of the step has 1, 2, ... kernels in flight, and which kernels run ALONE (serialisation points)
iv = sorted((e.time_range.start, e.time_range.end, e.name.split("(")[0][:60]) for e in ev)
pts = []
for s, t, n in iv:
    pts.append((s, 1, n))
    pts.append((t, -1, n))
pts.sort(key=lambda p: (p[0], p[1]))
active, last, hist, alone = {}, None, {}, {}
for t, d, n in pts:
    k = sum(active.values())
    if last is not None and t > last and k > 0:
        hist[k] = hist.get(k, 0.0) + (t - last)
        if k == 1:
            nm = next(a for a, c in active.items() if c > 0)
            alone[nm] = alone.get(nm, 0.0) + (t - last)
    active[n] = active.get(n, 0) + d
    last = t
span = iv[-1][1] - iv[0][0] if iv else 0
print("\nstep span %.2f ms; kernels in flight -> ms: %s; idle %.2f ms" % (
    span / 1000, {k: round(v / 1000, 2) for k, v in sorted(hist.items())}, (span - sum(hist.values())) / 1000))
print("time with exactly ONE kernel in flight, by kernel:")
for n, t in sorted(alone.items(), key=lambda kv: -kv[1])[:14]:
    print("   %-60s %8.3f ms" % (n, t / 1000))

# ---- largest idle gaps (no kernel in flight) and what surrounds them
ends = []
cur_end, gaps = iv[0][1], []
last_name = iv[0][2]
for s, t, n in iv[1:]:
    if s > cur_end:
        gaps.append((s - cur_end, last_name, n, cur_end - iv[0][0]))
    if t > cur_end:
        cur_end, last_name = t, n
print("idle gaps > 20 us: %d, total %.2f ms" % (sum(1 for g in gaps if g[0] > 20), sum(g[0] for g in gaps if g[0] > 20) / 1000))
for g in sorted(gaps, reverse=True)[:14]:
    print("   %7.1f us at t=%8.2f ms   after %-40s before %s" % (g[0], g[3] / 1000, g[1][:40], g[2][:40]))
