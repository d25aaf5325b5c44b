#!/usr/bin/env python
"""bench.py -- train samples/s of the BPMulT fusion trunk on synthetic data (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                 # our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W         # the reference's own modules on the host cores (CPU)
    python bench.py --config cfg3|cfg4 ...                        # the other BASELINE.json configs (default cfg2 = configs[1])

  cfg2  mmtrvat,  CMU-MOSEI-unaligned shape: batch 64 per GPU; text 50x768, audio 500x74, vision 500x35 (zero-padded to 512 steps by
        the model); hidden 300, 12 heads, 8 layers; 6 classes                       (the configuration the metric is quoted on)
  cfg3  mmtrvapt, Moviescope shape: batch 8 per GPU (README.md:30); text 512x768, video 200x4096, audio 200x96 (post-encoder
        features), poster 4096; hidden 768, 6 heads (head dim 128), 5 layers; 13 classes
  cfg4  mmtrvapt, MM-IMDb bimodal shape: batch 6 per GPU (README.md:36); text 512x768, "video" = GloVe plot 200x300, "audio" = BoW
        200x1, poster 4096; hidden 768, 6 heads, 5 layers; 23 classes

A "step" = forward + BCEWithLogits + backward + gradient all-reduce (N > 1) + Adam on one batch per GPU, README dropouts.  `value`
times K steps with the batch already resident in HBM; `e2e` times K steps through the public `Trainer.step_async()` with HOST tensors
(H2D of every step's batch on a copy stream, D2H + host read of every step's loss one step later; the fully blocking `Trainer.step()`
figure is reported beside it as `e2e.blocking`).  Timing: CUDA events, barrier + synchronize on both sides, max over
ranks.  Prints ONE JSON line on rank 0."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from argparse import Namespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

README_DROPOUT = dict(attn_dropout=0.1, attn_dropout_v=0.0, attn_dropout_a=0.0, relu_dropout=0.1, res_dropout=0.1, out_dropout=0.0,
                      embed_dropout=0.25)                      # train.py:86-92 defaults (README.md:30,36,43 leave them unset)
NV4 = {"l": 512, "a": 200, "v": 200}                            # mmtr.py:371-373


def make_config(name, layers=None, batch=None):
    base = dict(vonly=True, lonly=True, aonly=True, attn_mask=True, hybrid=False, bert_model="none", orig_d_p=4096, **README_DROPOUT)
    if name == "cfg2":
        a = Namespace(model="mmtrvat", orig_d_l=768, orig_d_v=35, orig_d_a=74, hidden_sz=300, num_heads=12, layers=layers or 8, n_classes=6, **base)
        return dict(name=name, args=a, B=batch or 64, lens=dict(l=50, a=500, v=500), padded=dict(l=512, a=512, v=512),
                    workload="BPMulT mmtrvat train step, synthetic CMU-MOSEI-unaligned shape (BASELINE configs[1])")
    if name == "cfg3":
        a = Namespace(model="mmtrvapt", orig_d_l=768, orig_d_v=4096, orig_d_a=96, hidden_sz=768, num_heads=6, layers=layers or 5, n_classes=13, **base)
        return dict(name=name, args=a, B=batch or 8, lens=dict(NV4), padded=dict(NV4),
                    workload="BPMulT mmtrvapt train step, synthetic Moviescope shape (BASELINE configs[2])")
    if name == "cfg4":
        a = Namespace(model="mmtrvapt", orig_d_l=768, orig_d_v=300, orig_d_a=1, hidden_sz=768, num_heads=6, layers=layers or 5, n_classes=23, **base)
        return dict(name=name, args=a, B=batch or 6, lens=dict(NV4), padded=dict(NV4),
                    workload="BPMulT mmtrvapt train step, synthetic MM-IMDb bimodal shape: text + poster, GloVe plot as video, BoW as audio (BASELINE configs[3])")
    raise ValueError(name)


def synth_batch(cfg, B, seed, pad=False):
    """(txt, img, audio[, poster], targets): N(0,1) features, Bernoulli(0.3) targets.  pad=True appends the zero time steps the
    model would append itself (mmtr.py:722-732,756-761) -- the same computation; used for the CPU reference, whose own padding code
    calls .cuda() unconditionally."""
    a, L = cfg["args"], cfg["lens"]
    g = torch.Generator().manual_seed(seed)
    txt = torch.randn(B, L["l"], a.orig_d_l, generator=g)
    img = torch.randn(B, L["v"], a.orig_d_v, generator=g)
    audio = torch.randn(B, L["a"], a.orig_d_a, generator=g)
    out = [txt, img, audio]
    if pad:
        P = cfg["padded"]
        out = [torch.cat([t, torch.zeros(B, n - t.shape[1], t.shape[2])], 1) if t.shape[1] < n else t for t, n in zip(out, (P["l"], P["v"], P["a"]))]
    if a.model == "mmtrvapt":
        out.append(torch.randn(B, a.orig_d_p, generator=g))
    out.append((torch.rand(B, a.n_classes, generator=g) < 0.3).float())
    return out


def flops_per_sample_train(cfg):
    """SURVEY 8d algorithmic FLOPs (unpadded dims, multiply-add = 2, dense attention), x3 for forward + backward."""
    a = cfg["args"]
    D, L = a.hidden_sz, a.layers
    layer = lambda T, S: D * D * (20 * T + 4 * S) + 4 * T * S * D
    if a.model == "mmtrvat":
        T = 512
        return 3 * (12 * L * layer(T, T) + 6 * 2 * T * D * (4 * D))
    nv = NV4
    fl = 0
    for q, k in (("v", "a"), ("a", "v"), ("v", "l"), ("l", "v"), ("a", "l"), ("l", "a")):          # wave 1
        fl += L * layer(nv[q], nv[k])
    for q, src_q in (("l", "v"), ("l", "a"), ("a", "l"), ("a", "v"), ("v", "l"), ("v", "a")):      # wave 2: K/V = a wave-1 output of length nv[src_q]
        T = nv[q]
        fl += L * (layer(T, nv[src_q]) + 8 * T * D * D + 4 * T * T * D)                            # + the biprojection self-attention block
    fl += sum(16 * nv[m] * D * D for m in "lav")                                                   # 6 sequence GMUs
    fl += 2 * (2 * 200 * 512 * D) * 2                                                              # time-axis linears
    fl += sum(2 * nv[m] * o * D for m, o in (("l", a.orig_d_l), ("a", a.orig_d_a), ("v", a.orig_d_v)) if o != D) + 2 * a.orig_d_p * D
    return 3 * fl


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        clk = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": clk[len(clk) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(clk)}


# ------------------------------------------------------------------------------------------------ reference / CPU arm
def cpu_step_fn(cfg, B, threads):
    """fwd + BCE + bwd of the reference's own modules on the host, README dropouts live, train() mode: the shimmed reference
    (oracle/ref_shim.py: $BPMULT_REF, /root/reference or baseline/_ref -- scripts/install_reference.sh fills the latter so that it
    travels to the GPU box); only when no reference tree exists at all, the oracle port (dropout-free).  Returns (callable, kind)."""
    from oracle import functional as Fn
    from oracle import synth
    from oracle.ref_shim import load_reference
    torch.set_num_threads(threads)
    a = Namespace(**vars(cfg["args"]))
    four = a.model == "mmtrvapt"
    batch = synth_batch(cfg, B, 2024, pad=True)
    feats, tgt = batch[:-1], batch[-1]
    ref = load_reference()
    if ref is not None:
        torch.manual_seed(1234)
        cls = ref.mmtr.MultiprojectionMMTransformerGMUClf if four else ref.mmtr.MultiprojectionMMTransformer3DGMUClf
        model = cls(a)
        model.train()
        crit = torch.nn.BCEWithLogitsLoss()

        def step():
            model.zero_grad()
            out = model(feats[0], None, None, feats[1], feats[2], *feats[3:])
            loss = crit(out, tgt)
            loss.backward()
            return float(loss)
        return step, "reference"
    for k in README_DROPOUT:
        setattr(a, k, 0.0)
    shapes = synth.mmtrvapt_shapes(a) if four else synth.mmtrvat_shapes(a)
    sd = {k: v.requires_grad_() for k, v in synth.make_state_dict(shapes, 1234).items()}
    fwd = Fn.mmtrvapt_forward if four else Fn.mmtrvat_forward

    def step():
        for v in sd.values():
            v.grad = None
        logits, _ = fwd(sd, a, *feats)
        loss = Fn.bce_with_logits(logits, tgt)
        loss.backward()
        return float(loss)
    return step, "port"


def cpu_sample_text(cfg, B, kind):
    return "B=%d per step of the %s workload (fwd+BCE+bwd, fp32, train mode, %s)" % (
        B, cfg["name"], "unmodified reference modules (shimmed import), README dropouts live" if kind == "reference" else "oracle port, dropout-free")


def run_reference(opt):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = make_config(opt.config)
    threads = os.cpu_count() or 1
    # bounded sample: the per-step batch is sized from one probe step so that warm-up + K steps stay within ~3 minutes
    step, kind = cpu_step_fn(cfg, 1, threads)
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0
    n_steps = max(1, opt.steps + opt.warmup)
    B = 1
    while B < 8 and 2 * B * t1 * n_steps < 170.0:
        B *= 2
    if B > 1:
        step, kind = cpu_step_fn(cfg, B, threads)
    for _ in range(opt.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(opt.steps):
        step()
    dt = time.perf_counter() - t0
    v = B * opt.steps / dt
    print(json.dumps({"impl": "reference", "metric": "train samples/s", "value": v, "unit": "samples/s", "n_gpus": opt.gpus, "steps": opt.steps,
                      "warmup": opt.warmup, "ms_per_step": 1e3 * dt / opt.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": workload_config(cfg, cfg["B"], max(1, opt.gpus)),     # (our arm's config; the sample is below)
                      "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": kind, "sample": cpu_sample_text(cfg, B, kind)},
                      "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(cfg, B, world):
    a, L = cfg["args"], cfg["lens"]
    return {"workload": cfg["workload"], "name": cfg["name"], "batch_per_gpu": B, "global_batch": B * world,
            "text": "%dx%d" % (L["l"], a.orig_d_l), "audio": "%dx%d" % (L["a"], a.orig_d_a), "vision": "%dx%d" % (L["v"], a.orig_d_v),
            "poster": a.orig_d_p if a.model == "mmtrvapt" else None, "padded_len": cfg["padded"], "hidden": a.hidden_sz, "heads": a.num_heads,
            "layers": a.layers, "classes": a.n_classes, "parallelism": "dp%d" % world, "dropout": "README (embed .25, attn .1/0/0, relu .1, res .1)",
            "optimizer": "Adam lr 1e-3", "l2": "working set per step >> 126 MB L2 (GBs of activations): no flush needed"}


# ------------------------------------------------------------------------------------------------ isolated-kernel rooflines
def time_kernel(fn, n_rot, iters=20, warm=3):
    """Average launch duration from CUDA events around a captured CUDA graph of `iters` launches (so the Python / ctypes launch
    cost is not in the number).  `fn(i)` launches the kernel on buffer set i % n_rot: rotating operand sets larger than the
    126 MB L2 keep every launch cold, like inside the training step."""
    for i in range(warm):
        fn(i % n_rot)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i % n_rot)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


MUFU_PER_CLK_SM = 16          # ex2 throughput of one B200 SM (guides/B300_MICROARCH: B300 has 2x this)


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of these kernels at these shapes, written by
    scripts/ncu_traffic.py (which records the commit it profiled); absent => null."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return {}


def kernel_rooflines(ops, cfg, B, peaks):
    """Times the kernel families of one encoder layer in isolation at this config's shape and returns, per family, the achieved
    rate against its bound plus `launches_per_step` so that the dominant family can be named from measured time."""
    from bpmult_b200.engine import Dims
    from bpmult_b200.ops import Drop
    a = cfg["args"]
    d = Dims(a.hidden_sz, a.num_heads)
    D, L = a.hidden_sz, a.layers
    four = a.model == "mmtrvapt"
    T = 512
    M = B * T
    bf, f32 = torch.bfloat16, torch.float32
    dev = ops.device
    R = 4 if M * d.FP * 2 * 4 > 150e6 else 8
    out = {}
    n_layers = 12 * L
    # layer executions whose query stream has T = 512 rows per sample (all of them for mmtrvat; the text-target encoders of mmtrvapt)
    n512 = n_layers if not four else 4 * L
    tens = lambda fl, t, note, n: {"bound": "tensor", "achieved": fl / t / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                   "frac": fl / t / 1e12 / peaks["bf16_tflops"], "us": t * 1e6, "traffic": None, "note": note, "launches_per_step": n}
    hbm = lambda by, t, note, n: {"bound": "hbm", "achieved": by / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / t / 1e9 / peaks["hbm_gbs"],
                                  "us": t * 1e6, "traffic": None, "note": note, "launches_per_step": n}
    # ---- GEMMs: fc1 forward (bias + ReLU + dropout), fc2 forward (dropout + fp32 residual), fc1 weight gradient, Wq weight gradient
    A = [torch.randn(M, d.Dp, device=dev).to(bf) for _ in range(R)]
    Hs = [torch.randn(M, d.FP, device=dev).to(bf) for _ in range(R)]
    W1 = torch.randn(d.FP, d.Dp, device=dev).to(bf)
    W2 = torch.randn(d.Dp, d.FP, device=dev).to(bf)
    b1, b2 = torch.zeros(d.FP, device=dev), torch.zeros(d.Dp, device=dev)
    X32 = [torch.randn(M, d.Dp, device=dev) for _ in range(R)]
    t = time_kernel(lambda i: ops.gemm(A[i], W1, Hs[i], M, d.FP, d.Dp, bias=b1, act=1, drop=Drop(0.1, 1, None, 3)), R)
    out["gemm_fc1"] = tens(2.0 * M * D * 4 * D, t, "fc1 forward: algorithmic 2*M*D*4D, M=%d D=%d (executed on padded %dx%d)" % (M, D, d.Dp, d.FP), n512)
    t = time_kernel(lambda i: ops.gemm(Hs[i], W2, X32[i], M, d.Dp, d.FP, bias=b2, drop=Drop(0.1, 1, None, 4), residual=X32[i]), R)
    out["gemm_fc2"] = tens(2.0 * M * D * 4 * D, t, "fc2 forward with dropout + fp32 residual epilogue", n512)
    G1 = torch.zeros(d.FP, d.Dp, device=dev)
    gb1 = torch.zeros(d.FP, device=dev)
    t = time_kernel(lambda i: ops.gemm(Hs[i], A[i], G1, d.FP, d.Dp, M, ta=1, tb=1, accumulate=True, colsum=gb1), R)
    out["wgrad_fc1"] = tens(2.0 * M * D * 4 * D, t, "dW1 = dH^T X (+ bias gradient), K = M = %d, split-K fp32 reduce-add" % M, 2 * n512)
    Q = [torch.randn(M, d.HP, device=dev).to(bf) for _ in range(R)]
    Gq = torch.zeros(d.HP, d.Dp, device=dev)
    gbq = torch.zeros(d.HP, device=dev)
    t = time_kernel(lambda i: ops.gemm(Q[i], A[i], Gq, d.HP, d.Dp, M, ta=1, tb=1, accumulate=True, colsum=gbq), R)
    out["wgrad_qproj"] = tens(2.0 * M * D * D, t, "dWq = dQ^T X (+ bias gradient); the out-proj weight gradient has the same shape", 2 * n512)
    Wq = torch.randn(d.HP, d.Dp, device=dev).to(bf)
    bq = torch.zeros(d.HP, device=dev)
    t = time_kernel(lambda i: ops.gemm(A[i], Wq, Q[i], M, d.HP, d.Dp, bias=bq, alpha=d.scaling), R)
    out["gemm_qproj"] = tens(2.0 * M * D * D, t, "q projection (K = N = D); out-proj / their dgrads have the same shape", 4 * n512)
    del Hs, A
    # ---- attention forward + backward
    q = [torch.randn(M, d.HP, device=dev).to(bf) * 0.3 for _ in range(R)]
    k = [torch.randn(M, d.HP, device=dev).to(bf) * 0.3 for _ in range(R)]
    v = [torch.randn(M, d.HP, device=dev).to(bf) for _ in range(R)]
    o = torch.empty(M, d.HP, device=dev, dtype=bf)
    lse = torch.empty(B * d.H * T, device=dev)
    t = time_kernel(lambda i: ops.xattn_fwd(q[i], k[i], v[i], o, lse, B, T, T, d.H, d.dh, d.dhp, mask_off=0), R, iters=5, warm=2)
    rho = (T + 1) / (2.0 * T)
    fl = 4.0 * B * d.H * T * T * d.dh * rho
    n_attn = n512 if not four else 2 * L + 6 * L                              # mmtrvapt: 512x512 self-attention of the text-target biprojection layers
    out["xattn_fwd"] = tens(fl, t, "algorithmic 4*B*H*T*S*dh*rho, T=S=512, rho=(T+1)/2T causal, dh=%d (stored %d)" % (d.dh, d.dhp), n_attn)
    do = torch.randn(M, d.HP, device=dev).to(bf)
    dq, dk, dv = [torch.empty(M, d.HP, device=dev, dtype=bf) for _ in range(3)]
    delta = torch.empty(ops.xattn_bwd_workspace(bf, B, T, T, d.H, d.dh, d.dhp), device=dev)
    ops.xattn_fwd(q[0], k[0], v[0], o, lse, B, T, T, d.H, d.dh, d.dhp, mask_off=0)
    t = time_kernel(lambda i: ops.xattn_bwd(q[0], k[0], v[0], o, do, lse, delta, dq, d.scaling, dk, dv, B, T, T, d.H, d.dh, d.dhp, mask_off=0), 1, iters=3, warm=1)
    out["xattn_bwd"] = tens(2.5 * fl, t, "2.5x forward FLOPs (incl. its delta / dQ-cast helper kernels)", n_attn)
    del q, k, v
    # ---- LayerNorm forward / backward (fp32 residual stream in, bf16 out)
    ys = [torch.empty(M, d.Dp, device=dev, dtype=bf) for _ in range(R)]
    gam, bet = torch.ones(d.Dp, device=dev), torch.zeros(d.Dp, device=dev)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    t = time_kernel(lambda i: ops.layernorm_fwd(X32[i], gam, bet, D, ys[i], mean, rstd), R)
    out["layernorm_fwd"] = hbm(M * D * (4 + 2), t, "algorithmic rows*D*(4+2) B, rows=%d D=%d" % (M, D), 2 * n512)
    ops.layernorm_fwd(X32[0], gam, bet, D, ys[0], mean, rstd)
    dxs = [torch.zeros(M, d.Dp, device=dev) for _ in range(R)]
    dg, db = torch.zeros(d.Dp, device=dev), torch.zeros(d.Dp, device=dev)
    t = time_kernel(lambda i: ops.layernorm_bwd(ys[i], X32[i], mean, rstd, gam, D, dxs[i], True, dg, db, cast_out=ys[(i + 1) % R], cast_drop=Drop(0.1, 1, None, 5)), R)
    out["layernorm_bwd"] = hbm(M * D * (2 + 4 + 8 + 2), t, "rows*D*(dy 2 + x 4 + dx read+write 8 + cast out 2) B", 2 * n512)
    # ---- sequence GMU combine (tanh / sigmoid / gating + addend), and input staging + k=1 projection of the narrowest stream
    hs = [[torch.randn(M, d.Dp, device=dev).to(bf) for _ in range(6)] for _ in range(max(2, R // 2))]
    yg = torch.empty(M, d.Dp, device=dev, dtype=bf)
    t = time_kernel(lambda i: ops.gmu_fwd(1, hs[i][0], hs[i][1], hs[i][2], hs[i][3], hs[i][4], hs[i][5], yg), len(hs))
    out["gmu_combine"] = hbm(M * D * 7 * 2, t, "rows*D*(6 inputs + 1 output)*2 B (gate tensors are not written)", 6 if not four else 2)
    del hs
    Tin, Cin = (500, a.orig_d_v) if not four else (200, a.orig_d_v)
    Kp = (Cin + 63) // 64 * 64
    src = [torch.randn(B, Tin, Cin, device=dev) for _ in range(R)]
    Xs = torch.empty(B * (512 if not four else 200), Kp, device=dev, dtype=bf)
    t = time_kernel(lambda i: ops.stage_rows(src[i], Xs, 512 if not four else 200), R)
    out["stage_rows"] = hbm(B * Tin * Cin * 4 + Xs.numel() * 2, t, "input staging (transpose / zero-pad / cast) of the vision stream: fp32 in + bf16 out", 3)
    traffic = load_traffic()
    for kname, v_ in out.items():
        ent = traffic.get(cfg["name"], {}).get(kname)
        if ent:
            v_["traffic"], v_["traffic_src"] = ent["dram_bytes"], ent.get("src")
    # the attention kernels at head dim 32 are bounded by exponentials, not by the tensor pipe: report that bound beside the tensor one
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    mufu_peak = MUFU_PER_CLK_SM * sms * 1.965e9
    n_exp = B * d.H * T * T * rho
    for kname in ("xattn_fwd", "xattn_bwd"):
        tt = out[kname]["us"] * 1e-6
        out[kname]["exp_bound"] = {"achieved": n_exp / tt / 1e12, "peak": mufu_peak / 1e12, "unit": "Texp/s", "frac": n_exp / tt / mufu_peak,
                                   "note": "one ex2 per unmasked score; peak = 16 /clk/SM x %d SMs x 1.965 GHz" % sms}
    return out


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(opt):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from bpmult_b200 import MultiprojectionMMTransformer3DGMUClf, MultiprojectionMMTransformerGMUClf, Trainer
    cfg = make_config(opt.config, opt.layers, opt.batch)
    args, B = cfg["args"], cfg["B"]
    if opt.prune:
        os.environ["BPM_PRUNE"] = "1"
    torch.manual_seed(1234)                                   # reference default seed (train.py:61)
    cls = MultiprojectionMMTransformerGMUClf if args.model == "mmtrvapt" else MultiprojectionMMTransformer3DGMUClf
    model = cls(args, precision=opt.precision).to(dev)
    model.train()
    tr = Trainer(model, lr=1e-3, seed=1234)
    host = [t.pin_memory() for t in synth_batch(cfg, B, 2024 + rank)]     # as a DataLoader(pin_memory=True) hands them over
    devb = [t.to(dev) for t in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, fin=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if fin is not None:
            fin()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0])

    for _ in range(max(opt.warmup, 3)):                       # >= 3 warm-up steps (2 eager + graph capture + replay)
        tr.step_device(*devb)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(lambda: tr.step_device(*devb), opt.steps)
    loss_dev = float(tr.loss_dev[0])
    # e2e: public API with host tensors, H2D + D2H inside the timed region
    tr.step(*host)
    ms_blk = timed(lambda: tr.step(*host), opt.steps)           # blocking API: the loss of every step is read before the next is enqueued
    # pipelined API (what a training loop does): step k + 1 is enqueued -- its H2D runs on a copy stream under step k -- and then the loss
    # of step k is read; every step's inputs cross PCIe and every step's loss is read on the host inside the timed region
    pend = []

    def pipelined():
        pend.append(tr.step_async(*host))
        if len(pend) > 1:
            pend.pop(0).item()
    ms_e2e = timed(pipelined, opt.steps, fin=lambda: [h.item() for h in pend])
    pend.clear()
    ms_again = timed(lambda: tr.step_device(*devb), opt.steps)   # the device-resident loop once more: separates clock drift from copy cost
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    line = None
    if rank == 0:
        value = world * B * opt.steps / (ms * 1e-3)
        e2e = world * B * opt.steps / (ms_e2e * 1e-3)
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
        try:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            peaks = {"hbm_gbs": pk["hbm_gbs"], "bf16_tflops": pk["bf16_tflops"], "bf16_tflops_sustained": pk["bf16_tflops_sustained"], "src": "measured"}
        except Exception:
            pass
        line = {"metric": "train samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": opt.steps, "warmup": max(opt.warmup, 3),
                "ms_per_step": ms / opt.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if opt.precision == "bf16" else "f32", "data": "synthetic", "config": workload_config(cfg, B, world),
                "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": tr.bytes_in(), "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / opt.steps,
                        "api": "Trainer.step_async(pinned host tensors) -> handle, handle.item() one step later: H2D of every step's batch (copy "
                               "stream, under the previous step) and D2H + host read of every step's loss inside the timed region",
                        "blocking": {"value": world * B * opt.steps / (ms_blk * 1e-3), "ms_per_step": ms_blk / opt.steps,
                                     "api": "Trainer.step(pinned host tensors) -> float, nothing overlapped"}},
                "gpu_launches": int(getattr(tr, "launches_per_step", 0)) * opt.steps, "launches_per_step": int(getattr(tr, "launches_per_step", 0)),
                "pruned_rows": bool(getattr(tr.eng, "prune", False)),
                "cuda_graph": bool(tr.use_graph), "loss": loss_dev, "params": tr.n_params, "ms_per_step_after_e2e": ms_again / opt.steps,
                "clocks": sampler.summary() if sampler else None}
        fl = flops_per_sample_train(cfg) * B
        line["step_tflops_algorithmic"] = fl / (ms / opt.steps * 1e-3) / 1e12
        line["step_frac_of_bf16_sustained_peak"] = line["step_tflops_algorithmic"] / peaks["bf16_tflops_sustained"]
        line["peaks"] = peaks
        if not opt.no_kernels:
            ks = kernel_rooflines(tr.ops, cfg, B, peaks)
            step_us = ms / opt.steps * 1e3
            for v_ in ks.values():
                v_["share_of_step_est"] = v_["launches_per_step"] * v_["us"] / step_us
            dom = max(ks, key=lambda k_: ks[k_]["share_of_step_est"])          # the family with the largest measured share of the step
            line["roofline"] = dict(ks[dom], kernel=dom)
            line["kernels"] = ks
        if world == 1 and not opt.no_cpu:
            threads = os.cpu_count() or 1
            step, kind = cpu_step_fn(cfg, 1, threads)
            step()
            t0 = time.perf_counter()
            n = 2
            for _ in range(n):
                step()
            dt = (time.perf_counter() - t0) / n
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "samples/s", "cores": threads, "kind": kind,
                                    "sample": "1 warm-up + %d timed steps: " % n + cpu_sample_text(cfg, 1, kind)}
        print(json.dumps(line))
        sys.stdout.flush()
    # orderly teardown: drop the captured graphs (they hold the NCCL kernels) before the communicator goes away
    tr.close()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4"])
    ap.add_argument("--prune", action="store_true",
                    help="mmtrvat only, NOT the headline: wave-2 query side and gated units on the two time steps that reach the head "
                         "(identical logits and gradients; the reference computes all rows, and so does the default)")
    ap.add_argument("--batch", type=int, default=0, help="samples per GPU per step (0 = the config's own: 64 / 8 / 6)")
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-kernels", action="store_true", help="skip the isolated-kernel roofline timings")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    opt = ap.parse_args()
    if opt.impl == "reference":
        return run_reference(opt)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if opt.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(opt.gpus), "--master-addr", "127.0.0.1",
               "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(opt)


if __name__ == "__main__":
    main()
