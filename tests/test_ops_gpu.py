"""GPU parity tests, op level: every C-ABI kernel (called through bpmult_b200.ops.CudaOps) against the pure-torch statement
of the same contract (tests/emu_ops.py), on seeded inputs, in fp32 and bf16 storage.
Tolerances: fp32 paths 2e-5 (max-rel per tensor); bf16-storage paths 1e-2 (the stated bf16 bar), elementwise-exact where
both sides round the same fp32 value to bf16 (<= 1 bf16 ulp)."""
import pytest
import torch

from emu_ops import EmuOps
from bpmult_b200.ops import Drop

pytestmark = pytest.mark.gpu
F32, BF16 = torch.float32, torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from bpmult_b200.ops import CudaOps
    return CudaOps()


def rnd(shape, seed, dtype=F32, scale=1.0):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).to(dtype)


def max_rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def tol(dtype):
    return 2e-5 if dtype == F32 else 1e-2


def both(ops, fn, ins, outs):
    """run `fn(o, *tensors)` with EmuOps on CPU copies and CudaOps on device copies; returns (emu outs, cuda outs)"""
    emu = EmuOps()
    ci = [None if t is None else t.clone() for t in ins]
    co = [None if t is None else t.clone() for t in outs]
    fn(emu, *ci, *co)
    gi = [None if t is None else t.clone().cuda() for t in ins]
    go = [None if t is None else t.clone().cuda() for t in outs]
    fn(ops, *gi, *go)
    torch.cuda.synchronize()
    return co, [None if t is None else t.cpu() for t in go]


DROP = Drop(0.3, 1234567, None, 77)


# ------------------------------------------------------------------------------------------------ dropout bit-exactness
def test_philox_mask_matches_emulation(ops):
    x = torch.ones(37, 64)
    (e,), (c,) = both(ops, lambda o, x, y: o.cast_drop(x, y, DROP), [x], [torch.zeros(37, 64)])
    assert torch.equal(e, c)
    keep = (c != 0).float().mean().item()
    assert abs(keep - 0.7) < 0.05
    assert torch.allclose(c[c != 0], torch.tensor(1 / 0.7))


# ------------------------------------------------------------------------------------------------ pack / stage / embed
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_pack_unpack(ops, dtype):
    src = rnd((300, 300), 1)
    (e,), (c,) = both(ops, lambda o, s, d: o.pack_matrix(s, d, row_map=(25, 32), col_map=(0, 0)), [src], [torch.zeros(384, 320, dtype=dtype)])
    assert torch.equal(e, c)
    (e,), (c,) = both(ops, lambda o, s, d: o.pack_matrix(s, d, col_map=(25, 32)), [src], [torch.zeros(320, 384, dtype=dtype)])
    assert torch.equal(e, c)
    g = rnd((384, 320), 2)
    (e,), (c,) = both(ops, lambda o, s, d: o.unpack_matrix(s, d, row_map=(25, 32), accumulate=True, scale=0.5), [g], [rnd((300, 300), 3)])
    assert torch.allclose(e, c, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_batched_pack_unpack_equals_single_calls(ops, dtype):
    """one launch over tensors of very different sizes (work units by element count; vector path where both sides are 16-byte aligned,
    scalar path for head remaps in the columns, odd widths and misaligned sources) == the per-tensor kernels, bit for bit"""
    dev = ops.device
    flat = rnd((1 << 20,), 7).to(dev)                                    # parameters live at arbitrary offsets of one flat buffer
    specs = [  # (rows, cols, offset in flat, padded rows, padded cols, row_map, col_map)
        (900, 300, 0, 1152, 320, (25, 32), (0, 0)), (300, 300, 270000, 320, 384, (0, 0), (25, 32)), (1200, 300, 360000, 1216, 320, (0, 0), (0, 0)),
        (1, 300, 720000, 1, 320, (0, 0), (0, 0)), (6, 300, 720301, 8, 320, (0, 0), (0, 0)), (300, 35, 722200, 320, 64, (0, 0), (0, 0)),
        (1, 6, 732706, 1, 8, (0, 0), (0, 0)), (768, 200, 733000, 768, 256, (0, 0), (0, 0))]
    srcs = [flat[o:o + r * c].view(r, c) for r, c, o, *_ in specs]
    single = [torch.full((rp, cp), 3.0, dtype=dtype, device=dev) for _, _, _, rp, cp, *_ in specs]
    batched = [t.clone() for t in single]
    for s_, d_, (_, _, _, _, _, rm, cm) in zip(srcs, single, specs):
        ops.pack_matrix(s_, d_, row_map=rm, col_map=cm)
    ops.batch_begin("pack", "t")
    for s_, d_, (_, _, _, _, _, rm, cm) in zip(srcs, batched, specs):
        ops.pack_matrix(s_, d_, row_map=rm, col_map=cm)
    ops.batch_end()
    for a, b in zip(single, batched):
        assert torch.equal(a, b)
    # unpack (fp32 padded accumulators -> reference layout), accumulate and scale, into views of a flat gradient buffer
    accs = [rnd((rp, cp), 20 + i).to(dev) for i, (_, _, _, rp, cp, *_) in enumerate(specs)]
    g1, g2 = rnd((1 << 20,), 9).to(dev), None
    g2 = g1.clone()
    for g, use_batch in ((g1, False), (g2, True)):
        outs = [g[o:o + r * c].view(r, c) for r, c, o, *_ in specs]
        if use_batch:
            ops.batch_begin("unpack", "t")
        for i, (a_, o_, (_, _, _, _, _, rm, cm)) in enumerate(zip(accs, outs, specs)):
            ops.unpack_matrix(a_, o_, row_map=rm, col_map=cm, accumulate=(i % 2 == 0), scale=0.5 if i % 3 == 0 else 1.0)
        if use_batch:
            ops.batch_end()
    assert torch.equal(g1, g2)


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_ln_fold(ops, dtype):
    """LayerNorm affine folded into the K / V projection: packed operands and the gradient unfold"""
    D, H, dh, dhp, Dp = 300, 12, 25, 32, 320
    W, bias, gamma, beta = rnd((2 * D, D), 4, scale=0.05), rnd((2 * D,), 5), 1 + 0.1 * rnd((D,), 6), 0.1 * rnd((D,), 7)
    e, c = both(ops, lambda o, W, b, g, bt, Wp, bp: o.ln_fold_fwd(W, b, g, bt, Wp, bp, row_map=(dh, dhp)), [W, bias, gamma, beta],
                [torch.zeros(2 * H * dhp, Dp, dtype=dtype), torch.zeros(2 * H * dhp)])
    assert max_rel(c[0], e[0]) < (1e-6 if dtype == F32 else 4e-3) and max_rel(c[1], e[1]) < 1e-5
    gWf, gbf = rnd((2 * H * dhp, Dp), 8), rnd((2 * H * dhp,), 9)
    outs = [rnd((3 * H * dhp, Dp), 10), rnd((3 * H * dhp,), 11), rnd((Dp,), 12), rnd((Dp,), 13)]
    e, c = both(ops, lambda o, W, g, bt, gWf, gbf, gW, gb, dg, db: o.ln_fold_bwd(W, g, bt, gWf, gbf, gW[H * dhp:], gb[H * dhp:], dg, db,
                                                                                 row_map=(dh, dhp)), [W, gamma, beta, gWf, gbf], outs)
    for a, b in zip(c, e):
        assert max_rel(a, b) < 2e-5


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_stage_embed(ops, dtype):
    B, T, C, Tp, Cp = 3, 10, 35, 16, 64
    src = rnd((T, B, C), 4).permute(1, 0, 2)          # time-major storage, batch-major view
    (e,), (c,) = both(ops, lambda o, s, d: o.stage_rows(s, d, Tp, DROP), [src], [torch.zeros(B * Tp, Cp, dtype=dtype)])
    assert torch.equal(e, c)
    g = rnd((B * Tp, Cp), 5)
    (e,), (c,) = both(ops, lambda o, g, d: o.unstage_rows(g, d, Tp, False, DROP), [g], [torch.zeros(B, T, C)])
    assert torch.equal(e, c)
    # embed
    D, Dp, T = 45, 64, 12
    from bpmult_b200.engine import sinusoid_table
    pe = sinusoid_table(T + 1, D, Dp, "cpu")
    x = torch.zeros(B * T, Dp)
    x[:, :D] = rnd((B * T, D), 6)
    x[5] = 0
    x[7, 0] = 0
    x = x.to(dtype)
    for yd in (F32, dtype):
        (e,), (c,) = both(ops, lambda o, x, pe, y: o.embed_fwd(x, pe, B, T, D, D ** 0.5, y, DROP), [x, pe], [torch.zeros(B * T, Dp, dtype=yd)])
        assert max_rel(c, e) < (1e-6 if yd == F32 else 8e-3)
        assert (c[:, D:] == 0).all()
    dy = rnd((B * T, Dp), 7)
    (e,), (c,) = both(ops, lambda o, dy, dx: o.embed_bwd(dy, D, D ** 0.5, dx, True, DROP), [dy], [rnd((B * T, Dp), 8)])
    assert max_rel(c, e) < 1e-6


# ------------------------------------------------------------------------------------------------ layernorm
@pytest.mark.parametrize("D,Dp", [(300, 320), (40, 64), (768, 768), (45, 64)])
@pytest.mark.parametrize("xd,yd", [(F32, F32), (F32, BF16), (BF16, BF16)])
def test_layernorm(ops, D, Dp, xd, yd):
    rows = 77
    x = torch.zeros(rows, Dp)
    x[:, :D] = rnd((rows, D), 9) * 3 + 1
    x = x.to(xd)
    gam, bet = torch.zeros(Dp), torch.zeros(Dp)
    gam[:D] = 1 + 0.1 * rnd((D,), 10)
    bet[:D] = 0.1 * rnd((D,), 11)
    outs = [torch.zeros(rows, Dp, dtype=yd), torch.zeros(rows), torch.zeros(rows)]
    e, c = both(ops, lambda o, x, g, b, y, m, r: o.layernorm_fwd(x, g, b, D, y, m, r), [x, gam, bet], outs)
    assert max_rel(c[0], e[0]) < tol(yd)
    assert max_rel(c[1], e[1]) < 1e-5 and max_rel(c[2], e[2]) < 1e-5
    assert (c[0][:, D:] == 0).all()
    dy = torch.zeros(rows, Dp)
    dy[:, :D] = rnd((rows, D), 12)
    dy = dy.to(yd)
    outs = [rnd((rows, Dp), 13), torch.zeros(Dp), torch.zeros(Dp)]
    e2, c2 = both(ops, lambda o, dy, x, m, r, g, dx, dg, db: o.layernorm_bwd(dy, x, m, r, g, D, dx, True, dg, db),
                  [dy, x, e[1], e[2], gam], outs)
    for a, b in zip(c2, e2):
        assert max_rel(a, b) < 2e-5


@pytest.mark.parametrize("D,Dp", [(300, 320), (40, 64), (100, 128), (768, 768), (380, 384)])
@pytest.mark.parametrize("cd", [F32, BF16])
@pytest.mark.parametrize("rows", [77, 1, 512])
def test_layernorm_bwd_with_fused_cast_and_dropout(ops, D, Dp, cd, rows):
    """the backward's fused epilogue (next block's GEMM operand = dropout mask * dx_new, bit-exact mask) in overwrite mode, for the
    half-warp-per-row mapping (Dp = 64 * n, odd row counts, a single row) and the warp-per-row one (Dp = 768)"""
    x = torch.zeros(rows, Dp)
    x[:, :D] = rnd((rows, D), 31) * 2 - 0.5
    gam = torch.zeros(Dp)
    gam[:D] = 1 + 0.1 * rnd((D,), 32)
    mean = x[:, :D].mean(1)
    rstd = (x[:, :D].var(1, unbiased=False) + 1e-5).rsqrt()
    dy = torch.zeros(rows, Dp)
    dy[:, :D] = rnd((rows, D), 33)
    dy = dy.to(BF16)
    outs = [rnd((rows, Dp), 34), torch.zeros(Dp), torch.zeros(Dp), torch.full((rows, Dp), 7.0, dtype=cd)]
    e, c = both(ops, lambda o, dy, x, m, r, g, dx, dg, db, co: o.layernorm_bwd(dy, x, m, r, g, D, dx, False, dg, db, cast_out=co, cast_drop=DROP),
                [dy, x, mean, rstd, gam], outs)
    for a, b in zip(c[:3], e[:3]):
        assert max_rel(a, b) < 2e-5
    assert torch.equal(c[3] == 0, e[3] == 0)                              # the same elements dropped
    assert max_rel(c[3].float(), e[3].float()) < tol(cd)
    assert (c[0][:, D:] == 0).all() and (c[3][:, D:] == 0).all()


# ------------------------------------------------------------------------------------------------ GEMM (FFMA fp32 + tcgen05 bf16)
GEMM_SHAPES = [(128, 320, 320), (256, 384, 320), (40, 320, 384), (200, 1216, 320), (300, 320, 1216), (128, 64, 64), (130, 48, 72), (1024, 320, 320)]


def _gemm_case(ops, dtype, M, N, K, ta, tb, cdt=None, **epi):
    cdt = cdt or dtype
    A = rnd((K, M) if ta else (M, K), 20, dtype, 0.5)
    B = rnd((K, N) if tb else (N, K), 21, dtype, 0.5)
    ins = [A, B]
    kw = {}
    if epi.get("bias"):
        ins.append(rnd((N,), 22)); kw["bias"] = 2
    if epi.get("gate"):
        ins.append(rnd((M, N), 23, dtype)); kw["gate"] = len(ins) - 1
    if epi.get("residual") is not None:
        ins.append(rnd((M, N), 24, epi["residual"])); kw["residual"] = len(ins) - 1
    C0 = rnd((M, N), 25, cdt) if epi.get("accumulate") else torch.zeros(M, N, dtype=cdt)

    def fn(o, *t):
        tensors = list(t)
        Cout = tensors[-1]
        args = {k: tensors[v] for k, v in kw.items()}
        o.gemm(tensors[0], tensors[1], Cout, M, N, K, ta=ta, tb=tb, alpha=epi.get("alpha", 1.0), act=epi.get("act", 0),
               drop=epi.get("drop"), gate_scale=epi.get("gate_scale", 1.0), accumulate=epi.get("accumulate", False), **args)
    (e,), (c,) = both(ops, fn, ins, [C0])
    return max_rel(c, e)


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain_nt(ops, dtype, M, N, K):
    assert _gemm_case(ops, dtype, M, N, K, 0, 0, cdt=F32) < 5e-5   # bf16 inputs are exact on both sides => pure fp32-accumulate parity


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("ta,tb", [(0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(128, 320, 320), (384, 320, 1000), (200, 1216, 320), (320, 384, 136)])
def test_gemm_transposed_operands(ops, dtype, ta, tb, M, N, K):
    assert _gemm_case(ops, dtype, M, N, K, ta, tb, cdt=F32) < 2e-5


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_gemm_epilogues(ops, dtype):
    M, N, K = 200, 320, 384
    t = tol(dtype)
    assert _gemm_case(ops, dtype, M, N, K, 0, 0, bias=True, alpha=0.2) < t
    assert _gemm_case(ops, dtype, M, N, K, 0, 0, bias=True, act=1, drop=DROP) < t
    assert _gemm_case(ops, dtype, M, N, K, 0, 0, cdt=F32, bias=True, drop=DROP, residual=F32) < 2e-5
    assert _gemm_case(ops, dtype, M, N, K, 0, 1, gate=True, gate_scale=1.25) < t
    assert _gemm_case(ops, dtype, M, N, K, 0, 0, residual=dtype) < t
    assert _gemm_case(ops, dtype, 384, 320, 4096, 1, 1, cdt=F32, accumulate=True) < 2e-5       # split-K wgrad shape


def test_colsum(ops):
    for dtype in (F32, BF16):
        X = rnd((1000, 384), 30, dtype)
        (e,), (c,) = both(ops, lambda o, X, out: o.colsum(X, 380, out), [X], [rnd((384,), 31)])
        assert max_rel(c, e) < 1e-5


# ------------------------------------------------------------------------------------------------ attention
ATTN_CASES = [  # B, T, S, H, dh, dhp, mask_off, p
    (2, 12, 12, 4, 10, 32, 0, 0.0), (2, 6, 4, 4, 10, 32, 2, 0.0), (3, 4, 6, 4, 10, 32, 2, 0.0), (2, 9, 17, 4, 10, 32, -1, 0.0),
    (2, 130, 200, 12, 25, 32, 70, 0.1), (1, 256, 256, 12, 25, 32, 0, 0.0), (1, 200, 512, 6, 128, 128, -1, 0.1), (2, 512, 512, 2, 25, 32, 0, 0.2),
    # head dims without two free padding columns: the instantiations that do NOT fold lse / delta / row sums into the MMAs
    (2, 140, 140, 3, 30, 32, 0, 0.0), (1, 64, 200, 2, 32, 32, -1, 0.0),
    # ragged lengths / offset-causal masks on the folded no-dropout instantiations (T != S as in the 4-modality model)
    (2, 200, 512, 12, 25, 32, 312, 0.0), (2, 512, 200, 12, 25, 32, 312, 0.0), (3, 384, 384, 5, 25, 32, 0, 0.0), (1, 76, 333, 3, 25, 32, -1, 0.0),
    # more than 512 queries at head dim 32: the backward runs as chunks of 512 queries whose dK / dV shares are added by TMA reduce
    (1, 1024, 640, 2, 25, 32, 384, 0.0), (1, 700, 700, 3, 25, 32, 0, 0.1), (2, 1100, 300, 2, 25, 32, -1, 0.0), (1, 2048, 2048, 1, 30, 32, 0, 0.0),
]


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("case", ATTN_CASES)
def test_xattn_fwd_bwd(ops, dtype, case):
    B, T, S, H, dh, dhp, off, p = case
    HP = H * dhp

    def mk(rows, seed, scale):
        t = torch.zeros(rows, H, dhp)
        t[:, :, :dh] = rnd((rows, H, dh), seed) * scale
        return t.view(rows, HP).to(dtype)
    q, k, v, do = mk(B * T, 40, dh ** -0.25), mk(B * S, 41, dh ** -0.25), mk(B * S, 42, 1.0), mk(B * T, 43, 1.0)
    drop = Drop(p, 99, None, 5) if p > 0 else None
    W = (S + 31) // 32
    for use_bits in ([False, True] if p > 0 else [False]):
        outs = [torch.zeros(B * T, HP, dtype=dtype), torch.zeros(B * H * T), torch.zeros(B * H * T * W, dtype=torch.int32)]
        e, c = both(ops, lambda o, q, k, v, out, lse, bits: o.xattn_fwd(q, k, v, out, lse, B, T, S, H, dh, dhp, off, None, drop,
                                                                         drop_bits=bits if use_bits else None), [q, k, v], outs)
        assert max_rel(c[0], e[0]) < tol(dtype)
        assert max_rel(c[1], e[1]) < (1e-5 if dtype == F32 else 2e-3)
        nws = 2 * B * H * T                                   # workspace: delta | lse*log2e (| fp32 dQ accumulator of the head-dim-128 kernel)
        outs = [torch.zeros((nws + 31) // 32 * 32 + B * T * HP), torch.zeros(B * T, HP, dtype=dtype), torch.zeros(B * S, HP, dtype=dtype), torch.zeros(B * S, HP, dtype=dtype)]
        e2, c2 = both(ops, lambda o, q, k, v, out, do, lse, bits, dl, dq, dk, dv: o.xattn_bwd(q, k, v, out, do, lse, dl, dq, 0.2, dk, dv, B, T, S, H, dh, dhp, off,
                                                                                             None, drop, drop_bits=bits if use_bits else None),
                      [q, k, v, e[0], do, e[1], c[2]], outs)
        t = 2e-5 if dtype == F32 else 1.5e-2
        c2[0], e2[0] = c2[0][:nws], e2[0][:nws]
        for a, b, nm in zip(c2, e2, ["delta", "dq", "dk", "dv"]):
            assert max_rel(a, b) < t, (nm, use_bits)
    if dtype == F32:
        (ew,), (cw,) = both(ops, lambda o, q, k, lse, w: o.xattn_weights(q, k, lse, w, B, T, S, H, dh, dhp, off, None, drop), [q, k, e[1]], [torch.zeros(B, T, S)])
        assert max_rel(cw, ew) < 2e-5


# head dim 64 / 128 on the tensor-core kernels of attn_tc128.cu (the 4-modality model: hidden 768 / 6 heads, README.md:30,36)
ATTN_CASES_WIDE = [  # B, T, S, H, dh, dhp, mask_off, p
    (1, 128, 128, 1, 128, 128, -1, 0.0), (2, 256, 256, 2, 128, 128, 0, 0.0), (1, 512, 512, 6, 128, 128, 0, 0.0),
    (2, 200, 512, 6, 128, 128, 312, 0.0), (2, 512, 200, 6, 128, 128, 312, 0.0), (3, 200, 200, 6, 128, 128, 0, 0.1),
    (1, 76, 333, 3, 128, 128, -1, 0.2), (1, 1024, 640, 2, 128, 128, 384, 0.0), (2, 130, 70, 2, 100, 128, 60, 0.0),
    (1, 384, 384, 2, 64, 64, 0, 0.0), (2, 200, 512, 3, 64, 64, -1, 0.1), (1, 300, 150, 2, 50, 64, 150, 0.0),
]


@pytest.mark.parametrize("case", ATTN_CASES_WIDE)
def test_xattn_fwd_bwd_wide_heads(ops, case):
    test_xattn_fwd_bwd(ops, BF16, case)


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("dims", [(2, 8, 10, 2, 16, 32, -1), (2, 140, 200, 3, 25, 32, 60), (2, 200, 333, 2, 128, 128, -1), (1, 256, 256, 2, 128, 128, 0)])
def test_xattn_key_padding_mask(ops, dtype, dims):
    """north star (2): key-padding mask (superset feature -- the reference has none, multihead_attention.py:52), forward AND backward.
    Padded keys get probability 0 and zero dK / dV; every query keeps at least one visible key."""
    B, T, S, H, dh, dhp, off = dims
    HP = H * dhp

    def mk(rows, seed, scale):
        t = torch.zeros(rows, H, dhp)
        t[:, :, :dh] = rnd((rows, H, dh), seed) * scale
        return t.view(rows, HP).to(dtype)
    q, k, v, do = mk(B * T, 50, dh ** -0.25), mk(B * S, 51, dh ** -0.25), mk(B * S, 52, 1.0), mk(B * T, 53, 1.0)
    kp = torch.zeros(B, S, dtype=torch.uint8)
    kp[0, S - S // 3:] = 1                         # a padded tail
    kp[B - 1, 3] = 1                               # and a hole
    kp[B - 1, S // 2:S // 2 + 5] = 1
    e, c = both(ops, lambda o, q, k, v, kp, out, lse: o.xattn_fwd(q, k, v, out, lse, B, T, S, H, dh, dhp, off, kp, None), [q, k, v, kp],
                [torch.zeros(B * T, HP, dtype=dtype), torch.zeros(B * H * T)])
    assert max_rel(c[0], e[0]) < tol(dtype)
    assert max_rel(c[1], e[1]) < (1e-5 if dtype == F32 else 2e-3)
    outs = [torch.zeros((2 * B * H * T + 31) // 32 * 32 + B * T * HP), torch.zeros(B * T, HP, dtype=dtype), torch.zeros(B * S, HP, dtype=dtype), torch.zeros(B * S, HP, dtype=dtype)]
    e2, c2 = both(ops, lambda o, q, k, v, kp, out, do, lse, dl, dq, dk, dv: o.xattn_bwd(q, k, v, out, do, lse, dl, dq, 0.2, dk, dv, B, T, S, H, dh, dhp, off,
                                                                                       kp, None), [q, k, v, kp, e[0], do, e[1]], outs)
    t = 2e-5 if dtype == F32 else 1.5e-2
    for a, b, nm in zip(c2[1:], e2[1:], ["dq", "dk", "dv"]):
        assert max_rel(a, b) < t, nm
    padded = kp.bool().view(B * S)
    assert float(c2[2].float()[padded].abs().max()) == 0.0 and float(c2[3].float()[padded].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------ GMU / pooling / head pieces / loss / adam
@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("features", [1, 0])
def test_gmu_combine(ops, dtype, features):
    rows, Dp = 50, 64
    a1, a2, h1, h2, zp, ad = [rnd((rows, Dp), 60 + i, dtype) for i in range(6)]
    e, c = both(ops, lambda o, a1, a2, h1, h2, zp, ad, y, z: o.gmu_fwd(features, a1, a2, h1, h2, zp, ad, y, z), [a1, a2, h1, h2, zp, ad],
                [torch.zeros(rows, Dp, dtype=dtype), torch.zeros(rows, Dp, dtype=dtype)])
    assert max_rel(c[0], e[0]) < tol(dtype) and max_rel(c[1], e[1]) < tol(dtype)
    dy = rnd((rows, Dp), 70)
    outs = [torch.zeros(rows, Dp, dtype=dtype) for _ in range(3)] + [rnd((rows, Dp), 71), rnd((rows, Dp), 72)]
    e, c = both(ops, lambda o, a1, a2, h1, h2, zp, dy, d1, d2, dz, g1, g2: o.gmu_bwd(features, a1, a2, h1, h2, zp, dy, d1, d2, dz, g1, g2),
                [a1, a2, h1, h2, zp, dy], outs)
    for a, b in zip(c, e):
        assert max_rel(a, b) < tol(dtype)


def test_small_ops(ops):
    B, T, Dp = 3, 7, 64
    for dtype in (F32, BF16):
        x = rnd((B * T, Dp), 80, dtype)
        (e,), (c,) = both(ops, lambda o, x, out: o.pool_fwd(x, B, T, out, Dp), [x], [torch.zeros(B, 3 * Dp)])
        assert max_rel(c, e) < 1e-6
        a, b = rnd((40, 64), 81, dtype), rnd((40, 64), 82, dtype)
        (e,), (c,) = both(ops, lambda o, a, b, y: o.add(a, b, y), [a, b], [torch.zeros(40, 64, dtype=dtype)])
        assert torch.equal(e, c)
        (e,), (c,) = both(ops, lambda o, a, d: o.axpy_f32(a, d, True), [a], [rnd((40, 64), 83)])
        assert torch.equal(e, c)
    dout = rnd((B, 3 * Dp), 84)
    (e,), (c,) = both(ops, lambda o, d, dx: o.pool_bwd(d, Dp, B, T, dx), [dout], [rnd((B * T, Dp), 85)])
    assert torch.equal(e, c)
    n = 3
    hp, zp = rnd((n, B, Dp), 86), rnd((n, B, Dp), 87)
    e, c = both(ops, lambda o, h, z, f, zo: o.tsgate_fwd(h, z, n, B, Dp, f, zo), [hp, zp], [torch.zeros(B, Dp), torch.zeros(B, n * Dp)])
    assert max_rel(c[0], e[0]) < 1e-5 and max_rel(c[1], e[1]) < 1e-5
    df = rnd((B, Dp), 88)
    e, c = both(ops, lambda o, h, z, d, dh, dz: o.tsgate_bwd(h, z, d, n, B, Dp, dh, dz), [hp, zp, df], [torch.zeros(n, B, Dp), torch.zeros(n, B, Dp)])
    assert max_rel(c[0], e[0]) < 1e-5 and max_rel(c[1], e[1]) < 1e-5


def test_bce_matches_golden_and_torch(ops):
    from helpers import load_gold
    g = load_gold("modules.pt")["bce"]
    B, C = g["x"].shape
    logits = torch.zeros(B, 8)
    logits[:, :C] = g["x"]
    loss, dl = torch.zeros(1).cuda(), torch.zeros(B, 8).cuda()
    ops.bce_fwd_bwd(logits.cuda(), g["y"].cuda(), g["w"].cuda(), B, C, 1.0, loss, dl)
    assert abs(loss.item() - g["loss"].item()) < 1e-6
    assert max_rel(dl[:, :C], g["dx"]) < 1e-5


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("Tin,Tout", [(200, 512), (512, 200), (7, 5)])
def test_time_axis_linear(ops, dtype, Tin, Tout):
    """transfm_x2y of the 4-modality model (mmtr.py:507-508): Linear over the TIME axis, on batch-major padded rows"""
    B, D, ld = 3, 45, 64
    x = torch.zeros(B, Tin, ld)
    x[:, :, :D] = rnd((B, Tin, D), 70)
    W, bias = rnd((Tout, Tin), 71) * 0.1, rnd((Tout,), 72)
    xr, Wr, br = x[:, :, :D].clone().requires_grad_(), W.clone().requires_grad_(), bias.clone().requires_grad_()
    if dtype == BF16:
        xr = x[:, :, :D].to(BF16).float().requires_grad_()
    yr = torch.nn.functional.linear(xr.permute(2, 0, 1), Wr, br).permute(1, 2, 0)          # [D, B, T] -> Linear -> [B, T2, D]
    gy = rnd((B, Tout, D), 73)
    yr.backward(gy)
    xg = x.reshape(B * Tin, ld).to(dtype).cuda()
    y = torch.full((B * Tout, ld), 7.0, dtype=dtype, device="cuda")
    ops.timelin_fwd(xg, W.cuda(), bias.cuda(), y, B, Tin, Tout, D)
    yv = y.float().cpu().view(B, Tout, ld)
    assert max_rel(yv[:, :, :D], yr.detach()) < (1e-5 if dtype == F32 else 1e-2)
    assert float(yv[:, :, D:].abs().max()) == 0.0
    dy = torch.zeros(B * Tout, ld, device="cuda")
    dy.view(B, Tout, ld)[:, :, :D] = gy.cuda()
    dx = torch.ones(B * Tin, ld, device="cuda")
    dW, db = torch.zeros(Tout, Tin, device="cuda"), torch.zeros(Tout, device="cuda")
    ops.timelin_bwd(dy, xg, W.cuda(), dx, True, dW, db, B, Tin, Tout, D)
    assert max_rel(dx.cpu().view(B, Tin, ld)[:, :, :D] - 1.0, xr.grad) < 1e-4
    assert max_rel(dW.cpu(), Wr.grad) < 1e-4 and max_rel(db.cpu(), br.grad) < 1e-4


def test_adam_matches_torch(ops):
    p0, g = rnd((1000,), 90), rnd((1000,), 91)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    p, m, v = p0.clone().cuda(), torch.zeros(1000).cuda(), torch.zeros(1000).cuda()
    step = torch.zeros(1, dtype=torch.int64).cuda()
    for i in range(3):
        ref.grad = g * (i + 1)
        opt.step()
        step += 1
        ops.adam_step(p, (g * (i + 1)).cuda(), m, v, 1e-3, 0.9, 0.999, 1e-8, 1.0, step)
    assert max_rel(p, ref.detach()) < 1e-6
    # the device-resident rate overrides the scalar one
    lr_t = torch.tensor([2e-3], device="cuda")
    opt.param_groups[0]["lr"] = 2e-3
    ref.grad = g
    opt.step()
    step += 1
    ops.adam_step(p, g.cuda(), m, v, 123.0, 0.9, 0.999, 1e-8, 1.0, step, lr_t)
    assert max_rel(p, ref.detach()) < 1e-6


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("M,N,K", [(384, 320, 4096), (320, 1216, 2000), (1216, 320, 1000)])
def test_gemm_wgrad_with_fused_bias_grad(ops, dtype, M, N, K):
    """dW = dY^T X (split-K, fp32 accumulate) with the column sums of dY (= bias gradient) fused into the same launch"""
    A, Bm = rnd((K, M), 26, dtype, 0.5), rnd((K, N), 27, dtype, 0.5)
    C0, cs0 = rnd((M, N), 28), rnd((M,), 29)
    e, c = both(ops, lambda o, A, Bm, C, cs: o.gemm(A, Bm, C, M, N, K, ta=1, tb=1, accumulate=True, colsum=cs), [A, Bm], [C0, cs0])
    assert max_rel(c[0], e[0]) < 2e-5
    assert max_rel(c[1], e[1]) < 2e-5


@pytest.mark.parametrize("cdt", [F32, BF16])
def test_gemm_bf16_tails_are_clipped(ops, cdt):
    """M, N tails: the TMA store must not touch memory outside C[:M, :N] (C is a view inside a larger poisoned buffer)"""
    M, N, K = 200, 328, 320
    A, Bm = rnd((M, K), 30, BF16, 0.5), rnd((N, K), 31, BF16, 0.5)
    big = torch.full((256, 512), 7.0, dtype=cdt)
    e, c = both(ops, lambda o, A, Bm, big: o.gemm(A, Bm, big[:M, :N], M, N, K), [A, Bm], [big])
    assert max_rel(c[0][:M, :N], e[0][:M, :N]) < tol(cdt)
    assert (c[0][M:] == 7.0).all() and (c[0][:, N:] == 7.0).all()


# ------------------------------------------------------------------------------------------------ AudioEncoder layout kernels (mmtr.py:93-108)
@pytest.mark.parametrize("dtype", [F32, BF16])
def test_conv1d_im2col_col2im_pool(ops, dtype):
    B, Tin, C, KW, st = 2, 300, 96, 128, 2
    Tout = (Tin - KW) // st + 1
    x = rnd((B * Tin, C), 90, dtype)
    (e,), (c,) = both(ops, lambda o, x, col: o.conv1d_im2col(x, B, Tin, C, KW, st, col, Tout), [x], [torch.zeros(B * Tout, KW * C, dtype=dtype)])
    assert torch.equal(e, c)
    dcol = rnd((B * Tout, KW * C), 91, dtype)
    (e,), (c,) = both(ops, lambda o, d, dx: o.conv1d_col2im(d, B, Tin, C, KW, st, Tout, dx), [dcol], [torch.zeros(B * Tin, C)])
    assert max_rel(c, e) < 1e-5
    W = rnd((C, C, KW), 92)
    (e,), (c,) = both(ops, lambda o, W, Wp: o.conv1d_pack_weight(W, Wp), [W], [torch.zeros(C, KW * C, dtype=dtype)])
    assert torch.equal(e, c)
    gWp = rnd((C, KW * C), 93)
    (e,), (c,) = both(ops, lambda o, g, gW: o.conv1d_unpack_wgrad(g, gW, True), [gWp], [rnd((C, C, KW), 94)])
    assert torch.allclose(e, c, rtol=1e-6, atol=1e-6)
    for T, Tp in ((255, 200), (80, 200), (400, 200)):                  # down-sampling, up-sampling, exact halving
        xs = rnd((B * T, C), 95, dtype)
        (e,), (c,) = both(ops, lambda o, xs, y: o.adaptive_pool_fwd(xs, B, T, C, Tp, y), [xs], [torch.zeros(B * Tp, 128, dtype=dtype)])
        assert max_rel(c, e) < (1e-6 if dtype == F32 else 8e-3)
        dy = rnd((B * Tp, 128), 96)
        (e,), (c,) = both(ops, lambda o, dy, dx: o.adaptive_pool_bwd(dy, B, T, C, Tp, dx), [dy], [torch.zeros(B * T, C)])
        assert max_rel(c, e) < 1e-5


@pytest.mark.parametrize("M,N,K", [(64, 320, 320), (64, 320, 960), (8, 8, 320), (64, 3072, 768), (33, 45, 100)])
def test_gemm_skinny_rows(ops, M, N, K):
    """the [B, D] head's GEMMs (M <= 64 rows, exact fp32): the skinny kernel with every epilogue the head uses"""
    assert _gemm_case(ops, F32, M, N, K, 0, 0) < 2e-5
    assert _gemm_case(ops, F32, M, N, K, 0, 0, bias=True, act=1, drop=DROP) < 2e-5
    assert _gemm_case(ops, F32, M, N, K, 0, 0, bias=True, residual=F32) < 2e-5
    assert _gemm_case(ops, F32, M, N, K, 0, 1, gate=True, gate_scale=1.25) < 2e-5
    assert _gemm_case(ops, F32, M, N, K, 0, 1, residual=F32) < 2e-5
