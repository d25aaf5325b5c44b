"""Thin tensor-level wrappers over the C ABI (include/bpmult_b200.h).  torch is plumbing here: it owns device memory
and the stream; every op below is one or two launches of our own CUDA kernels on `torch.cuda.current_stream()`.

`CudaOps` is the only product implementation.  (tests/emu_ops.py holds a pure-torch emulation of the same contract
that is used ONLY to unit-test the host-side engine logic on machines without a GPU.)"""
import ctypes as C
from collections import namedtuple

import torch

from . import _lib
from ._lib import Attn, Dropout, Gemm

Drop = namedtuple("Drop", "p seed seed_ptr site")
NO_DROP = Drop(0.0, 0, None, 0)


def _dt(t):
    if t.dtype == torch.float32:
        return _lib.BPM_F32
    if t.dtype == torch.bfloat16:
        return _lib.BPM_BF16
    raise TypeError("bpmult_b200: unsupported dtype %s" % t.dtype)


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _drop(d):
    if d is None or d.p <= 0.0:
        return Dropout(0, None, 0, 0.0)
    return Dropout(int(d.seed) & 0xFFFFFFFFFFFFFFFF, None if d.seed_ptr is None else d.seed_ptr.data_ptr(), int(d.site), float(d.p))


class CudaOps:
    """All methods write into caller-provided output tensors (no allocation, no sync)."""
    name = "cuda"

    def __init__(self, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.BpmError("bpmult_b200: no CUDA device -- this package has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if not self.lib.bpm_device_ok(self.device.index or 0):
            raise _lib.BpmError("bpmult_b200: device %s is not sm_100 (B200)" % self.device)
        self.launches = 0
        self._depth = 0
        self._batch = None          # (mode, key, [descriptor tuples]) while pack / unpack calls are being collected
        self._zlist, self._zdepth = None, 0     # tensors collected between zero_begin() and zero_end()
        self._tables = {}           # descriptor-list signature -> [device table, pinned by a captured graph]  (see _table)
        self.max_tables = 256
        self.capture_owner = "anonymous"

    # ------------------------------------------------------------------ batched pack / unpack (one launch per table)
    def batch_begin(self, mode, key):
        if self._batch is not None:                 # nested (e.g. an encoder inside the model-level batch): join the outer one
            self._depth += 1
            return
        self._batch, self._depth = (mode, key, []), 0

    def batch_end(self):
        import numpy as np
        if self._depth > 0:
            self._depth -= 1
            return
        mode, key, items = self._batch
        self._batch = None
        if not items:
            return
        def build():
            dt = np.dtype([("src", "u8"), ("dst", "u8")] + [(n, "i4") for n in ("rows", "cols", "ld_src", "ld_dst", "rows_p", "cols_p", "row_dh",
                          "row_dhp", "col_dh", "col_dhp", "dst_dtype", "accumulate")] + [("scale", "f4"), ("pad", "i4")], align=True)
            assert dt.itemsize == 72
            arr = np.zeros(len(items), dtype=dt)
            for i, it in enumerate(items):
                arr[i] = it + (0,)
            # work units: runs of destination rows worth ~16 K elements each (padded rows when packing, reference rows when unpacking)
            units = []
            for i, it in enumerate(items):
                rows, cols = (it[6], it[7]) if mode == "pack" else (it[2], it[3])
                step = max(1, 16384 // max(1, cols))
                units += [(i, r0, min(step, rows - r0), 0) for r0 in range(0, rows, step)]
            return np.concatenate([arr.view("uint8"), np.asarray(units, dtype=np.int32).reshape(-1).view("uint8")]), len(units)
        tab, n_units = self._table((mode,) + tuple(items), build)
        self._ck(self.lib.bpm_remap_units(tab.data_ptr(), tab.data_ptr() + 72 * len(items), n_units, 0 if mode == "pack" else 1, self._s()),
                 "remap_units")

    def _table(self, sig, build):
        """Device descriptor table for a batched launch, cached by the full descriptor list (alternating engines never rebuild, and
        a CUDA-graph capture never sees a host->device copy: the eager warm-up steps have created every table it uses).
        A table that is used while the stream is capturing is baked into that graph's kernel arguments, so it is PINNED by the
        capturing owner (`capture_owner`, set by the Trainer): it is never evicted -- its memory must not return to the allocator while
        the graph can still be replayed -- until the owner calls release_tables().  Only unpinned tables (the eager autograd path,
        whose gradient tensors may have new addresses every call) are evicted, oldest first, beyond `max_tables`."""
        ent = self._tables.get(sig)
        capturing = torch.cuda.is_current_stream_capturing()
        if ent is None:
            if capturing:
                raise _lib.BpmError("bpmult_b200: a descriptor table is missing during CUDA-graph capture (run the step eagerly once first)")
            built, meta = build(), None
            if isinstance(built, tuple):                    # (table bytes, host-side metadata kept with the entry)
                built, meta = built
            ent = [torch.from_numpy(built.view("uint8").copy()).to(self.device), set(), meta]
            self._tables[sig] = ent
            unpinned = [k for k, e in self._tables.items() if not e[1]]
            for k in unpinned[:max(0, len(unpinned) - self.max_tables)]:
                del self._tables[k]
        else:
            self._tables[sig] = self._tables.pop(sig)       # most recently used last
        if capturing:
            ent[1].add(self.capture_owner)
        return ent[0] if ent[2] is None else (ent[0], ent[2])

    def release_tables(self, owner):
        """the graphs captured by `owner` are gone: their descriptor tables may be evicted again"""
        for e in self._tables.values():
            e[1].discard(owner)

    # ------------------------------------------------------------------ helpers
    def _s(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _ck(self, rc, what):
        self.launches += 1
        if rc != 0:
            _lib.check(rc, what)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ weight staging
    def pack_matrix(self, src, dst, row_map=(0, 0), col_map=(0, 0)):
        assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1 and dst.stride(-1) == 1
        if self._batch is not None and self._batch[0] == "pack":
            self._batch[2].append((src.data_ptr(), dst.data_ptr(), src.shape[0], src.shape[1], src.stride(0), dst.stride(0), dst.shape[0],
                                   dst.shape[1], row_map[0], row_map[1], col_map[0], col_map[1], _dt(dst), 0, 1.0))
            return
        assert dst.is_contiguous()
        self._ck(self.lib.bpm_pack_matrix(src.data_ptr(), src.shape[0], src.shape[1], src.stride(0), dst.data_ptr(), dst.shape[0],
                                          dst.shape[1], _dt(dst), row_map[0], row_map[1], col_map[0], col_map[1], self._s()), "pack_matrix")

    def unpack_matrix(self, src_p, dst, row_map=(0, 0), col_map=(0, 0), accumulate=False, scale=1.0):
        assert src_p.dtype == torch.float32 and dst.dtype == torch.float32 and dst.dim() == 2 and dst.stride(1) == 1
        if self._batch is not None and self._batch[0] == "unpack":
            self._batch[2].append((src_p.data_ptr(), dst.data_ptr(), dst.shape[0], dst.shape[1], src_p.stride(0), dst.stride(0), src_p.shape[0],
                                   src_p.shape[1], row_map[0], row_map[1], col_map[0], col_map[1], 0, int(accumulate), float(scale)))
            return
        assert src_p.is_contiguous()
        self._ck(self.lib.bpm_unpack_matrix(src_p.data_ptr(), src_p.shape[0], src_p.shape[1], dst.data_ptr(), dst.shape[0], dst.shape[1],
                                            dst.stride(0), row_map[0], row_map[1], col_map[0], col_map[1], int(accumulate), float(scale),
                                            self._s()), "unpack_matrix")

    # ------------------------------------------------------------------ staging / embed
    def stage_rows(self, src, dst, Tp, drop=None):
        """src fp32 (B, T, C) with arbitrary strides -> dst [B*Tp, Cp]"""
        B, T, Cc = src.shape
        assert src.dtype == torch.float32 and dst.shape[0] == B * Tp
        self._ck(self.lib.bpm_stage_rows(src.data_ptr(), B, T, Cc, src.stride(0), src.stride(1), src.stride(2), dst.data_ptr(), Tp,
                                         dst.shape[1], _dt(dst), _drop(drop), self._s()), "stage_rows")

    def unstage_rows(self, g, dsrc, Tp, accumulate=False, drop=None):
        B, T, Cc = dsrc.shape
        assert g.dtype == torch.float32 and dsrc.dtype == torch.float32
        self._ck(self.lib.bpm_unstage_rows(g.data_ptr(), B, T, Cc, Tp, g.shape[1], dsrc.data_ptr(), dsrc.stride(0), dsrc.stride(1),
                                           dsrc.stride(2), int(accumulate), _drop(drop), self._s()), "unstage_rows")

    def embed_fwd(self, x, pe, B, T, D, scale, y, drop=None):
        Dp = x.shape[1]
        assert pe.shape[0] >= T + 1 and pe.shape[1] == Dp and pe.dtype == torch.float32
        self._ck(self.lib.bpm_embed_fwd(x.data_ptr(), _dt(x), pe.data_ptr(), B, T, D, Dp, float(scale), y.data_ptr(), _dt(y), _drop(drop),
                                        self._s()), "embed_fwd")

    def embed_bwd(self, dy, D, scale, dx, accumulate, drop=None):
        assert dy.dtype == torch.float32 and dx.dtype == torch.float32
        self._ck(self.lib.bpm_embed_bwd(dy.data_ptr(), dy.shape[0], D, dy.shape[1], float(scale), dx.data_ptr(), int(accumulate), _drop(drop),
                                        self._s()), "embed_bwd")

    # ------------------------------------------------------------------ layernorm
    def layernorm_fwd(self, x, gamma, beta, D, y, mean, rstd, eps=1e-5):
        self._ck(self.lib.bpm_layernorm_fwd(x.data_ptr(), _dt(x), gamma.data_ptr(), beta.data_ptr(), x.shape[0], D, x.shape[1], float(eps),
                                            y.data_ptr(), _dt(y), mean.data_ptr(), rstd.data_ptr(), self._s()), "layernorm_fwd")

    def layernorm_bwd(self, dy, x, mean, rstd, gamma, D, dx, accumulate, dgamma, dbeta, cast_out=None, cast_drop=None):
        """cast_out (optional, storage type): dropmask(cast_drop) * dx_new, fused (replaces a following cast_drop(dx, cast_out, cast_drop))"""
        assert dx.dtype == torch.float32
        self._ck(self.lib.bpm_layernorm_bwd_cast(dy.data_ptr(), _dt(dy), x.data_ptr(), _dt(x), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                                 x.shape[0], D, x.shape[1], dx.data_ptr(), int(accumulate), dgamma.data_ptr(), dbeta.data_ptr(),
                                                 _ptr(cast_out), _dt(cast_out) if cast_out is not None else 0, _drop(cast_drop), self._s()),
                 "layernorm_bwd")

    def fold_batch_begin(self, mode):
        """collect ln_fold_fwd / ln_fold_bwd calls and run them as ONE launch at fold_batch_end() (mode "fwd" | "bwd")"""
        self._fold = (mode, [])

    def fold_batch_end(self):
        import numpy as np
        mode, items = self._fold
        self._fold = None
        if not items:
            return
        def build():
            dt = np.dtype([(n, "u8") for n in ("W", "bias", "gamma", "beta", "Wp", "bp", "gWf", "gbf", "gW", "gb", "dgamma", "dbeta")] +
                          [(n, "i4") for n in ("rows", "cols", "ldw", "ldp", "ldf", "ldg", "row_dh", "row_dhp", "wp_dtype", "pad")], align=True)
            assert dt.itemsize == 136
            arr = np.zeros(len(items), dtype=dt)
            for i, it in enumerate(items):
                arr[i] = it
            return arr
        tab = self._table(("fold", mode) + tuple(items), build)
        self._ck(self.lib.bpm_ln_fold_batch(tab.data_ptr(), len(items), max(it[12] for it in items), 0 if mode == "fwd" else 1, self._s()),
                 "ln_fold_batch")

    def ln_fold_fwd(self, W, bias, gamma, beta, Wp, bp, row_map=(0, 0)):
        """Wp[map(i), :cols] = W[i] * gamma;  bp[map(i)] = bias[i] + W[i] . beta   (W, bias, gamma, beta fp32 reference layout)"""
        assert W.dtype == torch.float32 and W.stride(1) == 1 and Wp.stride(1) == 1 and bp.dtype == torch.float32
        if getattr(self, "_fold", None) is not None:
            self._fold[1].append((W.data_ptr(), bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(), Wp.data_ptr(), bp.data_ptr(), 0, 0, 0, 0, 0, 0,
                                  W.shape[0], W.shape[1], W.stride(0), Wp.stride(0), 0, 0, row_map[0], row_map[1], _dt(Wp), 0))
            return
        self._ck(self.lib.bpm_ln_fold_fwd(W.data_ptr(), W.stride(0), bias.data_ptr(), gamma.data_ptr(), beta.data_ptr(), W.shape[0], W.shape[1],
                                          row_map[0], row_map[1], Wp.data_ptr(), _dt(Wp), Wp.stride(0), bp.data_ptr(), self._s()), "ln_fold_fwd")

    def ln_fold_bwd(self, W, gamma, beta, gWf, gbf, gW, gb, dgamma, dbeta, row_map=(0, 0)):
        """gW += gWf * gamma + gbf (x) beta, gb += gbf (padded fp32 accumulators); dgamma / dbeta (fp32) accumulated"""
        assert gW.dtype == torch.float32 and gW.stride(1) == 1 and gWf.dtype == torch.float32 and gWf.stride(1) == 1
        if getattr(self, "_fold", None) is not None:
            self._fold[1].append((W.data_ptr(), 0, gamma.data_ptr(), beta.data_ptr(), 0, 0, gWf.data_ptr(), gbf.data_ptr(), gW.data_ptr(), gb.data_ptr(),
                                  dgamma.data_ptr(), dbeta.data_ptr(), W.shape[0], W.shape[1], W.stride(0), 0, gWf.stride(0), gW.stride(0), row_map[0],
                                  row_map[1], 0, 0))
            return
        self._ck(self.lib.bpm_ln_fold_bwd(W.data_ptr(), W.stride(0), gamma.data_ptr(), beta.data_ptr(), W.shape[0], W.shape[1], row_map[0],
                                          row_map[1], gWf.data_ptr(), gWf.stride(0), gbf.data_ptr(), gW.data_ptr(), gW.stride(0), gb.data_ptr(),
                                          dgamma.data_ptr(), dbeta.data_ptr(), self._s()), "ln_fold_bwd")

    # ------------------------------------------------------------------ gemm
    def gemm(self, A, B, Cout, M, N, K, ta=0, tb=0, bias=None, alpha=1.0, act=0, drop=None, gate=None, gate_scale=1.0, residual=None,
             accumulate=False, split_k=0, colsum=None):
        assert A.dtype == B.dtype and A.stride(1) == 1 and B.stride(1) == 1 and Cout.stride(1) == 1
        g = Gemm()
        g.colsum_out = _ptr(colsum)
        g.ab_dtype, g.ta, g.tb, g.M, g.N, g.K = _dt(A), int(ta), int(tb), M, N, K
        g.A, g.lda, g.B, g.ldb = A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0)
        g.C, g.ldc, g.c_dtype = Cout.data_ptr(), Cout.stride(0), _dt(Cout)
        g.bias, g.alpha, g.act = _ptr(bias), float(alpha), int(act)
        g.drop = _drop(drop)
        g.gate, g.ldg, g.gate_dtype, g.gate_scale = _ptr(gate), (gate.stride(0) if gate is not None else 0), (_dt(gate) if gate is not None else 0), float(gate_scale)
        g.residual, g.ldr, g.res_dtype = _ptr(residual), (residual.stride(0) if residual is not None else 0), (_dt(residual) if residual is not None else 0)
        g.accumulate, g.split_k = int(accumulate), int(split_k)
        self._ck(self.lib.bpm_gemm(C.byref(g), self._s()), "gemm")

    def colsum(self, X, N, out):
        self._ck(self.lib.bpm_colsum(X.data_ptr(), _dt(X), X.shape[0], N, X.stride(0), out.data_ptr(), self._s()), "colsum")

    # ------------------------------------------------------------------ attention
    def _attn(self, q, B, T, S, H, dh, dhp, mask_off, key_pad, drop, drop_bits=None, k=None, v=None, dk=None, dv=None):
        a = Attn()
        a.drop_bits = _ptr(drop_bits)
        a.ld_kv = a.ld_dkv = 0
        if k is not None:                           # k / v (and dk / dv) may be column slices of wider row-major buffers
            assert k.stride(1) == 1 and (v is None or (v.stride(1) == 1 and v.stride(0) == k.stride(0)))
            a.ld_kv = k.stride(0)
        if dk is not None:
            assert dk.stride(1) == 1 and dv.stride(1) == 1 and dv.stride(0) == dk.stride(0)
            a.ld_dkv = dk.stride(0)
        a.dtype, a.B, a.T, a.S, a.H, a.dh, a.dhp, a.mask_off = _dt(q), B, T, S, H, dh, dhp, int(mask_off)
        a.key_pad = _ptr(key_pad)
        a.drop = _drop(drop)
        return a

    def xattn_fwd(self, q, k, v, out, lse, B, T, S, H, dh, dhp, mask_off=-1, key_pad=None, drop=None, drop_bits=None):
        a = self._attn(q, B, T, S, H, dh, dhp, mask_off, key_pad, drop, drop_bits, k=k, v=v)
        self._ck(self.lib.bpm_xattn_fwd(C.byref(a), q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), lse.data_ptr(), self._s()), "xattn_fwd")

    def xattn_bwd_workspace(self, dtype, B, T, S, H, dh, dhp):
        """floats the backward needs at `delta` (2*B*H*T, plus the fp32 dQ accumulator of the head-dim-128 tensor-core kernel)"""
        a = Attn()
        a.dtype, a.B, a.T, a.S, a.H, a.dh, a.dhp = (_lib.BPM_BF16 if dtype == torch.bfloat16 else _lib.BPM_F32), B, T, S, H, dh, dhp
        return int(self.lib.bpm_xattn_bwd_workspace(C.byref(a)))

    def xattn_bwd(self, q, k, v, out, dout, lse, delta, dq, dq_scale, dk, dv, B, T, S, H, dh, dhp, mask_off=-1, key_pad=None, drop=None,
                  drop_bits=None):
        a = self._attn(q, B, T, S, H, dh, dhp, mask_off, key_pad, drop, drop_bits, k=k, v=v, dk=dk, dv=dv)
        assert delta.numel() >= self.lib.bpm_xattn_bwd_workspace(C.byref(a)), "xattn_bwd: workspace too small (see xattn_bwd_workspace)"
        self._ck(self.lib.bpm_xattn_bwd(C.byref(a), q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(),
                                        delta.data_ptr(), dq.data_ptr(), float(dq_scale), dk.data_ptr(), dv.data_ptr(), self._s()), "xattn_bwd")

    def xattn_weights(self, q, k, lse, w, B, T, S, H, dh, dhp, mask_off=-1, key_pad=None, drop=None):
        a = self._attn(q, B, T, S, H, dh, dhp, mask_off, key_pad, drop, k=k)
        self._ck(self.lib.bpm_xattn_weights(C.byref(a), q.data_ptr(), k.data_ptr(), lse.data_ptr(), w.data_ptr(), self._s()), "xattn_weights")

    # ------------------------------------------------------------------ GMU / elementwise
    def gmu_fwd(self, features, a1, a2, h1p, h2p, zp, addend, y, z_out=None):
        self._ck(self.lib.bpm_gmu_fwd(_dt(h1p), int(features), _ptr(a1), _ptr(a2), h1p.data_ptr(), h2p.data_ptr(), zp.data_ptr(), _ptr(addend),
                                      h1p.shape[0], h1p.shape[1], y.data_ptr(), _ptr(z_out), self._s()), "gmu_fwd")

    def gmu_bwd(self, features, a1, a2, h1p, h2p, zp, dy, dh1, dh2, dz, da1, da2):
        self._ck(self.lib.bpm_gmu_bwd(_dt(h1p), int(features), _ptr(a1), _ptr(a2), h1p.data_ptr(), h2p.data_ptr(), zp.data_ptr(), dy.data_ptr(),
                                      h1p.shape[0], h1p.shape[1], dh1.data_ptr(), dh2.data_ptr(), dz.data_ptr(), _ptr(da1), _ptr(da2),
                                      self._s()), "gmu_bwd")

    def add(self, a, b, y):
        self._ck(self.lib.bpm_add(_dt(a), a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), self._s()), "add")

    def axpy_f32(self, src, dst, accumulate=True):
        self._ck(self.lib.bpm_axpy_f32(src.data_ptr(), _dt(src), dst.data_ptr(), src.numel(), int(accumulate), self._s()), "axpy_f32")

    def cast_drop(self, x, y, drop=None):
        self._ck(self.lib.bpm_cast_drop(x.data_ptr(), y.data_ptr(), _dt(y), x.shape[0], x.shape[1], _drop(drop), self._s()), "cast_drop")

    def pool_fwd(self, x, B, T, out, col_off):
        self._ck(self.lib.bpm_pool_fwd(x.data_ptr(), _dt(x), B, T, x.shape[1], out.data_ptr(), out.stride(0), col_off, self._s()), "pool_fwd")

    def pool_bwd(self, dout, col_off, B, T, dx):
        self._ck(self.lib.bpm_pool_bwd(dout.data_ptr(), dout.stride(0), col_off, B, T, dx.shape[1], dx.data_ptr(), self._s()), "pool_bwd")

    def tsgate_fwd(self, hpre, zpre, n_in, B, Dp, fused, z_out=None):
        self._ck(self.lib.bpm_tsgate_fwd(hpre.data_ptr(), zpre.data_ptr(), n_in, B, Dp, fused.data_ptr(), _ptr(z_out), self._s()), "tsgate_fwd")

    def tsgate_bwd(self, hpre, zpre, dfused, n_in, B, Dp, dhpre, dzpre):
        self._ck(self.lib.bpm_tsgate_bwd(hpre.data_ptr(), zpre.data_ptr(), dfused.data_ptr(), n_in, B, Dp, dhpre.data_ptr(), dzpre.data_ptr(),
                                         self._s()), "tsgate_bwd")

    def bce_fwd_bwd(self, logits, targets, pos_weight, B, Cc, grad_scale, loss, dlogits):
        self._ck(self.lib.bpm_bce_fwd_bwd(logits.data_ptr(), logits.stride(0), targets.data_ptr(), _ptr(pos_weight), B, Cc, float(grad_scale),
                                          loss.data_ptr(), dlogits.data_ptr(), self._s()), "bce_fwd_bwd")

    def timelin_fwd(self, x, W, bias, y, B, Tin, Tout, D):
        """y[b, t2, :] = bias[t2] + sum_t W[t2, t] x[b, t, :]   (x [B*Tin, ld], y [B*Tout, ld], W fp32 [Tout, Tin])"""
        self._ck(self.lib.bpm_timelin_fwd(_dt(x), x.data_ptr(), W.data_ptr(), 0 if bias is None else bias.data_ptr(), y.data_ptr(), B, Tin, Tout, D,
                                          x.shape[1], self._s()), "timelin_fwd")

    def timelin_bwd(self, dy, x, W, dx, accumulate_dx, dW, db, B, Tin, Tout, D):
        self._ck(self.lib.bpm_timelin_bwd(_dt(x), dy.data_ptr(), x.data_ptr(), W.data_ptr(), 0 if dx is None else dx.data_ptr(), int(accumulate_dx),
                                          0 if dW is None else dW.data_ptr(), 0 if db is None else db.data_ptr(), B, Tin, Tout, D, x.shape[1],
                                          self._s()), "timelin_bwd")

    # ------------------------------------------------------------------ AudioEncoder layout kernels (mmtr.py:93-108)
    def conv1d_im2col(self, x, B, Tin, C, KW, stride, col, Tout):
        """x [B*Tin, ldx] rows -> col [B*Tout, KW*C] (tap-major K)"""
        assert x.dtype == col.dtype and col.shape[1] == KW * C and col.is_contiguous()
        self._ck(self.lib.bpm_conv1d_im2col(x.data_ptr(), _dt(x), B, Tin, C, x.stride(0), KW, stride, col.data_ptr(), Tout, self._s()), "conv1d_im2col")

    def conv1d_col2im(self, dcol, B, Tin, C, KW, stride, Tout, dx):
        assert dx.dtype == torch.float32 and dcol.is_contiguous()
        self._ck(self.lib.bpm_conv1d_col2im(dcol.data_ptr(), _dt(dcol), B, Tin, C, KW, stride, Tout, dx.data_ptr(), dx.stride(0), self._s()), "conv1d_col2im")

    def conv1d_pack_weight(self, W, Wp):
        """W fp32 (Cout, Cin, KW) -> Wp (Cout, KW*Cin), K tap-major"""
        Cout, Cin, KW = W.shape
        assert W.dtype == torch.float32 and W.is_contiguous() and Wp.is_contiguous()
        self._ck(self.lib.bpm_conv1d_pack_weight(W.data_ptr(), Cout, Cin, KW, Wp.data_ptr(), _dt(Wp), self._s()), "conv1d_pack_weight")

    def conv1d_unpack_wgrad(self, gWp, gW, accumulate=False):
        Cout, Cin, KW = gW.shape
        assert gWp.dtype == torch.float32 and gW.dtype == torch.float32 and gW.is_contiguous() and gWp.is_contiguous()
        self._ck(self.lib.bpm_conv1d_unpack_wgrad(gWp.data_ptr(), Cout, Cin, KW, gW.data_ptr(), int(accumulate), self._s()), "conv1d_unpack_wgrad")

    def adaptive_pool_fwd(self, x, B, T, C, Tp, y):
        self._ck(self.lib.bpm_adaptive_pool_fwd(x.data_ptr(), _dt(x), B, T, C, x.stride(0), Tp, y.data_ptr(), _dt(y), y.stride(0), self._s()), "adaptive_pool_fwd")

    def adaptive_pool_bwd(self, dy, B, T, C, Tp, dx):
        assert dy.dtype == torch.float32 and dx.dtype == torch.float32
        self._ck(self.lib.bpm_adaptive_pool_bwd(dy.data_ptr(), B, T, C, Tp, dy.stride(0), dx.data_ptr(), dx.stride(0), self._s()), "adaptive_pool_bwd")

    def adam_step(self, param, grad, m, v, lr, beta1, beta2, eps, grad_scale, step_t, lr_t=None):
        self._ck(self.lib.bpm_adam_step(param.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), param.numel(), float(lr), float(beta1),
                                        float(beta2), float(eps), float(grad_scale), step_t.data_ptr(),
                                        0 if lr_t is None else lr_t.data_ptr(), self._s()), "adam_step")

    def zero_(self, t):
        """memset on the current stream (cudaMemsetAsync through torch; capturable); inside zero_begin() / zero_end() the tensors are
        collected and cleared by ONE multi-tensor launch (the ~100 gradient accumulators of a model: 0.6 ms of 6 us fills per step)."""
        if self._zlist is not None:
            self._zlist.append(t)
        else:
            t.zero_()

    def zero_begin(self):
        self._zdepth += 1
        if self._zlist is None:
            self._zlist = []

    def zero_end(self):
        self._zdepth -= 1
        if self._zdepth == 0:
            lst, self._zlist = self._zlist, None
            if lst:
                torch._foreach_zero_(lst)
