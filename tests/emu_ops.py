"""TEST INFRASTRUCTURE ONLY: a pure-torch (CPU) emulation of the op contract of bpmult_b200.ops.CudaOps.

It exists so that the HOST-SIDE logic of the product (engine.py / model_engine.py: kernel sequencing, padded layouts,
hand-derived backward, dropout-site bookkeeping) can be checked against the oracle on machines without a GPU, and so
that GPU kernel tests have a second, op-level statement of each contract.  The product never imports this file."""
import torch

MASK32 = 0xFFFFFFFF


def mix32(x):
    """vectorised over int64 tensors holding uint32 values; mirrors mix32() in csrc/bpm_common.cuh"""
    x = x ^ (x >> 16)
    x = (x * 0x7feb352d) & MASK32
    x = x ^ (x >> 15)
    x = (x * 0x846ca68b) & MASK32
    x = x ^ (x >> 16)
    return x


def _mix_int(x):
    return int(mix32(torch.tensor([x & MASK32], dtype=torch.int64))[0])


def drop_mult(drop, idx):
    """multiplier (0 or 1/(1-p)) for int64 element indices `idx` (any shape): 16-bit decisions from a multiply-xorshift-multiply
    hash of the element-pair counter (see csrc/bpm_common.cuh)"""
    if drop is None or drop.p <= 0:
        return torch.ones(idx.shape, dtype=torch.float32)
    seed = int(drop.seed_ptr.item()) if drop.seed_ptr is not None else int(drop.seed)
    seed &= 0xFFFFFFFFFFFFFFFF
    site = int(drop.site)
    k0 = _mix_int((seed & MASK32) ^ _mix_int((site & MASK32) + 0x9E3779B9))
    k1 = _mix_int(((seed >> 32) & MASK32) ^ _mix_int(((site >> 32) & MASK32) + 0x85EBCA6B) ^ 0xC2B2AE35)
    pair = idx >> 1
    lo, hi = pair & MASK32, (pair >> 32) & MASK32
    x = (lo * 0x7feb352d + (k0 ^ ((hi * 0x85ebca77) & MASK32))) & MASK32
    x = x ^ (x >> 16)
    r = (x * 0x846ca68b + k1) & MASK32
    hw = torch.where((idx & 1) == 1, r >> 16, r & 0xFFFF)
    p32 = float(torch.tensor(drop.p, dtype=torch.float32))
    thresh = 65536 if p32 >= 1 else int(p32 * 65536.0 + 0.5)
    inv = torch.tensor(1.0, dtype=torch.float32) / (torch.tensor(1.0, dtype=torch.float32) - torch.tensor(drop.p, dtype=torch.float32))
    return torch.where(hw >= thresh, inv, torch.zeros((), dtype=torch.float32))


def _idx2d(rows, ld, cols):
    return torch.arange(rows).unsqueeze(1) * ld + torch.arange(cols).unsqueeze(0)


def _remap(n, dh, dhp):
    i = torch.arange(n)
    return (i // dh) * dhp + (i % dh) if dh > 0 else i


class EmuOps:
    name = "emu"

    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0

    def empty(self, shape, dtype):
        if not dtype.is_floating_point:
            return torch.zeros(shape, dtype=dtype)
        return torch.full(shape, float("nan"), dtype=dtype)          # poison: catches reads of unwritten buffers

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def zero_(self, t):
        t.zero_()

    def zero_begin(self):
        pass

    def zero_end(self):
        pass

    def batch_begin(self, mode, key):
        pass

    def batch_end(self):
        pass

    # ------------------------------------------------------------------ weight staging
    def fold_batch_begin(self, mode):
        pass

    def fold_batch_end(self):
        pass

    def ln_fold_fwd(self, W, bias, gamma, beta, Wp, bp, row_map=(0, 0)):
        r = _remap(W.shape[0], *row_map)
        Wp[r, :W.shape[1]] = (W * gamma.unsqueeze(0)).to(Wp.dtype)
        bp[r] = bias + W @ beta

    def ln_fold_bwd(self, W, gamma, beta, gWf, gbf, gW, gb, dgamma, dbeta, row_map=(0, 0)):
        r = _remap(W.shape[0], *row_map)
        c = W.shape[1]
        t = gWf[r, :c]
        g = gbf[r]
        dgamma[:c] += (t * W).sum(0)
        dbeta[:c] += (g.unsqueeze(1) * W).sum(0)
        gW[r, :c] += t * gamma.unsqueeze(0) + g.unsqueeze(1) * beta.unsqueeze(0)
        gb[r] += g

    def pack_matrix(self, src, dst, row_map=(0, 0), col_map=(0, 0)):
        dst.zero_()
        r, c = _remap(src.shape[0], *row_map), _remap(src.shape[1], *col_map)
        dst[r.unsqueeze(1), c.unsqueeze(0)] = src.to(dst.dtype)

    def unpack_matrix(self, src_p, dst, row_map=(0, 0), col_map=(0, 0), accumulate=False, scale=1.0):
        r, c = _remap(dst.shape[0], *row_map), _remap(dst.shape[1], *col_map)
        v = src_p[r.unsqueeze(1), c.unsqueeze(0)] * scale
        if accumulate:
            dst += v
        else:
            dst.copy_(v)

    # ------------------------------------------------------------------ staging / embed
    def stage_rows(self, src, dst, Tp, drop=None):
        B, T, C = src.shape
        Cp = dst.shape[1]
        d3 = dst.view(B, Tp, Cp)
        d3.zero_()
        idx = (torch.arange(B).view(B, 1, 1) * Tp + torch.arange(T).view(1, T, 1)) * Cp + torch.arange(C).view(1, 1, C)
        d3[:, :T, :C] = (src * drop_mult(drop, idx)).to(dst.dtype)

    def unstage_rows(self, g, dsrc, Tp, accumulate=False, drop=None):
        B, T, C = dsrc.shape
        Cp = g.shape[1]
        idx = (torch.arange(B).view(B, 1, 1) * Tp + torch.arange(T).view(1, T, 1)) * Cp + torch.arange(C).view(1, 1, C)
        v = g.view(B, Tp, Cp)[:, :T, :C] * drop_mult(drop, idx)
        if accumulate:
            dsrc += v
        else:
            dsrc.copy_(v)

    def embed_fwd(self, x, pe, B, T, D, scale, y, drop=None):
        rows, Dp = x.shape
        xf = x.float()
        t = torch.arange(rows) % T
        pos = torch.where(xf[:, 0] != 0, t + 1, torch.zeros_like(t))
        v = (scale * xf + pe[pos]) * drop_mult(drop, _idx2d(rows, Dp, Dp))
        v[:, D:] = 0
        y.copy_(v.to(y.dtype))

    def embed_bwd(self, dy, D, scale, dx, accumulate, drop=None):
        rows, Dp = dy.shape
        v = scale * dy * drop_mult(drop, _idx2d(rows, Dp, Dp))
        v[:, D:] = 0
        if accumulate:
            dx += v
        else:
            dx.copy_(v)

    # ------------------------------------------------------------------ layernorm
    def layernorm_fwd(self, x, gamma, beta, D, y, mean, rstd, eps=1e-5):
        xf = x.float()[:, :D]
        mu = xf.mean(1)
        var = ((xf - mu[:, None]) ** 2).mean(1)
        rs = torch.rsqrt(var + eps)
        out = torch.zeros(x.shape, dtype=torch.float32)
        out[:, :D] = (xf - mu[:, None]) * rs[:, None] * gamma[:D] + beta[:D]
        y.copy_(out.to(y.dtype))
        mean.copy_(mu)
        rstd.copy_(rs)

    def layernorm_bwd(self, dy, x, mean, rstd, gamma, D, dx, accumulate, dgamma, dbeta, cast_out=None, cast_drop=None):
        g = dy.float()[:, :D]
        xh = (x.float()[:, :D] - mean[:, None]) * rstd[:, None]
        gh = g * gamma[:D]
        s1 = gh.mean(1, keepdim=True)
        s2 = (gh * xh).mean(1, keepdim=True)
        v = torch.zeros(x.shape, dtype=torch.float32)
        v[:, :D] = rstd[:, None] * (gh - s1 - xh * s2)
        if accumulate:
            dx += v
        else:
            dx.copy_(v)
        dgamma[:D] += (g * xh).sum(0)
        dbeta[:D] += g.sum(0)
        if cast_out is not None:
            self.cast_drop(dx, cast_out, cast_drop)

    # ------------------------------------------------------------------ gemm
    def gemm(self, A, B, Cout, M, N, K, ta=0, tb=0, bias=None, alpha=1.0, act=0, drop=None, gate=None, gate_scale=1.0, residual=None,
             accumulate=False, split_k=0, colsum=None):
        a = (A[:K, :M].float().t() if ta else A[:M, :K].float())
        if colsum is not None:
            assert ta == 1
            colsum[:M] += a.sum(1)
        b = (B[:K, :N].float() if tb else B[:N, :K].float().t())
        v = a @ b
        if bias is not None:
            v = v + bias[:N]
        v = v * alpha
        if act == 1:
            v = torch.relu(v)
        if drop is not None and drop.p > 0:
            v = v * drop_mult(drop, _idx2d(M, Cout.stride(0), N))
        if gate is not None:
            v = torch.where(gate[:M, :N].float() > 0, v * gate_scale, torch.zeros_like(v))
        if residual is not None:
            v = v + residual[:M, :N].float()
        if accumulate:
            Cout[:M, :N] += v
        else:
            Cout[:M, :N] = v.to(Cout.dtype)

    def colsum(self, X, N, out):
        out[:N] += X[:, :N].float().sum(0)

    # ------------------------------------------------------------------ attention
    def _probs(self, q, k, B, T, S, H, dhp, mask_off, key_pad):
        qh = q.float().view(B, T, H, dhp).permute(0, 2, 1, 3)
        kh = k.float().reshape(B, S, H, dhp).permute(0, 2, 1, 3)
        s = qh @ kh.transpose(-1, -2)                                  # [B,H,T,S]
        if mask_off >= 0:
            i, j = torch.arange(T).view(T, 1), torch.arange(S).view(1, S)
            s = s.masked_fill(j > i + mask_off, float("-inf"))
        if key_pad is not None:
            s = s.masked_fill(key_pad.view(B, 1, 1, S).bool(), float("-inf"))
        return s

    def _attn_mult(self, B, T, S, H, drop):
        idx = torch.arange(B * H * T * S).view(B, H, T, S)
        return drop_mult(drop, idx)

    def xattn_fwd(self, q, k, v, out, lse, B, T, S, H, dh, dhp, mask_off=-1, key_pad=None, drop=None, drop_bits=None):
        s = self._probs(q, k, B, T, S, H, dhp, mask_off, key_pad)
        L = torch.logsumexp(s, -1)
        p = torch.exp(s - L.unsqueeze(-1)) * self._attn_mult(B, T, S, H, drop)
        vh = v.float().reshape(B, S, H, dhp).permute(0, 2, 1, 3)
        o = (p @ vh).permute(0, 2, 1, 3).reshape(B * T, H * dhp)
        out.copy_(o.to(out.dtype))
        lse.copy_(L.reshape(-1))

    def xattn_bwd_workspace(self, dtype, B, T, S, H, dh, dhp):
        return 2 * B * H * T

    def xattn_bwd(self, q, k, v, out, dout, lse, delta, dq, dq_scale, dk, dv, B, T, S, H, dh, dhp, mask_off=-1, key_pad=None, drop=None, drop_bits=None):
        s = self._probs(q, k, B, T, S, H, dhp, mask_off, key_pad)
        p = torch.exp(s - lse.view(B, H, T, 1))
        mult = self._attn_mult(B, T, S, H, drop)
        qh = q.float().view(B, T, H, dhp).permute(0, 2, 1, 3)
        kh = k.float().reshape(B, S, H, dhp).permute(0, 2, 1, 3)
        vh = v.float().reshape(B, S, H, dhp).permute(0, 2, 1, 3)
        go = dout.float().view(B, T, H, dhp).permute(0, 2, 1, 3)
        oh = out.float().view(B, T, H, dhp).permute(0, 2, 1, 3)
        dl = (go * oh).sum(-1, keepdim=True)
        dv_ = (p * mult).transpose(-1, -2) @ go
        dpt = go @ vh.transpose(-1, -2)
        ds = p * (dpt * mult - dl)
        dq_ = ds @ kh * dq_scale
        dk_ = ds.transpose(-1, -2) @ qh
        n = B * H * T
        delta[:n].copy_(dl.reshape(-1))
        delta[n:2 * n].copy_((lse.float() * 1.4426950408889634).reshape(-1))      # workspace [1]: lse * log2e
        dq.copy_(dq_.permute(0, 2, 1, 3).reshape(B * T, H * dhp).to(dq.dtype))
        dk.copy_(dk_.permute(0, 2, 1, 3).reshape(B * S, H * dhp).to(dk.dtype))
        dv.copy_(dv_.permute(0, 2, 1, 3).reshape(B * S, H * dhp).to(dv.dtype))

    def xattn_weights(self, q, k, lse, w, B, T, S, H, dh, dhp, mask_off=-1, key_pad=None, drop=None):
        s = self._probs(q, k, B, T, S, H, dhp, mask_off, key_pad)
        p = torch.exp(s - lse.view(B, H, T, 1)) * self._attn_mult(B, T, S, H, drop)
        w.copy_(p.sum(1) / H)

    # ------------------------------------------------------------------ GMU / elementwise
    def gmu_fwd(self, features, a1, a2, h1p, h2p, zp, addend, y, z_out=None):
        z, t1, t2 = torch.sigmoid(zp.float()), torch.tanh(h1p.float()), torch.tanh(h2p.float())
        v = z * t1 * a1.float() + (1 - z) * t2 * a2.float() if features else z * t1 + (1 - z) * t2
        if addend is not None:
            v = v + addend.float()
        y.copy_(v.to(y.dtype))
        if z_out is not None:
            z_out.copy_(z.to(z_out.dtype))

    def gmu_bwd(self, features, a1, a2, h1p, h2p, zp, dy, dh1, dh2, dz, da1, da2):
        z, t1, t2 = torch.sigmoid(zp.float()), torch.tanh(h1p.float()), torch.tanh(h2p.float())
        u1 = a1.float() if features else torch.ones_like(z)
        u2 = a2.float() if features else torch.ones_like(z)
        dh1.copy_((dy * z * u1 * (1 - t1 * t1)).to(dh1.dtype))
        dh2.copy_((dy * (1 - z) * u2 * (1 - t2 * t2)).to(dh2.dtype))
        dz.copy_((dy * (t1 * u1 - t2 * u2) * z * (1 - z)).to(dz.dtype))
        if features:
            da1 += dy * z * t1
            da2 += dy * (1 - z) * t2

    def add(self, a, b, y):
        y.copy_((a.float() + b.float()).to(y.dtype))

    def axpy_f32(self, src, dst, accumulate=True):
        if accumulate:
            dst += src.float()
        else:
            dst.copy_(src.float())

    def cast_drop(self, x, y, drop=None):
        rows, cols = x.shape
        y.copy_((x * drop_mult(drop, _idx2d(rows, cols, cols))).to(y.dtype))

    def pool_fwd(self, x, B, T, out, col_off):
        Dp = x.shape[1]
        x3 = x.float().view(B, T, Dp)
        out[:, col_off:col_off + Dp] = x3[:, 0] + x3[:, T - 1]

    def pool_bwd(self, dout, col_off, B, T, dx):
        Dp = dx.shape[1]
        g = dout[:, col_off:col_off + Dp]
        d3 = dx.view(B, T, Dp)
        d3[:, 0] += g
        d3[:, T - 1] += g

    def tsgate_fwd(self, hpre, zpre, n_in, B, Dp, fused, z_out=None):
        z = torch.sigmoid(zpre)
        fused.copy_((z * torch.tanh(hpre)).sum(0))
        if z_out is not None:
            z_out.copy_(z.permute(1, 0, 2).reshape(B, n_in * Dp))

    def tsgate_bwd(self, hpre, zpre, dfused, n_in, B, Dp, dhpre, dzpre):
        z, t = torch.sigmoid(zpre), torch.tanh(hpre)
        dhpre.copy_(dfused.unsqueeze(0) * z * (1 - t * t))
        dzpre.copy_(dfused.unsqueeze(0) * t * z * (1 - z))

    def bce_fwd_bwd(self, logits, targets, pos_weight, B, Cc, grad_scale, loss, dlogits):
        x = logits[:, :Cc].detach().clone().requires_grad_()
        l = torch.nn.functional.binary_cross_entropy_with_logits(x, targets, pos_weight=pos_weight)
        l.backward()
        loss.copy_(l.detach().view(1))
        dlogits.zero_()
        dlogits[:, :Cc] = x.grad * grad_scale

    def timelin_fwd(self, x, W, bias, y, B, Tin, Tout, D):
        xb = x.float().view(B, Tin, -1)
        r = torch.einsum("ot,btd->bod", W, xb)
        if bias is not None:
            r = r + bias.view(1, -1, 1)
        r[:, :, D:] = 0
        y.copy_(r.reshape(B * Tout, -1).to(y.dtype))

    def timelin_bwd(self, dy, x, W, dx, accumulate_dx, dW, db, B, Tin, Tout, D):
        g = dy.float().view(B, Tout, -1).clone()
        g[:, :, D:] = 0
        xb = x.float().view(B, Tin, -1)
        if dx is not None:
            r = torch.einsum("ot,bod->btd", W, g).reshape(B * Tin, -1)
            dx.copy_(dx + r if accumulate_dx else r)
        if dW is not None:
            dW += torch.einsum("bod,btd->ot", g, xb)
        if db is not None:
            db += g.sum(dim=(0, 2))

    # ------------------------------------------------------------------ AudioEncoder layout kernels
    def conv1d_im2col(self, x, B, Tin, C, KW, stride, col, Tout):
        xv = x.view(B, Tin, x.shape[1])[:, :, :C]
        for t in range(Tout):
            col.view(B, Tout, KW, C)[:, t] = xv[:, stride * t:stride * t + KW]

    def conv1d_col2im(self, dcol, B, Tin, C, KW, stride, Tout, dx):
        d = dcol.float().view(B, Tout, KW, C)
        out = torch.zeros(B, Tin, C)
        for t in range(Tout):
            out[:, stride * t:stride * t + KW] += d[:, t]
        dx.view(B, Tin, dx.shape[1])[:, :, :C] = out

    def conv1d_pack_weight(self, W, Wp):
        Wp.copy_(W.permute(0, 2, 1).reshape(W.shape[0], -1).to(Wp.dtype))

    def conv1d_unpack_wgrad(self, gWp, gW, accumulate=False):
        Cout, Cin, KW = gW.shape
        g = gWp.view(Cout, KW, Cin).permute(0, 2, 1)
        gW.copy_(gW + g if accumulate else g)

    def adaptive_pool_fwd(self, x, B, T, C, Tp, y):
        xv = x.float().view(B, T, x.shape[1])[:, :, :C].permute(0, 2, 1)
        y.view(B, Tp, y.shape[1])[:, :, :C] = torch.nn.functional.adaptive_avg_pool1d(xv, Tp).permute(0, 2, 1).to(y.dtype)

    def adaptive_pool_bwd(self, dy, B, T, C, Tp, dx):
        d = dy.view(B, Tp, dy.shape[1])[:, :, :C]
        out = torch.zeros(B, T, C)
        for i in range(Tp):                                    # torch's AdaptiveAvgPool1d windows: [floor(i*T/Tp), ceil((i+1)*T/Tp))
            t0, t1 = (i * T) // Tp, -((-(i + 1) * T) // Tp)
            out[:, t0:t1] += d[:, i:i + 1] / (t1 - t0)
        dx.view(B, T, dx.shape[1])[:, :, :C] = out

    def adam_step(self, param, grad, m, v, lr, beta1, beta2, eps, grad_scale, step_t, lr_t=None):
        step = float(step_t.item())
        if lr_t is not None:
            lr = float(lr_t.item())
        g = grad * grad_scale
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
        param.sub_((lr / bc1) * m / (v.sqrt() / bc2 ** 0.5 + eps))
