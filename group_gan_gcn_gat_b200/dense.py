"""Standalone dense-adjacency layers: GraphAttentionLayer.forward(h, adj) and one GCN layer relu((A H) W).

API parity for sgan/models.py:198-210 and 573-580 with an arbitrary dense ``adj``.  Every product runs
through ``sgx_gemm`` and the masked-softmax rows through ``sgx_dense_att_fwd/bwd``; torch is used only for
allocation and trivial elementwise glue (ELU / ReLU masks).  The encoders never take this path.
"""
import torch

from . import _lib
from .ops import _ptr, _stream


def _mm(a, b, relu=False):
    """a @ b for fp32 CUDA tensors with arbitrary 2-D strides (transposed views are fine)."""
    if not (a.is_cuda and b.is_cuda):
        raise RuntimeError('sgx dense layers need CUDA tensors (no CPU fallback)')
    a, b = a.float(), b.float()
    M, K = a.shape
    N = b.shape[1]
    c = torch.empty(M, N, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().sgx_gemm(_ptr(a), a.stride(0), a.stride(1), _ptr(b), b.stride(0), b.stride(1), _ptr(c), N,
                                       M, N, K, 0, int(relu), _stream(a)), 'sgx_gemm')
    return c


class _GatLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, adj, W, a, alpha, concat):
        h, adj, W, a = (t.contiguous().float() for t in (h, adj, W, a))
        n, f = h.shape[0], W.shape[1]
        wh = _mm(h, W)
        st = _mm(wh, a.reshape(2, f).t())               # [n,2]: s = Wh a[:F], t = Wh a[F:]
        att = torch.empty(n, n, dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            _lib.check(_lib.lib().sgx_dense_att_fwd(_ptr(st), _ptr(adj), n, alpha, _ptr(att), _stream(h)),
                       'sgx_dense_att_fwd')
        hp = _mm(att, wh)
        out = torch.nn.functional.elu(hp) if concat else hp
        ctx.save_for_backward(h, adj, W, a, wh, st, att, out)
        ctx.alpha, ctx.concat = alpha, concat
        return out

    @staticmethod
    def backward(ctx, gout):
        h, adj, W, a, wh, st, att, out = ctx.saved_tensors
        n, f = wh.shape
        dhp = gout.contiguous().float()
        if ctx.concat:
            dhp = dhp * torch.where(out > 0, torch.ones_like(out), out + 1.0)
        datt = _mm(dhp, wh.t())
        ds = torch.empty(n, dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            _lib.check(_lib.lib().sgx_dense_att_bwd(_ptr(st), _ptr(adj), _ptr(att), n, ctx.alpha, _ptr(datt), _ptr(ds),
                                                    _stream(h)), 'sgx_dense_att_bwd')
        ones = torch.ones(n, 1, dtype=torch.float32, device=h.device)
        dt = _mm(datt.t(), ones).reshape(n)             # column sums of d(pre)
        a1, a2 = a.reshape(2, f)[0], a.reshape(2, f)[1]
        dwh = _mm(att.t(), dhp) + ds[:, None] * a1[None, :] + dt[:, None] * a2[None, :]
        da = torch.cat([_mm(ds[None, :], wh), _mm(dt[None, :], wh)], dim=1).reshape(a.shape)
        dW = _mm(h.t(), dwh)
        dh = _mm(dwh, W.t())
        return dh, None, dW, da, None, None


def gat_layer(h, adj, W, a, alpha, concat):
    return _GatLayer.apply(h, adj, W, a, alpha, concat)


class _GcnLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, H, W):
        A, H, W = (t.contiguous().float() for t in (A, H, W))
        ah = _mm(A, H)
        out = _mm(ah, W, relu=True)
        ctx.save_for_backward(A, H, W, ah, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        A, H, W, ah, out = ctx.saved_tensors
        dz = gout.contiguous().float() * (out > 0).float()
        dW = _mm(ah.t(), dz)
        dah = _mm(dz, W.t())
        dH = _mm(A.t(), dah)
        dA = _mm(dah, H.t()) if ctx.needs_input_grad[0] else None
        return dA, dH, dW


def gcn_layer(A, H, W):
    return _GcnLayer.apply(A, H, W)
