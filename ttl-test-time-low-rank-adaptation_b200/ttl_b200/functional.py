"""The loss-head functions of the reference, backed by the CUDA library (no PyTorch arithmetic on the path):
    select_confident_samples (ttl.py:50-54), avg_entropy (ttl.py:56-61), softmax_entropy (deyo.py:85-90) and the
    default-flag weighted-entropy loss of forward_and_adapt_sar (deyo.py:97-181).
Each is a thin torch.autograd.Function over a C-ABI kernel, so the reference's own control flow can call them."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch

from . import _lib as L


def _st() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _need_cuda(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("ttl_b200 functions need CUDA tensors (sm_100a); there is no CPU fallback")
    return t.detach().to(torch.float32).contiguous()


def softmax_entropy(x: torch.Tensor) -> torch.Tensor:
    """Per-row entropy of softmax(x); no gradient (the fused losses below own the gradient path)."""
    lg = _need_cuda(x)
    ent = torch.empty(lg.shape[0], device=lg.device, dtype=torch.float32)
    L.check(L.load().ttl_op_entropy(_p(lg), _p(ent), lg.shape[0], lg.shape[1], _st()))
    return ent


def select_indices(entropy: torch.Tensor, k: int) -> torch.Tensor:
    """argsort(entropy)[:k] with lowest-index-first tie-break (the reference's unstable argsort leaves ties open)."""
    e = _need_cuda(entropy)
    idx = torch.empty(max(k, 1), device=e.device, dtype=torch.int32)
    if k > 0:
        L.check(L.load().ttl_op_select(_p(e), e.shape[0], k, _p(idx), _st()))
    return idx[:k].to(torch.int64)


def select_confident_samples(logits: torch.Tensor, top: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """ttl.py:50-54: (logits[idx], idx) for the int(V*top) lowest-entropy views."""
    ent = softmax_entropy(logits)
    idx = select_indices(ent, int(ent.size(0) * top))
    return logits[idx], idx


class _AvgEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs):
        x = _need_cuda(outputs)
        K, Cn = x.shape
        loss = torch.empty(1, device=x.device, dtype=torch.float32)
        dl = torch.empty(K, Cn, device=x.device, dtype=torch.float32)
        L.check(L.load().ttl_op_tpt_loss(_p(x), None, K, Cn, _p(loss), _p(dl), _st()))
        ctx.save_for_backward(dl)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g


def avg_entropy(outputs: torch.Tensor) -> torch.Tensor:
    """ttl.py:56-61: entropy of the view-averaged distribution; gradient from the same kernel."""
    return _AvgEntropy.apply(outputs)


class _DeyoLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, outputs, margin_e0):
        x = _need_cuda(outputs)
        V, Cn = x.shape
        loss = torch.empty(1, device=x.device, dtype=torch.float32)
        dl = torch.empty(V, Cn, device=x.device, dtype=torch.float32)
        L.check(L.load().ttl_op_deyo_loss(_p(x), V, Cn, float(margin_e0), _p(loss), _p(dl), _st()))
        ctx.save_for_backward(dl)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None


def deyo_weighted_entropy(outputs: torch.Tensor, margin_e0: float = 0.4) -> torch.Tensor:
    """deyo.py:102-181 with filter_ent=0, filter_plpd=0, reweight_ent=1: mean_v( H_v * exp(-(H_v - e0)) ), H <= ln 1000."""
    return _DeyoLoss.apply(outputs, margin_e0)
