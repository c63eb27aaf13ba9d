"""Helpers for the -m gpu parity tests: call the C ABI (ctypes) on torch CUDA tensors."""
import ctypes as C

import torch

from ttl_b200 import _lib as L


def lib():
    return L.load()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ok(rc):
    if rc != 0:
        raise RuntimeError(f"rc={rc}: {lib().ttl_last_error(None).decode()}")


def gemm(a, b, epi=L.EPI_BF16, bias=None, a2=None, b2=None, resid=None, aux=None, pos=None, tpv=0, block_n=0,
         want_out2=False, out_rows=None):
    """a [M,K] bf16, b [N,K] bf16 -> (out, out2)."""
    M, K = a.shape
    N = b.shape[0]
    f32_out = epi in (L.EPI_RESID_F32, L.EPI_PATCH_F32, L.EPI_F32)
    rows = out_rows or M
    out = torch.zeros(rows, N, device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    out2 = torch.zeros(M, N, device=a.device, dtype=torch.bfloat16) if want_out2 else None
    K2 = a2.shape[1] if a2 is not None else 0
    ok(lib().ttl_op_gemm(ptr(a), ptr(b), ptr(a2), ptr(b2), M, N, K, K2, epi, ptr(bias), ptr(out), ptr(out2), ptr(resid),
                         ptr(aux), ptr(pos), tpv, block_n, stream()))
    torch.cuda.synchronize()
    return out, out2


def rel_err(got, ref):
    got, ref = got.double(), ref.double()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))
