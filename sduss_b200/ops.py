"""Thin Python wrappers over the C ABI: they pass data_ptr()s and the current CUDA stream and
raise on a non-zero status. PyTorch is plumbing here (memory + streams); all arithmetic
happens inside libsduss_b200.so."""
import ctypes

import torch

from . import _lib
from ._lib import EpilogueDesc, check, lib

EPI_BIAS, EPI_GELU_TANH, EPI_GATE_RESID, EPI_QK_RMSNORM, EPI_GEGLU, EPI_ROWVEC = range(6)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req(t, dtype=torch.bfloat16):
    assert t.is_cuda and t.dtype == dtype, (t.device, t.dtype)
    return t


def gemm(a, w, out=None, *, epi=EPI_BIAS, bias=None, resid=None, gate=None, row_group=None,
         rowvec=None, rms_wq=None, rms_wk=None, rms_q_cols=0, rms_k_cols=0, rms_eps=1e-6,
         q_scale=1.0, out_fp32=False):
    """out = epilogue(a @ w.T). a: [M, K] bf16 (row stride may exceed K), w: [N, K] bf16."""
    _req(a), _req(w)
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    n_out = N // 2 if epi == EPI_GEGLU else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device,
                          dtype=torch.float32 if out_fp32 else torch.bfloat16)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    d = EpilogueDesc()
    d.C, d.ldc, d.out_fp32 = _ptr(out), out.stride(0), int(out.dtype == torch.float32)
    d.bias = _ptr(bias)
    d.resid, d.ldr = _ptr(resid), (resid.stride(0) if resid is not None else 0)
    d.gate, d.ldg = _ptr(gate), (gate.stride(0) if gate is not None else 0)
    d.row_group = _ptr(row_group)
    d.rowvec, d.ldv = _ptr(rowvec), (rowvec.stride(0) if rowvec is not None else 0)
    d.rms_wq, d.rms_wk = _ptr(rms_wq), _ptr(rms_wk)
    d.rms_q_cols, d.rms_k_cols = rms_q_cols, rms_k_cols
    d.rms_eps, d.q_scale = rms_eps, q_scale
    check(lib.b200_gemm_bf16(_ptr(a), a.stride(0), _ptr(w), w.stride(0), M, N, K, epi,
                             ctypes.byref(d), _stream()), "b200_gemm_bf16")
    return out
