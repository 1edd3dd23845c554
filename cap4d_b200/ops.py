"""torch-tensor wrappers of the single-kernel C-ABI entry points (building blocks of the U-Net
executor; used by the parity tests and micro-benchmarks).  Every function requires CUDA tensors and
runs on the current stream; `iters`/timing return the CUDA-event time per launch in ms.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib

OUT_F32, OUT_BF16, OUT_GEGLU = 0, 1, 2


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cap4d_b200.ops: CUDA tensors required (there is no CPU path)")


def gemm(a: torch.Tensor, w: torch.Tensor, bias=None, residual=None, out_mode: int = OUT_F32, time_iters: int = 0):
    """out = a[M,K] @ w[N,K]^T (+bias) (+residual); a, w bf16.  GEGLU: w/bias rows interleaved [x32|gate32]."""
    _check_cuda(a, w, bias, residual)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.is_contiguous() and w.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if out_mode == OUT_GEGLU else N
    out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_mode == OUT_F32 else torch.bfloat16)
    ms = ctypes.c_float(0)
    lib = _lib.load()
    _lib.check(
        lib.cap4d_b200_gemm_bf16(_ptr(a), _ptr(w), M, N, K, _ptr(bias), _ptr(residual), _ptr(out), out_mode,
                                 _stream(a.device), ctypes.byref(ms) if time_iters else None, max(1, time_iters)),
        "gemm_bf16",
    )
    return (out, ms.value) if time_iters else out


def gemm_mixed(a: torch.Tensor, w: torch.Tensor, bias=None, residual=None, out_mode: int = OUT_F32, time_iters: int = 0):
    """gemm() with each operand either bf16 or fp16 (the U-Net executor's combination is bf16 activations x fp16
    weights); the formats are taken from the tensors' dtypes."""
    _check_cuda(a, w, bias, residual)
    fmts = {torch.bfloat16: 0, torch.float16: 1}
    assert a.dtype in fmts and w.dtype in fmts and a.is_contiguous() and w.is_contiguous()
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if out_mode == OUT_GEGLU else N
    out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_mode == OUT_F32 else torch.bfloat16)
    ms = ctypes.c_float(0)
    lib = _lib.load()
    _lib.check(
        lib.cap4d_b200_gemm_mixed(_ptr(a), _ptr(w), M, N, K, fmts[a.dtype], fmts[w.dtype], _ptr(bias), _ptr(residual),
                                  _ptr(out), out_mode, _stream(a.device), ctypes.byref(ms) if time_iters else None,
                                  max(1, time_iters)),
        "gemm_mixed",
    )
    return (out, ms.value) if time_iters else out


def conv3x3(a_nhwc: torch.Tensor, w_packed: torch.Tensor, n_img: int, H_out: int, W_out: int, stride: int = 1,
            bias=None, rowbias=None, residual=None, time_iters: int = 0):
    """3x3 conv, pad 1.  a_nhwc: bf16 [n_img,H,W,Cin] (stride 2: parity planes [4,n_img,H/2,W/2,Cin]);
    w_packed: bf16 [Cout, 9*Cin] tap-major.  Returns fp32 [n_img*H_out*W_out, Cout]."""
    _check_cuda(a_nhwc, w_packed, bias, rowbias, residual)
    assert a_nhwc.dtype == torch.bfloat16 and w_packed.dtype == torch.bfloat16
    Cin = a_nhwc.shape[-1]
    Cout = w_packed.shape[0]
    out = torch.empty((n_img * H_out * W_out, Cout), device=a_nhwc.device, dtype=torch.float32)
    ms = ctypes.c_float(0)
    lib = _lib.load()
    _lib.check(
        lib.cap4d_b200_conv3x3_bf16(_ptr(a_nhwc), _ptr(w_packed), n_img, H_out, W_out, Cin, Cout, stride, _ptr(bias),
                                    _ptr(rowbias), _ptr(residual), _ptr(out), _stream(out.device),
                                    ctypes.byref(ms) if time_iters else None, max(1, time_iters)),
        "conv3x3_bf16",
    )
    return (out, ms.value) if time_iters else out


def upsample_conv3x3(a_nhwc: torch.Tensor, w_oihw: torch.Tensor, bias=None, time_iters: int = 0):
    """nearest-2x upsample + 3x3 conv (pad 1), folded into four 2x2-tap phase convs.
    a_nhwc: bf16 [n,H,W,Cin] (low resolution); w_oihw: fp32 [Cout,Cin,3,3].  Returns fp32 [n*2H*2W, Cout]."""
    _check_cuda(a_nhwc, w_oihw, bias)
    n, H, W, Cin = a_nhwc.shape
    Cout = w_oihw.shape[0]
    w32 = w_oihw.to(torch.float32).contiguous()
    out = torch.empty((n * 4 * H * W, Cout), device=a_nhwc.device, dtype=torch.float32)
    ms = ctypes.c_float(0)
    lib = _lib.load()
    _lib.check(
        lib.cap4d_b200_upsample_conv3x3_bf16(_ptr(a_nhwc), _ptr(w32), n, H, W, Cin, Cout, _ptr(bias), _ptr(out),
                                             _stream(out.device), ctypes.byref(ms) if time_iters else None,
                                             max(1, time_iters)),
        "upsample_conv3x3_bf16",
    )
    return (out, ms.value) if time_iters else out


def pack_conv_weight(w_oihw: torch.Tensor) -> torch.Tensor:
    """[O,I,3,3] fp32 -> bf16 [O, 9*I] with K index = (ky*3+kx)*I + i (what the conv kernel expects)."""
    O, I, KH, KW = w_oihw.shape
    return w_oihw.permute(0, 2, 3, 1).reshape(O, KH * KW * I).to(torch.bfloat16).contiguous()


def parity_planes(x_nhwc: torch.Tensor) -> torch.Tensor:
    """[n,H,W,C] -> [4,n,H/2,W/2,C], plane = (y&1)*2 + (x&1)."""
    return torch.stack([x_nhwc[:, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=0).contiguous()


def attention(qkv: torch.Tensor, C: int, L: int, scale: float = 0.125, time_iters: int = 0):
    """qkv bf16 [M, 3C]; sequences are L consecutive rows; returns bf16 [M, C]."""
    _check_cuda(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape[1] == 3 * C
    M = qkv.shape[0]
    out = torch.empty((M, C), device=qkv.device, dtype=torch.bfloat16)
    ms = ctypes.c_float(0)
    lib = _lib.load()
    _lib.check(
        lib.cap4d_b200_attention_bf16(_ptr(qkv), _ptr(out), M, C, L, float(scale), _stream(qkv.device),
                                      ctypes.byref(ms) if time_iters else None, max(1, time_iters)),
        "attention_bf16",
    )
    return (out, ms.value) if time_iters else out


def groupnorm(x1: torch.Tensor, x2: Optional[torch.Tensor], n_img: int, hw: int, gamma, beta, eps: float, silu: bool,
              want_raw: bool = False, time_iters: int = 0):
    """GroupNorm(32) (+SiLU) over the channel concat of NHWC fp32 [n_img*hw, C1] (and [.., C2]) -> bf16."""
    _check_cuda(x1, x2, gamma, beta)
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    out = torch.empty((n_img * hw, C1 + C2), device=x1.device, dtype=torch.bfloat16)
    raw = torch.empty_like(out) if want_raw else None
    ms = ctypes.c_float(0)
    lib = _lib.load()
    _lib.check(
        lib.cap4d_b200_groupnorm_bf16(_ptr(x1), C1, _ptr(x2), C2, n_img, hw, _ptr(gamma), _ptr(beta), float(eps),
                                      int(silu), _ptr(out), _ptr(raw), _stream(x1.device),
                                      ctypes.byref(ms) if time_iters else None, max(1, time_iters)),
        "groupnorm_bf16",
    )
    res: Tuple = (out, raw) if want_raw else (out,)
    if time_iters:
        res = res + (ms.value,)
    return res if len(res) > 1 else res[0]


def layernorm(x: torch.Tensor, gamma, beta, eps: float = 1e-5, time_iters: int = 0):
    _check_cuda(x, gamma, beta)
    M, C = x.shape
    out = torch.empty((M, C), device=x.device, dtype=torch.bfloat16)
    ms = ctypes.c_float(0)
    lib = _lib.load()
    _lib.check(
        lib.cap4d_b200_layernorm_bf16(_ptr(x), M, C, _ptr(gamma), _ptr(beta), float(eps), _ptr(out),
                                      _stream(x.device), ctypes.byref(ms) if time_iters else None, max(1, time_iters)),
        "layernorm_bf16",
    )
    return (out, ms.value) if time_iters else out


def cfg_ddim_update(latents, eps, gen_idx, n_groups, V, R, cfg_scale, x_coef, e_coef):
    _check_cuda(latents, eps, gen_idx)
    chw = latents[0].numel()
    lib = _lib.load()
    _lib.check(
        lib.cap4d_b200_cfg_ddim_update(_ptr(latents), _ptr(eps), _ptr(gen_idx), n_groups, V, R, chw, float(cfg_scale),
                                       float(x_coef), float(e_coef), _stream(latents.device)),
        "cfg_ddim_update",
    )
