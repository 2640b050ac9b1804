"""Reference-facing operators backed by the CUDA library (module mode, autograd-capable).

Mirrors: tcnn.NetworkWithInputEncoding (src/models/immoco.py:60-65), FFT / IFFT
(src/utils/data_utils.py:29-34), GradientEntropyLoss (src/utils/losses.py:20-40).
All operators require CUDA tensors and raise otherwise (no CPU fallback by design).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _native as nat
from .encoding import GridSpec, MlpSpec, grid_spec, mlp_spec, twiddles, OUT_PAD, N_ENCODED


def _stream(device=None) -> int:
    """Raw handle of torch's current stream on ``device`` (default: the current device)."""
    return torch.cuda.current_stream(device).cuda_stream


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (the B200 path has no CPU fallback)")


_TW_CACHE: Dict[Tuple[int, str], torch.Tensor] = {}


def twiddle_table(n: int, device) -> torch.Tensor:
    key = (n, str(device))
    t = _TW_CACHE.get(key)
    if t is None:
        t = torch.from_numpy(twiddles(n)).to(device)
        _TW_CACHE[key] = t
    return t


# ------------------------------------------------------------------------------------------------
# tcnn.NetworkWithInputEncoding
# ------------------------------------------------------------------------------------------------
def init_inr_params(grid: GridSpec, mlp: MlpSpec, seed: int, device="cpu") -> torch.Tensor:
    """[W1 | W2 | table] fp32: MLP Xavier-uniform, table U(-1e-4, 1e-4) (tiny-cuda-nn's init
    distributions; its RNG stream itself is not reproducible, so the seed is ours).  Generated on
    the target device, like tiny-cuda-nn does."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    b1 = math.sqrt(6.0 / (N_ENCODED + mlp.width))
    b2 = math.sqrt(6.0 / (mlp.width + OUT_PAD))
    out = torch.empty(mlp.n_params + grid.n_table_params, dtype=torch.float32, device=device)
    out[: mlp.n_w1].uniform_(-b1, b1, generator=g)
    out[mlp.n_w1: mlp.n_params].uniform_(-b2, b2, generator=g)
    out[mlp.n_params:].uniform_(-1e-4, 1e-4, generator=g)
    return out


class _InrFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, grid_desc, mlp, n_out):
        _need_cuda(x, "NetworkWithInputEncoding input")
        _need_cuda(params, "NetworkWithInputEncoding params")
        lib = nat.lib()
        x = x.detach().float().contiguous()
        p = params.detach()
        n = x.shape[0]
        enc = torch.empty((N_ENCODED // 2, n, 2), dtype=torch.float32, device=x.device)
        out = torch.empty((n, 2), dtype=torch.float32, device=x.device)
        w1 = p.data_ptr()
        w2 = w1 + 4 * mlp.n_w1
        table = w1 + 4 * mlp.n_params
        with torch.cuda.device(x.device):       # the library launches on the CURRENT device: make it the tensors'
            s = _stream(x.device)
            nat.check(lib.immoco_hashgrid_fwd(grid_desc, x.data_ptr(), table, enc.data_ptr(), n, s), "hashgrid_fwd")
            nat.check(lib.immoco_mlp_fwd(enc.data_ptr(), w1, w2, out.data_ptr(), n, mlp.width, mlp.act, 0, s), "mlp_fwd")
        ctx.save_for_backward(x, params, enc)
        ctx.grid_desc, ctx.mlp, ctx.n_out = grid_desc, mlp, n_out
        return out[:, :n_out]

    @staticmethod
    def backward(ctx, grad_out):
        x, params, enc = ctx.saved_tensors
        mlp, n_out = ctx.mlp, ctx.n_out
        lib = nat.lib()
        n = x.shape[0]
        d_out = torch.zeros((n, 2), dtype=torch.float32, device=x.device)
        d_out[:, :n_out] = grad_out.float()
        d_enc = torch.empty_like(enc)
        grads = torch.zeros_like(params)
        p = params.detach()
        w1 = p.data_ptr()
        w2 = w1 + 4 * mlp.n_w1
        g1 = grads.data_ptr()
        g2 = g1 + 4 * mlp.n_w1
        gt = g1 + 4 * mlp.n_params
        with torch.cuda.device(x.device):
            s = _stream(x.device)
            nat.check(lib.immoco_mlp_bwd(enc.data_ptr(), w1, w2, d_out.data_ptr(), d_enc.data_ptr(), g1, g2, n,
                                         mlp.width, mlp.act, s), "mlp_bwd")
            nat.check(lib.immoco_hashgrid_bwd(ctx.grid_desc, x.data_ptr(), d_enc.data_ptr(), gt, n, s), "hashgrid_bwd")
        return None, grads, None, None, None


class NetworkWithInputEncoding(nn.Module):
    """Drop-in for ``tinycudann.NetworkWithInputEncoding(n_in, n_out, encoding_config, network_config)``.

    One flat fp32 ``params`` Parameter laid out [W1 | W2 (16 rows, padded) | hash table] on the
    current CUDA device; ``forward((N, n_in)) -> (N, n_out)`` fp32.  Unknown encoding keys (e.g.
    ``fine_resolution``) are ignored like tiny-cuda-nn does.  Gradients w.r.t. the input
    coordinates are not provided (the reference never needs them: immoco.py:72-80 are buffers).
    """

    def __init__(self, n_input_dims: int, n_output_dims: int, encoding_config: dict,
                 network_config: dict, seed: int = 1337, device: Optional[torch.device] = None):
        super().__init__()
        if not 1 <= n_output_dims <= 2:
            raise NotImplementedError("n_output_dims must be 1 or 2 on the IM-MoCo path")
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.encoding_config = dict(encoding_config)
        self.network_config = dict(network_config)
        self.grid = grid_spec(n_input_dims, encoding_config)
        self.mlp = mlp_spec(network_config)
        self.seed = seed
        self._desc = self.grid.desc()
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("NetworkWithInputEncoding needs a CUDA device (no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        self.params = nn.Parameter(init_inr_params(self.grid, self.mlp, seed, device))

    @property
    def n_params(self) -> int:
        return self.mlp.n_params + self.grid.n_table_params

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 2 or x.shape[1] != self.n_input_dims:
            raise ValueError(f"expected input of shape (N, {self.n_input_dims})")
        if x.requires_grad:
            raise NotImplementedError("input-coordinate gradients are not implemented")
        return _InrFunction.apply(x, self.params, self._desc, self.mlp, self.n_output_dims)


# ------------------------------------------------------------------------------------------------
# FFT / IFFT
# ------------------------------------------------------------------------------------------------
def _fft2c_raw(x: torch.Tensor, inverse: bool, scale: float) -> torch.Tensor:
    _need_cuda(x, "FFT")
    if not x.is_complex():
        x = x.to(torch.complex64)
    if x.dtype != torch.complex64:
        raise NotImplementedError("FFT: complex64 only")
    if x.dim() < 2:
        raise ValueError("FFT needs at least 2 dims")
    h, w = x.shape[-2], x.shape[-1]
    xr = torch.view_as_real(x.contiguous())
    batch = xr.numel() // (h * w * 2) if h * w > 0 else 0
    out = torch.empty_like(xr)
    tmp = torch.empty_like(xr)
    with torch.cuda.device(x.device):
        nat.check(nat.lib().immoco_fft2c(xr.data_ptr(), out.data_ptr(), tmp.data_ptr(), batch, h, w,
                                         twiddle_table(h, x.device).data_ptr(),
                                         twiddle_table(w, x.device).data_ptr(), 1 if inverse else 0,
                                         float(scale), _stream(x.device)), "fft2c")
    return torch.view_as_complex(out)


class _FFTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inverse):
        ctx.inverse = inverse
        ctx.hw = x.shape[-2] * x.shape[-1]
        return _fft2c_raw(x, inverse, 1.0 / ctx.hw if inverse else 1.0)

    @staticmethod
    def backward(ctx, g):
        # adjoint of the un-normalised forward = conjugate transform; adjoint of IFFT = FFT / (HW)
        if ctx.inverse:
            return _fft2c_raw(g, False, 1.0 / ctx.hw), None
        return _fft2c_raw(g, True, 1.0), None


def FFT(x: torch.Tensor) -> torch.Tensor:
    """fftshift(fftn(ifftshift(x))) over the last two dims, un-normalised (data_utils.py:29-30)."""
    return _FFTFunction.apply(x, False)


def IFFT(x: torch.Tensor) -> torch.Tensor:
    """ifftshift(ifftn(fftshift(x))) over the last two dims, 1/(HW) (data_utils.py:33-34)."""
    return _FFTFunction.apply(x, True)


# ------------------------------------------------------------------------------------------------
# gradient entropy
# ------------------------------------------------------------------------------------------------
class _GradEntropyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _need_cuda(x, "GradientEntropyLoss")
        if x.dim() != 2 or x.dtype != torch.complex64:
            raise ValueError("GradientEntropyLoss expects a complex64 (H, W) image")
        xr = torch.view_as_real(x.detach().contiguous())
        h, w = x.shape
        acc = torch.zeros(1, dtype=torch.float64, device=x.device)
        grad = torch.empty_like(xr)
        with torch.cuda.device(x.device):
            nat.check(nat.lib().immoco_grad_entropy(xr.data_ptr(), 1.0, acc.data_ptr(), grad.data_ptr(), 0,
                                                    h, w, _stream(x.device)), "grad_entropy")
        ctx.save_for_backward(grad)
        return acc[0].float()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return torch.view_as_complex(grad * g)


class GradientEntropyLoss(nn.Module):
    """-sum(g * log(g + 1e-24)), g = |dx| + |dy| of a complex image (losses.py:20-40)."""

    def entropy(self, x):
        return -torch.sum(torch.mul(x, torch.log(x + 1e-24)))

    def forward(self, x):
        return _GradEntropyFunction.apply(x)
