"""Evaluation metrics of the step that follows the fit, on the device.

Host-side mirror of src/utils/evaluate.py: ``normalize`` (:19-29), ``rmse`` (:32-34), ``my_psnr``
(:37-47) and ``calmetric2D`` (:57-80), plus the central-half crop of src/test/test_immoco.py:74-85.
``piq.ssim(kernel_size=11, data_range=1.0)`` (piq 0.8.0, absent here: parity unpinned, restated from
its published algorithm -- Gaussian sigma 1.5, k1=0.01, k2=0.03, valid convolution, average-pool
down-sampling by max(1, round(min(H,W)/256))) runs as one CUDA kernel together with the squared error;
nothing is copied to the host: the four results are 0-dim CUDA tensors.  HaarPSI (piq.haarpsi(scales=3),
evaluate.py:76) is two more small kernels on the same normalised pairs -- restated from its published algorithm
like SSIM (parity unpinned: piq is absent).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from .ops import _need_cuda, _stream


def normalize(x: torch.Tensor) -> torch.Tensor:
    """Min-max normalise to [0, 1]; batch-wise when the leading dim is > 1 (evaluate.py:19-29)."""
    if x.shape[0] > 1:
        flat = x.reshape(x.shape[0], -1)
        lo, hi = flat.min(1).values, flat.max(1).values
        shape = (-1,) + (1,) * (x.dim() - 1)
        return (x - lo.view(shape)) / ((hi - lo).view(shape) + 1e-24)
    return (x - x.min()) / (x.max() - x.min() + 1e-24)


def rmse(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    return torch.sqrt(torch.mean((x - y) ** 2))


def my_psnr(img1, img2, data_range=None, reduction="mean"):
    mse = torch.mean((img1 - img2) ** 2, dim=(1, 2, 3))
    max_pixel = img2.reshape(img2.shape[0], -1).max(1).values if data_range is None else data_range
    out = 20 * torch.log10(max_pixel / torch.sqrt(mse))
    return out if reduction == "none" else out.mean()


def _view(t: torch.Tensor):
    """(tensor kept alive, data_ptr, image stride, row stride, is_complex) of a (B,1,H,W) tensor whose
    last dim is contiguous; strides in elements."""
    if t.dim() != 4 or t.shape[1] != 1:
        raise ValueError("expected a (B, 1, H, W) tensor")
    if t.is_complex():
        if t.dtype != torch.complex64:
            t = t.to(torch.complex64)
    elif t.dtype != torch.float32:
        t = t.float()
    if t.stride(3) != 1:
        t = t.contiguous()
    return t, t.data_ptr(), t.stride(0), t.stride(2), 1 if t.is_complex() else 0


def metric_sums(pred_recon: torch.Tensor, gt_recon: torch.Tensor, kernel_size: int = 11, haarpsi: bool = False):
    """(B, 4) float64 CUDA tensor: per image {sum of squared error of the min-max-normalised images,
    sum of the SSIM map, SSIM map size, 0}.  Inputs may be crops (strided views) and complex."""
    _need_cuda(pred_recon, "calmetric2D")
    _need_cuda(gt_recon, "calmetric2D")
    if pred_recon.shape != gt_recon.shape:
        raise ValueError("pred and gt must have the same shape")
    p, p_ptr, p_is, p_rs, p_c = _view(pred_recon)
    g, g_ptr, g_is, g_rs, g_c = _view(gt_recon)
    b, _, h, w = p.shape
    pool = max(1, round(min(h, w) / 256))
    minmax = torch.empty((b, 4), dtype=torch.float32, device=p.device)
    acc = torch.zeros((b, 4), dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        s = _stream(p.device)
        nat.check(nat.lib().immoco_metrics2d(p_ptr, p_is, p_rs, p_c, g_ptr, g_is, g_rs, g_c, b, h, w, kernel_size, pool,
                                             minmax.data_ptr(), acc.data_ptr(), s), "metrics2d")
        if haarpsi and h >= 16 and w >= 16:
            # piq.haarpsi(scales=3) on the same normalised pairs (evaluate.py:76); parity unpinned like SSIM
            dp = max(h % 2, w % 2)
            pooled = torch.empty((b, 2, (h + dp) // 2, (w + dp) // 2), dtype=torch.float32, device=p.device)
            hacc = torch.zeros((b, 2), dtype=torch.float64, device=p.device)
            nat.check(nat.lib().immoco_haarpsi(p_ptr, p_is, p_rs, p_c, g_ptr, g_is, g_rs, g_c, b, h, w, 30.0, 4.2,
                                               minmax.data_ptr(), pooled.data_ptr(), hacc.data_ptr(), s), "haarpsi")
            return acc, hacc
    return (acc, None) if haarpsi else acc


def calmetric2D(pred_recon: torch.Tensor, gt_recon: torch.Tensor):
    """(psnr, ssim, haar_psi, rmse) of (B, 1, H, W) reconstructions (evaluate.py:57-80): both inputs are
    min-max normalised per image, PSNR with data_range 1, SSIM 11x11 Gaussian, all 'mean' reductions."""
    if not pred_recon.ndim == 4 or not gt_recon.ndim == 4:
        raise ValueError("Input tensors must be 4D")
    h, w = pred_recon.shape[-2:]
    ssim_kernel = 11
    if w < ssim_kernel or h < ssim_kernel:
        ssim_kernel = min(w, h, ssim_kernel) - 1
    acc, hacc = metric_sums(pred_recon, gt_recon, ssim_kernel, haarpsi=True)
    n_px = float(h * w)
    mse = acc[:, 0] / n_px
    psnr = (20 * torch.log10(1.0 / torch.sqrt(mse))).mean().float()
    ssim = (acc[:, 1] / acc[:, 2]).mean().float()
    rmse_all = torch.sqrt(acc[:, 0].sum() / (n_px * acc.shape[0])).float()
    if hacc is None:          # images smaller than HaarPSI's 16-pixel kernel (piq raises there)
        haar = torch.full((), float("nan"), device=acc.device)
    else:
        eps = float(torch.finfo(torch.float32).eps)
        score = (hacc[:, 0] + eps) / (hacc[:, 1] + eps)
        haar = ((torch.log(score / (1.0 - score)) / 4.2) ** 2).mean().float()
    return psnr, ssim, haar, rmse_all


def crop_metrics(refined_image: torch.Tensor, image_gt: torch.Tensor):
    """test_immoco.py:74-85: |.| of both, central half [H/4:-H/4, W/4:-W/4], calmetric2D.  The crop
    and the magnitude are folded into the kernel's loads (no copies)."""
    h, w = image_gt.shape[-2:]
    ch, cw = int(h / 4), int(w / 4)
    p = refined_image[..., ch:-ch, cw:-cw]
    g = image_gt[..., ch:-ch, cw:-cw]
    while p.dim() < 4:
        p, g = p.unsqueeze(0), g.unsqueeze(0)
    return calmetric2D(p, g)
