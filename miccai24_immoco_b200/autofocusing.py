"""Autofocusing baseline on the IM-MoCo kernels (src/models/autofocusing.py:8-91; loop of
src/test/test_autofocusing.py:58-72): one rigid (rotation, x-shift, y-shift) per movement group,
bicubic resampling of the group's partial image, Adam(lr=1.0) x 60 on the gradient entropy alone.

Same constructor / parameter names / forward signature as the reference.  The heavy ops run on this
repo's kernels -- centred FFT / IFFT (forward_model.cu), rigid bicubic resampling with its gradient
reduced to d theta (autofocus.cu), gradient entropy (forward_model.cu); the 3 x M parameter algebra
(angle -> matrix -> affine) and the mask products are tiny torch ops on the device.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native as nat
from .ops import FFT, IFFT, GradientEntropyLoss, _need_cuda, _stream


class _RigidBicubic(torch.autograd.Function):
    """(M,H,W) complex images (constants) warped by theta (M,2,3); gradient w.r.t. theta only."""

    @staticmethod
    def forward(ctx, images_ri, theta):
        m, h, w, _ = images_ri.shape
        th = theta.detach().reshape(m, 6).float().contiguous()
        out = torch.empty_like(images_ri)
        nat.check(nat.lib().immoco_rigid_bicubic_fwd(images_ri.data_ptr(), th.data_ptr(), out.data_ptr(), m, h, w,
                                                     _stream()), "rigid_bicubic_fwd")
        ctx.save_for_backward(images_ri, th)
        return out

    @staticmethod
    def backward(ctx, d_out):
        images_ri, th = ctx.saved_tensors
        m, h, w, _ = images_ri.shape
        d_theta = torch.zeros((m, 6), dtype=torch.float64, device=images_ri.device)
        nat.check(nat.lib().immoco_rigid_bicubic_bwd_theta(images_ri.data_ptr(), th.data_ptr(),
                                                           d_out.contiguous().float().data_ptr(), d_theta.data_ptr(),
                                                           m, h, w, _stream()), "rigid_bicubic_bwd_theta")
        return None, d_theta.float().view(m, 2, 3)


class Autofocusing(nn.Module):
    """``Autofocusing(masks)``: masks (M, H, W) movement-group masks on the CUDA device."""

    def __init__(self, masks):
        super().__init__()
        _need_cuda(masks, "Autofocusing(masks)")
        self.num_movements = masks.shape[0]
        dev = masks.device
        self.motion_parameters = nn.ParameterDict(dict(
            rot_vector=nn.Parameter(torch.zeros(self.num_movements, device=dev)),
            x_shifts=nn.Parameter(torch.zeros(self.num_movements, device=dev)),
            y_shifts=nn.Parameter(torch.zeros(self.num_movements, device=dev)),
        ))
        self.device = dev
        self.masks = masks

    def affine(self, h: int, w: int) -> torch.Tensor:
        """(M, 2, 3) affine of autofocusing.py:31-67, quirks kept: the rotation matrix is transposed,
        shift[:, 1] starts from shift[:, 0] (:50-53), row 0 is divided by 2H-1 and row 1 by 2W-1."""
        p = self.motion_parameters
        a = torch.deg2rad(p["rot_vector"])
        c, s = torch.cos(a), torch.sin(a)
        r00, r01, r10, r11 = c, s, -s, c                  # permute(0, 2, 1) of [[c, -s], [s, c]]
        tx, ty = p["x_shifts"], p["y_shifts"]
        sh0 = -r00 * tx - r01 * ty
        sh1 = sh0 + (-r10 * tx - r11 * ty)
        t0 = sh0.float() / (2.0 * h - 1)
        t1 = sh1.float() / (2.0 * w - 1)
        return torch.stack([torch.stack([r00, r01, t0], -1), torch.stack([r10, r11, t1], -1)], 1)

    def forward(self, ks_input):
        _need_cuda(ks_input, "Autofocusing.forward")
        ks = ks_input.squeeze().to(torch.complex64)
        h, w = ks.shape
        masks_f = self.masks.float()
        with torch.no_grad():                              # constants of the optimisation (:29)
            images = IFFT(ks.unsqueeze(0) * masks_f)
            images_ri = torch.view_as_real(images).contiguous()
        moved = _RigidBicubic.apply(images_ri, self.affine(h, w))
        image_2d = torch.view_as_complex(moved)
        return ks * (1 - self.masks.sum(0)).float() + (FFT(image_2d) * masks_f).sum(0)


def autofocus_motion_correction(kspace, masks, iters: int = 60, learning_rate: float = 1.0, lambda_ge: float = 1e-4,
                                return_trace: bool = False):
    """The per-slice loop of test_autofocusing.py:58-76: k-space scaled by max|IFFT(k)|, Adam(lr=1.0) x 60
    on ``GradientEntropyLoss()(IFFT(k_refined)) * 1e-4``.  Returns (|IFFT(k_refined)|, k_refined[, trace])."""
    if not torch.cuda.is_available():
        raise RuntimeError("autofocus_motion_correction needs a CUDA device (no CPU fallback)")
    kspace = kspace.cuda().to(torch.complex64)
    masks = masks.cuda()
    k = kspace / IFFT(kspace).abs().max()
    model = Autofocusing(masks)
    opt = torch.optim.Adam(model.parameters(), lr=learning_rate)
    ge = GradientEntropyLoss()
    trace = []
    k_ref = k
    for _ in range(iters):
        opt.zero_grad()
        k_ref = model(k)
        loss = ge(IFFT(k_ref)) * lambda_ge
        loss.backward()
        opt.step()
        if return_trace:
            trace.append(loss.detach())
    image = IFFT(k_ref).abs().detach()
    if return_trace:
        return image, k_ref.detach(), torch.stack(trace).cpu().numpy() if trace else None
    return image, k_ref.detach()
