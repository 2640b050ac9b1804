"""ORACLE -- test infrastructure, NOT product code.

Plain torch (fp32) restatement of the autofocusing baseline: ``Autofocusing.forward``
(src/models/autofocusing.py:24-91) and the per-slice loop of src/test/test_autofocusing.py:58-72.
PINNED: ``oracle/gen_golden_autofocus.py`` imports the reference's own autofocusing.py unchanged and checks
this restatement bit-for-bit (forward and an Adam trajectory), then writes
``tests/golden/autofocus_small.npz``.  Only tests/ may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import immoco_oracle as orc


def affine_from_params(rot_deg: torch.Tensor, x_shifts: torch.Tensor, y_shifts: torch.Tensor, h: int, w: int):
    """(M, 2, 3) affine of autofocusing.py:31-67 (quirks kept: transposed rotation, shift[:, 1] builds on
    shift[:, 0], row 0 divided by 2H-1 and row 1 by 2W-1)."""
    a = torch.deg2rad(rot_deg)
    m = a.shape[0]
    rot = torch.zeros((m, 2, 2))
    rot[:, 0, 0] = rot[:, 0, 0] + torch.cos(a)
    rot[:, 0, 1] = rot[:, 0, 1] + -torch.sin(a)
    rot[:, 1, 0] = rot[:, 1, 0] + torch.sin(a)
    rot[:, 1, 1] = rot[:, 1, 1] + torch.cos(a)
    rot = rot.permute(0, 2, 1)
    shift = torch.zeros((m, 2))
    shift[:, 0] = shift[:, 0] + (-rot[:, 0, 0] * x_shifts - rot[:, 0, 1] * y_shifts)
    shift[:, 1] = shift[:, 0] + (-rot[:, 1, 0] * x_shifts - rot[:, 1, 1] * y_shifts)
    aff = torch.zeros((m, 2, 3))
    aff[:, 0, -1] = aff[:, 0, -1] + shift[:, 0].float()
    aff[:, 1, -1] = aff[:, 1, -1] + shift[:, 1].float()
    aff[:, 0, 0] = aff[:, 0, 0] + rot[:, 0, 0]
    aff[:, 0, 1] = aff[:, 0, 1] + rot[:, 0, 1]
    aff[:, 1, 0] = aff[:, 1, 0] + rot[:, 1, 0]
    aff[:, 1, 1] = aff[:, 1, 1] + rot[:, 1, 1]
    aff[:, :, -1] = aff[:, :, -1] / ((torch.tensor([h, w]) * 2.0) - 1)
    return aff


def rigid_bicubic(images: torch.Tensor, aff: torch.Tensor) -> torch.Tensor:
    """(M,H,W) complex images warped like autofocusing.py:69-84 (bicubic, zeros, align_corners mix)."""
    m, h, w = images.shape
    grid = F.affine_grid(aff, (m, 2, h, w), align_corners=True)
    out = F.grid_sample(torch.view_as_real(images).permute(0, 3, 1, 2), grid.float(), mode="bicubic",
                        align_corners=False)
    return torch.view_as_complex(out.permute(0, 2, 3, 1).contiguous())


def autofocus_forward(ks: torch.Tensor, masks: torch.Tensor, rot_deg, x_shifts, y_shifts) -> torch.Tensor:
    h, w = ks.shape
    images = orc.IFFT(ks.squeeze().unsqueeze(0) * masks.float())
    moved = rigid_bicubic(images, affine_from_params(rot_deg, x_shifts, y_shifts, h, w))
    return (ks.squeeze() * (1 - masks.sum(0)).float()) + (orc.FFT(moved) * masks.float()).sum(0)


def autofocus_loop(kspace: torch.Tensor, masks: torch.Tensor, iters: int = 60, lr: float = 1.0, lam: float = 1e-4):
    """test_autofocusing.py:58-72.  Returns (|IFFT(k_refined)|, k_refined, loss trace, final parameters)."""
    k = kspace / orc.IFFT(kspace).abs().max()
    m = masks.shape[0]
    params = [torch.zeros(m, requires_grad=True) for _ in range(3)]      # rot_vector, x_shifts, y_shifts
    opt = torch.optim.Adam(params, lr=lr)
    trace, k_ref = [], k
    for _ in range(iters):
        opt.zero_grad()
        k_ref = autofocus_forward(k, masks, *params)
        loss = orc.gradient_entropy(orc.IFFT(k_ref)) * lam
        loss.backward()
        opt.step()
        trace.append(float(loss))
    return orc.IFFT(k_ref).abs().detach(), k_ref.detach(), trace, [p.detach().clone() for p in params]
