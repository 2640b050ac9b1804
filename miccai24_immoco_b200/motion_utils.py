"""kld-net motion-mask -> movement-group interface (src/utils/motion_utils.py:56-109).

``extract_movement_groups`` keeps the reference's name, arguments and outputs; the per-line Python
loop (one device sync per phase-encode line) is replaced by a cumulative sum of run ends, and the
output is generalised from the reference's square (W, W) to (H, W) through ``height``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _native as nat


def extract_movement_groups(motionline_indcies: torch.Tensor, make_list: bool = False,
                            height: Optional[int] = None) -> torch.Tensor:
    """Label runs of detected phase-encode lines.

    motionline_indcies: (W,) bool / 0-1 tensor (test_immoco.py:59-61 passes ``column_vote > 0.2``).
    Returns (H, W) int64 labels (0 = static, 1..M = group) or, with ``make_list``, the (M, H, W)
    int64 one-hot list consumed by ``IMMoCo(masks)``.  ``height=None`` reproduces the reference's
    square output.  Element tests follow the reference literally: a line is labelled when it
    equals 1 and its right neighbour equals 1 or 0 (or it is the last line); the label increments
    after a labelled line whose right neighbour equals 0.
    """
    x = motionline_indcies
    if x.dim() != 1:
        raise ValueError("expected a 1-D tensor of per-line flags")
    w = x.shape[0]
    h = w if height is None else int(height)
    dev = x.device
    if w == 0:
        shape = (0, h, 0) if make_list else (h, 0)
        return torch.zeros(shape, dtype=torch.long, device=dev)
    labels = movement_group_labels(x)
    groups = labels.unsqueeze(0).expand(h, w).contiguous()
    if not make_list:
        return groups
    # number of distinct non-zero labels (one host sync: M sizes the output)
    present = torch.zeros(w + 2, dtype=torch.bool, device=dev)
    present[labels] = True
    n = int(present[1:].sum().item())
    ids = torch.arange(1, n + 1, device=dev, dtype=torch.long).view(n, 1, 1)
    return (groups.unsqueeze(0) == ids).to(torch.long)


def movement_group_labels(flags: torch.Tensor) -> torch.Tensor:
    """Run labels of per-line flags along the LAST dim ((W,) or (B, W)): 0 = static, 1..M = group, with the
    element tests of ``extract_movement_groups`` (motion_utils.py:56-109).  Labels are consecutive, so the
    number of groups of a slice is its largest label."""
    is1 = flags == 1
    is0 = flags == 0
    pad = flags.shape[:-1] + (1,)
    dev = flags.device
    nxt1 = torch.cat([is1[..., 1:], torch.ones(pad, dtype=torch.bool, device=dev)], dim=-1)    # last line: no test
    nxt0 = torch.cat([is0[..., 1:], torch.zeros(pad, dtype=torch.bool, device=dev)], dim=-1)
    labelled = is1 & (nxt1 | nxt0)
    run_end = (is1 & nxt0).to(torch.long)
    before = torch.cumsum(run_end, -1) - run_end        # run ends strictly before line i
    return (before + 1) * labelled.to(torch.long)


def lines_from_mask(mask: torch.Tensor) -> torch.Tensor:
    """Column vote of test_immoco.py:59-61: fraction of rows flagged per line > 0.2 -> (W,) bool."""
    m = mask.squeeze()
    return m.sum(0).div(m.shape[0]) > 0.2


# ------------------------------------------------------------------------------------------------
# synthetic rigid motion (src/utils/motion_utils.py:7-34,121-202)
# ------------------------------------------------------------------------------------------------
def generate_list(size: int, n: int, mingap: int, acs: int = 0) -> torch.Tensor:
    """Random window starts with a minimum gap (motion_utils.py:7-24; ``acs`` is unused there too).
    Consumes the global torch RNG in the reference's order."""
    slack = size - mingap * (n - 1)
    steps = int(torch.randint(0, slack, (1,))[0])
    inc = torch.hstack([torch.ones((steps,), dtype=torch.long), torch.zeros((n,), dtype=torch.long)])
    inc = inc[torch.randperm(inc.shape[0])]
    locs = torch.argwhere(inc == 0).flatten()
    return torch.cumsum(inc, dim=0)[locs] + mingap * torch.arange(0, n)


def get_rand_int(data_range, size=None) -> torch.Tensor:
    """torch.randint with 0 replaced by 1 (motion_utils.py:27-34)."""
    r = torch.randint(data_range[0], data_range[1], size=(1,) if size is None else size)
    return r + 1 if int(r) == 0 else r


def rotation_matrix_2d(angle: torch.Tensor) -> torch.Tensor:
    a = torch.deg2rad(angle)
    return torch.tensor([[torch.cos(a), -torch.sin(a)], [torch.sin(a), torch.cos(a)]])


def motion_simulation2D(image_2d: torch.Tensor, n_movements: Optional[int] = None):
    """Rigid per-movement corruption of k-space line windows (motion_utils.py:121-202), image work on the
    GPU.  ``image_2d``: (H, W) complex CUDA tensor.  The random draws are made on the host from the
    global torch RNG in the reference's order, so a common ``torch.manual_seed`` gives the reference's
    movements.  Returns (kspace (H,W) complex64, mask (H,W) int64, rotations (n,), translations (n,2))."""
    from .ops import FFT, _need_cuda, _stream
    _need_cuda(image_2d, "motion_simulation2D")
    if image_2d.dim() != 2:
        raise ValueError("expected a 2-D complex image")
    image_2d = image_2d.to(torch.complex64).contiguous()
    h, w = image_2d.shape
    dev = image_2d.device
    if n_movements is None:
        n_movements = int(get_rand_int([5, 20]))
    starts = generate_list(w, n_movements, w // n_movements, 0)
    theta = torch.zeros((n_movements, 2, 3))
    w0 = np.zeros(n_movements, np.int32)
    w1 = np.zeros(n_movements, np.int32)
    rot = torch.zeros((n_movements,))
    trans = torch.zeros((n_movements, 2))
    denom = torch.tensor([w, w], dtype=torch.float32) * 2.0 - 1     # reference: shape of image_2d[0, ...] = (W,)
    for m in range(n_movements):
        sx = int(get_rand_int([-10, 10]))
        sy = int(get_rand_int([-10, 10]))
        ang = get_rand_int([-10, 10])
        t = torch.tensor([[1, 0, sx], [0, 1, sy]]).float()
        t[:2, :2] = rotation_matrix_2d(ang)
        t[:, -1] /= denom
        theta[m] = t
        w0[m] = int(starts[m])
        w1[m] = w0[m] + int(get_rand_int([1, 10]))
        rot[m] = ang
        trans[m, :] = torch.tensor([sx, sy])
    lib = nat.lib()
    k = torch.view_as_real(FFT(image_2d)).contiguous()
    mask = torch.zeros((h, w), dtype=torch.long, device=dev)
    if n_movements > 0:
        theta_d = theta.reshape(n_movements, 6).contiguous().to(dev)
        moved = torch.empty((n_movements, h, w, 2), dtype=torch.float32, device=dev)
        nat.check(lib.immoco_rigid_resample(torch.view_as_real(image_2d).data_ptr(), theta_d.data_ptr(),
                                            moved.data_ptr(), n_movements, h, w, _stream()), "rigid_resample")
        k_moved = torch.view_as_real(FFT(torch.view_as_complex(moved))).contiguous()
        w0_d, w1_d = torch.from_numpy(w0).to(dev), torch.from_numpy(w1).to(dev)
        nat.check(lib.immoco_replace_lines(k.data_ptr(), k_moved.data_ptr(), mask.data_ptr(), w0_d.data_ptr(),
                                           w1_d.data_ptr(), n_movements, h, w, _stream()), "replace_lines")
    return torch.view_as_complex(k), mask, rot, trans
