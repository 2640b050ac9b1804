"""kld-net motion-mask -> movement-group interface (src/utils/motion_utils.py:56-109).

``extract_movement_groups`` keeps the reference's name, arguments and outputs; the per-line Python
loop (one device sync per phase-encode line) is replaced by a cumulative sum of run ends, and the
output is generalised from the reference's square (W, W) to (H, W) through ``height``.
"""
from __future__ import annotations

from typing import Optional

import torch


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
    is1 = x == 1
    is0 = x == 0
    nxt1 = torch.cat([is1[1:], torch.ones(1, dtype=torch.bool, device=dev)])    # last line: no test
    nxt0 = torch.cat([is0[1:], torch.zeros(1, dtype=torch.bool, device=dev)])
    labelled = is1 & (nxt1 | nxt0)
    run_end = is1 & nxt0
    # label of line i = 1 + number of run ends strictly before i
    before = torch.cumsum(run_end.to(torch.long), 0) - run_end.to(torch.long)
    labels = (before + 1) * labelled.to(torch.long)
    groups = labels.unsqueeze(0).expand(h, w).contiguous()
    if not make_list:
        return groups
    # number of distinct non-zero labels (one host sync: M sizes the output)
    present = torch.zeros(w + 2, dtype=torch.bool, device=dev)
    present[labels] = True
    n = int(present[1:].sum().item())
    ids = torch.arange(1, n + 1, device=dev, dtype=torch.long).view(n, 1, 1)
    return (groups.unsqueeze(0) == ids).to(torch.long)


def lines_from_mask(mask: torch.Tensor) -> torch.Tensor:
    """Column vote of test_immoco.py:59-61: fraction of rows flagged per line > 0.2 -> (W,) bool."""
    m = mask.squeeze()
    return m.sum(0).div(m.shape[0]) > 0.2
