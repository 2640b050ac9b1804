"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

from oracle import immoco_oracle as orc


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu() if not a.is_complex() else torch.view_as_real(a.detach().cpu()).double()
    b = b.detach().double().cpu() if not b.is_complex() else torch.view_as_real(b.detach().cpu()).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def case_params(seed, device="cpu"):
    lv2 = orc.make_grid_levels(2, orc.ENCODING_CONFIG)
    lv3 = orc.make_grid_levels(3, orc.ENCODING_CONFIG)
    return (orc.init_params(lv2, orc.IMAGE_NETWORK_CONFIG, 100 + seed).to(device),
            orc.init_params(lv3, orc.MOTION_NETWORK_CONFIG, 200 + seed).to(device))


def drift_band(golden) -> np.ndarray:
    """Running max of the relative loss difference between two exact restatements of the loop
    that differ only in rounding (oracle/gen_golden.py): what ANY implementation can hold."""
    want = golden["loss_trace"]
    return np.maximum.accumulate(np.abs(golden["loss_trace_perturbed"] - want) / np.abs(want))
