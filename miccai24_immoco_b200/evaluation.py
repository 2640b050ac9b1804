"""The data format and the evaluation loop either side of the IM-MoCo fit.

Host-side mirror of the two reference scripts that frame the hot path:

* ``src/utils/prepareData.py:144-216`` (``motion_test_data``) writes one ``.pth`` dictionary per motion scenario,
  ``{"kspace_motion": (B, H, W) complex64, "image_rss": (B, H, W) complex64, "rotation": [ (n_b,) ],
  "translation": [ (n_b, 2) ], "mask": (B, H, W) int64, "metrics": [ {ssim, psnr, haar_psi, rmse} ]}`` -- the file
  ``src/test/test_immoco.py:31-35`` loads.  ``make_test_set`` builds that dictionary from ground-truth images with the
  device-side ``motion_simulation2D`` / ``calmetric2D``; ``save_test_set`` / ``load_test_set`` are ``torch.save`` /
  ``torch.load`` with the layout checked.  (Reading fastMRI ``.h5`` volumes, ``prepareData.py:21-141``, needs the dataset
  and ``h5py``: out of scope -- the images are an argument here.)
* ``src/test/test_immoco.py:37-130``: per scenario and slice kld-net line detection -> movement groups -> IM-MoCo fit ->
  central-crop metrics, the list-of-lists ``immoco_metrics.pth`` and the mean / std table.  ``run_test_immoco`` is that
  loop over the CUDA path (batched group labelling, lock-step fits through ``reconstruct_batch``); ``summarize_metrics``
  is the table.  Figures and LaTeX output (``:96-108``, ``:131-160``) are not reproduced.

Every compute step is one of the package's CUDA entry points; the steps can be replaced through keyword arguments, which
is how the CPU tests exercise the control flow and the file formats without a GPU.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

TEST_SET_KEYS = ("kspace_motion", "image_rss", "rotation", "translation", "mask", "metrics")
METRIC_KEYS = ("ssim", "psnr", "haar_psi", "rmse")          # order of the reference's dictionaries (test_immoco.py:87-94)
# movement counts of the two scenarios (prepareData.py:148: np.arange(6, 10) / np.arange(16, 20))
SCENARIO_MOVEMENTS = {"light": (6, 10), "heavy": (16, 20)}


def _metric_dict(values) -> Dict[str, float]:
    """calmetric2D returns (psnr, ssim, haar_psi, rmse); the reference stores them under these keys."""
    psnr, ssim, haar, rmse = values
    return {"ssim": float(ssim), "psnr": float(psnr), "haar_psi": float(haar), "rmse": float(rmse)}


def validate_test_set(data: dict) -> dict:
    """Checks the dictionary layout ``motion_test_data`` writes; returns ``data``."""
    missing = [k for k in TEST_SET_KEYS if k not in data]
    if missing:
        raise KeyError(f"test-set dictionary lacks {missing}")
    k, img, mask = data["kspace_motion"], data["image_rss"], data["mask"]
    if k.dim() != 3 or not k.is_complex():
        raise ValueError("kspace_motion must be a (B, H, W) complex tensor")
    if tuple(img.shape) != tuple(k.shape) or tuple(mask.shape) != tuple(k.shape):
        raise ValueError("image_rss and mask must have the shape of kspace_motion")
    if mask.dtype != torch.int64:
        raise ValueError("mask must be int64 (line indicators as motion_simulation2D returns them)")
    n = k.shape[0]
    for name in ("rotation", "translation", "metrics"):
        if len(data[name]) != n:
            raise ValueError(f"{name} must hold one entry per slice")
    for m in data["metrics"]:
        if set(m) != set(METRIC_KEYS):
            raise ValueError(f"metric dictionaries must have the keys {METRIC_KEYS}")
    return data


def make_test_set(images: Sequence[torch.Tensor], movements=(6, 10), *, simulate: Optional[Callable] = None,
                  metrics: Optional[Callable] = None, ifft: Optional[Callable] = None, device="cuda") -> dict:
    """One scenario of ``motion_test_data`` (prepareData.py:150-214) from ground-truth images.

    ``images``: (H, W) complex images, all of one shape (the reference skips everything but 320 x 320, ``:165-168``).
    ``movements``: [low, high) of the per-slice movement count, drawn with ``np.random.choice`` like ``:162``;
    the motion parameters come from the global torch RNG in the reference's order (``motion_simulation2D``).
    Per slice the metrics of the CORRUPTED image against the ground truth on the central half are recorded
    (``:183-196``).  Tensors are returned on the host, like the file.  ``simulate`` / ``metrics`` / ``ifft`` default to
    the package's CUDA entry points (there is no CPU path; the arguments exist for the host-logic tests)."""
    from .metrics import crop_metrics
    from .motion_utils import motion_simulation2D
    from .ops import IFFT
    simulate = motion_simulation2D if simulate is None else simulate
    metrics = crop_metrics if metrics is None else metrics
    ifft = IFFT if ifft is None else ifft
    if len(images) == 0:
        raise ValueError("no images")
    shape = tuple(images[0].shape)
    ks, gts, rots, trs, masks, mets = [], [], [], [], [], []
    choices = np.arange(int(movements[0]), int(movements[1]))
    for img in images:
        if tuple(img.shape) != shape or img.dim() != 2:
            raise ValueError("all images must be 2-D and of one shape")
        n_mov = int(np.random.choice(choices, 1, replace=True)[0])
        img_d = img.to(device=device, dtype=torch.complex64)
        k_motion, mask, rot, trans = simulate(img_d, n_mov)
        mets.append(_metric_dict(metrics(ifft(k_motion).abs(), img_d.abs())))
        ks.append(k_motion.detach().cpu())
        gts.append(img_d.detach().cpu())
        rots.append(rot.detach().cpu())
        trs.append(trans.detach().cpu())
        masks.append(mask.detach().cpu())
    return validate_test_set({"kspace_motion": torch.stack(ks), "image_rss": torch.stack(gts), "rotation": rots,
                              "translation": trs, "mask": torch.stack(masks), "metrics": mets})


def save_test_set(data: dict, path: str) -> None:
    torch.save(validate_test_set(data), path)


def load_test_set(path: str) -> dict:
    """``torch.load(data_path)`` of test_immoco.py:33 with the layout checked."""
    return validate_test_set(torch.load(path, map_location="cpu", weights_only=False))


def run_test_immoco(data: Union[dict, str], net, iters: int = 200, learning_rate: float = 1e-2, lambda_ge: float = 1e-2,
                    *, device="cuda", chunk: int = 64, detect: Optional[Callable] = None, fit: Optional[Callable] = None,
                    metrics: Optional[Callable] = None, return_images: bool = False):
    """The per-scenario loop of src/test/test_immoco.py:37-94 over one test-set dictionary (or its path).

    Per slice: kld-net line detection and movement groups (``:47-61``), ``imcoco_motion_correction(k, masks,
    iters=200, learning_rate=1e-2, lambda_ge=1e-2)`` (``:65-72``), central-half crop metrics against ``image_rss``
    (``:77-94``).  Returns the list of ``{ssim, psnr, haar_psi, rmse}`` dictionaries in slice order (and the corrected
    magnitude images when asked).  Slices are processed ``chunk`` at a time: one batched kld-net pass + one host
    synchronisation for the group counts, then lock-step fits (``reconstruct_batch``); results equal the slice-by-slice
    loop (each slice is its own optimisation)."""
    from .batch import reconstruct_batch
    from .kld_net import movement_masks_from_kspace
    from .metrics import crop_metrics
    if isinstance(data, str):
        data = load_test_set(data)
    else:
        validate_test_set(data)
    detect = (lambda k: movement_masks_from_kspace(net, k)) if detect is None else detect
    if fit is None:
        def fit(ks, masks):
            return reconstruct_batch(ks, masks, iters, learning_rate, lambda_ge)
    metrics = crop_metrics if metrics is None else metrics
    kspaces, gts = data["kspace_motion"], data["image_rss"]
    out: List[Dict[str, float]] = []
    images: List[torch.Tensor] = []
    if net is not None and hasattr(net, "eval"):
        net.eval()
    for b0 in range(0, kspaces.shape[0], max(1, int(chunk))):
        k = kspaces[b0: b0 + chunk].to(device)
        with torch.no_grad():
            masks = detect(k)
        refined = fit([k[i] for i in range(k.shape[0])], masks)
        for i, img in enumerate(refined):
            gt = gts[b0 + i].to(device).abs()
            out.append(_metric_dict(metrics(img.abs(), gt)))
            if return_images:
                images.append(img.abs().detach().cpu())
    return (out, images) if return_images else out


def summarize_metrics(metrics_all: Sequence[Sequence[Dict[str, float]]], scenarios: Sequence[str] = ("light", "heavy")):
    """Mean and standard deviation of every metric per scenario (test_immoco.py:112-130: ``torch.mean`` /
    ``torch.std`` -- the unbiased estimator -- over the slices): ``{scenario: {metric: (mean, std)}}``."""
    table: Dict[str, Dict[str, tuple]] = {}
    for name, rows in zip(scenarios, metrics_all):
        table[name] = {}
        for key in (rows[0].keys() if rows else ()):
            v = torch.tensor([float(r[key]) for r in rows], dtype=torch.float32)
            table[name][key] = (float(v.mean()), float(v.std()) if v.numel() > 1 else float("nan"))
    return table


def evaluate_scenarios(paths: Dict[str, str], net, out_path: Optional[str] = None, **kw):
    """test_immoco.py:27-110 end to end: every scenario file -> metrics; the list of lists is stored like
    ``results/immoco/immoco_metrics.pth`` (``:110``) when ``out_path`` is given.  Returns (metrics_all, table)."""
    names = list(paths)
    metrics_all = [run_test_immoco(paths[n], net, **kw) for n in names]
    if out_path is not None:
        torch.save(metrics_all, out_path)
    return metrics_all, summarize_metrics(metrics_all, names)
