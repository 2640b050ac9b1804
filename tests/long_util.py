"""1000-iteration fits of the headline configuration (C2: 320x320, n_M=4) -- ours against the oracle loop
run on the same GPU -- shared by tests/test_gpu_zz_long_runs.py (asserts) and tools/long_run_stats.py (the table kept
under profiles/).

The loop is chaotic: a 1-ulp perturbation of the initial parameters changes the trajectory of
the ORACLE ITSELF by percents after a few dozen iterations.  A 1000-iteration comparison can therefore only be
statistical: per slice we record the tail loss level (10th percentile of the last 200 iterations -- the last
sample, and even the median of the last 50, may sit inside one of Adam's loss excursions), the final PSNR / SSIM
against the ground-truth phantom, and
compare ours - oracle with oracle(perturbed) - oracle over the same slices."""
import numpy as np
import torch

import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
from tests.gpu_util import case_params

DEV = "cuda"
H = W = 320
N_MOV = 4


def summarize(trace, image_abs, gt_abs):
    trace = np.asarray(trace, dtype=np.float64)
    met = orc.crop_metrics(image_abs.cpu(), gt_abs)
    # loss spikes: samples more than 3 x the median of their +-25-iteration neighbourhood (Adam at lr 1e-2, no decay)
    spikes = 0
    for t in range(100, len(trace)):
        lo, hi = max(0, t - 25), min(len(trace), t + 26)
        spikes += int(trace[t] > 3.0 * np.median(trace[lo:hi]))
    # "tail" = the level the fit has reached: 10th percentile of the last 200 losses.  The median of the last 50
    # (kept as tail_median50) is NOT robust: Adam's excursions last tens of iterations, and a run that ends inside
    # one (oracle and ours alike, 10-25 % of the runs) reports a tail 10-100 x above its own level.
    return {"tail": float(np.percentile(trace[-200:], 10)), "tail_median50": float(np.median(trace[-50:])),
            "last": float(trace[-1]), "max_tail": float(trace[-50:].max()), "spikes": spikes,
            "psnr": float(met["psnr"]), "ssim": float(met["ssim"]), "rmse": float(met["rmse"])}


def run_ours(seed, iters, deterministic):
    case = orc.make_case(H, W, N_MOV, seed)
    p_img, p_mot = case_params(seed, DEV)
    im, _, trace = mb.imcoco_motion_correction(case["kspace_motion"].to(DEV), case["masks"].to(DEV), iters=iters,
                                               image_params=p_img, motion_params=p_mot, return_trace=True,
                                               deterministic=deterministic)
    return summarize(trace, im.abs(), case["image"].abs())


def run_oracle(seed, iters, perturb=0.0):
    """The oracle loop (torch fp32, TF32 off) on the GPU; ``perturb`` > 0 multiplies the initial image-INR
    parameters by (1 + perturb * N(0,1)) -- 1e-7 flips the last bit of a fraction of them."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    case = orc.make_case(H, W, N_MOV, seed)
    p_img, p_mot = case_params(seed, DEV)
    if perturb > 0:
        g = torch.Generator(device=DEV).manual_seed(seed)
        p_img = p_img * (1.0 + perturb * torch.randn(p_img.shape, device=DEV, generator=g))
    im, _, trace = orc.imcoco_motion_correction(case["kspace_motion"].to(DEV), case["masks"].to(DEV), iters=iters,
                                                image_params=p_img, motion_params=p_mot, return_trace=True)
    return summarize(trace, im.detach().abs(), case["image"].abs())


def _perturbed(seed, iters, j):
    """Another 1-ulp perturbation of the SAME slice: different random signs."""
    torch.backends.cuda.matmul.allow_tf32 = False
    case = orc.make_case(H, W, N_MOV, seed)
    p_img, p_mot = case_params(seed, DEV)
    g = torch.Generator(device=DEV).manual_seed(seed + 7919 * j)
    p_img = p_img * (1.0 + 1e-7 * torch.randn(p_img.shape, device=DEV, generator=g))
    im, _, trace = orc.imcoco_motion_correction(case["kspace_motion"].to(DEV), case["masks"].to(DEV), iters=iters,
                                                image_params=p_img, motion_params=p_mot, return_trace=True)
    return summarize(trace, im.detach().abs(), case["image"].abs())


def compare(seeds, iters=1000, modes=("deterministic", "atomic"), log=print, n_perturbed=1):
    rows = []
    for seed in seeds:
        row = {"seed": seed, "oracle": run_oracle(seed, iters), "oracle_perturbed": run_oracle(seed, iters, 1e-7)}
        for j in range(1, n_perturbed):          # further 1-ulp perturbations (other random signs)
            row[f"oracle_perturbed{j + 1}"] = _perturbed(seed, iters, j)
        for mode in modes:
            row[mode] = run_ours(seed, iters, mode == "deterministic")
        rows.append(row)
        log("seed %d: " % seed + "; ".join(
            "%s tail %.5f (med50 %.5f) last %.5f psnr %.3f ssim %.4f spikes %d" % (k, v["tail"], v["tail_median50"], v["last"],
                                                                                   v["psnr"], v["ssim"], v["spikes"])
            for k, v in row.items() if k != "seed"))
    return rows


def spread(rows, a, b, key, relative=False):
    """|a - b| per slice for metric `key` (relative to b when asked)."""
    d = np.asarray([abs(r[a][key] - r[b][key]) / (abs(r[b][key]) if relative else 1.0) for r in rows])
    return d
