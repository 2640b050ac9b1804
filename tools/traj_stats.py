"""Run-to-run distribution of the distance between our C2 loss trajectory and the reference golden, in units of
the tolerance used by tests/test_gpu_loop.py (max(1e-3, 5 x the oracle's own rounding-drift band))."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_util import case_params, drift_band
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "loop_c2_i200.npz"))
h, n_mov, seed, iters = int(g["h"]), int(g["n_mov"]), int(g["seed"]), int(g["iters"])
case = orc.make_case(h, h, n_mov, seed)
p_img, p_mot = case_params(seed, "cuda")
want = g["loss_trace"][:50]; band = drift_band(g)[:50]; tol = np.maximum(1e-3, 5.0 * band)
worst = []
for r in range(16):
    im, k, trace = mb.imcoco_motion_correction(case["kspace_motion"].cuda(), case["masks"].cuda(), iters=iters,
                                               image_params=p_img, motion_params=p_mot, return_trace=True)
    rel = np.abs(trace[:50] - want) / np.abs(want)
    q = rel / tol
    met = orc.crop_metrics(im.abs().cpu(), case["image"].abs())
    worst.append(q.max())
    print(f"run {r:2d}: worst rel/tol {q.max():.3f} at it {int(q.argmax())}; its<10 {q[:10].max():.3f}, its<30 {q[:30].max():.3f}, its<45 {q[:45].max():.3f}; "
          f"psnr {met['psnr']:.3f} ssim {met['ssim']:.5f}", flush=True)
print("worst ratios sorted:", np.round(np.sort(worst), 3))
print("reference psnr/ssim", float(g["psnr_out"]), float(g["ssim_out"]), "perturbed", float(g["psnr_out_perturbed"]), float(g["ssim_out_perturbed"]))
