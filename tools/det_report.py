"""Deterministic mode on the GPU box: (1) cost -- us per C2 iteration for the float-atomic path, the
reproducible path with gather-then-Adam and with the fused gather + Adam kernel, with per-kernel durations;
(2) distance of the (single, reproducible) golden runs to the reference trajectories in units of the oracle's
own drift band.   python tools/det_report.py [--iters 300] [--golden]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai24_immoco_b200 as mb  # noqa: E402
from miccai24_immoco_b200 import _native as nat  # noqa: E402
from oracle import immoco_oracle as orc  # noqa: E402
from tests.gpu_util import case_params, drift_band  # noqa: E402

DEV = "cuda"


def timing(h, w, n_mov, iters):
    lib = mb.lib()
    case = orc.make_case(h, w, n_mov, 1000)
    p_img, p_mot = case_params(1000, DEV)
    lam = mb.lambda_schedule(max(iters, 10), 1e-2)[:iters]
    k = case["kspace_motion"]
    k_in = (k / k.abs().max() * 16000).to(DEV)
    for name, det, fuse in (("atomic", False, False), ("det gather->adam", True, False), ("det fused", True, True),
                            ("det fused, one stream", True, True)):
        lib.immoco_set_branch_overlap(0 if "one stream" in name else 1)
        model = mb.IMMoCo(case["masks"].to(DEV))
        eng = mb.FitEngine(model, iters, deterministic=det, fuse_adam=fuse)
        eng.set_kspace(k_in)
        eng.reset(p_img, p_mot)
        eng.run(lam, 1e-2, 0, min(20, iters))          # warm-up (tap list build, clocks)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            eng.reset(p_img, p_mot)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run(lam, 1e-2)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / iters * 1e3)
        prof = lib.immoco_profile_create(64)
        eng.reset(p_img, p_mot)
        eng.run(lam, 1e-2, 0, min(iters, 200), profile=prof, profile_every=10)
        torch.cuda.synchronize()
        ms = (C.c_float * len(nat.PROFILE_SLOTS))()
        n = lib.immoco_profile_read(prof, ms)
        lib.immoco_profile_destroy(prof)
        per = {s: round(ms[i] / n * 1e3, 1) for i, s in enumerate(nat.PROFILE_SLOTS) if ms[i] > 0}
        print(f"[{h}x{w} M={n_mov}] {name:18s}: {best:7.1f} us / iteration (two-stream); serial per-kernel us: {per}; "
              f"serial sum {sum(per.values()):.1f}", flush=True)
        del eng, model
    lib.immoco_set_branch_overlap(1)


def golden(tag):
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", f"loop_{tag}.npz"))
    h, n_mov, seed, iters = int(g["h"]), int(g["n_mov"]), int(g["seed"]), int(g["iters"])
    w = int(g["w"]) if "w" in g.files else h
    case = orc.make_case(h, w, n_mov, seed)
    p_img, p_mot = case_params(seed, DEV)
    want = g["loss_trace"]
    band = drift_band(g)
    rows = []
    for name, det, reps in (("deterministic", True, 2), ("atomic", False, 6)):
        for r in range(reps):
            im, k, trace = mb.imcoco_motion_correction(case["kspace_motion"].to(DEV), case["masks"].to(DEV), iters=iters,
                                                       image_params=p_img, motion_params=p_mot, return_trace=True,
                                                       deterministic=det)
            n = min(50, iters)
            rel = np.abs(trace[:n] - want[:n]) / np.abs(want[:n])
            ratio = rel / np.maximum(band[:n], 1e-12)
            met = orc.crop_metrics(im.abs().cpu(), case["image"].abs())
            rows.append((name, r, rel[:4].max(), rel[:10].max(), rel.max(), float(np.max(rel / np.maximum(1e-3, 5 * band[:n]))),
                         float(np.max(rel / np.maximum(1e-3, 3 * band[:n]))), met["psnr"], met["ssim"]))
    print(f"golden {tag}: {h}x{w} M={n_mov} iters={iters}; band end {band[min(49, iters - 1)]:.3e}; "
          f"reference psnr {float(g['psnr_out']) if 'psnr_out' in g.files else float('nan'):.3f}")
    for row in rows:
        print("   %-14s run %d: rel its<4 %.2e, its<10 %.2e, its<50 %.2e; worst rel/max(1e-3,5band) %.2f, /max(1e-3,3band) %.2f; "
              "psnr %.3f ssim %.4f" % row, flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--golden", action="store_true")
    ap.add_argument("--shapes", default="c2")
    a = ap.parse_args()
    mb.build()
    torch.backends.cuda.matmul.allow_tf32 = False
    if "c2" in a.shapes:
        timing(320, 320, 4, a.iters)
    if "c3" in a.shapes:
        timing(640, 368, 5, a.iters)
    if "m2" in a.shapes:
        timing(320, 320, 2, a.iters)
    if a.golden:
        for tag in ("s32_m1", "s64_m2", "c2_i200"):
            golden(tag)
