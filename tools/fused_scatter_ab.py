"""A/B: hashed-level scatter fused into the 64-wide MLP backward kernel (immoco_set_fused_scatter) vs the separate
scatter kernel, inside the C2 iteration (two-stream us / iteration, serial per-kernel us) and for other n_M.
python tools/fused_scatter_ab.py [--iters 300]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai24_immoco_b200 as mb  # noqa: E402
from miccai24_immoco_b200 import _native as nat  # noqa: E402
from oracle import immoco_oracle as orc  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=300)
    a = ap.parse_args()
    mb.build()
    lib = mb.lib()
    for h, w, m in ((320, 320, 4), (320, 320, 2), (320, 320, 8), (640, 368, 5)):
        case = orc.make_case(h, w, m, 1000)
        k = case["kspace_motion"]
        lam = mb.lambda_schedule(max(a.iters, 10), 1e-2)[:a.iters]
        for fused in (0, 1, 0, 1):
            lib.immoco_set_fused_scatter(fused)
            model = mb.IMMoCo(case["masks"].cuda())
            eng = mb.FitEngine(model, a.iters)
            eng.set_kspace((k / k.abs().max() * 16000).cuda())
            p_i, p_m = eng.image_params(), eng.motion_params()
            eng.run(lam, 1e-2, 0, 20)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                eng.reset(p_i, p_m)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.run(lam, 1e-2)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / a.iters * 1e3)
            prof = lib.immoco_profile_create(32)
            eng.reset(p_i, p_m)
            eng.run(lam, 1e-2, 0, min(a.iters, 100), profile=prof, profile_every=10)
            torch.cuda.synchronize()
            ms = (C.c_float * len(nat.PROFILE_SLOTS))()
            n = lib.immoco_profile_read(prof, ms)
            lib.immoco_profile_destroy(prof)
            per = {s: round(ms[i] / n * 1e3, 1) for i, s in enumerate(nat.PROFILE_SLOTS) if s in ("mlp_bwd_motion", "hashgrid_bwd_motion")}
            print(f"[{h}x{w} M={m}] fused_scatter={fused}: {best:7.1f} us / iteration; serial {per}; "
                  f"final loss {eng.loss_trace(lam)[-1]:.5f}", flush=True)
            del eng, model
    lib.immoco_set_fused_scatter(0)
