"""Tap-indexed image table (FitEngine(compact_image=...)) A/B: us / iteration and serial per-kernel times."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat
from oracle import immoco_oracle as orc
lib = mb.lib()
iters = 400
shapes = ((320, 320, 4), (320, 320, 2), (640, 368, 5)) if "--all" in sys.argv else ((320, 320, 4),)
for h, w, m in shapes:
    case = orc.make_case(h, w, m, 1000)
    model = mb.IMMoCo(case["masks"].cuda())
    p_img = model.image_inr.params.detach().clone(); p_mot = model.motion_inr.params.detach().clone()
    k = case["kspace_motion"]
    lam = mb.lambda_schedule(iters, 1e-2)
    for rep in range(2):
        for compact in (False, True):
            eng = mb.FitEngine(model, iters, compact_image=compact, deterministic=False)
            eng.set_kspace((k / k.abs().max() * 16000).cuda()); eng.reset(p_img, p_mot)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.run(lam, 1e-2, 0, 100); e0.record(); eng.run(lam, 1e-2, 100, iters); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / (iters - 100) * 1e3
            tr = eng.loss_trace(lam)
            line = f"{h}x{w} n_M={m} compact={int(compact)}: {us:7.1f} us / iteration  loss {tr[0]:.4f} -> {tr[99]:.5f} -> {tr[-1]:.5f}"
            if rep == 1:
                prof = lib.immoco_profile_create(16)
                eng.reset(p_img, p_mot)
                eng.run(lam, 1e-2, 0, 160, profile=prof, profile_every=10); torch.cuda.synchronize()
                ms = (C.c_float * len(nat.PROFILE_SLOTS))()
                cnt = lib.immoco_profile_read(prof, ms)
                per = {n: round(ms[i] / cnt * 1e3, 1) for i, n in enumerate(nat.PROFILE_SLOTS)}
                line += "\n    serial us: " + str({n: per[n] for n in ("hashgrid_fwd_image", "hashgrid_bwd_image", "adam_image", "adam_motion",
                                                                        "hashgrid_fwd_motion", "hashgrid_bwd_motion")})
                lib.immoco_profile_destroy(prof)
                if compact:
                    line += f"\n    live image rows {eng._taps.n_active_rows} of {model.image_inr.grid.n_rows}"
            print(line, flush=True)
            del eng
    del model
