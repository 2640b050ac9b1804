"""640x368 n_M=5 (config 3 shape): grouped layout on / off, us / iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
iters = 300
for h, w, m in ((640, 368, 5), (320, 320, 3)):
    case = orc.make_case(h, w, m, 1000)
    model = mb.IMMoCo(case["masks"].cuda())
    p_img = model.image_inr.params.detach().clone(); p_mot = model.motion_inr.params.detach().clone()
    k = case["kspace_motion"]; lam = mb.lambda_schedule(iters, 1e-2)
    for rep in range(2):
        for grouped in (False, True):
            eng = mb.FitEngine(model, iters, grouped_layout=grouped, deterministic=False)
            eng.set_kspace((k / k.abs().max() * 16000).cuda()); eng.reset(p_img, p_mot)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.run(lam, 1e-2, 0, 50); e0.record(); eng.run(lam, 1e-2, 50, iters); e1.record(); torch.cuda.synchronize()
            tr = eng.loss_trace(lam)
            print(f"{h}x{w} n_M={m} grouped={int(grouped)}: {e0.elapsed_time(e1) / (iters - 50) * 1e3:7.1f} us / iteration  "
                  f"loss {tr[0]:.4f} -> {tr[-1]:.5f}", flush=True)
            del eng
    del model
