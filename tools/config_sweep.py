"""ms / iteration of the fused loop at the other BASELINE.json configs (C3: 640x368 n_M=5; C5: 320x320, n_M in {2,4,8})."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
iters = 300
for h, w, m in ((320, 320, 2), (320, 320, 4), (320, 320, 8), (640, 368, 5), (640, 368, 2)):
    case = orc.make_case(h, w, m, 1000 + m)
    model = mb.IMMoCo(case["masks"].cuda())
    eng = mb.FitEngine(model, iters)
    k = case["kspace_motion"]; eng.set_kspace((k / k.abs().max() * 16000).cuda())
    lam = mb.lambda_schedule(iters, 1e-2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.run(lam, 1e-2, 0, 50); e0.record(); eng.run(lam, 1e-2, 50, iters); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (iters - 50) * 1e3
    tr = eng.loss_trace(lam)
    print(f"{h}x{w} n_M={model.num_movements}: {us:7.1f} us / iteration -> {1e6/us/1000:.3f} slices/s at 1000 its   "
          f"loss {tr[0]:.4f} -> {tr[-1]:.4f}", flush=True)
    del eng, model
