"""Steady-state iteration time with / without programmatic dependent launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
lib = mb.lib()
case = orc.make_case(320, 320, 4, 1000)
model = mb.IMMoCo(case["masks"].cuda())
eng = mb.FitEngine(model, 600)
k = case["kspace_motion"]; eng.set_kspace((k / k.abs().max() * 16000).cuda())
lam = mb.lambda_schedule(600, 1e-2)
for rep in range(2):
    for pdl in (0, 1):
        lib.immoco_set_pdl(pdl)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.run(lam, 1e-2, 0, 100); e0.record(); eng.run(lam, 1e-2, 100, 600); e1.record(); torch.cuda.synchronize()
        print(f"pdl={pdl}: {e0.elapsed_time(e1)/500*1e3:.1f} us / iteration   final loss {eng.loss_trace(lam)[599]:.5f}", flush=True)
