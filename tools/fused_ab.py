"""A/B of the fused row launches (immoco_set_fused_rows): loss-trace agreement over the first iterations and
steady-state iteration time, for the library variant selected by IMMOCO_LIB_PATH."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat
from oracle import immoco_oracle as orc
lib = mb.lib()
print("lib:", os.path.basename(nat.LIB_PATH), flush=True)
case = orc.make_case(320, 320, 4, 1000)
model = mb.IMMoCo(case["masks"].cuda())
p_img = model.image_inr.params.detach().clone(); p_mot = model.motion_inr.params.detach().clone()
eng = mb.FitEngine(model, 600)
k = case["kspace_motion"]; eng.set_kspace((k / k.abs().max() * 16000).cuda())
lam = mb.lambda_schedule(600, 1e-2)
traces = {}
for fused in (0, 1):
    lib.immoco_set_fused_rows(fused)
    eng.reset(p_img, p_mot); eng.run(lam, 1e-2, 0, 12); torch.cuda.synchronize()
    traces[fused] = eng.loss_trace(lam)[:12].copy()
    print(f"fused={fused} k_out norm {float(eng.k_out.norm()):.6e} image norm {float(eng.image.norm()):.6e}")
rel = np.abs(traces[1] - traces[0]) / np.abs(traces[0])
print("loss rel diff fused vs separate, its 0..11:", " ".join(f"{r:.1e}" for r in rel), flush=True)
for rep in range(2):
    for fused, dz in ((0, 0), (1, 0), (1, 1)):
        lib.immoco_set_fused_rows(fused); lib.immoco_set_deferred_zero(dz)
        eng.reset(p_img, p_mot)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.run(lam, 1e-2, 0, 100); e0.record(); eng.run(lam, 1e-2, 100, 600); e1.record(); torch.cuda.synchronize()
        print(f"fused rows={fused} deferred zero={dz}: {e0.elapsed_time(e1)/500*1e3:.1f} us / iteration   final loss {eng.loss_trace(lam)[599]:.5f}", flush=True)
