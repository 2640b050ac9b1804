"""Adam tuning knobs (unroll / streaming hints / CTAs per SM) measured INSIDE the iteration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
lib = mb.lib()
case = orc.make_case(320, 320, 4, 1000)
model = mb.IMMoCo(case["masks"].cuda())
p_img = model.image_inr.params.detach().clone(); p_mot = model.motion_inr.params.detach().clone()
eng = mb.FitEngine(model, 400)
k = case["kspace_motion"]; eng.set_kspace((k / k.abs().max() * 16000).cuda())
lam = mb.lambda_schedule(400, 1e-2)
for rep in range(2):
    for variant, ctas in ((0, 32), (1, 32), (2, 16), (3, 32), (4, 32), (5, 16), (0, 16), (0, 64), (1, 16)):
        lib.immoco_set_adam_tuning(variant, ctas)
        eng.reset(p_img, p_mot)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.run(lam, 1e-2, 0, 100); e0.record(); eng.run(lam, 1e-2, 100, 400); e1.record(); torch.cuda.synchronize()
        print(f"adam variant={variant} ctas/SM={ctas}: {e0.elapsed_time(e1)/300*1e3:.1f} us / iteration", flush=True)
lib.immoco_set_adam_tuning(0, 32)
