"""us / iteration of the fused loop at C2 (and 640x368 n_M=5 with --c3) for the library IMMOCO_LIB_PATH points at
(build-variant A/B: variants are compiled with IMMOCO_NVCC_FLAGS, see miccai24_immoco_b200/_native.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
iters = 500
tag = os.environ.get("IMMOCO_LIB_PATH", "default")
for h, w, m in ((320, 320, 4),) + (((640, 368, 5),) if "--c3" in sys.argv else ()):
    case = orc.make_case(h, w, m, 1000)
    model = mb.IMMoCo(case["masks"].cuda())
    p_img = model.image_inr.params.detach().clone(); p_mot = model.motion_inr.params.detach().clone()
    eng = mb.FitEngine(model, iters)
    k = case["kspace_motion"]; eng.set_kspace((k / k.abs().max() * 16000).cuda())
    lam = mb.lambda_schedule(iters, 1e-2)
    res = []
    for rep in range(3):
        eng.reset(p_img, p_mot)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.run(lam, 1e-2, 0, 100); e0.record(); eng.run(lam, 1e-2, 100, iters); e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / (iters - 100) * 1e3)
    print(f"{tag}: {h}x{w} n_M={m}: " + " ".join(f"{u:7.1f}" for u in res) + " us / iteration", flush=True)
    del eng, model
