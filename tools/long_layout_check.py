"""1000-iteration C2 fits (float atomics) per storage layout: final PSNR / SSIM / tail loss for the seeds of the long-run test.
Answers whether the tap-indexed image table / the grouped motion layout shift the END of the chaotic trajectory
statistically (they must not: they are storage permutations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
from tests.gpu_util import case_params
from tests.long_util import summarize
iters = 1000
lam = mb.lambda_schedule(iters, 1e-2)
for seed in (1000, 1001, 1002, 1003, 1004, 1005):
    case = orc.make_case(320, 320, 4, seed)
    p_img, p_mot = case_params(seed, "cuda")
    model = mb.IMMoCo(case["masks"].cuda())
    k = case["kspace_motion"].cuda()
    kin = k / k.abs().max() * 16000
    line = f"seed {seed}:"
    for name, kw in (("old", dict(compact_image=False, grouped_layout=False)), ("new", dict())):
        vals = []
        for rep in range(3):
            eng = mb.FitEngine(model, iters, deterministic=False, **kw)
            eng.set_kspace(kin); eng.reset(p_img, p_mot); eng.run(lam, 1e-2)
            s = summarize(eng.loss_trace(lam), torch.view_as_complex(eng.image).abs(), case["image"].abs())
            vals.append(s)
            del eng
        line += f"  {name}: psnr " + " ".join(f"{v['psnr']:.2f}" for v in vals) + " tail " + " ".join(f"{v['tail']:.5f}" for v in vals)
    print(line, flush=True)
