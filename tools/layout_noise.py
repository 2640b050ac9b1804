"""Noise floor of the engine-layout comparisons in tests/test_gpu_loop.py: two runs of the SAME configuration vs runs of
two storage layouts (float atomics make every run different).  Prints the statistics the tests assert on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
from tests.gpu_util import rel_l2
def run(model, k, p_img, p_mot, lam, **kw):
    eng = mb.FitEngine(model, 3, deterministic=False, **kw)
    eng.set_kspace((k / k.abs().max() * 16000).cuda()); eng.reset(p_img, p_mot)
    eng.run(lam, 1e-2, 0, 3); torch.cuda.synchronize()
    return dict(mot=eng.motion_params(), img=eng.image_params(), disp=eng.disp.clone(), image=eng.image.clone(),
                k=eng.k_out.clone(), trace=eng.loss_trace(lam))
def cmp(a, b):
    return (f"params>1e-3: mot {float(((a['mot']-b['mot']).abs()>1e-3).float().mean()):.2e} img {float(((a['img']-b['img']).abs()>1e-3).float().mean()):.2e} | "
            f"rel_l2 disp {rel_l2(a['disp'], b['disp']):.2e} image {rel_l2(a['image'], b['image']):.2e} k {rel_l2(a['k'], b['k']):.2e} | "
            f"trace rel {np.max(np.abs(a['trace']-b['trace'])/np.abs(b['trace'])):.2e}")
for h, w, m in ((40, 36, 8), (64, 48, 2), (64, 46, 5), (320, 320, 4)):
    case = orc.make_case(h, w, m, 1000)
    model = mb.IMMoCo(case["masks"].cuda()); k = case["kspace_motion"]
    p_img = model.image_inr.params.detach().clone(); p_mot = model.motion_inr.params.detach().clone()
    p_mot[2048:3072] *= 10.0; p_mot[3072:] *= 300.0
    lam = mb.lambda_schedule(10, 1e-2)[:3]
    for rep in range(3):
        a = run(model, k, p_img, p_mot, lam, grouped_layout=False); b = run(model, k, p_img, p_mot, lam, grouped_layout=False)
        c = run(model, k, p_img, p_mot, lam, grouped_layout=True); d = run(model, k, p_img, p_mot, lam, compact_image=False)
        print(f"{h}x{w} M={m} same-config : {cmp(a, b)}")
        print(f"{h}x{w} M={m} grouped/pair: {cmp(c, a)}")
        print(f"{h}x{w} M={m} taps on/off : {cmp(c, d)}", flush=True)
