"""Per-slice set-up cost of a fit (model construction, engine construction, parameter loading), ms, synchronised."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
case = orc.make_case(320, 320, 4, 1000)
masks = case["masks"].cuda()
def t(fn, n=6):
    out = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); out.append((time.perf_counter() - t0) * 1e3)
    return r, sorted(out)[len(out) // 2], out[0]
model, ms_model, first = t(lambda: mb.IMMoCo(masks))
print(f"IMMoCo(masks): {ms_model:.2f} ms (first {first:.1f})")
for kw in (dict(compact_image=False, grouped_layout=False), dict(compact_image=True, grouped_layout=False),
           dict(compact_image=False, grouped_layout=True), dict()):
    eng, ms_eng, first = t(lambda: mb.FitEngine(model, 1000, deterministic=False, **kw))
    p_img, p_mot = model.image_inr.params.detach(), model.motion_inr.params.detach()
    _, ms_reset, _ = t(lambda: eng.reset(p_img, p_mot))
    _, ms_out, _ = t(lambda: (eng.image_params(), eng.motion_params()))
    print(f"FitEngine({kw}): {ms_eng:.2f} ms (first {first:.1f}); reset {ms_reset:.2f} ms; params out {ms_out:.2f} ms", flush=True)
    del eng
