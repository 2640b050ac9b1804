"""Where do the SIMT and tcgen05 paths diverge inside ONE iteration on real C2 tensors?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
from tests.gpu_util import case_params

lib = mb.lib()
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
case = orc.make_case(320, 320, 4, 7)
p_img, p_mot = case_params(7, "cuda")
p_mot = p_mot.clone(); p_mot[2048:3072] *= 10.0; p_mot[3072:] *= 300.0
lib.immoco_set_branch_overlap(0)
out = {}
for impl in (0, 1):
    lib.immoco_set_mlp_impl(impl)
    model = mb.IMMoCo(case["masks"].cuda())
    with torch.no_grad():
        model.image_inr.params.copy_(p_img); model.motion_inr.params.copy_(p_mot)
    eng = mb.FitEngine(model, 10)
    k = case["kspace_motion"].cuda(); eng.set_kspace(k / k.abs().max() * 16000)
    eng.run(mb.lambda_schedule(10, 1e-2), 0.0, 0, 1)      # lr = 0: parameters stay put
    torch.cuda.synchronize()
    out[impl] = {n: getattr(eng, n).clone() for n in ("enc_image", "enc_motion", "image", "disp", "c_tmp", "k_out", "d_c", "d_image", "d_disp", "d_enc_motion", "d_enc_image")}
for n in out[0]:
    a, b = out[1][n], out[0][n]
    print(f"{n:14s} rel diff tc vs simt {rel(a, b):.3e}   max|b| {float(b.abs().max()):.3e}")
d0, d1 = out[0]["d_disp"].view(-1, 2), out[1]["d_disp"].view(-1, 2)
err = (d1 - d0).norm(dim=1); big = torch.topk(err, 5)
print("largest d_disp deviations at points", big.indices.tolist(), big.values.tolist(), "typical |d_disp|", float(d0.norm(dim=1).median()))
i0, i1 = out[0]["image"].view(-1, 2), out[1]["image"].view(-1, 2)
print("image abs err max", float((i1 - i0).abs().max()), "image |max|", float(i0.abs().max()), "image std", float(i0.std()))
lib.immoco_set_mlp_impl(1); lib.immoco_set_branch_overlap(1)
