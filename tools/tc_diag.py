"""Accuracy of the MLP backward kernels on REAL IM-MoCo tensors (C2, iteration 0) vs a float64 reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat
from oracle import immoco_oracle as orc
from tests.gpu_util import case_params

lib = mb.lib()
s = lambda: torch.cuda.current_stream().cuda_stream
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())

case = orc.make_case(320, 320, 4, 7)
p_img, p_mot = case_params(7, "cuda")
model = mb.IMMoCo(case["masks"].cuda())
with torch.no_grad():
    model.image_inr.params.copy_(p_img); model.motion_inr.params.copy_(p_mot)
    if len(sys.argv) > 1 and sys.argv[1] == "scaled":
        model.motion_inr.params[2048:3072] *= 10.0; model.motion_inr.params[3072:] *= 300.0
lib.immoco_set_mlp_impl(0)
eng = mb.FitEngine(model, 10)
k = case["kspace_motion"].cuda(); eng.set_kspace(k / k.abs().max() * 16000)
eng.run(mb.lambda_schedule(10, 1e-2), 1e-2, 0, 1)
torch.cuda.synchronize()
for name, enc, d_out, width, act, ofs in (("motion", eng.enc_motion, eng.d_disp.view(-1, 2), 64, nat.ACT_TANH, 0),
                                          ("image", eng.enc_image, eng.d_image.view(-1, 2), 256, nat.ACT_RELU, eng.n_motion)):
    n = enc.shape[1]
    # parameters BEFORE the step are gone (Adam ran); use the injected ones
    src = (p_mot if name == "motion" else p_img).clone()
    if name == "motion" and len(sys.argv) > 1 and sys.argv[1] == "scaled":
        src[2048:3072] *= 10.0
    w1 = src[: width * 32].view(width, 32).contiguous(); w2 = src[width * 32: width * 48].view(16, width).contiguous()
    e64 = enc.permute(1, 0, 2).reshape(n, 32).double().requires_grad_(True)
    w1d = w1.double().requires_grad_(True); w2d = w2.double().requires_grad_(True)
    f = torch.tanh if act == nat.ACT_TANH else torch.relu
    out = (f(e64 @ w1d.t()) @ w2d.t())[:, :2]
    (out * d_out.double()).sum().backward()
    print(f"[{name}] n={n} |enc| max {float(enc.abs().max()):.3e} |d_out| max {float(d_out.abs().max()):.3e} "
          f"|gW1| {float(w1d.grad.norm()):.3e} sum|terms| est {float((d_out.abs().sum()) * enc.abs().mean()):.3e}")
    for impl in (0, 1):
        lib.immoco_set_mlp_impl(impl)
        d_enc = torch.empty_like(enc); g1 = torch.zeros_like(w1); g2 = torch.zeros_like(w2)
        nat.check(lib.immoco_mlp_bwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.contiguous().data_ptr(),
                                     d_enc.data_ptr(), g1.data_ptr(), g2.data_ptr(), n, width, act, s()), "bwd")
        torch.cuda.synchronize()
        print(f"   impl={impl}: dE {rel(d_enc.permute(1, 0, 2).reshape(n, 32), e64.grad):.2e}  gW1 {rel(g1, w1d.grad):.2e}  "
              f"gW2 {rel(g2[:2], w2d.grad[:2]):.2e}")
    # torch fp32 autograd for comparison
    e32 = enc.permute(1, 0, 2).reshape(n, 32).clone().requires_grad_(True)
    w1f = w1.clone().requires_grad_(True); w2f = w2.clone().requires_grad_(True)
    ((f(e32 @ w1f.t()) @ w2f.t())[:, :2] * d_out).sum().backward()
    print(f"   torch fp32: dE {rel(e32.grad, e64.grad):.2e}  gW1 {rel(w1f.grad, w1d.grad):.2e}  gW2 {rel(w2f.grad[:2], w2d.grad[:2]):.2e}")
lib.immoco_set_mlp_impl(1)
