"""tanh_small (mlp_tc.cu) must equal CUDA's tanhf bit for bit below 0.6: compare the tensor-core MLP forward
(64 wide, tanh) on inputs scaled so that (a) all pre-activations are small and (b) some are large."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
lib = mb.lib(); s = lambda: torch.cuda.current_stream().cuda_stream
n = 40000
g = torch.Generator(device="cuda").manual_seed(1)
for scale in (1e-2, 3.0):
    enc = torch.randn(16, n, 2, device="cuda", generator=g) * scale
    w1 = torch.randn(64, 32, device="cuda", generator=g) * 0.2; w2 = torch.randn(16, 64, device="cuda", generator=g) * 0.2
    out = torch.empty(n, 2, device="cuda")
    lib.immoco_mlp_fwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), out.data_ptr(), n, 64, 2, 1, s())
    e = enc.permute(1, 0, 2).reshape(n, 32).double()
    z = e @ w1.double().t()
    ref = torch.tanh(torch.tanh(z) @ w2[:2].double().t())
    print(f"scale {scale}: max|z| {float(z.abs().max()):.3f}  rel-L2 vs float64 {float((out.double()-ref).norm()/ref.norm()):.2e}")
