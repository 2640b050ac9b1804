"""Runs the tensor-core MLP kernels at the C2 shapes a few times (target for an ncu capture)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
lib = mb.lib(); s = lambda: torch.cuda.current_stream().cuda_stream
for width, act, n in ((64, 2, 409600), (256, 1, 102400)):
    enc = torch.randn(16, n, 2, device="cuda") * 1e-2
    w1 = torch.randn(width, 32, device="cuda") * 0.2; w2 = torch.randn(16, width, device="cuda") * 0.2
    out = torch.empty(n, 2, device="cuda"); d_out = torch.randn(n, 2, device="cuda")
    d_enc = torch.empty_like(enc); g1 = torch.zeros_like(w1); g2 = torch.zeros_like(w2)
    for _ in range(3):
        lib.immoco_mlp_fwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), out.data_ptr(), n, width, act, 1, s())
        lib.immoco_mlp_bwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(), d_enc.data_ptr(),
                           g1.data_ptr(), g2.data_ptr(), n, width, act, s())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.immoco_mlp_bwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(), d_enc.data_ptr(),
                           g1.data_ptr(), g2.data_ptr(), n, width, act, s())
    e1.record(); torch.cuda.synchronize()
    print(f"bwd W={width} n={n}: {e0.elapsed_time(e1)/10*1e3:.1f} us")
