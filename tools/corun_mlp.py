"""Independent co-run on two streams: SM-bound MLP kernels (64-wide, 409,600 points) beside the L2-bound 3-D
hash-grid kernels.  Upper bound for what chunk-pipelining MLP -> hash-grid inside a branch could recover."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200.encoding import grid_spec
lib = mb.lib()
coords = mb.make_grids((4, 320, 320), "cuda"); gs = grid_spec(3, mb.encoding_config); d = gs.desc(); n = coords.shape[0]
table = (torch.rand(gs.n_rows, 2, device="cuda") - 0.5) * 1e-3
enc2 = torch.empty(16, n, 2, device="cuda"); d_enc2 = torch.randn(16, n, 2, device="cuda"); grad = torch.zeros_like(table)
width, act, nn = 64, 2, 409600
enc = torch.randn(16, nn, 2, device="cuda") * 1e-2
w1 = torch.randn(width, 32, device="cuda") * 0.2; w2 = torch.randn(16, width, device="cuda") * 0.2
out = torch.empty(nn, 2, device="cuda"); d_out = torch.randn(nn, 2, device="cuda")
d_enc = torch.empty_like(enc); g1 = torch.zeros_like(w1); g2 = torch.zeros_like(w2)
def mlp_bwd(s): lib.immoco_mlp_bwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), d_out.data_ptr(), d_enc.data_ptr(), g1.data_ptr(), g2.data_ptr(), nn, width, act, s.cuda_stream)
def mlp_fwd(s): lib.immoco_mlp_fwd(enc.data_ptr(), w1.data_ptr(), w2.data_ptr(), out.data_ptr(), nn, width, act, 1, s.cuda_stream)
def hgf(s): lib.immoco_hashgrid_fwd(C.byref(d), coords.data_ptr(), table.data_ptr(), enc2.data_ptr(), n, s.cuda_stream)
def hgb(s): lib.immoco_hashgrid_bwd(C.byref(d), coords.data_ptr(), d_enc2.data_ptr(), grad.data_ptr(), n, s.cuda_stream)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
def timed(fns, reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        for f, s in fns: f(s)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
timed([(hgf, s1)]); timed([(hgf, s1)])      # clocks up
for name, a, b, bn in (("mlp_bwd64", mlp_bwd, hgb, "hashgrid_bwd"), ("mlp_fwd64", mlp_fwd, hgf, "hashgrid_fwd"), ("mlp_bwd64", mlp_bwd, hgf, "hashgrid_fwd")):
    ta, tb = timed([(a, s2)]), timed([(b, s1)])
    both = timed([(b, s1), (a, s2)]); both_r = timed([(a, s1), (b, s2)])
    print(f"{name} alone {ta:6.1f} us, {bn} alone {tb:6.1f} us, sum {ta+tb:6.1f}; co-run {both:6.1f} us (MLP high priority) / {both_r:6.1f} us (hash-grid high priority)", flush=True)
