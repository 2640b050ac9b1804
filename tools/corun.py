"""Do the L2-bound hash-grid kernels and the HBM-bound Adam overlap when co-run on two streams?"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200.encoding import grid_spec
lib = mb.lib()
coords = mb.make_grids((4, 320, 320), "cuda"); gs = grid_spec(3, mb.encoding_config); d = gs.desc(); n = coords.shape[0]
table = (torch.rand(gs.n_rows, 2, device="cuda") - 0.5) * 1e-3
enc = torch.empty(16, n, 2, device="cuda"); d_enc = torch.randn(16, n, 2, device="cuda"); grad = torch.zeros_like(table)
na = 14232576
p = torch.randn(na, device="cuda"); st = torch.zeros(3, na, device="cuda"); st[0].normal_()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
def hgf(s): lib.immoco_hashgrid_fwd(C.byref(d), coords.data_ptr(), table.data_ptr(), enc.data_ptr(), n, s.cuda_stream)
def hgb(s): lib.immoco_hashgrid_bwd(C.byref(d), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), n, s.cuda_stream)
def adam(s): lib.immoco_adam_step(p.data_ptr(), st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), na, 1e-2, 0.9, 0.999, 1e-8, 5, 1, s.cuda_stream)
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
for name, a, b in (("hashgrid_fwd", hgf, adam), ("hashgrid_bwd", hgb, adam), ("hashgrid_fwd", hgf, hgb)):
    bn = "adam" if b is adam else "hashgrid_bwd"
    ta, tb = timed([(a, s1)]), timed([(b, s1)])
    both = timed([(a, s1), (b, s2)])
    print(f"{name} alone {ta:6.1f} us, {bn} alone {tb:6.1f} us, serial sum {ta+tb:6.1f}; co-run on two streams {both:6.1f} us", flush=True)
