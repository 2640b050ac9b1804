"""Per-level cost of the hash-grid kernels at the C2 shapes (CUDA events, L2 flushed between runs)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat
from miccai24_immoco_b200.encoding import grid_spec
lib = mb.lib(); s = lambda: torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for dims, coords in ((3, mb.make_grids((4, 320, 320), "cuda")), (2, mb.immoco._identity_grid(320, 320, "cuda").view(-1, 2).contiguous())):
    gs = grid_spec(dims, mb.encoding_config); d = gs.desc(); n = coords.shape[0]
    table = (torch.rand(gs.n_rows, 2, device="cuda") - 0.5) * 1e-3
    enc = torch.empty(16, n, 2, device="cuda"); d_enc = torch.randn(16, n, 2, device="cuda"); grad = torch.zeros_like(table)
    def t(fn, reps=5):
        ts = []
        for _ in range(reps):
            flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        return sorted(ts)[len(ts) // 2]
    print(f"dims={dims} n={n}")
    for l in (range(16) if '--levels' in sys.argv else ()):
        f = t(lambda: lib.immoco_hashgrid_fwd_levels(C.byref(d), coords.data_ptr(), table.data_ptr(), enc.data_ptr(), n, l, l + 1, s()))
        b = t(lambda: lib.immoco_hashgrid_bwd_levels(C.byref(d), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), n, l, l + 1, s()))
        print(f"  level {l:2d} res {gs.resolutions[l]:7d} entries {gs.entries[l]:7d} hashed {gs.hashed[l]}: fwd {f:7.1f} us  bwd {b:7.1f} us")
    for impl in (0, 1):
        lib.immoco_set_hashgrid_impl(impl)
        f = t(lambda: lib.immoco_hashgrid_fwd(C.byref(d), coords.data_ptr(), table.data_ptr(), enc.data_ptr(), n, s()), 9)
        b = t(lambda: lib.immoco_hashgrid_bwd(C.byref(d), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), n, s()), 9)
        print(f"  all levels impl={impl}: fwd {f:7.1f} us  bwd {b:7.1f} us")
        e_ = enc.clone(); grad.zero_()
        lib.immoco_hashgrid_bwd(C.byref(d), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), n, s())
        if impl == 0: e0_, g0_ = e_, grad.clone()
        else: print(f"  pair vs single: enc rel {float((e_-e0_).norm()/e0_.norm()):.2e}  grad rel {float((grad-g0_).norm()/g0_.norm()):.2e}")
