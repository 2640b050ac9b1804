"""ncu target: one launch each of the lane-pair and the grouped hash-grid kernels (forward and backward) at the C2 shape."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat
from miccai24_immoco_b200.encoding import grid_spec
lib = mb.lib(); s = lambda: torch.cuda.current_stream().cuda_stream
gs = grid_spec(3, mb.encoding_config)
m, h, w = 4, 320, 320
coords = mb.make_grids((m, h, w), "cuda").contiguous(); n = coords.shape[0]; p = h * w
u = torch.linspace(-1, 1, m).numpy()
lut = gs.linear_layout(u); words = tuple(nat.LAYOUT_LUT if lut[l].any() else 0 for l in range(16))
lut_t = torch.from_numpy(lut.view(np.int32)).cuda()
d_lut = gs.desc(words, lut_t.data_ptr()); d_swz = gs.desc(gs.row_swizzle(u))
table = (torch.rand(gs.n_rows, 2, device="cuda") - 0.5) * 1e-3
enc = torch.empty(16, n, 2, device="cuda"); d_enc = torch.randn(16, n, 2, device="cuda"); grad = torch.zeros_like(table)
for _ in range(2):
    lib.immoco_hashgrid_fwd(C.byref(d_swz), coords.data_ptr(), table.data_ptr(), enc.data_ptr(), n, s())
    lib.immoco_hashgrid_fwd_grouped(C.byref(d_lut), coords.data_ptr(), table.data_ptr(), enc.data_ptr(), p, m, s())
    lib.immoco_hashgrid_bwd(C.byref(d_swz), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), n, s())
    lib.immoco_hashgrid_bwd_grouped(C.byref(d_lut), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), p, m, s())
torch.cuda.synchronize()
print("done")
