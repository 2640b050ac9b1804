"""Grouped (bundle) hash-grid kernels + general linear row layout vs lane-pair kernels + Gray/exchange layout:
standalone kernel times (L2 flushed between runs) and us / iteration of the fused loop (FitEngine(grouped_layout=...))."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat
from miccai24_immoco_b200.encoding import grid_spec
from oracle import immoco_oracle as orc
lib = mb.lib(); s = lambda: torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
gs = grid_spec(3, mb.encoding_config)
for m in (4, 2, 8):
    h = w = 320
    coords = mb.make_grids((m, h, w), "cuda").contiguous(); n = coords.shape[0]; p = h * w
    u = torch.linspace(-1, 1, m).numpy()
    lut = gs.linear_layout(u); words = tuple(nat.LAYOUT_LUT if lut[l].any() else 0 for l in range(16))
    lut_t = torch.from_numpy(lut.view(np.int32)).cuda()
    d_lut = gs.desc(words, lut_t.data_ptr()); d_swz = gs.desc(gs.row_swizzle(u))
    table = (torch.rand(gs.n_rows, 2, device="cuda") - 0.5) * 1e-3
    enc = torch.empty(16, n, 2, device="cuda"); d_enc = torch.randn(16, n, 2, device="cuda"); grad = torch.zeros_like(table)
    f_pair = t(lambda: lib.immoco_hashgrid_fwd(C.byref(d_swz), coords.data_ptr(), table.data_ptr(), enc.data_ptr(), n, s()))
    b_pair = t(lambda: lib.immoco_hashgrid_bwd(C.byref(d_swz), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), n, s()))
    f_gs = t(lambda: lib.immoco_hashgrid_fwd_grouped(C.byref(d_swz), coords.data_ptr(), table.data_ptr(), enc.data_ptr(), p, m, s()))
    b_gs = t(lambda: lib.immoco_hashgrid_bwd_grouped(C.byref(d_swz), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), p, m, s()))
    f_gl = t(lambda: lib.immoco_hashgrid_fwd_grouped(C.byref(d_lut), coords.data_ptr(), table.data_ptr(), enc.data_ptr(), p, m, s()))
    b_gl = t(lambda: lib.immoco_hashgrid_bwd_grouped(C.byref(d_lut), coords.data_ptr(), d_enc.data_ptr(), grad.data_ptr(), p, m, s()))
    print(f"3-D grid {m}x{h}x{w}: lane-pair/swizzle fwd {f_pair:6.1f} bwd {b_pair:6.1f} | grouped/swizzle fwd {f_gs:6.1f} bwd {b_gs:6.1f} | "
          f"grouped/linear layout fwd {f_gl:6.1f} bwd {b_gl:6.1f} us", flush=True)
    del table, enc, d_enc, grad
iters = 400
for h, w, m in ((320, 320, 4), (320, 320, 2), (320, 320, 8)):
    case = orc.make_case(h, w, m, 1000)
    model = mb.IMMoCo(case["masks"].cuda())
    p_img = model.image_inr.params.detach().clone(); p_mot = model.motion_inr.params.detach().clone()
    k = case["kspace_motion"]; lam = mb.lambda_schedule(iters, 1e-2)
    for rep in range(2):
        for grouped in (False, True):
            eng = mb.FitEngine(model, iters, grouped_layout=grouped, deterministic=False)
            eng.set_kspace((k / k.abs().max() * 16000).cuda()); eng.reset(p_img, p_mot)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.run(lam, 1e-2, 0, 100); e0.record(); eng.run(lam, 1e-2, 100, iters); e1.record(); torch.cuda.synchronize()
            tr = eng.loss_trace(lam)
            print(f"{h}x{w} n_M={m} grouped={int(grouped)}: {e0.elapsed_time(e1) / (iters - 100) * 1e3:7.1f} us / iteration  "
                  f"loss {tr[0]:.4f} -> {tr[99]:.5f} -> {tr[-1]:.5f}", flush=True)
            del eng
    del model
