"""Steady-state iteration time vs the SM share of the backward hash-grid scatter (thin persistent grid) for
the library variant selected by IMMOCO_LIB_PATH (register budget of the MLP backward kernels)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat
from oracle import immoco_oracle as orc
lib = mb.lib()
print("lib:", os.path.basename(nat.LIB_PATH), flush=True)
case = orc.make_case(320, 320, 4, 1000)
model = mb.IMMoCo(case["masks"].cuda())
eng = mb.FitEngine(model, 600)
k = case["kspace_motion"]; eng.set_kspace((k / k.abs().max() * 16000).cuda())
lam = mb.lambda_schedule(600, 1e-2)
caps = [int(c) for c in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0,1,2,4".split(","))]
for rep in range(2):
    for cap in caps:
        lib.immoco_set_hashgrid_bwd_ctas_per_sm(cap)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.run(lam, 1e-2, 0, 100); e0.record(); eng.run(lam, 1e-2, 100, 600); e1.record(); torch.cuda.synchronize()
        print(f"bwd ctas/SM={cap}: {e0.elapsed_time(e1)/500*1e3:.1f} us / iteration   final loss {eng.loss_trace(lam)[599]:.5f}", flush=True)
if "--timeline" in sys.argv:
    for cap in caps:
        lib.immoco_set_hashgrid_bwd_ctas_per_sm(cap)
        lib.immoco_set_profile_overlap(1)
        prof = lib.immoco_profile_create(8)
        eng.run(lam, 1e-2, 0, 400, profile=prof, profile_every=100)
        torch.cuda.synchronize()
        b = (C.c_float * len(nat.PROFILE_SLOTS))(); e = (C.c_float * len(nat.PROFILE_SLOTS))()
        assert lib.immoco_profile_timeline(prof, 2, b, e) == 0
        print(f"--- two-stream timeline, bwd ctas/SM={cap} ---")
        for i in sorted(range(len(nat.PROFILE_SLOTS)), key=lambda i: b[i]):
            print(f"{nat.PROFILE_SLOTS[i]:22s} begin {b[i]*1e3:7.1f} us  end {e[i]*1e3:7.1f} us  dur {(e[i]-b[i])*1e3:6.1f}")
        lib.immoco_profile_destroy(prof)
