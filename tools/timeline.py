"""Two-stream timeline of one steady-state iteration (CUDA events around every kernel, overlap kept)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from miccai24_immoco_b200 import _native as nat
from oracle import immoco_oracle as orc
lib = mb.lib()
case = orc.make_case(320, 320, 4, 1000)
model = mb.IMMoCo(case["masks"].cuda())
eng = mb.FitEngine(model, 400)
k = case["kspace_motion"]; eng.set_kspace((k / k.abs().max() * 16000).cuda())
lam = mb.lambda_schedule(400, 1e-2)
for mode in (0, 1):
    lib.immoco_set_profile_overlap(mode)
    prof = lib.immoco_profile_create(8)
    eng.run(lam, 1e-2, 0, 400, profile=prof, profile_every=100)
    torch.cuda.synchronize()
    b = (C.c_float * len(nat.PROFILE_SLOTS))(); e = (C.c_float * len(nat.PROFILE_SLOTS))()
    assert lib.immoco_profile_timeline(prof, 2, b, e) == 0
    print(f"--- instrumented iteration, {'two-stream' if mode else 'serial'} ---")
    for i in sorted(range(len(nat.PROFILE_SLOTS)), key=lambda i: b[i]):
        print(f"{nat.PROFILE_SLOTS[i]:22s} begin {b[i]*1e3:7.1f} us  end {e[i]*1e3:7.1f} us  dur {(e[i]-b[i])*1e3:6.1f}")
    print(f"iteration span {max(e)*1e3:.1f} us")
    lib.immoco_profile_destroy(prof)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.run(lam, 1e-2, 0, 100); e0.record(); eng.run(lam, 1e-2, 100, 400); e1.record(); torch.cuda.synchronize()
print(f"steady state: {e0.elapsed_time(e1)/300*1e3:.1f} us / iteration")
