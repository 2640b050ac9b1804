"""Adam kernel variants at the C2 parameter count (CUDA events; state 407 MB >> L2, no flush needed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
lib = mb.lib(); s = lambda: torch.cuda.current_stream().cuda_stream
n = 25429504
p = torch.randn(n, device="cuda"); st = torch.zeros(3, n, device="cuda"); st[0].normal_()
def run(reps=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): lib.immoco_adam_step(p.data_ptr(), st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), n, 1e-2, 0.9, 0.999, 1e-8, 5, 1, s())
    e0.record()
    for _ in range(reps): lib.immoco_adam_step(p.data_ptr(), st[0].data_ptr(), st[1].data_ptr(), st[2].data_ptr(), n, 1e-2, 0.9, 0.999, 1e-8, 5, 1, s())
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for variant in range(6):
    for ctas in (4, 8, 16, 32):
        assert lib.immoco_set_adam_tuning(variant, ctas) == 0
        us = run()
        print(f"variant {variant} (U={[1,2,4][variant%3]}, hints={variant>=3}) ctas/SM {ctas:2d}: {us:7.1f} us  {28*n/us/1e3:7.1f} GB/s")
