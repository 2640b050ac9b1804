"""Small target for `ncu --set full`: three iterations of the C2 fit on the float-atomic path (16 kernels each);
capture the last iteration with `-s 32 -c 16`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai24_immoco_b200 as mb  # noqa: E402
from oracle import immoco_oracle as orc  # noqa: E402
mb.build()
case = orc.make_case(320, 320, 4, 1000)
model = mb.IMMoCo(case["masks"].cuda())
eng = mb.FitEngine(model, 10, deterministic=False)
k = case["kspace_motion"]; eng.set_kspace((k / k.abs().max() * 16000).cuda())
mb.lib().immoco_set_branch_overlap(0)          # serial: one stream, kernels in slot order
eng.run(mb.lambda_schedule(10, 1e-2)[:3], 1e-2)
torch.cuda.synchronize()
print("done")
