"""Short target for `ncu --set full`: a few iterations of the C2 fit in both accumulation modes and one kld-net
forward, so that one capture holds the dominant kernels of every row (hash-grid scatter / gather, row-sorted
gather, tensor-core convolution).   python tools/ncu_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai24_immoco_b200 as mb  # noqa: E402
from oracle import immoco_oracle as orc  # noqa: E402
from oracle import kld_net_oracle as ko  # noqa: E402

mb.build()
case = orc.make_case(320, 320, 4, 1000)
k, masks = case["kspace_motion"].cuda(), case["masks"].cuda()
for det in (False, True):
    im, _ = mb.imcoco_motion_correction(k, masks, iters=12, deterministic=det)
torch.cuda.synchronize()
net = mb.get_unet(2, 1, 32, 4, 0.0)
net.load_state_dict(ko.init_unet_state(3))
net = net.cuda()
y = net(torch.randn(4, 2, 320, 320, device="cuda"))
torch.cuda.synchronize()
print("ok", float(im.abs().mean()), float(y.abs().mean()))
