"""kld-net inference timing (config 4: batch of 64 synthetic 320x320 slices) vs torch/cuDNN fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from oracle import kld_net_oracle as ko
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
state = ko.init_unet_state(3); net = mb.get_unet(2, 1, 32, 4, 0.0); net.load_state_dict(state); net = net.cuda()
sd = {k: v.cuda() for k, v in state.items()}
for n, h, w in ((1, 320, 320), (16, 320, 320), (64, 320, 320), (16, 640, 368)):
    x = torch.randn(n, 2, h, w, device="cuda")
    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
    net.tensor_cores = True; ours = t(lambda: net(x))
    net.tensor_cores = False; simt = t(lambda: net(x)); net.tensor_cores = True
    ref = t(lambda: ko.unet_forward(sd, x))
    gf = 37.7e9 * n * (h * w) / (320 * 320)
    print(f"{n:3d} x {h}x{w}: tcgen05 3xTF32 {ours:8.2f} ms ({gf/ours/1e9:6.1f} TFLOP/s fp32-equivalent)   fp32 SIMT {simt:8.2f} ms   "
          f"torch/cuDNN fp32 {ref:8.2f} ms", flush=True)
