"""Per-layer timing of the kld-net 3x3 convolutions at config 4's batch (64 x 320x320): tcgen05 kernel vs the fp32
SIMT kernel vs cuDNN fp32, every distinct (cin, cout, resolution) of the network.   python tools/conv_layers_bench.py"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai24_immoco_b200 as mb  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
lib = mb.lib()
N = int(os.environ.get("N", "64"))
LAYERS = [(32, 0, 32, 320), (32, 32, 32, 320), (32, 0, 64, 160), (64, 0, 64, 160), (64, 64, 64, 160), (64, 0, 128, 80),
          (128, 0, 128, 80), (128, 128, 128, 80), (128, 0, 256, 40), (256, 0, 256, 40), (256, 256, 256, 40),
          (256, 0, 512, 20), (512, 0, 512, 20)]
COUNT = {(32, 0, 32, 320): 2, (64, 0, 64, 160): 2, (128, 0, 128, 80): 2, (256, 0, 256, 40): 2}


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tot = {"tc": 0.0, "simt": 0.0, "cudnn": 0.0}
for c0, c1, cout, hw in LAYERS:
    cin = c0 + c1
    x0 = torch.randn(N, c0, hw, hw, device="cuda")
    x1 = torch.randn(N, c1, hw, hw, device="cuda") if c1 else None
    wt = torch.randn(cout, cin, 3, 3, device="cuda") / (3 * cin ** 0.5)
    w_hi = torch.empty((cin // 4) * 9 * cout * 4, device="cuda")
    w_lo = torch.empty_like(w_hi)
    s = torch.cuda.current_stream().cuda_stream
    lib.immoco_unet_pack_conv3x3(wt.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), cout, cin, s)
    out = torch.empty(N, cout, hw, hw, device="cuda")
    stats = torch.zeros(N, cout, 2, dtype=torch.float64, device="cuda")
    p1 = 0 if x1 is None else x1.data_ptr()
    tc = t(lambda: lib.immoco_unet_conv3x3_tc(x0.data_ptr(), c0, p1, c1, w_hi.data_ptr(), w_lo.data_ptr(), out.data_ptr(),
                                              stats.data_ptr(), N, cout, hw, hw, s))
    simt = t(lambda: lib.immoco_unet_conv3x3(x0.data_ptr(), c0, p1, c1, wt.data_ptr(), out.data_ptr(), stats.data_ptr(), N,
                                             cout, hw, hw, s))
    xin = x0 if x1 is None else torch.cat([x0, x1], 1)
    cud = t(lambda: F.conv2d(xin, wt, padding=1))
    gf = 2.0 * N * hw * hw * cin * cout * 9 / 1e9
    k = COUNT.get((c0, c1, cout, hw), 1)
    tot["tc"] += k * tc
    tot["simt"] += k * simt
    tot["cudnn"] += k * cud
    print(f"{N} x {cin:3d}->{cout:3d} @{hw:3d}^2 (x{k}): {gf:7.1f} GFLOP  tcgen05 {tc:6.2f} ms ({gf / tc:6.1f} TF/s)  SIMT {simt:6.2f} ms  "
          f"cuDNN fp32 {cud:6.2f} ms", flush=True)
print(f"all 3x3 layers of the network: tcgen05 {tot['tc']:.1f} ms, SIMT {tot['simt']:.1f} ms, cuDNN fp32 {tot['cudnn']:.1f} ms")
