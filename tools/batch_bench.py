"""Throughput of reconstruct_batch vs slices in flight (C2 slices, pinned host inputs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import miccai24_immoco_b200 as mb
from oracle import immoco_oracle as orc
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n = 6
cases = [orc.make_case(320, 320, 4, 1000 + i) for i in range(n)]
ks = [c["kspace_motion"].to(torch.complex64).pin_memory() for c in cases]
ms = [c["masks"].pin_memory() for c in cases]
mb.reconstruct_batch(ks[:2], ms[:2], 20, in_flight=2)
torch.cuda.synchronize()
for in_flight in (1, 2, 3):
    for chunk in (10,):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = mb.reconstruct_batch(ks, ms, iters, in_flight=in_flight, chunk=chunk)
        host = [o.cpu() for o in out]
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"in_flight {in_flight} chunk {chunk:3d}: {dt/n/iters*1e6:7.1f} us per slice-iteration  "
              f"({n/dt*iters/1000:.3f} slices/s at 1000 its)  |img| {float(host[0].abs().mean()):.4f}", flush=True)
