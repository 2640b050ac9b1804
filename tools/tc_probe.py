import ctypes as C, os, sys
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(root, "tests", "hostcheck", "_tc_probe.so"))
lib.tc_probe.argtypes = [C.c_void_p, C.c_void_p]
out = torch.full((6, 128, 32), float("nan"), device="cuda")
rc = lib.tc_probe(out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("rc", rc)
m = np.arange(128)[:, None]; k = np.arange(8)[None, :]; n = np.arange(32)[:, None]
A = ((m % 7) - 3) + 0.25 * k; B = ((n % 5) - 2) + 0.5 * k
ref = A @ B.T
names = ["SS K/K", "SS K/MN(lbo=kgrp,sbo=mngrp)", "SS K/MN swapped", "TS K/K", "TS K/MN", "TS K/MN swapped"]
o = out.cpu().numpy()
for v in range(6):
    err = np.abs(o[v] - ref).max()
    print(f"{names[v]:32s} max|err| {err:.3g}  sample {o[v][1,:4]} ref {ref[1,:4]}")
