"""Instance-batched fit (immoco_fit_run_batched): us per SLICE-iteration for B = 1, 2, 4, 8 slices in lock step,
both accumulation modes, with the per-kernel durations of the batched launches.  Kept as
profiles/round2_batched_fit.txt.   python tools/batch_fit_bench.py [--shape 320 320 4] [--iters 200]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai24_immoco_b200 as mb  # noqa: E402
from miccai24_immoco_b200 import _native as nat  # noqa: E402
from oracle import immoco_oracle as orc  # noqa: E402
from tests.gpu_util import case_params  # noqa: E402

DEV = "cuda"

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=3, default=[320, 320, 4])
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--batches", type=int, nargs="+", default=[1, 2, 4, 8])
    a = ap.parse_args()
    h, w, m = a.shape
    mb.build()
    lib = mb.lib()
    lam = mb.lambda_schedule(max(a.iters, 10), 1e-2)[:a.iters]
    for det in (False, True):
        for nb in a.batches:
            engines = []
            for i in range(nb):
                case = orc.make_case(h, w, m, 1000 + i)
                model = mb.IMMoCo(case["masks"].to(DEV))
                eng = mb.FitEngine(model, a.iters, deterministic=det)
                k = case["kspace_motion"]
                eng.set_kspace((k / k.abs().max() * 16000).to(DEV))
                engines.append(eng)
            inits = [(e.image_params(), e.motion_params()) for e in engines]
            mb.run_batched(engines, lam, 1e-2, 0, min(20, a.iters))
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                for e, (pi, pm) in zip(engines, inits):
                    e.reset(pi, pm)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                mb.run_batched(engines, lam, 1e-2)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / (a.iters * nb) * 1e3)
            prof = lib.immoco_profile_create(32)
            for e, (pi, pm) in zip(engines, inits):
                e.reset(pi, pm)
            mb.run_batched(engines, lam, 1e-2, 0, min(a.iters, 100), profile=prof, profile_every=10)
            torch.cuda.synchronize()
            ms = (C.c_float * len(nat.PROFILE_SLOTS))()
            n = lib.immoco_profile_read(prof, ms)
            lib.immoco_profile_destroy(prof)
            per = {s: round(ms[i] / n / nb * 1e3, 1) for i, s in enumerate(nat.PROFILE_SLOTS) if ms[i] > 0}
            print(f"[{h}x{w} M={m}] {'deterministic' if det else 'atomic':13s} B={nb}: {best:7.1f} us per slice-iteration; "
                  f"serial per-kernel us per slice: {per}", flush=True)
            del engines
