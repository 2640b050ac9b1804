"""1000-iteration C2 fits, ours (deterministic + atomic) vs the oracle loop on the same GPU (+ its 1-ulp
perturbed self), per slice: tail loss, last loss, loss spikes, final PSNR / SSIM.  The table is kept under
profiles/ (round2_long_runs.txt).   python tools/long_run_stats.py [--seeds 8] [--iters 1000]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import miccai24_immoco_b200 as mb  # noqa: E402
from tests import long_util as lu  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=8)
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--perturbed", type=int, default=3, help="1-ulp-perturbed oracle runs per slice")
    a = ap.parse_args()
    mb.build()
    t0 = time.time()
    rows = lu.compare(range(1000, 1000 + a.seeds), iters=a.iters, log=lambda s: print(s, flush=True), n_perturbed=a.perturbed)
    print(f"[{time.time() - t0:.0f} s]")
    names = ["deterministic", "atomic", "oracle_perturbed"] + [f"oracle_perturbed{j + 1}" for j in range(1, a.perturbed)]
    for mode in names:
        for key, rel in (("psnr", False), ("rmse", False), ("ssim", False), ("tail", True), ("tail_median50", True), ("last", True)):
            d = lu.spread(rows, mode, "oracle", key, rel)
            print(f"{mode:17s} vs oracle, {key:13s}{' (relative)' if rel else ''}: median {np.median(d):.4g}  max {d.max():.4g}")
    for name in ["oracle"] + names:
        last = np.asarray([r[name]["last"] for r in rows])
        tail = np.asarray([r[name]["tail"] for r in rows])
        med50 = np.asarray([r[name]["tail_median50"] for r in rows])
        print(f"{name:17s}: last-iteration loss min {last.min():.5f} max {last.max():.5f}; tail level (p10 of last 200) min "
              f"{tail.min():.5f} max {tail.max():.5f}; runs ending inside an excursion (median of last 50 > 3 x level): "
              f"{(med50 > 3 * tail).sum()} of {len(rows)}; last sample > 10 x level: {(last > 10 * tail).sum()} of {len(rows)}; "
              f"spikes (> 3 x local median) per run: {[r[name]['spikes'] for r in rows]}")
