#!/usr/bin/env python
"""IM-MoCo hot-path benchmark (BASELINE.json metric: slices/s, 1000 iterations, 320x320, n_M=4).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU reference arm

A *step* is one pass of the hot path over one batch: every rank reconstructs ONE synthetic slice
(fresh INR parameters, `--iters` optimisation iterations).  Rank 0 prints one JSON line.

  value   : whole-job slices/s, inputs resident in HBM, native loop, CUDA-event timed, max over ranks
  e2e     : same metric through the public API ``imcoco_motion_correction`` with HOST (pinned)
            k-space + masks in, corrected image out, copies inside the timed region
  roofline: dominant kernel (largest share of the step, CUDA events recorded around every kernel
            of every 100th iteration inside the timed region), algorithmic bytes / measured time
  cpu_baseline: the oracle's torch-CPU loop (hash-grid INRs) on a bounded sample of iterations
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "IM-MoCo slices/sec (1000 iters, 320x320)"
H = W = 320
N_MOV = 4
POOL = 4                      # distinct synthetic slices per rank, cycled over the steps


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(slot: str, p: int, m: int, n_par: int, t_img: int, t_mot: int) -> float:
    """Algorithmic bytes per launch of each kernel of one iteration (sums to
    SURVEY 8(d)'s B_iter = 28 N_par + 16 (T_img + T_mot) + 64 P (M+1) plus the feature planes the
    un-fused round-1 kernels still round-trip)."""
    mp = m * p
    table = {
        "adam_motion": 28.0 * n_par[0], "adam_image": 28.0 * n_par[1],
        "hashgrid_fwd_image": 8.0 * t_img + 8.0 * p + 128.0 * p,
        "hashgrid_bwd_image": 8.0 * t_img + 8.0 * p + 128.0 * p,
        "hashgrid_fwd_motion": 8.0 * t_mot + 12.0 * mp + 128.0 * mp,
        "hashgrid_bwd_motion": 8.0 * t_mot + 12.0 * mp + 128.0 * mp,
        "mlp_fwd_image": 128.0 * p + 8.0 * p,
        "mlp_fwd_motion": 128.0 * mp + 8.0 * mp,
        "mlp_bwd_image": 256.0 * p + 8.0 * p,
        "mlp_bwd_motion": 256.0 * mp + 8.0 * mp,
        # the static row passes are folded into the fused row launches (slots motion_rows_fwd / _bwd);
        # their own slots bracket nothing and are dropped from the report
        "fft_rows": 0.0, "fft_rows_adj": 0.0,
        "motion_rows_fwd": 16.0 * p + 8.0 * p + 16.0 * mp, "motion_rows_bwd": 24.0 * p + 24.0 * p + 24.0 * mp,
        "colpass_loss": 32.0 * p, "grad_entropy": 16.0 * p,
    }
    return table[slot]


# what actually paces each kernel (ncu --set full, profiles/round1_v8_ncu_full.txt / round1_v9_ncu_full.txt)
LIMITERS = {
    "hashgrid_bwd_motion": "L1 wavefronts + L2 atomic throughput (ncu v9: l1tex 78 %, lts 74 %, DRAM 10 %): 52.4 M 8-byte "
                           "reductions on hashed rows with no locality across pixels; DRAM traffic equals the algorithmic bytes",
    "hashgrid_fwd_motion": "L1 wavefronts (ncu v9: l1tex 75 %, lts 57 %, DRAM 10 %): 52.4 M 8-byte gathers on hashed rows, "
                           "one 128-byte line per lane pair",
    "hashgrid_bwd_image": "L2 atomic throughput", "hashgrid_fwd_image": "L1 wavefronts / L2 gather rate",
    "adam_motion": "HBM (28 B / parameter at ~5.9 TB/s; the zeroing of the gradients runs as a memset on a third stream)",
    "adam_image": "HBM",
    "mlp_bwd_motion": "barrier / MMA-chain latency + SIMT epilogue (ncu v8: issue active 36 %, tensor pipe 27 %)",
    "mlp_bwd_image": "MMA-chain waits + SIMT epilogue (ncu v8: issue active 34 %, tensor pipe 27 %)",
    "mlp_fwd_motion": "latency, 16 warps per SM (ncu v8: issue active 40 %)", "mlp_fwd_image": "latency (2.7 tiles per CTA)",
}

# SURVEY 8(d) / oracle.touched_entries(): distinct table rows touched at 320x320
T_IMG_320 = 3041608
T_MOT_320 = {2: 5917982, 4: 6513775, 8: 6756779}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import miccai24_immoco_b200 as mb
    from miccai24_immoco_b200 import _native as nat
    from oracle import immoco_oracle as orc   # synthetic-input generator + cpu_baseline leg only

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        mb.build()
    if world > 1:
        dist.barrier()
    lib = mb.lib()
    iters = args.iters
    lambdas = mb.lambda_schedule(iters, 1e-2)

    # ---- synthetic slices: seed 1000 + global slice index -------------------------
    cases, models, engines, k_dev, k_host, masks_host = [], [], [], [], [], []
    for i in range(POOL):
        case = orc.make_case(H, W, N_MOV, 1000 + rank * POOL + i)
        masks = case["masks"]
        model = mb.IMMoCo(masks.to(dev), image_seed=11 + i, motion_seed=101 + i)
        eng = mb.FitEngine(model, iters)
        k = case["kspace_motion"]
        k_norm = (k / k.abs().max() * 16000).to(torch.complex64)
        cases.append(case)
        models.append(model)
        engines.append(eng)
        k_dev.append(k_norm.to(dev))
        k_host.append(k.to(torch.complex64).pin_memory())
        masks_host.append(masks.pin_memory())
    init_img = models[0].image_inr.params.detach().clone()
    init_mot = models[0].motion_inr.params.detach().clone()
    gathered = [torch.empty((H, W, 2), device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    prof = lib.immoco_profile_create(max(1, (iters // 100 + 1) * args.steps))

    def step_resident(i, profile=None):
        eng = engines[i % POOL]
        eng.reset(init_img, init_mot)
        eng.k_in.copy_(torch.view_as_real(k_dev[i % POOL]))
        eng.run(lambdas, 1e-2, profile=profile, profile_every=100 if profile else 0)
        if world > 1:   # the path's only collective: corrected images to rank 0 (NCCL gather)
            dist.gather(eng.image, gathered, dst=0)
        return eng

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: HBM-resident inputs, native loop ------------------------------------------------
    for i in range(args.warmup):
        step_resident(i)
    sync_all()
    launches0 = sum(e.launches for e in engines)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step_resident(args.warmup + i, profile=prof)
    ev1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total)
    launches = sum(e.launches for e in engines) - launches0
    last = engines[(args.warmup + args.steps - 1) % POOL]
    final_loss = float(last.loss_trace(lambdas)[-1])
    import ctypes as C
    ms_sum = (C.c_float * len(nat.PROFILE_SLOTS))()
    n_prof = lib.immoco_profile_read(prof, ms_sum)
    lib.immoco_profile_destroy(prof)

    # ---- e2e: public API, host buffers in, host image out ---------------------------------------
    def step_e2e(i):
        j = i % POOL
        im, _ = mb.imcoco_motion_correction(k_host[j], masks_host[j], iters, 1e-2, 1e-2, False)
        return im.cpu()

    for e in engines:          # free the resident engines' HBM before the API path allocates its own
        del e
    engines.clear()
    n_e2e = max(1, min(args.steps, args.e2e_steps))
    step_e2e(0)
    sync_all()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        out_img = step_e2e(1 + i)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel --------------------------------------------------------
    p = H * W
    n_par2 = (models[0].motion_inr.n_params, models[0].image_inr.n_params)
    n_par = sum(n_par2)
    per_slot = {s: (ms_sum[k] / n_prof if n_prof else 0.0) for k, s in enumerate(nat.PROFILE_SLOTS)
                if algorithmic_bytes(s, p, N_MOV, n_par2, T_IMG_320, T_MOT_320[N_MOV]) > 0}
    iter_ms = sum(per_slot.values())
    dom = max(per_slot, key=per_slot.get) if n_prof else "adam_motion"
    peak, peak_src = measured_peak_gbs()
    dom_bytes = algorithmic_bytes(dom, p, N_MOV, n_par2, T_IMG_320, T_MOT_320[N_MOV])
    achieved = dom_bytes / (per_slot[dom] * 1e-3) / 1e9 if per_slot.get(dom) else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(dom)
    except Exception:
        pass
    b_iter = 28.0 * n_par + 16.0 * (T_IMG_320 + T_MOT_320[N_MOV]) + 64.0 * p * (N_MOV + 1)
    ms_per_iter = ms_total / (args.steps * iters)
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
        "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": round(per_slot[dom], 5),
        "kernel_share_of_iteration": round(per_slot[dom] / iter_ms, 4) if iter_ms else None,
        "per_kernel_ms": {k: round(v, 5) for k, v in per_slot.items()},
        "per_kernel_gbs": {k: round(algorithmic_bytes(k, p, N_MOV, n_par2, T_IMG_320, T_MOT_320[N_MOV]) / (v * 1e-3) / 1e9, 1)
                           for k, v in per_slot.items() if v > 0},
        "instrumented_iterations": n_prof,
        "limiter": LIMITERS.get(dom, ""),
        "iteration": {"algorithmic_bytes": b_iter, "achieved": round(b_iter / (ms_per_iter * 1e-3) / 1e9, 1),
                      "frac": round(b_iter / (ms_per_iter * 1e-3) / 1e9 / peak, 4)},
    }

    # ---- cpu_baseline: oracle loop (torch CPU, hash-grid INRs) on a bounded sample ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(orc, torch, n_iters=args.cpu_iters)

    value = world * args.steps / (ms_total * 1e-3)
    line = {
        "metric": METRIC, "value": round(value, 4), "unit": "slices/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3),
        "ms_per_iter": round(ms_per_iter, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2: one {H}x{W} single-coil slice per GPU per step, n_M={N_MOV}, hash-grid "
                               f"Image INR + Motion INR, {iters} iterations, fresh parameters per slice",
                   "iters": iters, "slices_per_step": world, "parallelism": f"instance-sharded x{world}, "
                   "one NCCL gather of corrected images per step" if world > 1 else "single GPU",
                   "l2_policy": "per-slice optimiser state (407 MB fp32) is streamed every iteration and exceeds the "
                                "126 MB L2; no explicit flush", "final_loss": final_loss},
        "clocks": clocks,
        "e2e": {"value": round(world * n_e2e / e2e_s, 4), "unit": "slices/s", "steps": n_e2e,
                "h2d_bytes_per_step": int(k_host[0].numel() * 8 + masks_host[0].numel() * 8),
                "d2h_bytes_per_step": int(out_img.numel() * 8),
                "api": "imcoco_motion_correction(kspace_host, masks_host, iters, lr, lambda_ge, debug) -> image.cpu()"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_sample(orc, torch, n_iters: int, warm: int = 1):
    """Oracle loop on the host cores for a bounded number of iterations of the C2 workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    case = orc.make_case(H, W, N_MOV, 1000)
    st = orc.LoopState(case["kspace_motion"], case["masks"], 1000)
    for j in range(warm):
        st.step(j)
    t0 = time.perf_counter()
    for j in range(warm, warm + n_iters):
        st.step(j)
    dt = (time.perf_counter() - t0) / n_iters
    return {"value": round(1.0 / (1000 * dt), 6), "unit": "slices/s", "cores": torch.get_num_threads(),
            "kind": "port", "ms_per_iter": round(dt * 1e3, 1),
            "sample": f"{n_iters} timed iterations (after {warm} warm-up) of the same C2 slice with the oracle's "
                      f"torch-CPU loop (hash-grid INRs, torch.optim.Adam), extrapolated to 1000 iterations"}


def run_reference(args):
    """CPU reference arm: the reference is pure Python (torch) + tiny-cuda-nn; its own files are not
    on the GPU box and tiny-cuda-nn is absent everywhere, so this times the oracle port (the
    reference's loop semantics on torch CPU with the torch fp32 hash-grid stand-in)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import torch

    from oracle import immoco_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    case = orc.make_case(H, W, N_MOV, 1000)
    st = orc.LoopState(case["kspace_motion"], case["masks"], 1000)
    per_step = args.ref_iters
    j = 0
    for _ in range(args.warmup):
        for _ in range(per_step):
            st.step(j)
            j += 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step):
            st.step(j)
            j += 1
    dt = time.perf_counter() - t0
    ms_iter = dt / (args.steps * per_step) * 1e3
    value = 1.0 / (args.iters * ms_iter * 1e-3)
    sample = (f"each step = {per_step} iterations of the C2 slice (oracle port of immoco.py:116-206 on torch CPU, "
              f"hash-grid INRs), extrapolated to {args.iters} iterations per slice")
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": "slices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 1), "ms_per_iter": round(ms_iter, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2: one {H}x{W} single-coil slice, n_M={N_MOV}, hash-grid Image INR + Motion INR, "
                               f"{args.iters} iterations", "iters": args.iters},
        "cpu_baseline": {"value": round(value, 6), "unit": "slices/s", "cores": torch.get_num_threads(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 6), "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--iters", type=int, default=1000, help="optimisation iterations per slice (metric: 1000)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-iters", type=int, default=15)
    ap.add_argument("--ref-iters", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
