#!/usr/bin/env python
"""IM-MoCo hot-path benchmark (BASELINE.json metric: slices/s, 1000 iterations, 320x320, n_M=4).

    python bench.py --gpus N --steps K --warmup W                       # configs[1] (C2), the headline line
    python bench.py --impl reference --gpus N --steps K --warmup W      # the CPU reference arm (same config)
    python bench.py --config c3|c4|c5 [--n-mov M] ...                   # the other configurations BASELINE.json names
    python bench.py --config c1                                         # configs[0]: the Fourier-feature CPU case

A *step* is one pass of the hot path over one batch.  Rank 0 prints ONE JSON line.

  c2 (default): every rank reconstructs ONE 320x320 slice per step (n_M=4, fresh INR parameters, `--iters`
       iterations).  value = slices/s with k-space resident in HBM (native loop, CUDA events, max over ranks);
       e2e = the same through ``imcoco_motion_correction`` with HOST (pinned) k-space + masks in and the image
       copied back; roofline = the dominant kernel (CUDA events around every kernel of every 100th iteration of
       the timed region); cpu_baseline = the oracle's torch-CPU loop (hash-grid INRs) on a bounded sample;
       cpu_baseline_c1 = configs[0] (Fourier-feature INRs, n_M=2, torch CPU, cores stated);
       torch_gpu_baseline = the oracle's torch loop on the SAME GPU (BASELINE.md section 4's last column).
  c3:  a step = one volume of 16 slices of 640x368, n_M=5, dealt over the ranks by ``reconstruct_slices``
       (round-robin instance sharding, one NCCL gather of the corrected images) -- scaling "strong".
  c5:  a step = 32 slices of 320x320 with n_M = --n-mov in {2,4,8} through the same path; 8 steps = the
       256-slice sweep -- scaling "strong".
  c4:  a step = a batch of 64 slices: kld-net inference + movement-group extraction (``movement_masks_from_kspace``)
       feeding the fits of the batch (``reconstruct_slices``); value = slices/s of the whole pipeline.
Warm-up steps of c3 / c4 / c5 reconstruct ONE slice per rank (clocks, allocator pools, per-shape tap lists);
the timed steps run the full batch.
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
POOL = 4                      # distinct synthetic slices per rank, cycled over the steps

# BASELINE.json configs[] (index = position in the list)
CONFIGS = {
    "c1": dict(index=0, h=320, w=320, n_mov=2, slices=1, step_slices=1),
    "c2": dict(index=1, h=320, w=320, n_mov=4, slices=1, step_slices=1),
    "c3": dict(index=2, h=640, w=368, n_mov=5, slices=16, step_slices=16),
    "c4": dict(index=3, h=320, w=320, n_mov=4, slices=64, step_slices=64),
    "c5": dict(index=4, h=320, w=320, n_mov=4, slices=256, step_slices=32),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def workload_string(name: str, n_mov: int, iters: int) -> str:
    """ONE wording per configuration, shared by both arms (the driver compares the strings)."""
    c = CONFIGS[name]
    shape = f"{c['h']}x{c['w']}"
    if name == "c1":
        return (f"C1: one {shape} single-coil slice, n_M={n_mov}, Fourier-feature Image INR + Motion INR on the host "
                f"CPU, {iters} iterations")
    if name == "c2":
        return (f"C2: one {shape} single-coil slice per GPU per step, n_M={n_mov}, hash-grid Image INR + Motion INR, "
                f"{iters} iterations, fresh parameters per slice")
    if name == "c3":
        return (f"C3: one volume of {c['slices']} slices of {shape} per step, n_M={n_mov}, sharded per slice over the "
                f"ranks, {iters} iterations per slice")
    if name == "c4":
        return (f"C4: batch of {c['slices']} slices of {shape} per step: kld-net mask inference + movement-group "
                f"extraction feeding the IM-MoCo fits ({iters} iterations per slice)")
    return (f"C5: {c['step_slices']} slices of {shape} per step (8 steps = the {c['slices']}-slice sweep), n_M={n_mov}, "
            f"sharded per slice over the ranks, {iters} iterations per slice")


def config_dict(name: str, n_mov: int, iters: int, world: int) -> dict:
    c = CONFIGS[name]
    per_step = world if name == "c2" else c["step_slices"]
    return {"workload": workload_string(name, n_mov, iters), "baseline_config": c["index"], "iters": iters,
            "slices_per_step": per_step, "n_mov": n_mov, "shape": [c["h"], c["w"]],
            "l2_policy": "per-slice optimiser state (>= 407 MB fp32) is streamed every iteration and exceeds the "
                         "126 MB L2; no explicit flush"}


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


def live_image_params(eng) -> int:
    """Image-INR parameters one Adam launch reads and writes (FitEngine: MLP block + the table rows some pixel touches
    when the hashed levels are stored tap-indexed, everything otherwise)."""
    taps = getattr(eng, "_taps", None)
    if taps is None:
        return eng.n_image
    return min(eng.n_image, (eng._n_mlp_image + 2 * taps.n_active_rows + 3) // 4 * 4)


def algorithmic_bytes(slot: str, p: int, m: int, n_par, t_img: int, t_mot: int) -> float:
    """Algorithmic bytes per launch of each kernel of one iteration (sums to
    SURVEY 8(d)'s B_iter = 28 N_par + 16 (T_img + T_mot) + 64 P (M+1) plus the feature planes the
    kernels still round-trip between the hash-grid and MLP kernels)."""
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
        # their own slots bracket nothing (atomic mode) and are dropped from the report
        "fft_rows": 0.0, "fft_rows_adj": 0.0,
        "motion_rows_fwd": 16.0 * p + 8.0 * p + 16.0 * mp, "motion_rows_bwd": 24.0 * p + 24.0 * p + 24.0 * mp,
        "colpass_loss": 32.0 * p, "grad_entropy": 16.0 * p,
    }
    return table[slot]


# what actually paces each kernel (ncu --set full, profiles/)
LIMITERS = {
    "hashgrid_bwd_motion": "SM-side reduction issue, 1.29 cycles per lane and RED (ncu: l1tex 78 %, lts 71 %, DRAM 10 %): 52.4 M "
                           "8-byte reductions on hashed rows with no locality across pixels, half of the dim-0 pairs merged into "
                           "one RED.ADD.F32x4; DRAM traffic equals the algorithmic bytes",
    "hashgrid_fwd_motion": "L1 data stage, two 32-byte sectors per wavefront (ncu: l1tex data-pipe wavefronts 76 %, lts 38 %, "
                           "DRAM 10 %): 52.4 M 8-byte gathers on hashed rows; the grouped kernels + linear row layout halved "
                           "the lines and L2 requests per pixel corner (profiles/round2_grouped_kernels.txt)",
    "hashgrid_bwd_image": "L2 atomic throughput", "hashgrid_fwd_image": "L1 wavefronts / L2 gather rate",
    "adam_motion": "HBM (28 B / parameter; the zeroing of the gradients runs as a memset on a third stream)",
    "adam_image": "HBM",
    "mlp_bwd_motion": "barrier / MMA-chain latency + SIMT epilogue (ncu: issue active 36 %, tensor pipe 27 %)",
    "mlp_bwd_image": "MMA-chain waits + SIMT epilogue (ncu: issue active 34 %, tensor pipe 27 %)",
    "mlp_fwd_motion": "latency, 16 warps per SM (ncu: issue active 40 %)", "mlp_fwd_image": "latency (2.7 tiles per CTA)",
}

# SURVEY 8(d) / oracle.touched_entries(): distinct table rows touched (recomputed for other shapes at run time)
T_IMG = {(320, 320): 3041608}
T_MOT = {(320, 320, 2): 5917982, (320, 320, 4): 6513775, (320, 320, 8): 6756779}


def touched_rows(orc, torch, h, w, m):
    if (h, w) not in T_IMG:
        lv2 = orc.make_grid_levels(2, orc.ENCODING_CONFIG)
        T_IMG[(h, w)] = orc.touched_entries(orc.identity_grid(h, w).view(-1, 2), lv2)
    if (h, w, m) not in T_MOT:
        lv3 = orc.make_grid_levels(3, orc.ENCODING_CONFIG)
        T_MOT[(h, w, m)] = orc.touched_entries(orc.make_grids((m, h, w)), lv3)
    return T_IMG[(h, w)], T_MOT[(h, w, m)]


def init_dist(torch, dist):
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, local, dev


def roofline_block(nat, ms_sum, n_prof, h, w, m, n_par2, t_img, t_mot, ms_per_iter):
    p = h * w
    n_par = sum(n_par2)
    per_slot = {s: (ms_sum[k] / n_prof if n_prof else 0.0) for k, s in enumerate(nat.PROFILE_SLOTS)
                if algorithmic_bytes(s, p, m, n_par2, t_img, t_mot) > 0}
    iter_ms = sum(per_slot.values())
    dom = max(per_slot, key=per_slot.get) if n_prof else "adam_motion"
    peak, peak_src = measured_peak_gbs()
    dom_bytes = algorithmic_bytes(dom, p, m, n_par2, t_img, t_mot)
    achieved = dom_bytes / (per_slot[dom] * 1e-3) / 1e9 if per_slot.get(dom) else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(dom) if (h, w, m) == (320, 320, 4) else None
    except Exception:
        pass
    b_iter = 28.0 * n_par + 16.0 * (t_img + t_mot) + 64.0 * p * (m + 1)
    # the hash-grid kernels against the roofline that actually binds them: L2 random 8-byte accesses, peaks measured
    # on this pool by tools/l2_peaks.cu; accesses = points x 2^D corners x 16 levels per launch
    l2 = None
    try:
        with open(os.path.join(ROOT, "profiles", "round2_l2_peaks.json")) as f:
            pk = json.load(f)
        l2 = {"unit": "G table rows / s (one row = 8 bytes = one 32-byte L2 sector)", "source": "profiles/round2_l2_peaks.json"}
        # the grouped kernels read a pixel corner's rows of ALL groups with adjacent lanes (M = 2, 4, 8, 16): their
        # reference pattern is the micro-benchmark's 8-lane bundle (8 rows in two lines), not the lane pair
        fwd_m_key = ("gather_8B_bundle8_two_lines_Gps" if m in (2, 4, 8, 16) and "gather_8B_bundle8_two_lines_Gps" in pk
                     else "gather_8B_pair_same_line_Gps")
        for slot, taps, key in (("hashgrid_fwd_motion", m * p * 8 * 16, fwd_m_key),
                                ("hashgrid_bwd_motion", m * p * 8 * 16, "red_f32x2_pair_same_slot_Gps"),
                                ("hashgrid_fwd_image", p * 4 * 16, "gather_8B_pair_same_line_Gps"),
                                ("hashgrid_bwd_image", p * 4 * 16, "red_f32x2_pair_same_slot_Gps")):
            if per_slot.get(slot):
                ach = taps / (per_slot[slot] * 1e-3) / 1e9
                l2[slot] = {"achieved": round(ach, 1), "peak": pk[key], "frac": round(ach / pk[key], 3), "peak_pattern": key}
    except Exception:
        pass
    return {
        "bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
        "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": round(per_slot[dom], 5),
        "kernel_share_of_iteration": round(per_slot[dom] / iter_ms, 4) if iter_ms else None,
        "per_kernel_ms": {k: round(v, 5) for k, v in per_slot.items()},
        "per_kernel_gbs": {k: round(algorithmic_bytes(k, p, m, n_par2, t_img, t_mot) / (v * 1e-3) / 1e9, 1)
                           for k, v in per_slot.items() if v > 0},
        "instrumented_iterations": n_prof,
        "limiter": LIMITERS.get(dom, ""),
        "l2": l2,
        "iteration": {"algorithmic_bytes": b_iter, "achieved": round(b_iter / (ms_per_iter * 1e-3) / 1e9, 1),
                      "frac": round(b_iter / (ms_per_iter * 1e-3) / 1e9 / peak, 4)},
    }


# ---------------------------------------------------------------------------------------------------------
# C2: the headline line
# ---------------------------------------------------------------------------------------------------------
def run_c2(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import miccai24_immoco_b200 as mb
    from miccai24_immoco_b200 import _native as nat
    from oracle import immoco_oracle as orc   # synthetic-input generator + the baseline legs only

    c = CONFIGS["c2"]
    h, w, n_mov = c["h"], c["w"], c["n_mov"]
    rank, world, local, dev = init_dist(torch, dist)
    if rank == 0:
        mb.build()
    if world > 1:
        dist.barrier()
    lib = mb.lib()
    iters = args.iters
    lambdas = mb.lambda_schedule(iters, 1e-2)
    det = bool(args.deterministic)

    # ---- synthetic slices: seed 1000 + global slice index -------------------------
    models, engines, k_dev, k_host, masks_host = [], [], [], [], []
    for i in range(POOL):
        case = orc.make_case(h, w, n_mov, 1000 + rank * POOL + i)
        masks = case["masks"]
        model = mb.IMMoCo(masks.to(dev), image_seed=11 + i, motion_seed=101 + i)
        eng = mb.FitEngine(model, iters, deterministic=det)
        k = case["kspace_motion"]
        k_norm = (k / k.abs().max() * 16000).to(torch.complex64)
        models.append(model)
        engines.append(eng)
        k_dev.append(k_norm.to(dev))
        k_host.append(k.to(torch.complex64).pin_memory())
        masks_host.append(masks.pin_memory())
    init_img = models[0].image_inr.params.detach().clone()
    init_mot = models[0].motion_inr.params.detach().clone()
    gathered = [torch.empty((h, w, 2), device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
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
    trace = last.loss_trace(lambdas)
    final_loss = float(trace[-1])
    import numpy as np
    tail_loss = float(np.percentile(trace[-200:], 10)) if len(trace) >= 200 else float(np.min(trace))
    ms_sum = (C.c_float * len(nat.PROFILE_SLOTS))()
    n_prof = lib.immoco_profile_read(prof, ms_sum)
    lib.immoco_profile_destroy(prof)

    # ---- e2e: public API, host buffers in, host image out ---------------------------------------
    def step_e2e(i):
        j = i % POOL
        im, _ = mb.imcoco_motion_correction(k_host[j], masks_host[j], iters, 1e-2, 1e-2, False, deterministic=det)
        return im.cpu()

    # parameters Adam actually visits: with the image table stored tap-indexed the rows no pixel touches are skipped
    n_par2 = (models[0].motion_inr.n_params, live_image_params(engines[0]))
    engines.clear()        # free the resident engines' HBM before the API path allocates its own
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

    t_img, t_mot = touched_rows(orc, torch, h, w, n_mov)
    ms_per_iter = ms_total / (args.steps * iters)
    roofline = roofline_block(nat, ms_sum, n_prof, h, w, n_mov, n_par2, t_img, t_mot, ms_per_iter)

    # ---- baselines (rank 0, N=1 only): bounded samples on the box's host cores / the same GPU ----------
    cpu = cpu_c1 = torch_gpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(orc, torch, "c2", n_iters=args.cpu_iters)
        cpu_c1 = cpu_sample(orc, torch, "c1", n_iters=args.cpu_iters)
        torch_gpu = torch_gpu_sample(orc, torch, "c2", n_iters=args.torch_gpu_iters)

    value = world * args.steps / (ms_total * 1e-3)
    cfg = config_dict("c2", n_mov, iters, world)
    cfg["parallelism"] = (f"instance-sharded x{world}, one NCCL gather of corrected images per step"
                          if world > 1 else "single GPU")
    line = {
        "metric": METRIC, "value": round(value, 4), "unit": "slices/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3),
        "ms_per_iter": round(ms_per_iter, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "accumulation": "deterministic (row-sorted gather, fixed-point image cotangent)" if det else
                        "float atomics (default; the bit-reproducible mode is immoco_set_deterministic / --deterministic)",
        "final_loss": final_loss, "tail_loss_p10_last_200": tail_loss,
        "clocks": clocks,
        "e2e": {"value": round(world * n_e2e / e2e_s, 4), "unit": "slices/s", "steps": n_e2e,
                "h2d_bytes_per_step": int(k_host[0].numel() * 8 + masks_host[0].numel() * 8),
                "d2h_bytes_per_step": int(out_img.numel() * 8),
                "api": "imcoco_motion_correction(kspace_host, masks_host, iters, lr, lambda_ge, debug) -> image.cpu()"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "cpu_baseline_c1": cpu_c1,
        "torch_gpu_baseline": torch_gpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# C3 / C5: stacks of slices through reconstruct_slices (instance sharding + one NCCL gather per step)
# ---------------------------------------------------------------------------------------------------------
def run_stack(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import miccai24_immoco_b200 as mb
    from miccai24_immoco_b200 import _native as nat
    from oracle import immoco_oracle as orc

    name = args.config
    c = CONFIGS[name]
    h, w = c["h"], c["w"]
    n_mov = args.n_mov if args.n_mov else c["n_mov"]
    rank, world, local, dev = init_dist(torch, dist)
    if rank == 0:
        mb.build()
    if world > 1:
        dist.barrier()
    lib = mb.lib()
    iters = args.iters
    step_slices = c["step_slices"]
    det = bool(args.deterministic)

    # POOL distinct synthetic slices (seed 1000 + i), cycled over the stack; host (pinned) and device copies
    pool_k, pool_m = [], []
    for i in range(min(8, step_slices)):
        case = orc.make_case(h, w, n_mov, 1000 + i)
        pool_k.append(case["kspace_motion"].to(torch.complex64).pin_memory())
        pool_m.append(case["masks"].pin_memory())
    host_k = [pool_k[s % len(pool_k)] for s in range(step_slices)]
    host_m = [pool_m[s % len(pool_m)] for s in range(step_slices)]
    dev_k = [k.to(dev) for k in pool_k]
    dev_m = [m.to(dev) for m in pool_m]
    res_k = [dev_k[s % len(dev_k)] for s in range(step_slices)]
    res_m = [dev_m[s % len(dev_m)] for s in range(step_slices)]

    def step(ks, ms, n):
        return mb.reconstruct_slices(ks[:n], ms[:n], iters=iters, learning_rate=1e-2, lambda_ge=1e-2,
                                     deterministic=det)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):          # light warm-up: one slice per rank
        step(res_k, res_m, world)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = step(res_k, res_m, step_slices)
    ev1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total)
    if rank == 0:
        assert out.shape == (step_slices, h, w) and bool(torch.isfinite(torch.view_as_real(out)).all())

    # ---- e2e: the same call with HOST k-space / masks, gathered stack copied back to the host ------------
    n_e2e = max(1, min(args.steps, args.e2e_steps))
    sync_all()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        o = step(host_k, host_m, step_slices)
        if rank == 0:
            o_host = o.cpu()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s)

    # ---- per-kernel durations of this shape: one instrumented fit on rank 0 -------------------------------
    roofline = None
    if rank == 0:
        model = mb.IMMoCo(dev_m[0])
        n_it = min(iters, 200)
        eng = mb.FitEngine(model, n_it, deterministic=det)
        eng.set_kspace(dev_k[0] / dev_k[0].abs().max() * 16000)
        prof = lib.immoco_profile_create(32)
        lam = mb.lambda_schedule(max(n_it, 10), 1e-2)[:n_it]
        eng.run(lam, 1e-2, profile=prof, profile_every=10)
        torch.cuda.synchronize()
        ms_sum = (C.c_float * len(nat.PROFILE_SLOTS))()
        n_prof = lib.immoco_profile_read(prof, ms_sum)
        lib.immoco_profile_destroy(prof)
        n_par2 = (model.motion_inr.n_params, live_image_params(eng))
        t_img, t_mot = touched_rows(orc, torch, h, w, n_mov)
        per_rank = (step_slices + world - 1) // world
        ms_per_iter = ms_total / (args.steps * per_rank * iters)
        roofline = roofline_block(nat, ms_sum, n_prof, h, w, n_mov, n_par2, t_img, t_mot, ms_per_iter)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_sample(orc, torch, name, n_iters=max(2, args.cpu_iters // 5), n_mov=n_mov)
    cfg = config_dict(name, n_mov, iters, world)
    cfg["parallelism"] = (f"reconstruct_slices: slices dealt round-robin over {world} rank(s), no data-path collective, "
                          "one NCCL gather of the corrected images per step")
    cfg["warmup_steps"] = "one slice per rank each"
    line = {
        "metric": METRIC.replace("320x320", f"{h}x{w}"), "value": round(args.steps * step_slices / (ms_total * 1e-3), 4),
        "unit": "slices/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_total / args.steps, 3), "ms_per_iter": round(ms_per_iter, 5),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "clocks": clocks,
        "e2e": {"value": round(n_e2e * step_slices / e2e_s, 4), "unit": "slices/s", "steps": n_e2e,
                "h2d_bytes_per_step": int(sum(k.numel() * 8 for k in host_k) + sum(m.numel() * 8 for m in host_m)),
                "d2h_bytes_per_step": int(o_host.numel() * 8),
                "api": "reconstruct_slices(kspaces_host, masks_host, iters=...) -> stack.cpu() on rank 0"},
        "gpu_launches": int(args.steps * step_slices * iters * lib.immoco_launches_per_iteration_mode(n_mov, int(det), int(det))),
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# C4: kld-net mask inference + movement groups feeding the fits
# ---------------------------------------------------------------------------------------------------------
def run_c4(args):
    import torch
    import torch.distributed as dist

    import miccai24_immoco_b200 as mb
    from oracle import immoco_oracle as orc

    c = CONFIGS["c4"]
    h, w, n_mov = c["h"], c["w"], c["n_mov"]
    rank, world, local, dev = init_dist(torch, dist)
    if rank == 0:
        mb.build()
    if world > 1:
        dist.barrier()
    iters = args.iters
    batch = c["step_slices"]
    torch.manual_seed(0)
    net = mb.get_unet(2, 1, 32, 4, 0.0).to(dev)           # seeded random weights: the trained kLDNet.pth is not available
    pool = [orc.make_case(h, w, n_mov, 1000 + i) for i in range(8)]
    k_host = torch.stack([pool[s % 8]["kspace_motion"].to(torch.complex64) for s in range(batch)]).pin_memory()
    k_dev = k_host.to(dev)
    # random weights detect arbitrary lines; the fits below use the masks of the simulator (what a trained net
    # would return), so their cost is the cost of the reference pipeline on n_M = 4 slices
    sim_masks = [pool[s % 8]["masks"].pin_memory() for s in range(batch)]

    def pipeline(k, n, host):
        masks_net = mb.movement_masks_from_kspace(net, k[:n].to(dev, non_blocking=True))
        ks = [k[s] for s in range(n)]
        out = mb.reconstruct_slices(ks, sim_masks[:n], iters=iters, learning_rate=1e-2, lambda_ge=1e-2)
        return masks_net, out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        pipeline(k_dev, world, False)
    sync_all()
    # kld-net + group extraction alone (every rank runs the whole batch: the net is batch-parallel, 64 slices
    # take ~0.1 s against ~40 s of fitting)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    mb.movement_masks_from_kspace(net, k_dev)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        masks_net = mb.movement_masks_from_kspace(net, k_dev)
    e1.record()
    torch.cuda.synchronize()
    kld_ms = e0.elapsed_time(e1) / 3
    x_in = mb.kld_net_input(k_dev)
    net(x_in)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        net(x_in)
    e1.record()
    torch.cuda.synchronize()
    unet_ms = e0.elapsed_time(e1) / 3
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        _, out = pipeline(k_dev, batch, False)
    ev1.record()
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total)
    n_e2e = 1
    sync_all()
    t0 = time.perf_counter()
    _, o = pipeline(k_host, batch, True)
    if rank == 0:
        o_host = o.cpu()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    flops = 37.7e9 * batch            # U-Net 37.7 GFLOP per 320x320 slice
    cfg = config_dict("c4", n_mov, iters, world)
    cfg["parallelism"] = f"kld-net on the whole batch per rank; fits dealt round-robin over {world} rank(s) (reconstruct_slices)"
    cfg["warmup_steps"] = "one slice per rank each"
    line = {
        "metric": METRIC, "value": round(args.steps * batch / (ms_total * 1e-3), 4), "unit": "slices/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "clocks": clocks,
        "kld_net": {"ms_per_batch": round(kld_ms, 3), "slices_per_s": round(batch / (kld_ms * 1e-3), 1),
                    "what": "k-space -> network input -> U-Net -> sigmoid / column vote -> movement-group masks of the batch",
                    "unet_ms_per_batch": round(unet_ms, 3), "unet_tflops_fp32_equivalent": round(flops / (unet_ms * 1e-3) / 1e12, 2),
                    "unet_kernels": "3x3 convolutions: tcgen05 kind::tf32, 3xTF32 split (csrc/unet_tc.cu)",
                    "weights": "seeded random (kLDNet.pth unavailable)",
                    "groups_found_first_slices": [int(m.shape[0]) for m in masks_net[:8]]},
        "e2e": {"value": round(n_e2e * batch / e2e_s, 4), "unit": "slices/s", "steps": n_e2e,
                "h2d_bytes_per_step": int(k_host.numel() * 8 + sum(m.numel() * 8 for m in sim_masks)),
                "d2h_bytes_per_step": int(o_host.numel() * 8),
                "api": "movement_masks_from_kspace(net, kspace_host) + reconstruct_slices(...) -> stack.cpu()"},
        "gpu_launches": int(args.steps * batch * iters * 16),
        "roofline": None, "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# baselines
# ---------------------------------------------------------------------------------------------------------
def _loop_state(orc, torch, name, n_mov=None, device="cpu"):
    c = CONFIGS[name]
    m = n_mov if n_mov else c["n_mov"]
    case = orc.make_case(c["h"], c["w"], m, 1000)
    inr = orc.FourierNetworkWithInputEncoding if name == "c1" else orc.NetworkWithInputEncoding
    return orc.LoopState(case["kspace_motion"].to(device), case["masks"].to(device), 1000, inr_cls=inr), m


def cpu_sample(orc, torch, name: str, n_iters: int, warm: int = 1, n_mov=None):
    """The oracle's torch-CPU loop on the host cores for a bounded number of iterations of configuration
    `name` (c1: Fourier-feature INRs -- BASELINE.json configs[0] / north_star's CPU path; others: hash-grid)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st, m = _loop_state(orc, torch, name, n_mov)
    for j in range(warm):
        st.step(j)
    t0 = time.perf_counter()
    for j in range(warm, warm + n_iters):
        st.step(j)
    dt = (time.perf_counter() - t0) / n_iters
    enc = "Fourier-feature INRs (gamma(x) = [sin, cos](2 pi B x), 32 features, same MLP widths)" if name == "c1" else "hash-grid INRs"
    return {"value": round(1.0 / (1000 * dt), 6), "unit": "slices/s", "cores": torch.get_num_threads(),
            "kind": "port", "ms_per_iter": round(dt * 1e3, 1),
            "workload": workload_string(name, m, 1000),
            "sample": f"{n_iters} timed iterations (after {warm} warm-up) of one slice with the oracle's torch-CPU loop "
                      f"({enc}, torch.optim.Adam), extrapolated to 1000 iterations"}


def torch_gpu_sample(orc, torch, name: str, n_iters: int, warm: int = 3):
    """The oracle's torch loop (immoco.py:116-206 semantics, fp32, TF32 off, hash-grid taps cached) on the
    same GPU: the 'reference torch path on GPU' column of BASELINE.md section 4 (tiny-cuda-nn itself does
    not exist on this box)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    st, m = _loop_state(orc, torch, name, device="cuda")
    for j in range(warm):
        st.step(j)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for j in range(warm, warm + n_iters):
        st.step(j)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n_iters
    del st
    torch.cuda.empty_cache()
    return {"value": round(1.0 / (1000 * dt), 4), "unit": "slices/s", "ms_per_iter": round(dt * 1e3, 3), "kind": "port",
            "sample": f"{n_iters} timed iterations (after {warm} warm-up) of the same slice: oracle loop in torch fp32 on "
                      f"cuda:0 (ATen gather / index_add hash grid with cached taps, cuBLAS fp32 MLPs, cuFFT, "
                      f"torch.optim.Adam), extrapolated to 1000 iterations"}


def run_c1(args):
    import torch

    from oracle import immoco_oracle as orc
    if env_int("RANK", 0) != 0:
        return
    res = cpu_sample(orc, torch, "c1", n_iters=max(args.cpu_iters, 10))
    line = {"metric": METRIC, "value": res["value"], "unit": "slices/s", "n_gpus": 0, "steps": 1, "warmup": 1,
            "ms_per_step": res["ms_per_iter"] * 1000, "ms_per_iter": res["ms_per_iter"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict("c1", 2, 1000, 1), "cpu_baseline": res, "gpu_launches": 0,
            "note": "configs[0] is the reference's CPU-runnable case: a timing baseline only (no GPU, no parity target)"}
    print(json.dumps(line), flush=True)


def run_reference(args):
    """CPU reference arm: the reference is pure Python (torch) + tiny-cuda-nn; its own files are not on the
    GPU box and tiny-cuda-nn is absent everywhere, so this times the oracle port (the reference's loop
    semantics on torch CPU with the torch fp32 hash-grid stand-in; Fourier-feature INRs for --config c1)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    import torch

    from oracle import immoco_oracle as orc
    name = args.config
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    st, n_mov = _loop_state(orc, torch, name, args.n_mov)
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
    sample = (f"each step = {per_step} iterations of one slice of the configuration (oracle port of immoco.py:116-206 on "
              f"torch CPU, all {cores} host cores), extrapolated to {args.iters} iterations per slice")
    c = CONFIGS[name]
    line = {
        "impl": "reference", "metric": METRIC if name in ("c1", "c2", "c4", "c5") else METRIC.replace("320x320", f"{c['h']}x{c['w']}"),
        "value": round(value, 6), "unit": "slices/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 1), "ms_per_iter": round(ms_iter, 1),
        "higher_is_better": True, "scaling": "weak" if name in ("c1", "c2") else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(name, n_mov, args.iters, max(1, args.gpus)),
        "cpu_baseline": {"value": round(value, 6), "unit": "slices/s", "cores": torch.get_num_threads(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 6), "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_only": True,
    }
    if name == "c2":
        line["config"]["parallelism"] = (f"instance-sharded x{args.gpus}, one NCCL gather of corrected images per step"
                                         if args.gpus > 1 else "single GPU")
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--n-mov", type=int, default=0, help="movement groups (c5 sweep: 2, 4, 8); 0 = the configuration's own")
    ap.add_argument("--iters", type=int, default=1000, help="optimisation iterations per slice (metric: 1000)")
    ap.add_argument("--deterministic", action="store_true", help="time the bit-reproducible accumulation mode")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-iters", type=int, default=15)
    ap.add_argument("--torch-gpu-iters", type=int, default=30)
    ap.add_argument("--ref-iters", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c1":
        run_c1(args)
    elif args.config == "c2":
        run_c2(args)
    elif args.config == "c4":
        run_c4(args)
    else:
        run_stack(args)


if __name__ == "__main__":
    main()
