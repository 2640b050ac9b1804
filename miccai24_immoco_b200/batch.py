"""A stack of slices on one GPU ("within a GPU, batch B instances").

The reference reconstructs a stack slice by slice (src/test/test_immoco.py:45-72): every slice is an
independent optimisation.  ``reconstruct_batch`` runs that loop with nothing synchronising the host:
inputs are uploaded stream-ordered from pinned memory, the k-space scale stays on the device, every
slice's result is cloned on its stream, and the caller's stream waits on the slot streams at the end.
``in_flight`` slices can be kept going at once, each on its own CUDA stream (plus the library's
per-stream auxiliary stream), fed round-robin in chunks of iterations from ONE host thread.

Measured on B200 (profiles/round1_v4_slices_in_flight.txt): more than one slice in flight does NOT
raise throughput -- 854 us per slice-iteration with one, 916 / 925 with two / three -- because every
large kernel of an iteration already fills the GPU (hash-grid: all thread slots, L2-bound; MLP backward:
the whole register file; Adam: HBM through the same L2).  The default is therefore one slice at a time;
the knob stays for smaller shapes.

Per-slice results are what ``imcoco_motion_correction`` returns for the same inputs (same kernels,
same schedule; only the order of floating-point atomics differs, as between any two runs).
"""
from __future__ import annotations

from collections import deque
from typing import List, Optional, Sequence

import numpy as np
import torch

from .immoco import FitEngine, IMMoCo, lambda_schedule, run_batched

DEFAULT_IN_FLIGHT = 1
DEFAULT_CHUNK = 10
# slices of one shape fitted in lock step by immoco_fit_run_batched (the latency-bound kernels of an iteration
# are one launch for all of them); measured on B200 in profiles/round2_batched_fit.txt
DEFAULT_BATCH = 4


class _Slot:
    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.indices: List[int] = []     # slices being fitted in lock step, empty: idle
        self.done = 0                    # iterations issued
        self.models: list = []
        self.engines: list = []


def _as_pinned(t: torch.Tensor) -> torch.Tensor:
    if t.is_cuda or t.is_pinned():
        return t
    return t.pin_memory()


def reconstruct_batch(kspaces: Sequence[torch.Tensor], masks: Sequence[torch.Tensor], iters: int = 200,
                      learning_rate: float = 1e-2, lambda_ge: float = 1e-2, debug: bool = False, *,
                      in_flight: int = DEFAULT_IN_FLIGHT, chunk: int = DEFAULT_CHUNK,
                      image_params: Optional[Sequence[torch.Tensor]] = None,
                      motion_params: Optional[Sequence[torch.Tensor]] = None,
                      seeds: Optional[Sequence[int]] = None, kmax: float = 16000.0, variant: str = "main",
                      return_kspace: bool = False, return_traces: bool = False, device=None,
                      deterministic: Optional[bool] = None, batch: int = DEFAULT_BATCH):
    """``imcoco_motion_correction`` over a stack of slices, ``in_flight`` of them concurrently.

    kspaces[i]: (H, W) complex k-space, masks[i]: (M_i, H, W) movement-group masks (host or device;
    host tensors are uploaded stream-ordered).  Slices may differ in shape and in M.  Returns the list
    of corrected images (complex64 CUDA tensors, 16000-normalised scale); with
    ``return_kspace`` / ``return_traces`` a tuple (images, kspaces_fwd, traces) with None for the parts
    not asked for.  Loss traces are read back once at the end.  ``batch``: up to this many pending slices of
    the SAME shape and movement-group count are fitted in lock step through ``immoco_fit_run_batched`` (each
    slice still gets exactly the result of its own fit).
    """
    if not torch.cuda.is_available():
        raise RuntimeError("reconstruct_batch needs a CUDA device (no CPU fallback)")
    n = len(kspaces)
    if len(masks) != n:
        raise ValueError("kspaces and masks must have the same length")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lambdas = lambda_schedule(iters, lambda_ge, variant)
    caller = torch.cuda.current_stream(dev)
    slots = [_Slot(dev) for _ in range(max(1, min(int(in_flight), n)))]
    for s in slots:
        s.stream.wait_stream(caller)        # inputs produced on the caller's stream are visible
    pending = deque(range(n))
    images: List[Optional[torch.Tensor]] = [None] * n
    ksp_out: List[Optional[torch.Tensor]] = [None] * n
    losses: List[Optional[torch.Tensor]] = [None] * n
    chunk = max(1, int(chunk)) if len(slots) > 1 else max(1, iters)     # one slot: nothing to interleave with

    def shape_key(i: int):
        return (tuple(kspaces[i].shape[-2:]), int(masks[i].shape[0]))

    def start_one(i: int):
        m_in = masks[i]
        k_in = kspaces[i]
        host_masks = None if m_in.is_cuda else m_in
        m_dev = _as_pinned(m_in).to(dev, non_blocking=True)
        seed = 1337 + 2 * i if seeds is None else int(seeds[i])
        model = IMMoCo(m_dev, image_seed=seed, motion_seed=seed + 1, host_masks=host_masks)
        with torch.no_grad():
            if image_params is not None:
                model.image_inr.params.copy_(image_params[i].to(dev, non_blocking=True))
            if motion_params is not None:
                model.motion_inr.params.copy_(motion_params[i].to(dev, non_blocking=True))
        k_dev = _as_pinned(k_in).to(dev, non_blocking=True).to(torch.complex64)
        scale = k_dev.abs().max()                      # stays on the device (immoco.py:137-139)
        engine = FitEngine(model, max(iters, 1), deterministic=deterministic)
        engine.set_kspace(k_dev.div(scale).mul(kmax))
        return model, engine

    def start(slot: _Slot) -> None:
        """Next pending slice plus up to batch - 1 further pending slices of the same shape and group count."""
        first = pending.popleft()
        group = [first]
        key = shape_key(first)
        for i in list(pending):
            if len(group) >= max(1, int(batch)):
                break
            if shape_key(i) == key:
                pending.remove(i)
                group.append(i)
        built = [start_one(i) for i in group]
        slot.indices, slot.done = group, 0
        slot.models = [b[0] for b in built]
        slot.engines = [b[1] for b in built]

    def finish(slot: _Slot) -> None:
        for i, eng in zip(slot.indices, slot.engines):
            img = torch.view_as_complex(eng.image.clone())
            img.record_stream(caller)
            images[i] = img
            if return_kspace:
                k = torch.view_as_complex(eng.k_out.clone())
                k.record_stream(caller)
                ksp_out[i] = k
            if return_traces or debug:
                acc = eng.loss[:iters].clone()
                acc.record_stream(caller)
                losses[i] = acc
        slot.indices, slot.models, slot.engines = [], [], []

    active = True
    while active:
        active = False
        for slot in slots:
            with torch.cuda.stream(slot.stream):
                if not slot.indices:
                    if not pending:
                        continue
                    start(slot)
                end = min(iters, slot.done + chunk)
                run_batched(slot.engines, lambdas, learning_rate, slot.done, end)
                slot.done = end
                if slot.done >= iters:
                    finish(slot)
                active = True
    for s in slots:
        caller.wait_stream(s.stream)

    traces = None
    if return_traces or debug:
        traces = []
        lam = np.asarray(lambdas, dtype=np.float64).astype(np.float32)
        for i in range(n):
            acc = losses[i].cpu().numpy()
            h, w = images[i].shape
            traces.append((acc[:, 0] / (2.0 * h * w)).astype(np.float32) + lam * acc[:, 1].astype(np.float32))
            if debug:
                print(f"slice {i}: DC_Loss first {traces[-1][0]:.4f} last {traces[-1][-1]:.4f}")
    if return_kspace or return_traces:
        return images, (ksp_out if return_kspace else None), traces
    return images
