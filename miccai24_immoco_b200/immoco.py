"""IM-MoCo model and per-instance optimisation loop on the B200 CUDA path.

Host-side mirror of the reference's src/models/immoco.py: same names, call signatures and
argument meaning (``IMMoCo(masks)``, ``imcoco_motion_correction(kspace_corr, masks, iters,
learning_rate, lambda_ge, debug)``, ``make_grids``, the three config dicts, ``ClearCache``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as nat
from .encoding import twiddles
from .ops import (FFT, IFFT, GradientEntropyLoss, NetworkWithInputEncoding, _need_cuda, _stream,
                  twiddle_table)

# src/models/immoco.py:11-37 (same keys / values; passed straight to the INR constructor)
network_config = {
    "otype": "CutLassMLP",
    "activation": "ReLU",
    "output_activation": "None",
    "n_neurons": 256,
    "n_hidden_layers": 1,
}

mot_network_config = {
    "otype": "FullyFusedMLP",
    "activation": "Tanh",
    "output_activation": "None",
    "n_neurons": 64,
    "n_hidden_layers": 1,
}

encoding_config = {
    "otype": "Grid",
    "type": "Hash",
    "n_levels": 16,
    "n_features_per_level": 2,
    "log2_hashmap_size": 19,
    "base_resolution": 16,
    "fine_resolution": 320,
    "per_level_scale": 2,
    "interpolation": "Linear",
}


class ClearCache:
    """Context manager of immoco.py:40-45."""

    def __enter__(self):
        torch.cuda.empty_cache()

    def __exit__(self, exc_type, exc_val, exc_tb):
        torch.cuda.empty_cache()


def make_grids(sizes, device="cpu"):
    """(prod(sizes), len(sizes)) coordinates in [-1, 1], "ij" order (immoco.py:48-53).

    The linspace values are produced on the CPU and copied, so they are bit-identical on every
    device (the kernels read coordinates from memory, they do not regenerate them)."""
    axes = [torch.linspace(-1, 1, int(s)) for s in sizes]
    grid = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1).view(-1, len(sizes))
    return grid.to(device)


def _identity_grid(h: int, w: int, device) -> torch.Tensor:
    theta = torch.eye(2, 3).unsqueeze(0)
    return F.affine_grid(theta, torch.Size((1, 1, h, w)), align_corners=True).to(device)


# The coordinate buffers depend on the shape only (immoco.py:72-80 rebuilds them per instance); one
# device copy per (shape, device) is shared by every instance: they are read-only on the data path.
_COORD_CACHE: dict = {}


def _cached_coords(kind: str, shape, device) -> torch.Tensor:
    key = (kind, tuple(int(v) for v in shape), str(device))
    t = _COORD_CACHE.get(key)
    if t is None:
        t = _identity_grid(shape[0], shape[1], device) if kind == "identity" else make_grids(shape, device)
        _COORD_CACHE[key] = t
    return t


def _upload(array: np.ndarray, device) -> torch.Tensor:
    """Small host array -> device through pinned memory, stream-ordered, no host synchronisation."""
    t = torch.from_numpy(np.ascontiguousarray(array))
    if torch.device(device).type != "cuda":
        return t
    return t.pin_memory().to(device, non_blocking=True)


class LineStructure:
    """Column structure of (M, H, W) movement-group masks, uploaded once per instance.

    The reference builds masks with extract_movement_groups (motion_utils.py:56-109): every mask
    is constant along rows.  Masks that vary along rows are rejected (documented limitation).
    ``masks`` may live on the host (no device synchronisation at all) or on the device (one D2H)."""

    def __init__(self, masks: torch.Tensor, device=None):
        if masks.dim() != 3:
            raise ValueError("masks must have shape (num_movements, H, num_lines)")
        m, h, w = masks.shape
        dev = masks.device if device is None else torch.device(device)
        mf = masks.detach().to("cpu").to(torch.float32).numpy()
        if m > 0 and h > 1 and not bool((mf == mf[:, :1, :]).all()):
            raise NotImplementedError(
                "movement-group masks must be constant along dim -2 (column indicators, as produced "
                "by extract_movement_groups); row-varying masks are not supported by the CUDA path")
        cols = mf[:, 0, :] if m > 0 else np.zeros((0, w), np.float32)   # (M, W)
        ofs, idx, wt = [0], [], []
        for g in range(m):
            nz = np.nonzero(cols[g])[0]
            idx.extend(int(v) for v in nz)
            wt.extend(float(cols[g, v]) for v in nz)
            ofs.append(len(idx))
        static = (1.0 - cols.sum(0)).astype(np.float32) if m > 0 else np.ones(w, np.float32)
        self.m, self.h, self.w = m, h, w
        self.group_ofs = _upload(np.asarray(ofs, np.int32), dev)
        self.line_idx = _upload(np.asarray(idx if idx else [0], np.int32), dev)
        self.line_w = _upload(np.asarray(wt if wt else [0.0], np.float32), dev)
        self.static_w = _upload(static, dev)
        self.max_lines = max([ofs[i + 1] - ofs[i] for i in range(m)] + [1])
        self.n_lines = len(idx)

    def struct(self) -> nat.Lines:
        s = nat.Lines()
        s.n_groups = self.m
        s.group_ofs = self.group_ofs.data_ptr()
        s.line_idx = self.line_idx.data_ptr()
        s.line_w = self.line_w.data_ptr()
        s.static_w = self.static_w.data_ptr()
        s.max_lines = self.max_lines
        return s


class _ForwardModelFunction(torch.autograd.Function):
    """k = F(I) * (1 - sum_m S_m) + sum_m F(warp_m(I)) * S_m   (immoco.py:91-111), one fused op."""

    @staticmethod
    def forward(ctx, image_ri, disp, model):
        lib = nat.lib()
        h, w = model.x, model.num_lines
        image_ri = image_ri.detach().contiguous()
        disp = disp.detach().contiguous()
        dev = image_ri.device
        c_tmp = torch.empty((h, w, 2), dtype=torch.float32, device=dev)
        k = torch.empty((h, w, 2), dtype=torch.float32, device=dev)
        lines = model._lines.struct()
        with torch.cuda.device(dev):
            nat.check(lib.immoco_forward_model(image_ri.data_ptr(), disp.data_ptr(), model._ident.data_ptr(),
                                               C.byref(lines), twiddle_table(h, dev).data_ptr(),
                                               twiddle_table(w, dev).data_ptr(), c_tmp.data_ptr(),
                                               k.data_ptr(), h, w, _stream(dev)), "forward_model")
        ctx.save_for_backward(image_ri, disp)
        ctx.model = model
        return k

    @staticmethod
    def backward(ctx, d_k):
        image_ri, disp = ctx.saved_tensors
        model = ctx.model
        lib = nat.lib()
        h, w = model.x, model.num_lines
        dev = image_ri.device
        d_k = d_k.contiguous().float()
        c_tmp = torch.empty((h, w, 2), dtype=torch.float32, device=dev)
        d_image = torch.zeros_like(image_ri)
        d_disp = torch.zeros_like(disp)
        lines = model._lines.struct()
        with torch.cuda.device(dev):
            nat.check(lib.immoco_forward_model_bwd(d_k.data_ptr(), image_ri.data_ptr(), disp.data_ptr(),
                                                   model._ident.data_ptr(), C.byref(lines),
                                                   twiddle_table(h, dev).data_ptr(),
                                                   twiddle_table(w, dev).data_ptr(), c_tmp.data_ptr(),
                                                   d_image.data_ptr(), d_disp.data_ptr(), 0, h, w, _stream(dev)),
                      "forward_model_bwd")
        return d_image, d_disp, None


class IMMoCo(nn.Module):
    """Image INR + Motion INR + motion forward model (immoco.py:56-113).

    Attributes kept from the reference: image_inr, motion_inr, masks, num_movements, x, num_lines,
    device, identy_grid, input_grid.  ``forward() -> (kspace_out, image_prior)``."""

    def __init__(self, masks, image_seed: int = 1337, motion_seed: int = 1338, host_masks=None):
        super().__init__()
        _need_cuda(masks, "IMMoCo(masks)")
        dev = masks.device
        with torch.cuda.device(dev):
            self.image_inr = NetworkWithInputEncoding(2, 2, encoding_config, network_config,
                                                      seed=image_seed, device=dev)
            self.motion_inr = NetworkWithInputEncoding(3, 2, encoding_config, mot_network_config,
                                                       seed=motion_seed, device=dev)
        self.masks = masks
        self.num_movements, self.x, self.num_lines = masks.shape
        self.device = dev
        self.identy_grid = _cached_coords("identity", (self.x, self.num_lines), dev)
        self.input_grid = _cached_coords("motion", (self.num_movements, self.x, self.num_lines), dev)
        # host_masks: the same masks still on the host (batch driver) -> no device synchronisation
        self._lines = LineStructure(masks if host_masks is None else host_masks, device=dev)
        self._ident = self.identy_grid.view(-1, 2).contiguous()

    def forward(self):
        h, w, m = self.x, self.num_lines, self.num_movements
        out = self.image_inr(self._ident).float().view(h, w, 2)
        image_prior = torch.view_as_complex(out.contiguous())
        if m > 0:
            disp = self.motion_inr(self.input_grid).float().tanh().view(m, h, w, 2)
        else:   # undefined in the reference: static branch only
            disp = torch.zeros((0, h, w, 2), dtype=torch.float32, device=self.device)
        k = _ForwardModelFunction.apply(out, disp, self)
        return torch.view_as_complex(k), image_prior


def lambda_schedule(iters: int, lambda_ge: float, variant: str = "main") -> List[float]:
    """lambda used by each iteration.  main: immoco.py:180-181 (halved on every iteration that is
    NOT a multiple of iters//10 after the midpoint; iters < 10 raises
    ZeroDivisionError like the reference).  downstream: test_immoco_downstream.py:189-190."""
    lams, lam = [], float(lambda_ge)
    for j in range(iters):
        lams.append(lam)
        if variant == "main":
            if j % (iters // 10) and j > (iters // 2):
                lam *= 0.5
        elif variant == "downstream":
            if j % 10 == 0 and j > 80:
                lam *= 0.5
        else:
            raise ValueError(f"unknown schedule variant {variant!r}")
    return lams


_ROW_PERM_CACHE: dict = {}


def _row_permutation(grid, swizzle, dev) -> torch.Tensor:
    """Device copy of ``GridSpec.row_permutation`` (one per device and layout, shared by all engines)."""
    key = (str(dev), grid.n_dims, grid.offsets, swizzle)
    if key not in _ROW_PERM_CACHE:
        _ROW_PERM_CACHE[key] = torch.from_numpy(grid.row_permutation(swizzle)).to(dev)
    return _ROW_PERM_CACHE[key]


_LAYOUT_CACHE: dict = {}


def _linear_layout(grid, m: int, dev):
    """Chunk tables of ``GridSpec.linear_layout`` for m movement groups on the device, the layout words and the row
    permutation they imply (one per device, grid and group count; shared by all engines)."""
    key = (str(dev), grid.n_dims, grid.offsets, grid.hashed, int(m))
    if key not in _LAYOUT_CACHE:
        lut = grid.linear_layout(torch.linspace(-1, 1, m).numpy())     # dim-0 values of make_grids: no device sync
        words = tuple(nat.LAYOUT_LUT if lut[lvl].any() else 0 for lvl in range(grid.n_levels))
        perm = torch.from_numpy(grid.row_permutation_lut(lut)).to(dev)
        _LAYOUT_CACHE[key] = (_upload(lut.view(np.int32), dev), words, perm)
    return _LAYOUT_CACHE[key]


_CSR_CACHE: dict = {}


class GridCsr:
    """Row-sorted tap list of one hash grid over one constant coordinate set (include/immoco_b200.h, section
    1b): built once by ``immoco_hashgrid_csr_build`` and shared by every fit of the same shape -- the gather
    form of the hash-grid backward pass that makes a fit bit-reproducible."""

    def __init__(self, grid, desc, coords: torch.Tensor):
        lib = nat.lib()
        dev = coords.device
        n = int(coords.shape[0])
        self.n_points = n
        self.n_taps = n * (1 << grid.n_dims) * grid.n_levels
        self.row_ptr = torch.empty(grid.n_rows + 1, dtype=torch.int32, device=dev)
        self.taps = torch.empty((max(self.n_taps, 1), 2), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            need = int(lib.immoco_hashgrid_csr_workspace_bytes(C.byref(desc), n))
            if need < 0:
                raise nat.NativeError("hashgrid_csr_workspace_bytes: rejected arguments")
            work = torch.empty(need, dtype=torch.uint8, device=dev)
            nat.check(lib.immoco_hashgrid_csr_build(C.byref(desc), coords.data_ptr(), n, self.row_ptr.data_ptr(),
                                                    self.taps.data_ptr(), work.data_ptr(), need,
                                                    torch.cuda.current_stream(dev).cuda_stream), "hashgrid_csr_build")
            work.record_stream(torch.cuda.current_stream(dev))
            # engines on other streams (reconstruct_batch slots) wait for the build before their first gather
            self.ready = torch.cuda.Event()
            self.ready.record(torch.cuda.current_stream(dev))

    def struct(self) -> nat.GridCsr:
        c = nat.GridCsr()
        c.row_ptr = self.row_ptr.data_ptr()
        c.taps = self.taps.data_ptr()
        c.n_taps = self.n_taps
        c.n_points = self.n_points
        return c


def _cached_csr(kind: str, shape, grid, swizzle, coords: torch.Tensor) -> GridCsr:
    key = (kind, tuple(int(v) for v in shape), str(coords.device), grid.n_dims, grid.offsets, grid.hashed,
           tuple(swizzle))
    csr = _CSR_CACHE.get(key)
    if csr is None:
        csr = GridCsr(grid, grid.desc(swizzle), coords)
        _CSR_CACHE[key] = csr
    torch.cuda.current_stream(coords.device).wait_event(csr.ready)
    return csr


_TAPS_CACHE: dict = {}


class GridTaps:
    """Tap-indexed storage of a 2-D grid's hashed levels over one constant coordinate set (include/immoco_b200.h,
    section 1c), built once per shape and shared by every fit of that shape.

    ``immoco_hashgrid_tap_rows`` lists the row of every (level, point, corner); the rows of the hashed levels are
    then ranked by their FIRST touch in that order.  ``perm[r]`` = physical row of reference-layout row r (the
    identity below ``first_level``), ``rows`` = the physical rows of every tap, ``n_active_rows`` = rows that any
    point touches (they come first; the others keep g = m = v = 0 for ever and are skipped by Adam)."""

    def __init__(self, grid, desc, coords: torch.Tensor):
        lib = nat.lib()
        dev = coords.device
        n = int(coords.shape[0])
        first = grid.n_levels
        while first > 0 and grid.hashed[first - 1] and (grid.entries[first - 1] & (grid.entries[first - 1] - 1)) == 0:
            first -= 1
        if grid.n_dims != 2 or first == grid.n_levels:
            raise ValueError("tap-indexed storage needs a 2-D grid whose last levels are hashed")
        self.first_level, self.n_points = first, n
        n_lv = grid.n_levels - first
        base = grid.offsets[first]
        n_hashed_rows = grid.n_rows - base
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            raw = torch.empty((n_lv, n, 4), dtype=torch.int32, device=dev)
            nat.check(lib.immoco_hashgrid_tap_rows(C.byref(desc), coords.data_ptr(), n, first, grid.n_levels,
                                                   raw.data_ptr(), stream.cuda_stream), "hashgrid_tap_rows")
            lvl_ofs = torch.tensor([grid.offsets[l] - base for l in range(first, grid.n_levels)], dtype=torch.int64,
                                   device=dev)
            key = (raw.to(torch.int64) + lvl_ofs[:, None, None]).reshape(-1)         # (level, point, corner) order
            n_taps = key.numel()
            first_touch = torch.full((n_hashed_rows,), n_taps, dtype=torch.int64, device=dev)
            first_touch.scatter_reduce_(0, key, torch.arange(n_taps, dtype=torch.int64, device=dev), "amin")
            order = torch.argsort(first_touch, stable=True)      # touched rows by first touch, then the others
            phys = torch.empty_like(order)
            phys[order] = torch.arange(n_hashed_rows, dtype=torch.int64, device=dev)
            self.rows = (phys[key] + base).to(torch.int32).view(n_lv, n, 4).contiguous()
            self.perm = torch.arange(grid.n_rows, dtype=torch.int64, device=dev)
            self.perm[base:] = phys + base
            # one host read per shape (cached): how many rows are live
            self.n_active_rows = base + int((first_touch < n_taps).sum())
            self.ready = torch.cuda.Event()
            self.ready.record(stream)

    def struct(self) -> nat.GridTaps:
        t = nat.GridTaps()
        t.rows = self.rows.data_ptr()
        t.first_level = self.first_level
        t.n_points = self.n_points
        t.n_active_rows = self.n_active_rows
        return t


def _cached_taps(kind: str, shape, grid, coords: torch.Tensor) -> GridTaps:
    key = (kind, tuple(int(v) for v in shape), str(coords.device), grid.n_dims, grid.offsets, grid.hashed)
    taps = _TAPS_CACHE.get(key)
    if taps is None:
        taps = GridTaps(grid, grid.desc(), coords)
        _TAPS_CACHE[key] = taps
    torch.cuda.current_stream(coords.device).wait_event(taps.ready)
    return taps


def _taps_supported(grid) -> bool:
    last = grid.n_levels - 1
    return grid.n_dims == 2 and bool(grid.hashed[last]) and (grid.entries[last] & (grid.entries[last] - 1)) == 0


def _taps_pay(grid, n_points: int) -> bool:
    """Whether tap-indexed storage is expected to pay for ``n_points`` pixels: 4 n taps thrown into a level's
    ``entries`` rows leave exp(-4 n / entries) of them untouched.  Measured on B200 (tools/taps_ab.py,
    profiles/round2_tap_indexed_image_table.txt): 320 x 320 (47 % untouched) 621 -> 584 us per iteration;
    640 x 368 (18 % untouched) 1390 -> 1410 us -- the indexed scatter issues more reductions than the lane-pair
    kernel there (fewer aligned row pairs) and the shorter Adam pass no longer makes up for it."""
    return _taps_supported(grid) and 4.0 * n_points <= 1.2 * grid.entries[grid.n_levels - 1]


def clear_caches() -> None:
    """Drops the per-shape device caches (coordinates, row permutations, tap lists)."""
    _CSR_CACHE.clear()
    _TAPS_CACHE.clear()
    _LAYOUT_CACHE.clear()
    _ROW_PERM_CACHE.clear()
    _COORD_CACHE.clear()


class FitEngine:
    """Device state + native loop for ONE slice: both INRs' parameters, gradients and Adam moments
    live in one flat fp32 vector [motion | image]; every iteration is 16 kernel launches issued by
    ``immoco_fit_run`` with no host synchronisation.

    The motion grid's hashed levels are stored in a permuted ROW LAYOUT (``GridSpec.row_swizzle``: both
    dim-0 corners of every lane pair in one 128-byte line; Adam is element-wise, so it does not care).
    Parameters are permuted on the way in (constructor, ``reset``) and back in ``write_back``; everything
    outside the engine sees the reference's layout.  ``row_swizzle=False`` keeps the reference layout.

    The image grid's hashed levels are stored TAP-INDEXED on the float-atomic path (``GridTaps``: rows ranked by
    first touch, the never-touched 47 % of them at 320 x 320 behind the live ones where Adam and the gradient
    memset do not go) whenever that is expected to pay (``_taps_pay``: at most ~70 % of the rows touched);
    ``compact_image=False`` / ``True`` forces the reference layout / the indexed one."""

    def __init__(self, model: IMMoCo, max_iters: int, row_swizzle: bool = True,
                 deterministic: Optional[bool] = None, fuse_adam: Optional[bool] = None,
                 compact_image: Optional[bool] = None, grouped_layout: Optional[bool] = None):
        """``deterministic`` (default: the library-wide ``immoco_get_deterministic()``): bit-reproducible fit --
        hash-grid backward as a row-sorted gather over a tap list built once per shape, MLP weight gradients
        as per-CTA blocks added in CTA order, image cotangent in 64-bit fixed point.  ``fuse_adam`` (default:
        on in deterministic mode): the table rows are updated inside the gather kernel."""
        self.model = model
        dev = model.device
        lib = nat.lib()
        self.deterministic = bool(lib.immoco_get_deterministic()) if deterministic is None else bool(deterministic)
        self.fuse_adam = self.deterministic and (True if fuse_adam is None else bool(fuse_adam))
        h, w, m = model.x, model.num_lines, model.num_movements
        p, mp = h * w, h * w * m
        img, mot = model.image_inr, model.motion_inr
        self.n_motion, self.n_image = mot.n_params, img.n_params
        n = self.n_motion + self.n_image
        self.params = torch.empty(n, dtype=torch.float32, device=dev)
        # physical row layout of the motion table (first coordinate = one value per movement group)
        self._swizzle: Tuple[int, ...] = ()
        self._perm: Optional[torch.Tensor] = None
        self._n_mlp_motion = mot.mlp.n_params
        # 2 .. 16 groups on the float-atomic path: general linear layout (chunk tables) + the grouped kernels,
        # which put the rows of all groups of a pixel corner into one or two 128-byte lines
        self._lut: Optional[torch.Tensor] = None
        # (default: power-of-two group counts only -- with 3 or 5 groups a bundle's surplus lanes idle and the kernels
        # are slower than the lane-pair ones: 640 x 368, n_M = 5: 1395 -> 1554 us per iteration, gpurun_out/r308)
        want_grouped = (2 <= m <= 16 and not self.deterministic and not lib.immoco_get_fused_scatter()
                        and (m in (2, 4, 8, 16) if grouped_layout is None else bool(grouped_layout)))
        if row_swizzle and want_grouped:
            self._lut, self._swizzle, self._perm = _linear_layout(mot.grid, m, dev)
        elif row_swizzle and m > 0:
            u = torch.linspace(-1, 1, m).numpy()        # dim-0 values of make_grids, known without a device sync
            swz = mot.grid.row_swizzle(u) if u.size <= 64 else ()
            if any(swz):
                self._swizzle = swz
                self._perm = _row_permutation(mot.grid, swz, dev)
        self._n_mlp_image = img.mlp.n_params
        self._taps: Optional[GridTaps] = None
        want_taps = _taps_pay(img.grid, p) if compact_image is None else (bool(compact_image) and _taps_supported(img.grid))
        if want_taps and not self.deterministic:
            with torch.cuda.device(dev):
                self._taps = _cached_taps("identity", (h, w), img.grid, model._ident)
        self._load_motion(mot.params.detach())
        self._load_image(img.params.detach())
        self.state = torch.zeros((3, n), dtype=torch.float32, device=dev)   # grads, exp_avg, exp_avg_sq
        f32 = dict(dtype=torch.float32, device=dev)
        self.enc_image = torch.empty((16, p, 2), **f32)
        self.d_enc_image = torch.empty((16, p, 2), **f32)
        self.enc_motion = torch.empty((16, max(mp, 1), 2), **f32)
        self.d_enc_motion = torch.empty((16, max(mp, 1), 2), **f32)
        self.image = torch.zeros((h, w, 2), **f32)
        self.d_image = torch.zeros((h, w, 2), **f32)
        self.disp = torch.zeros((max(m, 1), h, w, 2), **f32)
        self.d_disp = torch.zeros((max(m, 1), h, w, 2), **f32)
        self.c_tmp = torch.empty((h, w, 2), **f32)
        self.d_c = torch.empty((h, w, 2), **f32)
        self.k_out = torch.zeros((h, w, 2), **f32)
        self.k_in = torch.zeros((h, w, 2), **f32)
        self.max_iters = max_iters
        self.loss = torch.zeros((max_iters, 2), dtype=torch.float64, device=dev)
        self.tw_h = twiddle_table(h, dev)
        self.tw_w = twiddle_table(w, dev)
        self.coords_motion = model.input_grid.contiguous() if m > 0 else torch.zeros((1, 3), **f32)
        f = nat.Fit()
        f.h, f.w, f.m = h, w, m
        f.grid_image = img.grid.desc()
        f.grid_motion = mot.grid.desc(self._swizzle, 0 if self._lut is None else self._lut.data_ptr())
        f.width_image, f.act_image = img.mlp.width, img.mlp.act
        f.width_motion, f.act_motion = mot.mlp.width, mot.mlp.act
        f.n_motion, f.n_image = self.n_motion, self.n_image
        f.params = self.params.data_ptr()
        f.grads = self.state[0].data_ptr()
        f.exp_avg = self.state[1].data_ptr()
        f.exp_avg_sq = self.state[2].data_ptr()
        f.coords_image = model._ident.data_ptr()
        f.coords_motion = self.coords_motion.data_ptr()
        f.lines = model._lines.struct()
        f.tw_h, f.tw_w = self.tw_h.data_ptr(), self.tw_w.data_ptr()
        f.k_in = self.k_in.data_ptr()
        f.enc_image, f.d_enc_image = self.enc_image.data_ptr(), self.d_enc_image.data_ptr()
        f.enc_motion, f.d_enc_motion = self.enc_motion.data_ptr(), self.d_enc_motion.data_ptr()
        f.image, f.d_image = self.image.data_ptr(), self.d_image.data_ptr()
        f.disp, f.d_disp = self.disp.data_ptr(), self.d_disp.data_ptr()
        f.c_tmp, f.d_c, f.k_out = self.c_tmp.data_ptr(), self.d_c.data_ptr(), self.k_out.data_ptr()
        f.loss = self.loss.data_ptr()
        f.lr, f.beta1, f.beta2, f.eps = 1e-2, 0.9, 0.999, 1e-8
        # per-CTA loss slots (both modes: no floating-point atomics on the loss trace)
        slots = (C.c_int32 * 2)()
        nat.check(lib.immoco_fit_loss_slots(h, w, slots), "fit_loss_slots")
        self.loss_slots = torch.zeros((max_iters, slots[0] + slots[1]), dtype=torch.float64, device=dev)
        f.loss_slots = self.loss_slots.data_ptr()
        f.deterministic, f.fuse_adam = int(self.deterministic), int(self.fuse_adam)
        if self.deterministic:
            with torch.cuda.device(dev):
                self.csr_image = _cached_csr("identity", (h, w), img.grid, (), model._ident)
                f.csr_image = self.csr_image.struct()
                n_part_i = lib.immoco_mlp_bwd_partial_count(p)
                self.mlp_part_image = torch.zeros((n_part_i, img.mlp.n_params), **f32)
                f.mlp_part_image = self.mlp_part_image.data_ptr()
                if m > 0:
                    self.csr_motion = _cached_csr("motion", (m, h, w), mot.grid, self._swizzle, self.coords_motion)
                    f.csr_motion = self.csr_motion.struct()
                    n_part_m = lib.immoco_mlp_bwd_partial_count(mp)
                    self.mlp_part_motion = torch.zeros((n_part_m, mot.mlp.n_params), **f32)
                    f.mlp_part_motion = self.mlp_part_motion.data_ptr()
            self.d_image_fx = torch.zeros((h, w, 2), dtype=torch.int64, device=dev)
            self.dc_max_bits = torch.zeros(max_iters, dtype=torch.int32, device=dev)
            f.d_image_fx = self.d_image_fx.data_ptr()
            f.dc_max_bits = self.dc_max_bits.data_ptr()
        if self._taps is not None:
            f.taps_image = self._taps.struct()
        self.fit = f
        self.launches = 0

    def set_kspace(self, k_in: torch.Tensor) -> None:
        self.k_in.copy_(torch.view_as_real(k_in.to(torch.complex64)))

    @staticmethod
    def _load(dst: torch.Tensor, src: torch.Tensor, k: int, perm: Optional[torch.Tensor]) -> None:
        """Reference-layout INR parameters [W1 | W2 | table] -> the engine's (permuted-row) storage."""
        if perm is None:
            dst.copy_(src)
            return
        dst[:k].copy_(src[:k])
        dst[k:].view(-1, 2).index_copy_(0, perm, src[k:].to(dst.device).view(-1, 2))

    @staticmethod
    def _unload(src: torch.Tensor, k: int, perm: Optional[torch.Tensor]) -> torch.Tensor:
        if perm is None:
            return src.clone()
        return torch.cat([src[:k], src[k:].view(-1, 2)[perm].reshape(-1)])

    def _load_motion(self, motion_params: torch.Tensor) -> None:
        self._load(self.params[: self.n_motion], motion_params, self._n_mlp_motion, self._perm)

    def _load_image(self, image_params: torch.Tensor) -> None:
        self._load(self.params[self.n_motion:], image_params, self._n_mlp_image,
                   None if self._taps is None else self._taps.perm)

    def motion_params(self) -> torch.Tensor:
        """The motion INR's parameters in the reference layout (a copy)."""
        return self._unload(self.params[: self.n_motion], self._n_mlp_motion, self._perm)

    def image_params(self) -> torch.Tensor:
        """The image INR's parameters in the reference layout (a copy)."""
        return self._unload(self.params[self.n_motion:], self._n_mlp_image,
                            None if self._taps is None else self._taps.perm)

    def reset(self, image_params: torch.Tensor, motion_params: torch.Tensor) -> None:
        """Fresh instance: initial INR parameters, zero gradients / Adam moments / loss trace."""
        self._load_motion(motion_params)
        self._load_image(image_params)
        self.state.zero_()
        self.loss.zero_()
        if self.deterministic:
            self.d_image_fx.zero_()

    def run(self, lambdas: List[float], learning_rate: float, it_begin: int = 0,
            it_end: Optional[int] = None, profile=None, profile_every: int = 0) -> None:
        it_end = len(lambdas) if it_end is None else it_end
        if it_end > self.max_iters:
            raise ValueError("more iterations than the loss buffer holds")
        self.fit.lr = float(learning_rate)
        lam = (C.c_float * len(lambdas))(*[float(v) for v in lambdas])
        dev = self.model.device
        with torch.cuda.device(dev):        # the library keys its auxiliary streams on the current device
            nat.check(nat.lib().immoco_fit_run(C.byref(self.fit), it_begin, it_end, lam,
                                               torch.cuda.current_stream(dev).cuda_stream, profile, profile_every),
                      "fit_run")
        self.launches += (it_end - it_begin) * nat.lib().immoco_launches_per_iteration_mode(
            self.fit.m, int(self.deterministic), int(self.fuse_adam))

    def loss_trace(self, lambdas: List[float]) -> np.ndarray:
        """fp32 loss of every iteration: mse + fp32(lambda) * GE, assembled like the reference
        (F.mse_loss mean over 2HW reals; GE.mul(lambda_ge) with a python-float lambda)."""
        acc = self.loss[: len(lambdas)].cpu().numpy()
        n = 2.0 * self.fit.h * self.fit.w
        dc = (acc[:, 0] / n).astype(np.float32)
        ge = acc[:, 1].astype(np.float32)
        lam = np.asarray(lambdas, dtype=np.float64).astype(np.float32)
        return dc + lam * ge

    def write_back(self) -> None:
        with torch.no_grad():
            self.model.motion_inr.params.copy_(self.motion_params())
            self.model.image_inr.params.copy_(self.image_params())


def run_batched(engines, lambdas: List[float], learning_rate: float, it_begin: int = 0,
                it_end: Optional[int] = None, profile=None, profile_every: int = 0) -> None:
    """Iterations [it_begin, it_end) of SEVERAL fits of one shape in lock step (``immoco_fit_run_batched``): the
    latency-bound kernels of an iteration are issued once for all instances, the GPU-filling ones per instance.
    Every engine ends up exactly where its own ``run`` would have put it (bit for bit in deterministic mode)."""
    engines = list(engines)
    if not engines:
        return
    if len(engines) == 1:
        engines[0].run(lambdas, learning_rate, it_begin, it_end, profile, profile_every)
        return
    lib = nat.lib()
    if len(engines) > lib.immoco_max_fit_batch():
        raise ValueError(f"at most {lib.immoco_max_fit_batch()} fits per batch")
    it_end = len(lambdas) if it_end is None else it_end
    e0 = engines[0]
    for e in engines:
        if it_end > e.max_iters:
            raise ValueError("more iterations than the loss buffer holds")
        if (e.fit.h, e.fit.w, e.fit.m, e.deterministic, e.fuse_adam, e.model.device) != \
                (e0.fit.h, e0.fit.w, e0.fit.m, e0.deterministic, e0.fuse_adam, e0.model.device):
            raise ValueError("batched fits must share shape, movement-group count, device and accumulation mode")
        e.fit.lr = float(learning_rate)
    lam = (C.c_float * len(lambdas))(*[float(v) for v in lambdas])
    fits = (C.POINTER(nat.Fit) * len(engines))(*[C.pointer(e.fit) for e in engines])
    dev = e0.model.device
    with torch.cuda.device(dev):
        nat.check(lib.immoco_fit_run_batched(fits, len(engines), it_begin, it_end, lam,
                                             torch.cuda.current_stream(dev).cuda_stream, profile, profile_every),
                  "fit_run_batched")
    single = lib.immoco_launches_per_iteration_mode(e0.fit.m, int(e0.deterministic), int(e0.fuse_adam))
    shared = (2 if e0.fit.m > 0 else 1) + 4 + (1 if e0.deterministic else 0)   # MLP fwd, GE, rows x2, colpass(, finalize)
    per_iter = shared + len(engines) * (single - shared)
    e0.launches += (it_end - it_begin) * per_iter


def imcoco_motion_correction(kspace_corr, masks, iters=200, learning_rate=1e-2, lambda_ge=1e-2,
                             debug=False, *, image_params=None, motion_params=None,
                             kmax: float = 16000.0, variant: str = "main", return_trace: bool = False,
                             deterministic: Optional[bool] = None):
    """Fit Image INR + Motion INR to one motion-corrupted k-space (immoco.py:116-206).

    Positional signature and defaults are the reference's.  Returns ``(image_prior,
    kspace_foward_model)`` of the LAST iteration's forward pass, i.e. before the last Adam step
    (SURVEY Q4), as detached complex64 CUDA tensors in the 16000-normalised scale (Q5).
    Keyword-only extras: injected initial parameters (tests), ``kmax``/``variant`` for the
    downstream copy's constants (test_immoco_downstream.py:152,189), ``return_trace``, ``deterministic``
    (bit-reproducible fit; default = ``immoco_get_deterministic()``, see ``FitEngine``).
    """
    if not torch.cuda.is_available():
        raise RuntimeError("imcoco_motion_correction needs a CUDA device (no CPU fallback)")
    masks = masks.cuda()
    model = IMMoCo(masks)
    with torch.no_grad():
        if image_params is not None:
            model.image_inr.params.copy_(image_params.to(model.device))
        if motion_params is not None:
            model.motion_inr.params.copy_(motion_params.to(model.device))
    kspace_corr = kspace_corr.to(model.device)
    scale = kspace_corr.abs().max()
    kspace_input = kspace_corr.div(scale).mul(kmax).clone().detach()
    if debug:
        print(f"Scale: {scale:.4f}")
        print(f"Kspace input: {kspace_input.abs().min().item():.4f}, {kspace_input.abs().max().item():.4f}")
    lambdas = lambda_schedule(iters, lambda_ge, variant)
    engine = FitEngine(model, max(iters, 1), deterministic=deterministic)
    engine.set_kspace(kspace_input)
    engine.run(lambdas, learning_rate)
    image_prior = torch.view_as_complex(engine.image.clone())
    kspace_foward_model = torch.view_as_complex(engine.k_out.clone())
    trace = engine.loss_trace(lambdas) if (return_trace or debug) else None
    if debug:
        for j in range(0, iters, 20):
            print(f"iter: {j}, DC_Loss: {trace[j]:.4f}")
    # The reference calls torch.cuda.empty_cache() here (immoco.py:203).  Releasing ~1.4 GB of
    # cached blocks costs 0.2-0.9 s per slice on B200 (profiles/round1_v1_e2e_breakdown.txt) and
    # buys nothing when the next slice re-allocates the same buffers, so the blocks stay with the
    # caching allocator; wrap the call in ``ClearCache()`` to get the reference behaviour.
    del engine, model
    if return_trace:
        return image_prior, kspace_foward_model, trace
    return image_prior, kspace_foward_model
